// Stand-alone plane kernels: rfft2 (real NCHW planes -> planar half spectrum) and irfft2.
//
// Replaces, for the general ("spectrum staged through L2") form of the Fourier unit, the
// reference's  torch.fft.rfftn + stack/permute/contiguous/view   (layers/ffc/fourier_unity.py:38-42)
// and          view/permute/contiguous/complex + torch.fft.irfftn (layers/ffc/fourier_unity.py:51-56).
// The spectrum tensor uses the reference's channel convention directly: (B, 2C, H, Wf) float32,
// channel 2c = Re, 2c+1 = Im, so the 1x1 channel mix and the BatchNorm see an ordinary NCHW tensor
// and no layout copy is ever made.  Along u the order is FftSplit<H>::pos() (identity for H <= 32).
#include "ffc_fft2.cuh"

template <int H, int W>
struct Rfft2Kernel {
    struct Params {
        const float* x;      // (nplanes, H, W)
        float* spec;         // (nplanes, 2, H, Wf)
        int nplanes, P;
        int colscale;        // 0: plain rfft2;  1: multiply interior columns (0 < v < W/2) by 2
        float scale;         // 1/sqrt(H*W)
    };
    static constexpr int kThreads = 512;
    typedef Fft2Plan<H, W> PL;
    static size_t smem_bytes(int P) { return (size_t)(2 * P * PL::REGION) * 4 + FFC_TW_N * 8; }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        const int plane0 = ctx.bx * p.P;
        const int np = (p.nplanes - plane0) < p.P ? (p.nplanes - plane0) : p.P;
        float* A = smem;
        float* B = smem + p.P * PL::REGION;
        float2* tw = reinterpret_cast<float2*>(B + p.P * PL::REGION);
        FFC_PHASE {
            for (int i = tid; i < FFC_TW_N; i += ctx.nt) tw[i] = c_tw128[i];
            // batched loads: LDU independent float4 loads are in flight before the first shared-memory store
            constexpr int per = H * (W / 4);
            constexpr int LDU = 8;
            const int total = np * per;
            const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)plane0 * H * W);
            for (int i0 = tid; i0 < total; i0 += ctx.nt * LDU) {
                float4 v[LDU];
                FFC_UNROLL
                for (int u = 0; u < LDU; ++u) {
                    const int i = i0 + u * ctx.nt;
                    if (i < total) v[u] = FFC_LDG(src + i);
                }
                FFC_UNROLL
                for (int u = 0; u < LDU; ++u) {
                    const int i = i0 + u * ctx.nt;
                    if (i < total) {
                        const int pl = i / per, rem = i % per, h = rem / (W / 4), j = rem % (W / 4);
                        *reinterpret_cast<float4*>(A + pl * PL::REGION + h * PL::RS + 4 * j) = v[u];
                    }
                }
            }
        } FFC_SYNC;
        float* S = nullptr;
        FFC_FFT2_FORWARD(H, W, np, A, B, tw, S);
        FFC_PHASE {
            const int per = H * PL::Wf;
            for (int i = tid; i < np * per; i += ctx.nt) {
                const int pl = i / per, rem = i % per, v = rem % PL::Wf;
                const float2 s = reinterpret_cast<const float2*>(S + pl * PL::REGION)[rem];
                const float a = (p.colscale && v != 0 && v != W / 2) ? 2.0f * p.scale : p.scale;
                float* o = p.spec + (size_t)(plane0 + pl) * 2 * per + rem;
                o[0] = s.x * a;
                o[per] = s.y * a;
            }
        } FFC_SYNC;
    }
};

template <int H, int W>
struct Irfft2Kernel {
    struct Params {
        const float* spec;       // (nplanes, 2, H, Wf)
        const float* residual;   // (nplanes, H, W) or null: out = residual + irfft2(spec)
        float* out;              // (nplanes, H, W)
        int nplanes, P;
        int colscale;            // 0: torch c2r semantics;  1: interior columns pre-multiplied by 1/2
        float scale;
        // optional BatchNorm + ReLU applied to the spectrum as it is loaded (fourier_unity.py:49 folded into :51-56):
        // per spectrum channel 2c / 2c+1 of plane (b, c), c = plane % cout;  null mean = none
        const float* mean; const float* invstd; const float* gamma; const float* beta; int cout;
    };
    static constexpr int kThreads = 512;
    typedef Fft2Plan<H, W> PL;
    static size_t smem_bytes(int P) { return (size_t)(2 * P * PL::REGION) * 4 + FFC_TW_N * 8; }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        const int plane0 = ctx.bx * p.P;
        const int np = (p.nplanes - plane0) < p.P ? (p.nplanes - plane0) : p.P;
        float* R0 = smem;
        float* R1 = smem + p.P * PL::REGION;
        float2* tw = reinterpret_cast<float2*>(R1 + p.P * PL::REGION);
        float* S = R0;   // spectrum region
        float* O = R1;   // other region
        FFC_PHASE {
            for (int i = tid; i < FFC_TW_N; i += ctx.nt) tw[i] = c_tw128[i];
            constexpr int per = H * PL::Wf;
            constexpr int LDU = 4;
            const int total = np * per;
            for (int i0 = tid; i0 < total; i0 += ctx.nt * LDU) {
                float re[LDU], im[LDU];
                FFC_UNROLL
                for (int u = 0; u < LDU; ++u) {
                    const int i = i0 + u * ctx.nt;
                    if (i < total) {
                        const float* s = p.spec + (size_t)(plane0 + i / per) * 2 * per + i % per;
                        re[u] = FFC_LDG(s); im[u] = FFC_LDG(s + per);
                    }
                }
                FFC_UNROLL
                for (int u = 0; u < LDU; ++u) {
                    const int i = i0 + u * ctx.nt;
                    if (i < total) {
                        const int pl = i / per, rem = i % per, v = rem % PL::Wf;
                        const float a = (p.colscale && v != 0 && v != W / 2) ? 0.5f : 1.0f;
                        float xr = re[u], xi = im[u];
                        if (p.mean) {            // same arithmetic as BnApplyKernel: relu((x - mean) * (invstd * gamma) + beta)
                            const int c0 = 2 * ((plane0 + pl) % p.cout), c1 = c0 + 1;
                            xr = (xr - FFC_LDG(p.mean + c0)) * (FFC_LDG(p.invstd + c0) * FFC_LDG(p.gamma + c0)) + FFC_LDG(p.beta + c0);
                            xi = (xi - FFC_LDG(p.mean + c1)) * (FFC_LDG(p.invstd + c1) * FFC_LDG(p.gamma + c1)) + FFC_LDG(p.beta + c1);
                            xr = xr > 0.f ? xr : 0.f; xi = xi > 0.f ? xi : 0.f;
                        }
                        reinterpret_cast<float2*>(S + pl * PL::REGION)[rem] = make_float2(xr * a, xi * a);
                    }
                }
            }
        } FFC_SYNC;
        float* R = nullptr;
        FFC_FFT2_INVERSE(H, W, np, S, O, tw, p.scale, R);
        FFC_PHASE {
            constexpr int per = H * (W / 4);
            constexpr int LDU = 8;
            const int total = np * per;
            const size_t g0 = (size_t)plane0 * H * W;
            float4* dst = reinterpret_cast<float4*>(p.out + g0);
            const float4* res = p.residual ? reinterpret_cast<const float4*>(p.residual + g0) : nullptr;
            for (int i0 = tid; i0 < total; i0 += ctx.nt * LDU) {
                float4 q[LDU];
                if (res) {
                    FFC_UNROLL
                    for (int u = 0; u < LDU; ++u) {
                        const int i = i0 + u * ctx.nt;
                        if (i < total) q[u] = FFC_LDG(res + i);
                    }
                }
                FFC_UNROLL
                for (int u = 0; u < LDU; ++u) {
                    const int i = i0 + u * ctx.nt;
                    if (i < total) {
                        const int pl = i / per, rem = i % per, h = rem / (W / 4), j = rem % (W / 4);
                        float4 v = *reinterpret_cast<const float4*>(R + pl * PL::REGION + h * PL::RS + 4 * j);
                        if (res) { v.x += q[u].x; v.y += q[u].y; v.z += q[u].z; v.w += q[u].w; }
                        dst[i] = v;
                    }
                }
            }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int H, int W>
static void fft2_tile_config(int nplanes, int* P, int* nt) {
    typedef Fft2Plan<H, W> PL;
    const size_t per_plane = (size_t)2 * PL::REGION * 4;
    int p = 256 / PL::kMaxItems;
    if (p < 1) p = 1;
    const int pmax_smem = (int)((160 * 1024) / per_plane);
    if (p > pmax_smem) p = pmax_smem;
    if (p < 1) p = 1;
    if (p > nplanes) p = nplanes;
    int t = p * PL::kMaxItems;
    if (t > 512) t = 512;
    t = (t + 31) / 32 * 32;
    if (t < 64) t = 64;
    *P = p; *nt = t;
}

template <int N>
static int rfft2_launch(const float* x, float* spec, int nplanes, int colscale, ffc_stream_t st) {
    typedef Rfft2Kernel<N, N> K;
    int P, nt; fft2_tile_config<N, N>(nplanes, &P, &nt);
    typename K::Params p{x, spec, nplanes, P, colscale, 1.0f / sqrtf((float)N * (float)N)};
    return ffc_launch<K>(ffc_cdiv(nplanes, P), 1, 1, nt, K::smem_bytes(P), st, p);
}
struct IrfftBn { const float* mean; const float* invstd; const float* gamma; const float* beta; int cout; };
template <int N>
static int irfft2_launch(const float* spec, const float* residual, float* out, int nplanes, int colscale, ffc_stream_t st,
                         IrfftBn bn = IrfftBn{nullptr, nullptr, nullptr, nullptr, 1}) {
    typedef Irfft2Kernel<N, N> K;
    int P, nt; fft2_tile_config<N, N>(nplanes, &P, &nt);
    typename K::Params p{spec, residual, out, nplanes, P, colscale, 1.0f / sqrtf((float)N * (float)N),
                         bn.mean, bn.invstd, bn.gamma, bn.beta, bn.cout};
    return ffc_launch<K>(ffc_cdiv(nplanes, P), 1, 1, nt, K::smem_bytes(P), st, p);
}

static bool fft2_supported(int H, int W) { return H == W && ffc_is_pow2(H) && H >= 4 && H <= 128; }

// every other plane up to 128x128 (odd, non-square, 48x48 ...): direct DFT kernels (ffc_dft2.cu), natural order along u
bool ffc_dft2_supported(int H, int W);
int ffc_dft2_fwd(const float* x, float* spec, int nplanes, int H, int W, int colscale, ffc_stream_t st);
int ffc_dft2_inv(const float* spec, const float* residual, float* out, int nplanes, int H, int W, int colscale,
                 const float* mean, const float* invstd, const float* gamma, const float* beta, int cout, ffc_stream_t st);

// 2: tuned power-of-two kernels, 1: direct DFT kernels, 0: unsupported
extern "C" int ffc_fft2_supported(int H, int W) { return fft2_supported(H, W) ? 2 : (ffc_dft2_supported(H, W) ? 1 : 0); }

extern "C" int ffc_rfft2(const float* x, float* spec, int nplanes, int H, int W, int colscale, void* stream) {
    FFC_REQUIRE(x && spec, "ffc_rfft2: null pointer");
    FFC_REQUIRE(nplanes >= 0, "ffc_rfft2: negative plane count");
    if (!fft2_supported(H, W)) {
        FFC_REQUIRE(ffc_dft2_supported(H, W), "ffc_rfft2: unsupported plane %dx%d (1..128 in both dimensions)", H, W);
        return nplanes == 0 ? FFC_OK : ffc_dft2_fwd(x, spec, nplanes, H, W, colscale, (ffc_stream_t)stream);
    }
    FFC_REQUIRE(((uintptr_t)x & 15) == 0, "ffc_rfft2: x must be 16-byte aligned");
    FFC_REQUIRE(nplanes >= 0, "ffc_rfft2: negative plane count");
    if (nplanes == 0) return FFC_OK;
    ffc_stream_t st = (ffc_stream_t)stream;
    switch (H) {
        case 4: return rfft2_launch<4>(x, spec, nplanes, colscale, st);
        case 8: return rfft2_launch<8>(x, spec, nplanes, colscale, st);
        case 16: return rfft2_launch<16>(x, spec, nplanes, colscale, st);
        case 32: return rfft2_launch<32>(x, spec, nplanes, colscale, st);
        case 64: return rfft2_launch<64>(x, spec, nplanes, colscale, st);
        default: return rfft2_launch<128>(x, spec, nplanes, colscale, st);
    }
}

extern "C" int ffc_irfft2(const float* spec, const float* residual, float* out, int nplanes, int H, int W,
                          int colscale, void* stream) {
    FFC_REQUIRE(spec && out, "ffc_irfft2: null pointer");
    FFC_REQUIRE(nplanes >= 0, "ffc_irfft2: negative plane count");
    if (!fft2_supported(H, W)) {
        FFC_REQUIRE(ffc_dft2_supported(H, W), "ffc_irfft2: unsupported plane %dx%d (1..128 in both dimensions)", H, W);
        return nplanes == 0 ? FFC_OK : ffc_dft2_inv(spec, residual, out, nplanes, H, W, colscale, nullptr, nullptr, nullptr, nullptr, 1, (ffc_stream_t)stream);
    }
    FFC_REQUIRE(((uintptr_t)out & 15) == 0 && ((uintptr_t)residual & 15) == 0, "ffc_irfft2: out/residual must be 16-byte aligned");
    FFC_REQUIRE(nplanes >= 0, "ffc_irfft2: negative plane count");
    if (nplanes == 0) return FFC_OK;
    ffc_stream_t st = (ffc_stream_t)stream;
    switch (H) {
        case 4: return irfft2_launch<4>(spec, residual, out, nplanes, colscale, st);
        case 8: return irfft2_launch<8>(spec, residual, out, nplanes, colscale, st);
        case 16: return irfft2_launch<16>(spec, residual, out, nplanes, colscale, st);
        case 32: return irfft2_launch<32>(spec, residual, out, nplanes, colscale, st);
        case 64: return irfft2_launch<64>(spec, residual, out, nplanes, colscale, st);
        default: return irfft2_launch<128>(spec, residual, out, nplanes, colscale, st);
    }
}

// irfft2 of relu(batchnorm(spec)) [+ residual]: the BatchNorm + ReLU of FourierUnitSN.forward (fourier_unity.py:49) applied
// while the spectrum is loaded, so the normalised spectrum never exists in memory (one read + one write of the spectrum
// less than bn_act followed by irfft2).  spec (nplanes, 2, H, W/2+1) with nplanes = B * cout; mean / invstd / gamma / beta
// over the 2*cout spectrum channels (mean / invstd as written by ffc_bn_stats).
extern "C" int ffc_irfft2_bn_relu(const float* spec, const float* residual, float* out, int nplanes, int cout, int H, int W,
                                  const float* mean, const float* invstd, const float* gamma, const float* beta, void* stream) {
    FFC_REQUIRE(spec && out && mean && invstd && gamma && beta, "ffc_irfft2_bn_relu: null pointer");
    if (!fft2_supported(H, W)) {
        FFC_REQUIRE(ffc_dft2_supported(H, W), "ffc_irfft2_bn_relu: unsupported plane %dx%d (1..128 in both dimensions)", H, W);
        FFC_REQUIRE(nplanes >= 0 && cout > 0 && nplanes % cout == 0, "ffc_irfft2_bn_relu: plane count must be a multiple of cout");
        return nplanes == 0 ? FFC_OK : ffc_dft2_inv(spec, residual, out, nplanes, H, W, 0, mean, invstd, gamma, beta, cout, (ffc_stream_t)stream);
    }
    FFC_REQUIRE(((uintptr_t)out & 15) == 0 && ((uintptr_t)residual & 15) == 0, "ffc_irfft2_bn_relu: out/residual must be 16-byte aligned");
    FFC_REQUIRE(nplanes >= 0 && cout > 0 && nplanes % cout == 0, "ffc_irfft2_bn_relu: plane count must be a multiple of cout");
    if (nplanes == 0) return FFC_OK;
    ffc_stream_t st = (ffc_stream_t)stream;
    const IrfftBn bn{mean, invstd, gamma, beta, cout};
    switch (H) {
        case 4: return irfft2_launch<4>(spec, residual, out, nplanes, 0, st, bn);
        case 8: return irfft2_launch<8>(spec, residual, out, nplanes, 0, st, bn);
        case 16: return irfft2_launch<16>(spec, residual, out, nplanes, 0, st, bn);
        case 32: return irfft2_launch<32>(spec, residual, out, nplanes, 0, st, bn);
        case 64: return irfft2_launch<64>(spec, residual, out, nplanes, 0, st, bn);
        default: return irfft2_launch<128>(spec, residual, out, nplanes, 0, st, bn);
    }
}
