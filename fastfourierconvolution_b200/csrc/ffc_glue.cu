// Glue between the FFC layers of the reference's generators, bandwidth-bound, one kernel each:
//   NoiseInjection   out = x + weight[c] * noise[b, 0, h, w]           (layers/noise_injection.py:20-32; fgan_complete.py:122-131)
//                    dweight[c] = sum_{b,h,w} dy * noise                (dx = dy passes through)
//   uint8 epilogue   out = uint8(255 * (clamp(x, lo, hi) * 0.5 + 0.5))  (fgan_complete.py:136-138, eval mode)
#include "ffc_common.cuh"

struct NoiseAddParams {
    const float* x; const float* w; const float* noise; float* out;
    int B, C, HW;
};
struct NoiseAddKernel {
    typedef NoiseAddParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const long long n = (long long)p.B * p.C * p.HW;
            if (p.HW % 4 == 0 && ((((uintptr_t)p.x | (uintptr_t)p.out | (uintptr_t)p.noise) & 15) == 0)) {
                const int q = p.HW / 4;
                const long long n4 = n / 4;
                for (long long i = (long long)ctx.bx * kThreads + tid; i < n4; i += (long long)ctx.gx * kThreads) {
                    const int r = (int)(i % q);
                    const long long bc = i / q;
                    const int c = (int)(bc % p.C), b = (int)(bc / p.C);
                    const float4 v = FFC_LDG(reinterpret_cast<const float4*>(p.x) + i);
                    const float4 z = FFC_LDG(reinterpret_cast<const float4*>(p.noise) + (size_t)b * q + r);
                    const float w = FFC_LDG(p.w + c);
                    reinterpret_cast<float4*>(p.out)[i] = make_float4(fmaf(w, z.x, v.x), fmaf(w, z.y, v.y), fmaf(w, z.z, v.z), fmaf(w, z.w, v.w));
                }
            } else {
                for (long long i = (long long)ctx.bx * kThreads + tid; i < n; i += (long long)ctx.gx * kThreads) {
                    const int r = (int)(i % p.HW);
                    const long long bc = i / p.HW;
                    const int c = (int)(bc % p.C), b = (int)(bc / p.C);
                    p.out[i] = fmaf(FFC_LDG(p.w + c), FFC_LDG(p.noise + (size_t)b * p.HW + r), FFC_LDG(p.x + i));
                }
            }
        } FFC_SYNC;
    }
};

struct NoiseWgradParams {
    const float* dy; const float* noise; float* dw;     // dw zeroed by the host wrapper
    int B, C, HW, nsplit;
};
struct NoiseWgradKernel {
    typedef NoiseWgradParams Params;
    static constexpr int kThreads = 256;
    static size_t smem_bytes() { return (size_t)kThreads * sizeof(double); }
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        double* red = reinterpret_cast<double*>(smem);
        const int c = ctx.bx, split = ctx.by;
        FFC_PHASE {
            const long long total = (long long)p.B * p.HW;
            const long long per = (total + p.nsplit - 1) / p.nsplit;
            const long long beg = split * per, end = (beg + per) < total ? (beg + per) : total;
            float a0 = 0.f, a1 = 0.f;
            long long i = beg + tid;
            for (; i + kThreads < end; i += 2 * kThreads) {
                const long long j = i + kThreads;
                const int b0 = (int)(i / p.HW), r0 = (int)(i % p.HW), b1 = (int)(j / p.HW), r1 = (int)(j % p.HW);
                a0 = fmaf(FFC_LDG(p.dy + ((size_t)b0 * p.C + c) * p.HW + r0), FFC_LDG(p.noise + i), a0);
                a1 = fmaf(FFC_LDG(p.dy + ((size_t)b1 * p.C + c) * p.HW + r1), FFC_LDG(p.noise + j), a1);
            }
            if (i < end) {
                const int b0 = (int)(i / p.HW), r0 = (int)(i % p.HW);
                a0 = fmaf(FFC_LDG(p.dy + ((size_t)b0 * p.C + c) * p.HW + r0), FFC_LDG(p.noise + i), a0);
            }
            red[tid] = (double)a0 + (double)a1;
        } FFC_SYNC;
        for (int s = kThreads / 2; s > 0; s >>= 1) {
            FFC_PHASE { if (tid < s) red[tid] += red[tid + s]; } FFC_SYNC;
        }
        FFC_PHASE { if (tid == 0) ffc_atomic_add(p.dw + c, (float)red[0]); } FFC_SYNC;
    }
};

struct ToU8Params {
    const float* x; unsigned char* out; long long n; float lo, hi;
};
struct ToU8Kernel {
    typedef ToU8Params Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            for (long long i = (long long)ctx.bx * kThreads + tid; i < p.n; i += (long long)ctx.gx * kThreads) {
                float v = FFC_LDG(p.x + i);
                v = v < p.lo ? p.lo : (v > p.hi ? p.hi : v);
                p.out[i] = (unsigned char)(255.0f * (v * 0.5f + 0.5f));          // float -> uint8 truncates, like Tensor.to(torch.uint8)
            }
        } FFC_SYNC;
    }
};

static int glue_grid(long long items, int per_block) {
    long long g = (items + per_block - 1) / per_block;
    const long long cap = (long long)ffc_sm_count() * 8;
    if (g > cap) g = cap;
    return g < 1 ? 1 : (int)g;
}

// out = x + w[c] * noise[b, hw]; x / out (B, C, HW), noise (B, HW), w (C)
extern "C" int ffc_noise_add_fwd(const float* x, const float* w, const float* noise, float* out, int B, int C, int HW, void* stream) {
    FFC_REQUIRE(x && w && noise && out, "ffc_noise_add_fwd: null pointer");
    FFC_REQUIRE(B >= 0 && C > 0 && HW > 0, "ffc_noise_add_fwd: bad sizes");
    if (B == 0) return FFC_OK;
    NoiseAddParams p{x, w, noise, out, B, C, HW};
    return ffc_launch<NoiseAddKernel>(glue_grid((long long)B * C * HW / 4 + 1, 256), 1, 1, 256, 0, (ffc_stream_t)stream, p);
}

// dw[c] = sum over (b, hw) of dy[b, c, hw] * noise[b, hw]
extern "C" int ffc_noise_add_bwd_w(const float* dy, const float* noise, float* dw, int B, int C, int HW, void* stream) {
    FFC_REQUIRE(dy && noise && dw, "ffc_noise_add_bwd_w: null pointer");
    FFC_REQUIRE(B >= 0 && C > 0 && HW > 0, "ffc_noise_add_bwd_w: bad sizes");
    ffc_stream_t st = (ffc_stream_t)stream;
    FFC_CHECK(ffc_memset_async(dw, 0, (size_t)C * sizeof(float), st));
    if (B == 0) return FFC_OK;
    int ns = (2 * ffc_sm_count() + C - 1) / C;
    const long long total = (long long)B * HW;
    if ((long long)ns * 1024 > total) ns = (int)((total + 1023) / 1024);
    if (ns < 1) ns = 1;
    NoiseWgradParams p{dy, noise, dw, B, C, HW, ns};
    return ffc_launch<NoiseWgradKernel>(C, ns, 1, 256, NoiseWgradKernel::smem_bytes(), st, p);
}

// out[i] = uint8(255 * (clamp(x[i], lo, hi) * 0.5 + 0.5)); lo > hi disables the clamp (fgan64 / fgan128 clamp to their own range)
extern "C" int ffc_to_uint8(const float* x, unsigned char* out, long long n, float lo, float hi, void* stream) {
    FFC_REQUIRE(x && out && n >= 0, "ffc_to_uint8: null pointer / negative size");
    if (n == 0) return FFC_OK;
    ToU8Params p{x, out, n, lo, hi};
    if (lo > hi) { p.lo = -3.0e38f; p.hi = 3.0e38f; }
    return ffc_launch<ToU8Kernel>(glue_grid(n, 256 * 4), 1, 1, 256, 0, (ffc_stream_t)stream, p);
}
