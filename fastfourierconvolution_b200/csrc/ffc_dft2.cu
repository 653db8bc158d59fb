// rfft2 / irfft2 for planes the power-of-two kernels do not cover: any H, W in 1..128 (odd, non-square, 48x48 of the
// mg = 6 scripts, fgan_cond_complete.py:325).  Direct DFT by two passes of dot products with twiddle tables in shared
// memory: O(H*W*(H+W)) per plane instead of O(H*W*log) -- the reference accepts every size (torch.fft.rfftn / irfftn,
// layers/ffc/fourier_unity.py:38, 56) and so does this library; the tuned kernels (ffc_fft2.cu, ffc_fu2/3/4.cu) stay the
// path for the 4..128 square power-of-two planes every BASELINE config uses.
// Same contract as Rfft2Kernel / Irfft2Kernel: planar spectrum (nplanes, 2, H, Wf), Wf = W/2 + 1, natural order along u;
// colscale: forward 1 = interior columns (those c2r weights with 2) times 2, inverse 1 = interior columns times 1/2;
// optional BatchNorm + ReLU while the spectrum is loaded.
#include "ffc_common.cuh"

// (cos, sin)(2 pi k / n)
FFC_DEVICE float2 dft2_twiddle(int k, int n) {
#ifdef FFC_EMU
    const double a = 2.0 * 3.14159265358979323846 * (double)k / (double)n;
    return make_float2((float)cos(a), (float)sin(a));
#else
    float s, c;
    sincospif(2.0f * (float)k / (float)n, &s, &c);
    return make_float2(c, s);
#endif
}
FFC_DEVICE bool dft2_interior(int v, int W) { return v != 0 && 2 * v != W; }

struct Dft2Fwd {
    struct Params {
        const float* x;      // (nplanes, H, W)
        float* spec;         // (nplanes, 2, H, Wf)
        int nplanes, H, W, colscale;
        float scale;         // 1/sqrt(H*W)
    };
    static constexpr int kThreads = 256;
    static size_t smem_bytes(int H, int W) { const int Wf = W / 2 + 1; return ((size_t)H * W + 2 * (size_t)H * Wf + 2 * (size_t)(H + W)) * 4 + 16; }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        const int H = p.H, W = p.W, Wf = W / 2 + 1;
        float* xs = smem;
        float2* T = reinterpret_cast<float2*>(smem + ((H * W + 1) & ~1));
        float2* twW = T + H * Wf;
        float2* twH = twW + W;
        for (int plane = ctx.bx; plane < p.nplanes; plane += ctx.gx) {
            FFC_PHASE {
                const float* src = p.x + (size_t)plane * H * W;
                for (int i = tid; i < H * W; i += ctx.nt) xs[i] = FFC_LDG(src + i);
                for (int i = tid; i < W; i += ctx.nt) twW[i] = dft2_twiddle(i, W);
                for (int i = tid; i < H; i += ctx.nt) twH[i] = dft2_twiddle(i, H);
            } FFC_SYNC;
            FFC_PHASE {                                   // rows: T[h][v] = sum_w x[h][w] e^{-2 pi i w v / W}
                for (int it = tid; it < H * Wf; it += ctx.nt) {
                    const int h = it / Wf, v = it % Wf;
                    const float* row = xs + h * W;
                    float re = 0.f, im = 0.f;
                    int k = 0;
                    for (int w = 0; w < W; ++w) {
                        const float2 t = twW[k];
                        re = fmaf(row[w], t.x, re);
                        im = fmaf(-row[w], t.y, im);
                        k += v; if (k >= W) k -= W;
                    }
                    T[it] = make_float2(re, im);
                }
            } FFC_SYNC;
            FFC_PHASE {                                   // columns: S[u][v] = sum_h T[h][v] e^{-2 pi i u h / H}
                float* o = p.spec + (size_t)plane * 2 * H * Wf;
                for (int it = tid; it < H * Wf; it += ctx.nt) {
                    const int u = it / Wf, v = it % Wf;
                    float re = 0.f, im = 0.f;
                    int k = 0;
                    for (int h = 0; h < H; ++h) {
                        const float2 t = twH[k], a = T[h * Wf + v];
                        re = fmaf(a.x, t.x, fmaf(a.y, t.y, re));            // (a.x + i a.y)(c - i s)
                        im = fmaf(a.y, t.x, fmaf(-a.x, t.y, im));
                        k += u; if (k >= H) k -= H;
                    }
                    const float a = (p.colscale && dft2_interior(v, W)) ? 2.0f * p.scale : p.scale;
                    o[it] = re * a;
                    o[H * Wf + it] = im * a;
                }
            } FFC_SYNC;
        }
    }
};

struct Dft2Inv {
    struct Params {
        const float* spec;       // (nplanes, 2, H, Wf)
        const float* residual;   // (nplanes, H, W) or null
        float* out;              // (nplanes, H, W)
        int nplanes, H, W, colscale;
        float scale;
        const float* mean; const float* invstd; const float* gamma; const float* beta; int cout;   // BatchNorm + ReLU on load (null mean: none)
    };
    static constexpr int kThreads = 256;
    static size_t smem_bytes(int H, int W) { const int Wf = W / 2 + 1; return (4 * (size_t)H * Wf + 2 * (size_t)(H + W)) * 4 + 16; }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        const int H = p.H, W = p.W, Wf = W / 2 + 1;
        float2* S = reinterpret_cast<float2*>(smem);
        float2* T = S + H * Wf;
        float2* twW = T + H * Wf;
        float2* twH = twW + W;
        for (int plane = ctx.bx; plane < p.nplanes; plane += ctx.gx) {
            FFC_PHASE {
                const float* src = p.spec + (size_t)plane * 2 * H * Wf;
                float m0 = 0.f, m1 = 0.f, a0 = 1.f, a1 = 1.f, b0 = 0.f, b1 = 0.f;
                if (p.mean) {            // same arithmetic as Irfft2Kernel: relu((x - mean) * (invstd * gamma) + beta)
                    const int c0 = 2 * (plane % p.cout), c1 = c0 + 1;
                    m0 = FFC_LDG(p.mean + c0); a0 = FFC_LDG(p.invstd + c0) * FFC_LDG(p.gamma + c0); b0 = FFC_LDG(p.beta + c0);
                    m1 = FFC_LDG(p.mean + c1); a1 = FFC_LDG(p.invstd + c1) * FFC_LDG(p.gamma + c1); b1 = FFC_LDG(p.beta + c1);
                }
                for (int i = tid; i < H * Wf; i += ctx.nt) {
                    float xr = FFC_LDG(src + i), xi = FFC_LDG(src + H * Wf + i);
                    if (p.mean) {
                        xr = (xr - m0) * a0 + b0; xi = (xi - m1) * a1 + b1;
                        xr = xr > 0.f ? xr : 0.f; xi = xi > 0.f ? xi : 0.f;
                    }
                    // c2r weights folded in: interior columns count twice (once when colscale asks for the r2c adjoint)
                    const int v = i % Wf;
                    const float c = (dft2_interior(v, W) && !p.colscale) ? 2.0f : 1.0f;
                    S[i] = make_float2(xr * c, xi * c);
                }
                for (int i = tid; i < W; i += ctx.nt) twW[i] = dft2_twiddle(i, W);
                for (int i = tid; i < H; i += ctx.nt) twH[i] = dft2_twiddle(i, H);
            } FFC_SYNC;
            FFC_PHASE {                                   // columns: T[h][v] = sum_u S[u][v] e^{+2 pi i u h / H}
                for (int it = tid; it < H * Wf; it += ctx.nt) {
                    const int h = it / Wf, v = it % Wf;
                    float re = 0.f, im = 0.f;
                    int k = 0;
                    for (int u = 0; u < H; ++u) {
                        const float2 t = twH[k], a = S[u * Wf + v];
                        re = fmaf(a.x, t.x, fmaf(-a.y, t.y, re));           // (a.x + i a.y)(c + i s)
                        im = fmaf(a.y, t.x, fmaf(a.x, t.y, im));
                        k += h; if (k >= H) k -= H;
                    }
                    T[it] = make_float2(re, im);
                }
            } FFC_SYNC;
            FFC_PHASE {                                   // rows: out[h][w] = sum_v Re(T[h][v] e^{+2 pi i v w / W})
                const size_t g0 = (size_t)plane * H * W;
                for (int it = tid; it < H * W; it += ctx.nt) {
                    const int h = it / W, w = it % W;
                    const float2* row = T + h * Wf;
                    float acc = 0.f;
                    int k = 0;
                    for (int v = 0; v < Wf; ++v) {
                        const float2 t = twW[k], a = row[v];
                        acc = fmaf(a.x, t.x, fmaf(-a.y, t.y, acc));
                        k += w; if (k >= W) k -= W;
                    }
                    float r = acc * p.scale;
                    if (p.residual) r += FFC_LDG(p.residual + g0 + it);
                    p.out[g0 + it] = r;
                }
            } FFC_SYNC;
        }
    }
};

bool ffc_dft2_supported(int H, int W) { return H >= 1 && W >= 1 && H <= 128 && W <= 128; }

static int dft2_grid(int nplanes) {
    const int cap = 4 * ffc_sm_count();
    return nplanes < cap ? nplanes : cap;
}

int ffc_dft2_fwd(const float* x, float* spec, int nplanes, int H, int W, int colscale, ffc_stream_t st) {
    Dft2Fwd::Params p{x, spec, nplanes, H, W, colscale, 1.0f / sqrtf((float)H * (float)W)};
    return ffc_launch<Dft2Fwd>(dft2_grid(nplanes), 1, 1, Dft2Fwd::kThreads, Dft2Fwd::smem_bytes(H, W), st, p);
}

int ffc_dft2_inv(const float* spec, const float* residual, float* out, int nplanes, int H, int W, int colscale,
                 const float* mean, const float* invstd, const float* gamma, const float* beta, int cout, ffc_stream_t st) {
    Dft2Inv::Params p{spec, residual, out, nplanes, H, W, colscale, 1.0f / sqrtf((float)H * (float)W), mean, invstd, gamma, beta, cout};
    return ffc_launch<Dft2Inv>(dft2_grid(nplanes), 1, 1, Dft2Inv::kThreads, Dft2Inv::smem_bytes(H, W), st, p);
}
