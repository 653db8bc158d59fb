// Spectral normalisation of a convolution weight in three small kernels ("a fused
// spectral-norm power iteration").
//
// torch.nn.utils.spectral_norm (layers/snffc/snffc.py:8, 23-33; fgan_complete.py:147-156) runs, before every forward,
//     v = normalize(W^T u);  u = normalize(W v)        (one power iteration, training mode only, no gradient)
//     sigma = u . (W v);     weight = W / sigma
// on W = weight_orig viewed as (h = out channels, w = the rest).  As PyTorch ops that is ~14 launches of a few
// microseconds per layer and forward (3 discriminator forwards x 7-9 layers per GAN step); here it is
//     SnColKernel    t = W^T u                       (column sums, split over rows with atomics)
//     SnRowKernel    v = t / max(|t|, eps);  s = W v  (one 32-thread group per row)
//     SnScaleKernel  u = s / max(|s|, eps);  sigma = u . s;  weight = W / sigma
// The matrix (<= 16 MB) stays in L2 between the three passes.  Without a power iteration (eval mode) the first kernel
// is skipped and u, v are read as they are.
#include "ffc_common.cuh"

struct SnParams {
    const float* W;        // (h, w) row-major
    float* u;              // (h)   in/out
    float* v;              // (w)   in/out
    float* u_save;         // (h)   copy of the u used for sigma (for the backward), or null
    float* v_save;         // (w)
    float* w_eff;          // (h, w)
    float* sigma;          // (1)
    float* t;              // workspace (w)
    float* s;              // workspace (h)
    int h, w, power_iteration, rows_per_split;
    int kk;                // 0: W is (h, w) row-major (SpectralNorm.dim == 0, nn.Conv2d / nn.Linear);
                           // > 0: the weight is stored (w / kk, h, kk) (dim == 1, nn.ConvTranspose2d with kk = k * k)
    float eps;
};

// element (i, j) of the matrix view
FFC_HD size_t sn_at(const SnParams& p, int i, int j) {
    if (p.kk == 0) return (size_t)i * p.w + j;
    return ((size_t)(j / p.kk) * p.h + i) * p.kk + j % p.kk;
}

static constexpr int SN_THREADS = 256;

// t[j] += sum over the rows of this split of W[i][j] * u[i]
struct SnColKernel {
    typedef SnParams Params;
    static constexpr int kThreads = SN_THREADS;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const int j = ctx.bx * kThreads + tid;
            if (j < p.w) {
                const int i0 = ctx.by * p.rows_per_split;
                const int i1 = (i0 + p.rows_per_split) < p.h ? (i0 + p.rows_per_split) : p.h;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                int i = i0;
                // column j of the matrix view: rows are `rs` floats apart
                const size_t rs = p.kk == 0 ? (size_t)p.w : (size_t)p.kk;
                const float* col = p.W + sn_at(p, 0, j);
                for (; i + 4 <= i1; i += 4) {
                    const float* r = col + (size_t)i * rs;
                    a0 = fmaf(FFC_LDG(r), FFC_LDG(p.u + i), a0);
                    a1 = fmaf(FFC_LDG(r + rs), FFC_LDG(p.u + i + 1), a1);
                    a2 = fmaf(FFC_LDG(r + 2 * rs), FFC_LDG(p.u + i + 2), a2);
                    a3 = fmaf(FFC_LDG(r + 3 * rs), FFC_LDG(p.u + i + 3), a3);
                }
                for (; i < i1; ++i) a0 = fmaf(FFC_LDG(col + (size_t)i * rs), FFC_LDG(p.u + i), a0);
                ffc_atomic_add(p.t + j, (a0 + a1) + (a2 + a3));
            }
        } FFC_SYNC;
    }
};

// every CTA: inv = 1 / max(|t|, eps) (recomputed per CTA, t is <= 32 KB in L2); rows bx*8 .. +7: s[r] = W[r] . v
struct SnRowKernel {
    typedef SnParams Params;
    static constexpr int kThreads = SN_THREADS;
    static constexpr int kRows = SN_THREADS / 32;
    static size_t smem_bytes() { return (size_t)SN_THREADS * sizeof(double); }
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        double* red = reinterpret_cast<double*>(smem);
        const float* vin = p.power_iteration ? p.t : p.v;
        FFC_PHASE {
            double a = 0.0;
            if (p.power_iteration)
                for (int j = tid; j < p.w; j += kThreads) { const float x = FFC_LDG(p.t + j); a += (double)x * x; }
            red[tid] = a;
        } FFC_SYNC;
        for (int wd = kThreads / 2; wd >= 1; wd >>= 1) {
            FFC_PHASE { if (tid < wd) red[tid] += red[tid + wd]; } FFC_SYNC;
        }
        FFC_TLS(float, inv);
        FFC_PHASE {
            FFC_TLS_REF(float, inv);
            inv = 1.f;
            if (p.power_iteration) { const float n = sqrtf((float)red[0]); inv = 1.f / (n > p.eps ? n : p.eps); }
        } FFC_SYNC;
        FFC_PHASE {
            FFC_TLS_REF(float, inv);
            // this CTA's slice of v (normalised t) goes back to the caller's buffer and to the saved copy
            const int per = (p.w + ctx.gx - 1) / ctx.gx;
            const int j0 = ctx.bx * per, j1 = (j0 + per) < p.w ? (j0 + per) : p.w;
            for (int j = j0 + tid; j < j1; j += kThreads) {
                const float x = FFC_LDG(vin + j) * inv;
                if (p.power_iteration) p.v[j] = x;
                if (p.v_save) p.v_save[j] = x;
            }
            const int r = ctx.bx * kRows + tid / 32, lane = tid % 32;
            float a0 = 0.f, a1 = 0.f;
            if (r < p.h && p.kk == 0) {
                const float* row = p.W + (size_t)r * p.w;
                int j = lane;
                for (; j + 32 < p.w; j += 64) {
                    a0 = fmaf(FFC_LDG(row + j), FFC_LDG(vin + j) * inv, a0);
                    a1 = fmaf(FFC_LDG(row + j + 32), FFC_LDG(vin + j + 32) * inv, a1);
                }
                for (; j < p.w; j += 32) a0 = fmaf(FFC_LDG(row + j), FFC_LDG(vin + j) * inv, a0);
            } else if (r < p.h) {
                for (int j = lane; j < p.w; j += 32) a0 = fmaf(FFC_LDG(p.W + sn_at(p, r, j)), FFC_LDG(vin + j) * inv, a0);
            }
            red[tid] = (double)(a0 + a1);
        } FFC_SYNC;
        for (int wd = 16; wd >= 1; wd >>= 1) {
            FFC_PHASE { if (tid % 32 < wd) red[tid] += red[tid + wd]; } FFC_SYNC;
        }
        FFC_PHASE {
            const int r = ctx.bx * kRows + tid / 32;
            if (tid % 32 == 0 && r < p.h) p.s[r] = (float)red[tid];
        } FFC_SYNC;
    }
};

// every CTA: sigma from s (and u); CTA 0 publishes u / sigma; all: w_eff = W / sigma
struct SnScaleKernel {
    typedef SnParams Params;
    static constexpr int kThreads = SN_THREADS;
    static size_t smem_bytes() { return (size_t)SN_THREADS * sizeof(double); }
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        double* red = reinterpret_cast<double*>(smem);
        FFC_PHASE {
            double a = 0.0;
            for (int i = tid; i < p.h; i += kThreads) {
                const float x = FFC_LDG(p.s + i);
                a += p.power_iteration ? (double)x * x : (double)x * FFC_LDG(p.u + i);
            }
            red[tid] = a;
        } FFC_SYNC;
        for (int wd = kThreads / 2; wd >= 1; wd >>= 1) {
            FFC_PHASE { if (tid < wd) red[tid] += red[tid + wd]; } FFC_SYNC;
        }
        FFC_TLS(float, sg);
        FFC_PHASE {
            FFC_TLS_REF(float, sg);
            float inv = 1.f;
            if (p.power_iteration) {
                const float n = sqrtf((float)red[0]);
                inv = 1.f / (n > p.eps ? n : p.eps);
                sg = (float)red[0] * inv;                    // u . s with u = s * inv
            } else {
                sg = (float)red[0];
            }
            if (ctx.bx == 0) {
                for (int i = tid; i < p.h; i += kThreads) {
                    const float x = p.power_iteration ? FFC_LDG(p.s + i) * inv : FFC_LDG(p.u + i);
                    if (p.u_save) p.u_save[i] = x;
                    if (p.power_iteration) p.u[i] = x;
                }
                if (tid == 0) p.sigma[0] = sg;
            }
        } FFC_SYNC;
        FFC_PHASE {
            FFC_TLS_REF(float, sg);
            const long long n = (long long)p.h * p.w;
            for (long long i = (long long)ctx.bx * kThreads + tid; i < n; i += (long long)ctx.gx * kThreads)
                p.w_eff[i] = FFC_LDG(p.W + i) / sg;
        } FFC_SYNC;
    }
};

extern "C" size_t ffc_spectral_norm_workspace_bytes(int h, int w) { return (size_t)(h + w) * sizeof(float) + 64; }

// W = weight_orig viewed as (h = out channels, w = the rest): row-major (h, w) when kk == 0 (SpectralNorm.dim == 0:
// nn.Conv2d, nn.Linear), stored as (w / kk, h, kk) when kk = k * k > 0 (dim == 1: nn.ConvTranspose2d, whose weight is
// (in, out, k, k) and whose matrix view is weight.permute(1, 0, 2, 3).reshape(out, -1)).  u (h), v (w): the module's
// weight_u / weight_v buffers, updated in place when power_iteration != 0.  u_save / v_save (nullable) receive the
// vectors sigma was computed with, sigma (1 float) the value itself: ffc consumers keep them for the backward
//     dW = g / sigma - (sum(g * W) / sigma^2) * u v^T.
extern "C" int ffc_spectral_norm_fwd(const float* w_orig, float* u, float* v, float* u_save, float* v_save,
                                     float* w_eff, float* sigma, int h, int w, int kk, int power_iteration, float eps,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(w_orig && u && v && w_eff && sigma, "ffc_spectral_norm_fwd: null pointer");
    FFC_REQUIRE(h > 0 && w > 0 && kk >= 0 && (kk == 0 || w % kk == 0), "ffc_spectral_norm_fwd: bad sizes");
    FFC_REQUIRE(workspace && workspace_bytes >= ffc_spectral_norm_workspace_bytes(h, w), "ffc_spectral_norm_fwd: workspace too small");
    ffc_stream_t st = (ffc_stream_t)stream;
    SnParams p;
    p.W = w_orig; p.u = u; p.v = v; p.u_save = u_save; p.v_save = v_save; p.w_eff = w_eff; p.sigma = sigma;
    p.t = (float*)(((uintptr_t)workspace + 15) & ~(uintptr_t)15); p.s = p.t + w;
    p.h = h; p.w = w; p.kk = kk; p.power_iteration = power_iteration ? 1 : 0; p.eps = eps;
    const int colblocks = ffc_cdiv(w, SN_THREADS);
    int rsplit = ffc_cdiv(2 * ffc_sm_count(), colblocks);
    if (rsplit > ffc_cdiv(h, 16)) rsplit = ffc_cdiv(h, 16);
    if (rsplit < 1) rsplit = 1;
    p.rows_per_split = ffc_cdiv(h, rsplit);
    rsplit = ffc_cdiv(h, p.rows_per_split);
    if (p.power_iteration) {
        FFC_CHECK(ffc_memset_async(p.t, 0, (size_t)w * sizeof(float), st));
        FFC_CHECK((ffc_launch<SnColKernel>(colblocks, rsplit, 1, SN_THREADS, 0, st, p)));
    }
    FFC_CHECK((ffc_launch<SnRowKernel>(ffc_cdiv(h, SnRowKernel::kRows), 1, 1, SN_THREADS, SnRowKernel::smem_bytes(), st, p)));
    long long items = ((long long)h * w + SN_THREADS - 1) / SN_THREADS;
    if (items > ffc_sm_count() * 8) items = ffc_sm_count() * 8;
    return ffc_launch<SnScaleKernel>((int)items, 1, 1, SN_THREADS, SnScaleKernel::smem_bytes(), st, p);
}


// ---------------------------------------------------------------------------------------------------------------------
// backward through weight = W / sigma, sigma = u^T W v (u, v constants of the forward's power iteration):
//     dW = g / sigma - (sum(g * W) / sigma^2) * u v^T          (torch.nn.utils.spectral_norm's autograd, restated)
// two kernels: the dot product (double atomics), then the element-wise combination.
// ---------------------------------------------------------------------------------------------------------------------
struct SnBwdParams {
    const float* g; const float* W; const float* u; const float* v; const float* sigma;
    float* dW; double* dot;
    int h, w, kk;
};
struct SnDotKernel {
    typedef SnBwdParams Params;
    static constexpr int kThreads = SN_THREADS;
    static size_t smem_bytes() { return (size_t)SN_THREADS * sizeof(double); }
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        double* red = reinterpret_cast<double*>(smem);
        FFC_PHASE {
            const long long n = (long long)p.h * p.w;
            float a0 = 0.f, a1 = 0.f;
            long long i = (long long)ctx.bx * kThreads + tid;
            const long long step = (long long)ctx.gx * kThreads;
            for (; i + step < n; i += 2 * step) {
                a0 = fmaf(FFC_LDG(p.g + i), FFC_LDG(p.W + i), a0);
                a1 = fmaf(FFC_LDG(p.g + i + step), FFC_LDG(p.W + i + step), a1);
            }
            if (i < n) a0 = fmaf(FFC_LDG(p.g + i), FFC_LDG(p.W + i), a0);
            red[tid] = (double)a0 + (double)a1;
        } FFC_SYNC;
        for (int s = kThreads / 2; s > 0; s >>= 1) {
            FFC_PHASE { if (tid < s) red[tid] += red[tid + s]; } FFC_SYNC;
        }
        FFC_PHASE { if (tid == 0) ffc_atomic_add(p.dot, red[0]); } FFC_SYNC;
    }
};
struct SnBwdKernel {
    typedef SnBwdParams Params;
    static constexpr int kThreads = SN_THREADS;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const float sg = FFC_LDG(p.sigma);
            const float inv = 1.0f / sg;
            const float coef = (float)(p.dot[0] / ((double)sg * (double)sg));
            const long long n = (long long)p.h * p.w;
            for (long long e = (long long)ctx.bx * kThreads + tid; e < n; e += (long long)ctx.gx * kThreads) {
                int i, j;                                   // storage index e -> element (i, j) of the matrix view
                if (p.kk == 0) { i = (int)(e / p.w); j = (int)(e % p.w); }
                else { const int r = (int)(e % p.kk); const long long q = e / p.kk; i = (int)(q % p.h); j = (int)(q / p.h) * p.kk + r; }
                p.dW[e] = FFC_LDG(p.g + e) * inv - coef * FFC_LDG(p.u + i) * FFC_LDG(p.v + j);
            }
        } FFC_SYNC;
    }
};

// g, w_orig, dw: the weight tensor's own storage order (same kk convention as the forward); u (h), v (w), sigma (1): what
// ffc_spectral_norm_fwd saved.  workspace >= 8 bytes.
extern "C" int ffc_spectral_norm_bwd(const float* g, const float* w_orig, const float* u, const float* v, const float* sigma,
                                     float* dw, int h, int w, int kk, void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(g && w_orig && u && v && sigma && dw, "ffc_spectral_norm_bwd: null pointer");
    FFC_REQUIRE(h > 0 && w > 0 && kk >= 0 && (kk == 0 || w % kk == 0), "ffc_spectral_norm_bwd: bad sizes");
    FFC_REQUIRE(workspace && workspace_bytes >= 16, "ffc_spectral_norm_bwd: workspace too small");
    ffc_stream_t st = (ffc_stream_t)stream;
    SnBwdParams p;
    p.g = g; p.W = w_orig; p.u = u; p.v = v; p.sigma = sigma; p.dW = dw; p.h = h; p.w = w; p.kk = kk;
    p.dot = (double*)(((uintptr_t)workspace + 7) & ~(uintptr_t)7);
    FFC_CHECK(ffc_memset_async(p.dot, 0, sizeof(double), st));
    long long items = ((long long)h * w + 2 * SN_THREADS - 1) / (2 * SN_THREADS);
    if (items > ffc_sm_count() * 4) items = ffc_sm_count() * 4;
    if (items < 1) items = 1;
    FFC_CHECK((ffc_launch<SnDotKernel>((int)items, 1, 1, SN_THREADS, SnDotKernel::smem_bytes(), st, p)));
    long long items2 = ((long long)h * w + SN_THREADS - 1) / SN_THREADS;
    if (items2 > ffc_sm_count() * 8) items2 = ffc_sm_count() * 8;
    return ffc_launch<SnBwdKernel>((int)items2, 1, 1, SN_THREADS, 0, st, p);
}
