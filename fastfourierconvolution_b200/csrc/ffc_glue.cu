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
    // ~8 CTAs per SM: a CTA's loop is a chain of dependent-latency loads (two independent accumulators), so the kernel is
    // latency bound and wants many short loops rather than few long ones (26 -> ~8 us on the 192-channel 8x8 layers)
    int ns = (8 * ffc_sm_count() + C - 1) / C;
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

// ---------------------------------------------------------------------------------------------------------------------
// Small FP32 matrix products: the Linear stem of the generators (fgan_complete.py:92-95, 117-119), the SN Linear head of
// the discriminators (:160-170) and their gradients.  C[m][n] = sum_k A(m, k) * B(k, n) [+ bias[n]] with arbitrary
// element strides, so x W^T, dy^T x and dy W are the same kernel.  64 x 64 tile per CTA, 16-deep chunks staged k-major in
// shared memory, 4 x 4 outputs per thread; skinny problems split K over gridDim.z (float atomics into a zeroed C).
// ---------------------------------------------------------------------------------------------------------------------
struct GemmParams {
    const float* A; const float* B; const float* bias; float* C;
    int M, N, K, kper;                       // kper: K range per gridDim.z slice
    long long sam, sak, sbk, sbn, scm, scn;  // element strides
};
struct GemmAcc { float v[16]; };
struct GemmKernel {
    typedef GemmParams Params;
    static constexpr int kThreads = 256, BK = 16, LD = 68;
    static size_t smem_bytes() { return (size_t)2 * BK * LD * 4; }
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        float* As = smem;                // [BK][LD]: A(m0 + i, k) at As[k][i]
        float* Bs = smem + BK * LD;      // [BK][LD]: B(k, n0 + i) at Bs[k][i]
        const int m0 = ctx.bx * 64, n0 = ctx.by * 64;
        const int kbeg = ctx.bz * p.kper, kend = (kbeg + p.kper) < p.K ? (kbeg + p.kper) : p.K;
        FFC_TLS(GemmAcc, acc);
        FFC_PHASE {
            FFC_TLS_REF(GemmAcc, acc);
            (void)tid;
            FFC_UNROLL
            for (int i = 0; i < 16; ++i) acc.v[i] = 0.f;
        } FFC_SYNC;
        for (int k0 = kbeg; k0 < kend; k0 += BK) {
            FFC_PHASE {
                // the faster-varying index of the thread -> element map follows the unit stride of each operand
                for (int idx = tid; idx < 64 * BK; idx += ctx.nt) {
                    int i, k;
                    if (p.sak == 1) { k = idx % BK; i = idx / BK; } else { i = idx % 64; k = idx / 64; }
                    const int m = m0 + i, kk = k0 + k;
                    As[k * LD + i] = (m < p.M && kk < kend) ? FFC_LDG(p.A + (long long)m * p.sam + (long long)kk * p.sak) : 0.f;
                }
                for (int idx = tid; idx < 64 * BK; idx += ctx.nt) {
                    int i, k;
                    if (p.sbk == 1) { k = idx % BK; i = idx / BK; } else { i = idx % 64; k = idx / 64; }
                    const int n = n0 + i, kk = k0 + k;
                    Bs[k * LD + i] = (n < p.N && kk < kend) ? FFC_LDG(p.B + (long long)kk * p.sbk + (long long)n * p.sbn) : 0.f;
                }
            } FFC_SYNC;
            FFC_PHASE {
                FFC_TLS_REF(GemmAcc, acc);
                // consecutive threads run along the unit stride of C, so that a warp's stores are whole 256-byte runs
                // (with m fastest and a row-major C every thread wrote 16 bytes of its own row: 40 us for the 8.4 MB stem output)
                const int tm = p.scn == 1 ? tid / 16 : tid % 16, tn = p.scn == 1 ? tid % 16 : tid / 16;
                FFC_UNROLL
                for (int k = 0; k < BK; ++k) {
                    const float4 a = *reinterpret_cast<const float4*>(As + k * LD + 4 * tm);
                    const float4 b = *reinterpret_cast<const float4*>(Bs + k * LD + 4 * tn);
                    acc.v[0] = fmaf(a.x, b.x, acc.v[0]); acc.v[1] = fmaf(a.x, b.y, acc.v[1]); acc.v[2] = fmaf(a.x, b.z, acc.v[2]); acc.v[3] = fmaf(a.x, b.w, acc.v[3]);
                    acc.v[4] = fmaf(a.y, b.x, acc.v[4]); acc.v[5] = fmaf(a.y, b.y, acc.v[5]); acc.v[6] = fmaf(a.y, b.z, acc.v[6]); acc.v[7] = fmaf(a.y, b.w, acc.v[7]);
                    acc.v[8] = fmaf(a.z, b.x, acc.v[8]); acc.v[9] = fmaf(a.z, b.y, acc.v[9]); acc.v[10] = fmaf(a.z, b.z, acc.v[10]); acc.v[11] = fmaf(a.z, b.w, acc.v[11]);
                    acc.v[12] = fmaf(a.w, b.x, acc.v[12]); acc.v[13] = fmaf(a.w, b.y, acc.v[13]); acc.v[14] = fmaf(a.w, b.z, acc.v[14]); acc.v[15] = fmaf(a.w, b.w, acc.v[15]);
                }
            } FFC_SYNC;
        }
        FFC_PHASE {
            FFC_TLS_REF(GemmAcc, acc);
            const int tm = p.scn == 1 ? tid / 16 : tid % 16, tn = p.scn == 1 ? tid % 16 : tid / 16;
            FFC_UNROLL
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + 4 * tm + i;
                if (p.scn == 1 && ctx.gz == 1 && m < p.M && n0 + 4 * tn + 3 < p.N &&
                    (((uintptr_t)(p.C + (long long)m * p.scm + n0 + 4 * tn)) & 15) == 0) {        // one 16-byte store per row
                    float4 r = make_float4(acc.v[4 * i], acc.v[4 * i + 1], acc.v[4 * i + 2], acc.v[4 * i + 3]);
                    if (p.bias) {
                        const int n = n0 + 4 * tn;
                        r.x += FFC_LDG(p.bias + n); r.y += FFC_LDG(p.bias + n + 1); r.z += FFC_LDG(p.bias + n + 2); r.w += FFC_LDG(p.bias + n + 3);
                    }
                    *reinterpret_cast<float4*>(p.C + (long long)m * p.scm + n0 + 4 * tn) = r;
                    continue;
                }
                FFC_UNROLL
                for (int l = 0; l < 4; ++l) {
                    const int n = n0 + 4 * tn + l;
                    if (m < p.M && n < p.N) {
                        float r = acc.v[4 * i + l];
                        if (p.bias && ctx.bz == 0) r += FFC_LDG(p.bias + n);
                        float* c = p.C + (long long)m * p.scm + (long long)n * p.scn;
                        if (ctx.gz > 1) ffc_atomic_add(c, r); else *c = r;
                    }
                }
            }
        } FFC_SYNC;
    }
};

// C (M x N, strides scm / scn) = A (M x K, strides sam / sak) * B (K x N, strides sbk / sbn) [+ bias[n]]
extern "C" int ffc_gemm_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                            long long sam, long long sak, long long sbk, long long sbn, long long scm, long long scn, void* stream) {
    FFC_REQUIRE(A && B && C, "ffc_gemm_f32: null pointer");
    FFC_REQUIRE(M >= 0 && N >= 0 && K >= 0, "ffc_gemm_f32: negative size");
    if (M == 0 || N == 0) return FFC_OK;
    ffc_stream_t st = (ffc_stream_t)stream;
    const int gx = ffc_cdiv(M, 64), gy = ffc_cdiv(N, 64);
    int split = 1;
    if ((long long)gx * gy < ffc_sm_count() && K > 256) {            // too few tiles to fill the SMs: split the contraction
        split = (int)(ffc_sm_count() / ((long long)gx * gy));
        const int most = ffc_cdiv(K, 128);
        if (split > most) split = most;
        if (split < 1) split = 1;
    }
    int kper = ffc_cdiv(ffc_cdiv(K > 0 ? K : 1, split), GemmKernel::BK) * GemmKernel::BK;
    split = K > 0 ? ffc_cdiv(K, kper) : 1;
    if (split > 1) {
        FFC_REQUIRE(scn == 1 && scm == N, "ffc_gemm_f32: split-K needs a dense row-major C");
        FFC_CHECK(ffc_memset_async(C, 0, (size_t)M * N * sizeof(float), st));
    }
    GemmParams p{A, B, bias, C, M, N, K, kper, sam, sak, sbk, sbn, scm, scn};
    return ffc_launch<GemmKernel>(gx, gy, split, 256, GemmKernel::smem_bytes(), st, p);
}

// out[n] = sum_m x[m][n] (bias gradient of a Linear layer): one thread per column, coalesced over the columns
struct ColSumParams { const float* x; float* out; int M, N; };
struct ColSumKernel {
    typedef ColSumParams Params;
    static constexpr int kThreads = 128;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const int n = ctx.bx * ctx.nt + tid;
            if (n < p.N) {
                float a0 = 0.f, a1 = 0.f;
                int m = 0;
                for (; m + 1 < p.M; m += 2) { a0 += FFC_LDG(p.x + (size_t)m * p.N + n); a1 += FFC_LDG(p.x + (size_t)(m + 1) * p.N + n); }
                if (m < p.M) a0 += FFC_LDG(p.x + (size_t)m * p.N + n);
                p.out[n] = a0 + a1;
            }
        } FFC_SYNC;
    }
};
extern "C" int ffc_colsum_f32(const float* x, float* out, int M, int N, void* stream) {
    FFC_REQUIRE(x && out && M >= 0 && N >= 0, "ffc_colsum_f32: null pointer / negative size");
    if (N == 0) return FFC_OK;
    ColSumParams p{x, out, M, N};
    return ffc_launch<ColSumKernel>(ffc_cdiv(N, 128), 1, 1, 128, 0, (ffc_stream_t)stream, p);
}

// ---------------------------------------------------------------------------------------------------------------------
// AdamW / Adam over FLAT parameter, gradient and moment buffers (fgan_complete.py:315-319: optim.AdamW(lr 2e-4, betas
// (0.5, 0.999)); sngan_complete.py:247-248: optim.Adam): one kernel per optimiser step whatever the number of parameter
// tensors.  lr and the step count live on the device (the step is replayed from a CUDA graph; the LR schedule writes lr
// between replays).  gscale multiplies the gradient first (1 / world size after a SUM all-reduce).
//   decoupled (AdamW): p *= 1 - lr * wd;       else (Adam): g += wd * p
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// ---------------------------------------------------------------------------------------------------------------------
struct AdamTickParams { float* step; };
struct AdamTickKernel {
    typedef AdamTickParams Params;
    static constexpr int kThreads = 32;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE { if (tid == 0) p.step[0] += 1.0f; } FFC_SYNC;
    }
};
struct AdamParams {
    float* p; const float* g; float* m; float* v; long long n;
    const float* lr; const float* step;
    float b1, b2, eps, wd, gscale; int decoupled;
};
struct AdamKernel {
    typedef AdamParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const float lr = FFC_LDG(p.lr), t = FFC_LDG(p.step);
            const float bc1 = 1.0f - powf(p.b1, t), bc2 = 1.0f - powf(p.b2, t);
            const float step_size = lr / bc1, inv_bc2_sqrt = 1.0f / sqrtf(bc2), decay = 1.0f - lr * p.wd;
            const long long n4 = p.n / 4;
            for (long long i = (long long)ctx.bx * ctx.nt + tid; i < n4; i += (long long)ctx.gx * ctx.nt) {
                float4 w = reinterpret_cast<float4*>(p.p)[i], m = reinterpret_cast<float4*>(p.m)[i], v = reinterpret_cast<float4*>(p.v)[i];
                const float4 g4 = FFC_LDG(reinterpret_cast<const float4*>(p.g) + i);
                float* wf = &w.x; float* mf = &m.x; float* vf = &v.x; const float* gf = &g4.x;
                FFC_UNROLL
                for (int j = 0; j < 4; ++j) {
                    float g = gf[j] * p.gscale;
                    if (p.decoupled) wf[j] *= decay; else g = fmaf(p.wd, wf[j], g);
                    mf[j] = fmaf(p.b1, mf[j], (1.0f - p.b1) * g);
                    vf[j] = fmaf(p.b2, vf[j], (1.0f - p.b2) * g * g);
                    wf[j] -= step_size * mf[j] / (sqrtf(vf[j]) * inv_bc2_sqrt + p.eps);
                }
                reinterpret_cast<float4*>(p.p)[i] = w; reinterpret_cast<float4*>(p.m)[i] = m; reinterpret_cast<float4*>(p.v)[i] = v;
            }
            for (long long i = 4 * n4 + (long long)ctx.bx * ctx.nt + tid; i < p.n; i += (long long)ctx.gx * ctx.nt) {
                float g = FFC_LDG(p.g + i) * p.gscale, w = p.p[i];
                if (p.decoupled) w *= decay; else g = fmaf(p.wd, w, g);
                const float m = fmaf(p.b1, p.m[i], (1.0f - p.b1) * g), v = fmaf(p.b2, p.v[i], (1.0f - p.b2) * g * g);
                p.m[i] = m; p.v[i] = v;
                p.p[i] = w - step_size * m / (sqrtf(v) * inv_bc2_sqrt + p.eps);
            }
        } FFC_SYNC;
    }
};
// One optimiser step over n contiguous FP32 values; step[0] (device) is incremented first and is the t of the bias corrections.
extern "C" int ffc_adam_step(float* p, const float* g, float* m, float* v, long long n, const float* lr, float* step,
                             float beta1, float beta2, float eps, float weight_decay, float grad_scale, int decoupled, void* stream) {
    FFC_REQUIRE(p && g && m && v && lr && step && n >= 0, "ffc_adam_step: null pointer / negative size");
    FFC_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "ffc_adam_step: buffers must be 16-byte aligned");
    ffc_stream_t st = (ffc_stream_t)stream;
    AdamTickParams tp{step};
    FFC_CHECK((ffc_launch<AdamTickKernel>(1, 1, 1, 32, 0, st, tp)));
    if (n == 0) return FFC_OK;
    AdamParams ap{p, g, m, v, n, lr, step, beta1, beta2, eps, weight_decay, grad_scale, decoupled};
    return ffc_launch<AdamKernel>(glue_grid(n / 4 + 1, 256), 1, 1, 256, 0, st, ap);
}

// The same step with the GRADIENTS left where autograd put them: parameters and both moments are flat, the gradient of
// tensor t is read through gptr[t] (null: no gradient this step, the tensor is skipped like torch skips grad == None).
// Block i handles the 4096-element piece blk_piece[i] of tensor blk_tensor[i] (a host-built table, fixed for a model).
// gather_dst != null turns the kernel into the packing pass of a multi-GPU step: gradients are copied into the flat
// buffer the all-reduce runs on (zeros where a tensor has none) and nothing else is touched.
static constexpr int kAdamPiece = 4096;
struct AdamTableParams {
    float* p; float* m; float* v; float* gather_dst;
    const float* const* gptr; const long long* offs; const long long* sizes; const int* blk_tensor; const int* blk_piece;
    const float* lr; const float* step;
    float b1, b2, eps, wd, gscale; int decoupled;
};
// step counts are per tensor, like torch's state["step"]: a tensor without a gradient does not age
struct AdamTableTickParams { float* step; const float* const* gptr; int T; };
struct AdamTableTickKernel {
    typedef AdamTableTickParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const int t = ctx.bx * ctx.nt + tid;
            if (t < p.T && p.gptr[t]) p.step[t] += 1.0f;
        } FFC_SYNC;
    }
};
struct AdamTableKernel {
    typedef AdamTableParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const int t = p.blk_tensor[ctx.bx];
            const long long beg = (long long)p.blk_piece[ctx.bx] * kAdamPiece, n = p.sizes[t], base = p.offs[t];
            const long long end = (beg + kAdamPiece) < n ? (beg + kAdamPiece) : n;
            const float* g = p.gptr[t];
            if (p.gather_dst) {
                for (long long i = beg + tid; i < end; i += ctx.nt) p.gather_dst[base + i] = g ? FFC_LDG(g + i) : 0.f;
            } else if (g) {
                const float lr = FFC_LDG(p.lr), st = FFC_LDG(p.step + t);
                const float bc1 = 1.0f - powf(p.b1, st), bc2 = 1.0f - powf(p.b2, st);
                const float step_size = lr / bc1, inv_bc2_sqrt = 1.0f / sqrtf(bc2), decay = 1.0f - lr * p.wd;
                for (long long i = beg + tid; i < end; i += ctx.nt) {
                    float gi = FFC_LDG(g + i) * p.gscale, w = p.p[base + i];
                    if (p.decoupled) w *= decay; else gi = fmaf(p.wd, w, gi);
                    const float m = fmaf(p.b1, p.m[base + i], (1.0f - p.b1) * gi), v = fmaf(p.b2, p.v[base + i], (1.0f - p.b2) * gi * gi);
                    p.m[base + i] = m; p.v[base + i] = v;
                    p.p[base + i] = w - step_size * m / (sqrtf(v) * inv_bc2_sqrt + p.eps);
                }
            }
        } FFC_SYNC;
    }
};
extern "C" int ffc_adam_step_table(float* p, float* m, float* v, const float* const* gptr, const long long* offs, const long long* sizes,
                                   const int* blk_tensor, const int* blk_piece, int ntensors, int nblocks, const float* lr, float* step,
                                   float beta1, float beta2, float eps, float weight_decay, float grad_scale, int decoupled, void* stream) {
    FFC_REQUIRE(p && m && v && gptr && offs && sizes && blk_tensor && blk_piece && lr && step && nblocks >= 0 && ntensors >= 0, "ffc_adam_step_table: null pointer");
    ffc_stream_t st = (ffc_stream_t)stream;
    if (nblocks == 0 || ntensors == 0) return FFC_OK;
    AdamTableTickParams tp{step, gptr, ntensors};
    FFC_CHECK((ffc_launch<AdamTableTickKernel>(ffc_cdiv(ntensors, 256), 1, 1, 256, 0, st, tp)));
    AdamTableParams ap{p, m, v, nullptr, gptr, offs, sizes, blk_tensor, blk_piece, lr, step, beta1, beta2, eps, weight_decay, grad_scale, decoupled};
    return ffc_launch<AdamTableKernel>(nblocks, 1, 1, 256, 0, st, ap);
}
extern "C" int ffc_gather_table(float* dst, const float* const* gptr, const long long* offs, const long long* sizes,
                                const int* blk_tensor, const int* blk_piece, int nblocks, void* stream) {
    FFC_REQUIRE(dst && gptr && offs && sizes && blk_tensor && blk_piece && nblocks >= 0, "ffc_gather_table: null pointer");
    if (nblocks == 0) return FFC_OK;
    AdamTableParams ap{nullptr, nullptr, nullptr, dst, gptr, offs, sizes, blk_tensor, blk_piece, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, 1.f, 0};
    return ffc_launch<AdamTableKernel>(nblocks, 1, 1, 256, 0, (ffc_stream_t)stream, ap);
}
