// BatchNorm2d (+ activation) forward / backward, bias gradient, and the SE gate of
// SpectralTransform.  Replaces nn.BatchNorm2d + activation at layers/ffc/ffc_bn_act.py:73-81,
// bn1/act1 at layers/ffc/spectral_transform.py:89, bn/relu at layers/ffc/fourier_unity.py:49 and
// SELayer at layers/ffc/spectral_transform.py:12-28.
//
// Batch statistics are reduced in double precision (per-thread double accumulators, shared-memory
// tree, one atomicAdd(double) pair per CTA), so that E[x^2] - E[x]^2 is safe in FP32 data.
#include "ffc_common.cuh"

#define FFC_RED_THREADS 256

// ---------------------------------------------------------------------------------------------
// per-channel sums over (B, HW):  out[c] += sum f0,  out[C + c] += sum f1    (double)
// MODE 0: f0 = x, f1 = x*x                         (BN forward statistics)
// MODE 1: f0 = g, f1 = g * xhat  with g = dy * act'(z)   (BN+act backward reduction)
// MODE 2: f0 = x only                              (bias gradient)
// ---------------------------------------------------------------------------------------------
struct ChanReduceParams {
    const float* x;        // (B, C, HW)  BN input (MODE 0/1) or dy (MODE 2)
    const float* dy;       // MODE 1
    const float* mean;     // MODE 1: [C] or null (identity norm)
    const float* invstd;   // MODE 1
    const float* gamma;    // MODE 1
    const float* beta;     // MODE 1
    double* out;           // [2C] accumulated with atomics (zeroed by the caller), or, with `partial`, [nsplit][2C] written once
    int B, C, HW, act, nsplit;
    float slope;
    int partial;           // 1: every (channel, split) CTA stores its own sums -- no zero fill in front, no atomics; the consumer adds them
};

template <int MODE>
struct ChanReduceKernel {
    typedef ChanReduceParams Params;
    static constexpr int kThreads = FFC_RED_THREADS;
    static size_t smem_bytes() { return (size_t)2 * kThreads * sizeof(double); }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        double* red = reinterpret_cast<double*>(smem);
        const int c = ctx.bx, split = ctx.by;
        const long long total = (long long)p.B * p.HW;
        const long long per = ((total + p.nsplit - 1) / p.nsplit + 3) / 4 * 4;       // multiple of 4: float4 items never straddle splits
        const long long beg = split * per, end = (beg + per) < total ? (beg + per) : total;
        FFC_PHASE {
            double s0 = 0.0, s1 = 0.0;
            float mu = 0.f, is = 1.f, ga = 1.f, be = 0.f;
            if (MODE == 1 && p.mean) { mu = p.mean[c]; is = p.invstd[c]; ga = p.gamma[c]; be = p.beta[c]; }
            const bool vec = (p.HW % 4 == 0) && (((uintptr_t)p.x | (uintptr_t)p.dy) & 15) == 0;
            if (vec) {
                // float4 items of channel c, numbered over (image, quarter-row); the (image, offset) pair of a thread is
                // advanced incrementally (no per-element division); FP32 partial sums of 4 values feed the doubles
                const int q = p.HW / 4;
                const long long beg4 = beg / 4, end4 = (end + 3) / 4 < (total / 4) ? (end + 3) / 4 : total / 4;
                long long i = beg4 + tid;
                int b = (int)(i / q), r = (int)(i % q);
                const int db = kThreads / q, dr = kThreads % q;          // one thread-block stride in (image, offset) form
                // four independent float4 loads per thread and iteration (one load in flight per thread left the kernel
                // latency bound at ~30 % of the HBM rate); out-of-range slots read as zeros, which add nothing
                constexpr int U = 4;
                for (; i < end4; i += (long long)U * kThreads) {
                    float4 v[U], d[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f); d[u] = v[u];
                        if (i + (long long)u * kThreads < end4) {
                            const size_t o4 = ((size_t)b * p.C + c) * q + r;
                            v[u] = FFC_LDG(reinterpret_cast<const float4*>(p.x) + o4);
                            if (MODE == 1) d[u] = FFC_LDG(reinterpret_cast<const float4*>(p.dy) + o4);
                        }
                        b += db; r += dr;
                        if (r >= q) { r -= q; ++b; }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (MODE == 0) {
                            s0 += (double)((v[u].x + v[u].y) + (v[u].z + v[u].w));
                            s1 += (double)(fmaf(v[u].x, v[u].x, v[u].y * v[u].y) + fmaf(v[u].z, v[u].z, v[u].w * v[u].w));
                        } else if (MODE == 2) {
                            s0 += (double)((v[u].x + v[u].y) + (v[u].z + v[u].w));
                        } else {
                            const float x0 = (v[u].x - mu) * is, x1 = (v[u].y - mu) * is, x2 = (v[u].z - mu) * is, x3 = (v[u].w - mu) * is;
                            const float g0 = d[u].x * ffc_act_bwd(x0 * ga + be, p.act, p.slope), g1 = d[u].y * ffc_act_bwd(x1 * ga + be, p.act, p.slope);
                            const float g2 = d[u].z * ffc_act_bwd(x2 * ga + be, p.act, p.slope), g3 = d[u].w * ffc_act_bwd(x3 * ga + be, p.act, p.slope);
                            s0 += (double)((g0 + g1) + (g2 + g3));
                            s1 += (double)(fmaf(g0, x0, g1 * x1) + fmaf(g2, x2, g3 * x3));
                        }
                    }
                }
            } else {
                for (long long i = beg + tid; i < end; i += kThreads) {
                    const int b = (int)(i / p.HW), r = (int)(i % p.HW);
                    const size_t o = ((size_t)b * p.C + c) * p.HW + r;
                    if (MODE == 0) { const float v = FFC_LDG(p.x + o); s0 += v; s1 += (double)v * v; }
                    else if (MODE == 2) { s0 += FFC_LDG(p.x + o); }
                    else {
                        const float xh = (FFC_LDG(p.x + o) - mu) * is;
                        const float z = xh * ga + be;
                        const float g = FFC_LDG(p.dy + o) * ffc_act_bwd(z, p.act, p.slope);
                        s0 += g; s1 += (double)g * xh;
                    }
                }
            }
            red[tid] = s0; red[kThreads + tid] = s1;
        } FFC_SYNC;
        for (int w = kThreads / 2; w >= 1; w >>= 1) {
            FFC_PHASE {
                if (tid < w) { red[tid] += red[tid + w]; red[kThreads + tid] += red[kThreads + tid + w]; }
            } FFC_SYNC;
        }
        FFC_PHASE {
            if (tid == 0) {
                if (p.partial) {
                    double* o = p.out + (size_t)split * 2 * p.C;
                    o[c] = red[0];
                    if (MODE != 2) o[p.C + c] = red[kThreads];
                } else {
                    ffc_atomic_add(p.out + c, red[0]);
                    if (MODE != 2) ffc_atomic_add(p.out + p.C + c, red[kThreads]);
                }
            }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
// finalize: per-channel mean / invstd from the sums, running-stat update (momentum, unbiased var)
// ---------------------------------------------------------------------------------------------
struct BnFinalizeParams {
    const double* sums;    // [nparts][2C] (nparts = 1: the atomically accumulated sums)
    float* mean; float* invstd;           // [C] saved for backward / used by apply
    float* running_mean; float* running_var;   // [C] or null
    int C; double count; float eps, momentum;
    int nparts;
};
struct BnFinalizeKernel {
    typedef BnFinalizeParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const int c = ctx.bx * kThreads + tid;
            if (c < p.C) {
                double s0 = 0.0, s1 = 0.0;
                for (int q = 0; q < p.nparts; ++q) { s0 += p.sums[(size_t)q * 2 * p.C + c]; s1 += p.sums[(size_t)q * 2 * p.C + p.C + c]; }
                const double m = s0 / p.count;
                double var = s1 / p.count - m * m;
                if (var < 0.0) var = 0.0;
                p.mean[c] = (float)m;
                p.invstd[c] = (float)(1.0 / sqrt(var + (double)p.eps));
                if (p.running_mean) {
                    const double unb = p.count > 1.0 ? var * p.count / (p.count - 1.0) : var;
                    p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * (float)m;
                    p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * (float)unb;
                }
            }
        } FFC_SYNC;
    }
};

// eval mode: mean / invstd from the running statistics
struct BnEvalStatsParams { const float* rm; const float* rv; float* m; float* is; int C; float eps; };
struct BnEvalStatsKernel {
    typedef BnEvalStatsParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& q, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const int c = ctx.bx * kThreads + tid;
            if (c < q.C) { q.m[c] = q.rm[c]; q.is[c] = (float)(1.0 / sqrt((double)q.rv[c] + (double)q.eps)); }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
// apply: y = act((x - mean) * invstd * gamma + beta)   (mean == null: y = act(x))
// ---------------------------------------------------------------------------------------------
struct BnApplyParams {
    const float* x; float* y;
    const float* mean; const float* invstd; const float* gamma; const float* beta;   // [C] or null
    int C, HW; long long total; int act; float slope;
};
struct BnApplyKernel {
    typedef BnApplyParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const bool vec = (p.HW % 4 == 0);
            const long long n = vec ? p.total / 4 : p.total;
            for (long long i = (long long)ctx.bx * kThreads + tid; i < n; i += (long long)ctx.gx * kThreads) {
                const int c = vec ? (int)((unsigned long long)i / (unsigned)(p.HW / 4) % (unsigned)p.C)
                                  : (int)((unsigned long long)i / (unsigned)p.HW % (unsigned)p.C);
                float mu = 0.f, a = 1.f, be = 0.f;
                if (p.mean) { mu = FFC_LDG(p.mean + c); a = FFC_LDG(p.invstd + c) * FFC_LDG(p.gamma + c); be = FFC_LDG(p.beta + c); }
                if (vec) {
                    float4 v = FFC_LDG(reinterpret_cast<const float4*>(p.x) + i);
                    v.x = ffc_act_fwd((v.x - mu) * a + be, p.act, p.slope);
                    v.y = ffc_act_fwd((v.y - mu) * a + be, p.act, p.slope);
                    v.z = ffc_act_fwd((v.z - mu) * a + be, p.act, p.slope);
                    v.w = ffc_act_fwd((v.w - mu) * a + be, p.act, p.slope);
                    reinterpret_cast<float4*>(p.y)[i] = v;
                } else {
                    p.y[i] = ffc_act_fwd((FFC_LDG(p.x + i) - mu) * a + be, p.act, p.slope);
                }
            }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
// backward apply:
//   training BN: dx = gamma*invstd * (g - sum_g/N - xhat * sum_gx/N)
//   eval BN    : dx = gamma*invstd * g
//   identity   : dx = g                       with g = dy * act'(z)
// and dgamma[c] = sum_gx, dbeta[c] = sum_g (written by the thread that owns element 0 of (b=0,c)).
// ---------------------------------------------------------------------------------------------
struct BnBwdApplyParams {
    const float* x; const float* dy; float* dx;
    const float* mean; const float* invstd; const float* gamma; const float* beta;  // null: identity norm
    const double* sums;    // [2C] from ChanReduce MODE 1 (null for identity norm)
    float* dgamma; float* dbeta;   // [C] or null
    int C, HW; long long total; int act, training; float slope; double count;
};
struct BnBwdApplyKernel {
    typedef BnBwdApplyParams Params;
    static constexpr int kThreads = 256;
    // gradient of one element; k1 = sum_g / N, k2 = sum_gx / N (0 in eval mode), a = gamma * invstd
    static FFC_DEVICE float one(const Params& p, float xv, float g0, float mu, float is, float ga, float be, float a, float k1, float k2) {
        const float xh = (xv - mu) * is;
        const float g = g0 * ffc_act_bwd(xh * ga + be, p.act, p.slope);
        return a * (g - k1 - xh * k2);
    }
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const bool vec = (p.HW % 4 == 0) && (((uintptr_t)p.x | (uintptr_t)p.dy | (uintptr_t)p.dx) & 15) == 0;
            const long long n = vec ? p.total / 4 : p.total;
            const unsigned per_plane = vec ? (unsigned)(p.HW / 4) : (unsigned)p.HW;
            const float inv_count = (float)(1.0 / p.count);
            for (long long i = (long long)ctx.bx * kThreads + tid; i < n; i += (long long)ctx.gx * kThreads) {
                const unsigned long long plane = (unsigned long long)i / per_plane;
                const int c = (int)(plane % (unsigned)p.C);
                if (!p.mean) {
                    if (vec) {
                        const float4 xv = FFC_LDG(reinterpret_cast<const float4*>(p.x) + i), g = FFC_LDG(reinterpret_cast<const float4*>(p.dy) + i);
                        reinterpret_cast<float4*>(p.dx)[i] = make_float4(g.x * ffc_act_bwd(xv.x, p.act, p.slope), g.y * ffc_act_bwd(xv.y, p.act, p.slope),
                                                                         g.z * ffc_act_bwd(xv.z, p.act, p.slope), g.w * ffc_act_bwd(xv.w, p.act, p.slope));
                    } else {
                        p.dx[i] = FFC_LDG(p.dy + i) * ffc_act_bwd(FFC_LDG(p.x + i), p.act, p.slope);
                    }
                    continue;
                }
                const float mu = FFC_LDG(p.mean + c), is = FFC_LDG(p.invstd + c), ga = FFC_LDG(p.gamma + c), be = FFC_LDG(p.beta + c);
                const double sg = p.sums[c], sgx = p.sums[p.C + c];
                const float k1 = p.training ? (float)sg * inv_count : 0.f, k2 = p.training ? (float)sgx * inv_count : 0.f;
                const float a = ga * is;
                if (vec) {
                    const float4 xv = FFC_LDG(reinterpret_cast<const float4*>(p.x) + i), g = FFC_LDG(reinterpret_cast<const float4*>(p.dy) + i);
                    reinterpret_cast<float4*>(p.dx)[i] = make_float4(one(p, xv.x, g.x, mu, is, ga, be, a, k1, k2), one(p, xv.y, g.y, mu, is, ga, be, a, k1, k2),
                                                                     one(p, xv.z, g.z, mu, is, ga, be, a, k1, k2), one(p, xv.w, g.w, mu, is, ga, be, a, k1, k2));
                } else {
                    p.dx[i] = one(p, FFC_LDG(p.x + i), FFC_LDG(p.dy + i), mu, is, ga, be, a, k1, k2);
                }
                if ((unsigned long long)i == (unsigned long long)c * per_plane) {       // first item of plane (b = 0, c)
                    if (p.dgamma) p.dgamma[c] = (float)sgx;
                    if (p.dbeta) p.dbeta[c] = (float)sg;
                }
            }
        } FFC_SYNC;
    }
};

// double -> float copy of the first n sums (bias gradient)
struct D2FParams { const double* in; float* out; int n; int nparts; int stride; };    // out[i] = sum over the parts of in[part * stride + i]
struct D2FKernel {
    typedef D2FParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const int i = ctx.bx * kThreads + tid;
            if (i < p.n) {
                double s = 0.0;
                for (int q = 0; q < p.nparts; ++q) s += p.in[(size_t)q * p.stride + i];
                p.out[i] = (float)s;
            }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
// SE layer (spectral_transform.py:23-28):  gate = sigmoid(W2 relu(W1 mean_hw(x))),  y = r(x) * gate
// where r is the SpectralTransform resampling that precedes it (:44-47, :79): identity,
// nearest x2 upsampling or 2x2 average pooling.  mean_hw(r(x)) == mean_hw(x) for both.
// ---------------------------------------------------------------------------------------------
struct SeGateParams {
    const double* pooled;   // [B*C] sums over HW (ChanReduce-free path: computed by SePoolKernel)
    const float* w1;        // [hid][C]
    const float* w2;        // [C][hid]
    float* mean;            // [B][C]   saved
    float* hidden;          // [B][hid] saved (post-ReLU)
    float* gate;            // [B][C]
    int B, C, hid; float inv_hw;
};
struct SePoolParams { const float* x; double* pooled; int planes, HW; };
struct SePoolKernel {     // one CTA per plane
    typedef SePoolParams Params;
    static constexpr int kThreads = 128;
    static size_t smem_bytes() { return kThreads * sizeof(double); }
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        double* red = reinterpret_cast<double*>(smem);
        const size_t base = (size_t)ctx.bx * p.HW;
        FFC_PHASE {
            double s = 0.0;
            for (int i = tid; i < p.HW; i += kThreads) s += FFC_LDG(p.x + base + i);
            red[tid] = s;
        } FFC_SYNC;
        for (int w = kThreads / 2; w >= 1; w >>= 1) {
            FFC_PHASE { if (tid < w) red[tid] += red[tid + w]; } FFC_SYNC;
        }
        FFC_PHASE { if (tid == 0) p.pooled[ctx.bx] = red[0]; } FFC_SYNC;
    }
};
struct SeGateKernel {     // one CTA per image
    typedef SeGateParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        const int b = ctx.bx;
        FFC_PHASE {
            for (int c = tid; c < p.C; c += kThreads) p.mean[b * p.C + c] = (float)(p.pooled[b * p.C + c] * (double)p.inv_hw);
        } FFC_SYNC;
        FFC_PHASE {
            for (int j = tid; j < p.hid; j += kThreads) {
                float s = 0.f;
                for (int c = 0; c < p.C; ++c) s = fmaf(FFC_LDG(p.w1 + j * p.C + c), p.mean[b * p.C + c], s);
                p.hidden[b * p.hid + j] = s > 0.f ? s : 0.f;
            }
        } FFC_SYNC;
        FFC_PHASE {
            for (int c = tid; c < p.C; c += kThreads) {
                float s = 0.f;
                for (int j = 0; j < p.hid; ++j) s = fmaf(FFC_LDG(p.w2 + c * p.hid + j), p.hidden[b * p.hid + j], s);
                p.gate[b * p.C + c] = 1.0f / (1.0f + expf(-s));
            }
        } FFC_SYNC;
    }
};

// y = gate[plane] * r(x);  mode 0 identity, 1 nearest x2, 2 avgpool 2x2
struct SeScaleParams { const float* x; const float* gate; float* y; int planes, Hi, Wi, mode; };
struct SeScaleKernel {
    typedef SeScaleParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        const int Ho = p.mode == 1 ? p.Hi * 2 : (p.mode == 2 ? p.Hi / 2 : p.Hi);
        const int Wo = p.mode == 1 ? p.Wi * 2 : (p.mode == 2 ? p.Wi / 2 : p.Wi);
        const long long total = (long long)p.planes * Ho * Wo;
        // four outputs of one row per thread, 32-bit index math (the element-per-thread form with 64-bit divisions ran at an
        // eighth of the memory rate: 51 us for the 42 MB of the 16x16 -> 32x32 stage of the fgan32 generator at batch 256)
        const bool vec = Wo % 4 == 0 && total < (1LL << 31) && ((((uintptr_t)p.x) | ((uintptr_t)p.y)) & 15) == 0;
        if (vec) {
            FFC_PHASE {
                const unsigned q = (unsigned)(Wo / 4), n4 = (unsigned)(total / 4);
                for (unsigned i = (unsigned)ctx.bx * kThreads + tid; i < n4; i += (unsigned)ctx.gx * kThreads) {
                    const unsigned ox4 = i % q, t = i / q, oy = t % (unsigned)Ho, pl = t / (unsigned)Ho;
                    const float* xp = p.x + (size_t)pl * p.Hi * p.Wi;
                    const float g = FFC_LDG(p.gate + pl);
                    float4 v;
                    if (p.mode == 1) {
                        const float2 a = FFC_LDG(reinterpret_cast<const float2*>(xp + (oy >> 1) * p.Wi + 2 * ox4));
                        v = make_float4(a.x, a.x, a.y, a.y);
                    } else if (p.mode == 2) {
                        const float* r0 = xp + (2 * oy) * p.Wi + 8 * ox4;
                        const float4 a0 = FFC_LDG(reinterpret_cast<const float4*>(r0)), a1 = FFC_LDG(reinterpret_cast<const float4*>(r0) + 1);
                        const float4 b0 = FFC_LDG(reinterpret_cast<const float4*>(r0 + p.Wi)), b1 = FFC_LDG(reinterpret_cast<const float4*>(r0 + p.Wi) + 1);
                        v = make_float4(0.25f * (a0.x + a0.y + b0.x + b0.y), 0.25f * (a0.z + a0.w + b0.z + b0.w),
                                        0.25f * (a1.x + a1.y + b1.x + b1.y), 0.25f * (a1.z + a1.w + b1.z + b1.w));
                    } else {
                        v = FFC_LDG(reinterpret_cast<const float4*>(xp + oy * p.Wi) + ox4);
                    }
                    reinterpret_cast<float4*>(p.y)[i] = make_float4(v.x * g, v.y * g, v.z * g, v.w * g);
                }
            } FFC_SYNC;
            return;
        }
        FFC_PHASE {
            for (long long i = (long long)ctx.bx * kThreads + tid; i < total; i += (long long)ctx.gx * kThreads) {
                const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
                const long long pl = i / ((long long)Wo * Ho);
                const float* xp = p.x + pl * p.Hi * p.Wi;
                float v;
                if (p.mode == 1) v = FFC_LDG(xp + (oy >> 1) * p.Wi + (ox >> 1));
                else if (p.mode == 2) {
                    const float* q = xp + (2 * oy) * p.Wi + 2 * ox;
                    v = 0.25f * (FFC_LDG(q) + FFC_LDG(q + 1) + FFC_LDG(q + p.Wi) + FFC_LDG(q + p.Wi + 1));
                } else v = FFC_LDG(xp + oy * p.Wi + ox);
                p.y[i] = v * FFC_LDG(p.gate + pl);
            }
        } FFC_SYNC;
    }
};

// backward step 1: dgate[plane] = sum_i x_i * t_i,  t = r^T(dy)   (one CTA per plane)
struct SeDgateParams { const float* x; const float* dy; float* dgate; int planes, Hi, Wi, mode; };
FFC_HD float se_rT(const float* dyp, int iy, int ix, int Wi, int mode) {
    if (mode == 1) {
        const int Wo = 2 * Wi;
        const float* q = dyp + (2 * iy) * Wo + 2 * ix;
        return q[0] + q[1] + q[Wo] + q[Wo + 1];
    } else if (mode == 2) {
        const int Wo = Wi / 2;
        return 0.25f * dyp[(iy >> 1) * Wo + (ix >> 1)];
    }
    return dyp[iy * Wi + ix];
}
struct SeDgateKernel {
    typedef SeDgateParams Params;
    static constexpr int kThreads = 128;
    static size_t smem_bytes() { return kThreads * sizeof(double); }
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        double* red = reinterpret_cast<double*>(smem);
        const int HWi = p.Hi * p.Wi;
        const int HWo = p.mode == 1 ? HWi * 4 : (p.mode == 2 ? HWi / 4 : HWi);
        const float* xp = p.x + (size_t)ctx.bx * HWi;
        const float* dyp = p.dy + (size_t)ctx.bx * HWo;
        FFC_PHASE {
            double s = 0.0;
            for (int i = tid; i < HWi; i += kThreads) s += (double)(xp[i] * se_rT(dyp, i / p.Wi, i % p.Wi, p.Wi, p.mode));
            red[tid] = s;
        } FFC_SYNC;
        for (int w = kThreads / 2; w >= 1; w >>= 1) {
            FFC_PHASE { if (tid < w) red[tid] += red[tid + w]; } FFC_SYNC;
        }
        FFC_PHASE { if (tid == 0) p.dgate[ctx.bx] = (float)red[0]; } FFC_SYNC;
    }
};
// backward step 2 (one CTA per image): MLP backward; dW1/dW2 accumulated with atomics; dmean out
struct SeGateBwdParams {
    const float* dgate; const float* gate; const float* hidden; const float* mean;
    const float* w1; const float* w2;
    float* dw1; float* dw2;      // accumulated (caller zeroes)
    float* dmean;                // [B][C]
    float* scratch;              // [B][C + hid] work area
    int B, C, hid;
};
struct SeGateBwdKernel {
    typedef SeGateBwdParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        const int b = ctx.bx;
        float* dpre2 = p.scratch + (size_t)b * (p.C + p.hid);   // [C]
        float* dhid = dpre2 + p.C;                              // [hid]
        FFC_PHASE {
            for (int c = tid; c < p.C; c += kThreads) {
                const float g = p.gate[b * p.C + c];
                dpre2[c] = p.dgate[b * p.C + c] * g * (1.f - g);
            }
        } FFC_SYNC;
        FFC_PHASE {
            for (int j = tid; j < p.hid; j += kThreads) {
                float s = 0.f;
                for (int c = 0; c < p.C; ++c) s = fmaf(FFC_LDG(p.w2 + c * p.hid + j), dpre2[c], s);
                dhid[j] = p.hidden[b * p.hid + j] > 0.f ? s : 0.f;
            }
            for (int e = tid; e < p.C * p.hid; e += kThreads) {      // dW2[c][j] += dpre2[c] * hidden[j]
                const int c = e / p.hid, j = e % p.hid;
                ffc_atomic_add(p.dw2 + e, dpre2[c] * p.hidden[b * p.hid + j]);
            }
        } FFC_SYNC;
        FFC_PHASE {
            for (int c = tid; c < p.C; c += kThreads) {
                float s = 0.f;
                for (int j = 0; j < p.hid; ++j) s = fmaf(FFC_LDG(p.w1 + j * p.C + c), dhid[j], s);
                p.dmean[b * p.C + c] = s;
            }
            for (int e = tid; e < p.C * p.hid; e += kThreads) {      // dW1[j][c] += dhid[j] * mean[c]
                const int j = e / p.C, c = e % p.C;
                ffc_atomic_add(p.dw1 + e, dhid[j] * p.mean[b * p.C + c]);
            }
        } FFC_SYNC;
    }
};
// backward step 3: dx = gate * r^T(dy) + dmean / (Hi*Wi)
struct SeDxParams { const float* dy; const float* gate; const float* dmean; float* dx; int planes, Hi, Wi, mode; };
struct SeDxKernel {
    typedef SeDxParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        const int HWi = p.Hi * p.Wi;
        const int HWo = p.mode == 1 ? HWi * 4 : (p.mode == 2 ? HWi / 4 : HWi);
        const long long total = (long long)p.planes * HWi;
        const float inv = 1.0f / (float)HWi;
        FFC_PHASE {
            for (long long i = (long long)ctx.bx * kThreads + tid; i < total; i += (long long)ctx.gx * kThreads) {
                const long long pl = i / HWi; const int r = (int)(i % HWi);
                const float t = se_rT(p.dy + pl * HWo, r / p.Wi, r % p.Wi, p.Wi, p.mode);
                p.dx[i] = FFC_LDG(p.gate + pl) * t + FFC_LDG(p.dmean + pl) * inv;
            }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int ew_grid(long long work_items, int threads) {
    long long g = (work_items + threads - 1) / threads;
    const long long cap = (long long)ffc_sm_count() * 16;            // grid-stride beyond 16 CTAs per SM
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
static int reduce_split(int C, long long per_channel) {
    int ns = (int)((2 * ffc_sm_count() + C - 1) / C);       // aim at >= 2 CTAs per SM in total
    const long long maxs = (per_channel + 4 * FFC_RED_THREADS - 1) / (4 * FFC_RED_THREADS);
    if (ns > maxs) ns = (int)maxs;
    if (ns < 1) ns = 1;
    if (ns > 65535) ns = 65535;
    return ns;
}

// Workspace: 2*C doubles (zeroed here).
extern "C" int ffc_bn_act_fwd(const float* x, float* y, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                              int B, int C, int HW, int norm, int training, float eps, float momentum,
                              int act, float slope, void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(x && y, "ffc_bn_act_fwd: null pointer");
    FFC_REQUIRE(B >= 0 && C > 0 && HW > 0, "ffc_bn_act_fwd: bad sizes");
    FFC_REQUIRE(act >= FFC_ACT_IDENTITY && act <= FFC_ACT_SIGMOID, "ffc_bn_act_fwd: bad activation code %d", act);
    ffc_stream_t st = (ffc_stream_t)stream;
    const long long total = (long long)B * C * HW;
    if (total == 0) return FFC_OK;
    FFC_REQUIRE(HW % 4 != 0 || ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0, "ffc_bn_act_fwd: x/y must be 16-byte aligned");
    BnApplyParams ap{x, y, nullptr, nullptr, nullptr, nullptr, C, HW, total, act, slope};
    if (norm) {
        FFC_REQUIRE(gamma && beta && save_mean && save_invstd, "ffc_bn_act_fwd: BN needs gamma/beta/save buffers");
        if (training) {
            FFC_REQUIRE(workspace && workspace_bytes >= (size_t)2 * C * sizeof(double), "ffc_bn_act_fwd: workspace too small");
            double* sums = (double*)workspace;
            ChanReduceParams rp{x, nullptr, nullptr, nullptr, nullptr, nullptr, sums, B, C, HW, 0, reduce_split(C, (long long)B * HW), 0.f, 0};
            // room for one pair of sums per (channel, split): no zero-fill launch in front, no atomics, bitwise reproducible
            rp.partial = workspace_bytes >= (size_t)2 * C * rp.nsplit * sizeof(double) ? 1 : 0;
            if (!rp.partial) FFC_CHECK(ffc_memset_async(sums, 0, (size_t)2 * C * sizeof(double), st));
            FFC_CHECK((ffc_launch<ChanReduceKernel<0>>(C, rp.nsplit, 1, FFC_RED_THREADS, ChanReduceKernel<0>::smem_bytes(), st, rp)));
            BnFinalizeParams fp{sums, save_mean, save_invstd, running_mean, running_var, C, (double)B * HW, eps, momentum, rp.partial ? rp.nsplit : 1};
            FFC_CHECK((ffc_launch<BnFinalizeKernel>(ffc_cdiv(C, 256), 1, 1, 256, 0, st, fp)));
        } else {
            FFC_REQUIRE(running_mean && running_var, "ffc_bn_act_fwd: eval mode needs running statistics");
            BnEvalStatsParams ep{running_mean, running_var, save_mean, save_invstd, C, eps};
            FFC_CHECK((ffc_launch<BnEvalStatsKernel>(ffc_cdiv(C, 256), 1, 1, 256, 0, st, ep)));
        }
        ap.mean = save_mean; ap.invstd = save_invstd; ap.gamma = gamma; ap.beta = beta;
    }
    const long long items = (HW % 4 == 0) ? total / 4 : total;
    return ffc_launch<BnApplyKernel>(ew_grid(items, 256), 1, 1, 256, 0, st, ap);
}

// The statistics half of ffc_bn_act_fwd on its own: save_mean / save_invstd from the batch (training, running statistics
// updated in place) or from the running statistics (eval).  Consumers that apply the normalisation themselves
// (ffc_irfft2_bn_relu) call this first.  Workspace: 2*C doubles.
extern "C" int ffc_bn_stats(const float* x, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                            int B, int C, int HW, int training, float eps, float momentum,
                            void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(x && save_mean && save_invstd, "ffc_bn_stats: null pointer");
    FFC_REQUIRE(B >= 0 && C > 0 && HW > 0, "ffc_bn_stats: bad sizes");
    ffc_stream_t st = (ffc_stream_t)stream;
    if ((long long)B * C * HW == 0) return FFC_OK;
    if (training) {
        FFC_REQUIRE(workspace && workspace_bytes >= (size_t)2 * C * sizeof(double), "ffc_bn_stats: workspace too small");
        double* sums = (double*)workspace;
        ChanReduceParams rp{x, nullptr, nullptr, nullptr, nullptr, nullptr, sums, B, C, HW, 0, reduce_split(C, (long long)B * HW), 0.f, 0};
        rp.partial = workspace_bytes >= (size_t)2 * C * rp.nsplit * sizeof(double) ? 1 : 0;
        if (!rp.partial) FFC_CHECK(ffc_memset_async(sums, 0, (size_t)2 * C * sizeof(double), st));
        FFC_CHECK((ffc_launch<ChanReduceKernel<0>>(C, rp.nsplit, 1, FFC_RED_THREADS, ChanReduceKernel<0>::smem_bytes(), st, rp)));
        BnFinalizeParams fp{sums, save_mean, save_invstd, running_mean, running_var, C, (double)B * HW, eps, momentum, rp.partial ? rp.nsplit : 1};
        return ffc_launch<BnFinalizeKernel>(ffc_cdiv(C, 256), 1, 1, 256, 0, st, fp);
    }
    FFC_REQUIRE(running_mean && running_var, "ffc_bn_stats: eval mode needs running statistics");
    BnEvalStatsParams ep{running_mean, running_var, save_mean, save_invstd, C, eps};
    return ffc_launch<BnEvalStatsKernel>(ffc_cdiv(C, 256), 1, 1, 256, 0, st, ep);
}

// Workspace: 2*C doubles.
extern "C" int ffc_bn_act_bwd(const float* x, const float* dy, float* dx, const float* gamma, const float* beta,
                              const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta,
                              int B, int C, int HW, int norm, int training, int act, float slope,
                              void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(x && dy && dx, "ffc_bn_act_bwd: null pointer");
    FFC_REQUIRE(B >= 0 && C > 0 && HW > 0, "ffc_bn_act_bwd: bad sizes");
    ffc_stream_t st = (ffc_stream_t)stream;
    const long long total = (long long)B * C * HW;
    if (total == 0) return FFC_OK;
    BnBwdApplyParams ap{x, dy, dx, nullptr, nullptr, nullptr, nullptr, nullptr, dgamma, dbeta, C, HW, total, act, training, slope, (double)B * HW};
    if (norm) {
        FFC_REQUIRE(gamma && beta && save_mean && save_invstd, "ffc_bn_act_bwd: BN needs gamma/beta/saved statistics");
        FFC_REQUIRE(workspace && workspace_bytes >= (size_t)2 * C * sizeof(double), "ffc_bn_act_bwd: workspace too small");
        double* sums = (double*)workspace;
        FFC_CHECK(ffc_memset_async(sums, 0, (size_t)2 * C * sizeof(double), st));
        ChanReduceParams rp{x, dy, save_mean, save_invstd, gamma, beta, sums, B, C, HW, act, reduce_split(C, (long long)B * HW), slope, 0};
        FFC_CHECK((ffc_launch<ChanReduceKernel<1>>(C, rp.nsplit, 1, FFC_RED_THREADS, ChanReduceKernel<1>::smem_bytes(), st, rp)));
        ap.mean = save_mean; ap.invstd = save_invstd; ap.gamma = gamma; ap.beta = beta; ap.sums = sums;
    }
    const bool vec = (HW % 4 == 0) && ((((uintptr_t)x) | ((uintptr_t)dy) | ((uintptr_t)dx)) & 15) == 0;
    return ffc_launch<BnBwdApplyKernel>(ew_grid(vec ? total / 4 : total, 256), 1, 1, 256, 0, st, ap);
}

// db[c] = sum_{b,hw} dy.  Workspace: C doubles.
extern "C" int ffc_bias_grad(const float* dy, float* db, int B, int C, int HW,
                             void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(dy && db && B >= 0 && C > 0 && HW > 0, "ffc_bias_grad: bad arguments");
    FFC_REQUIRE(workspace && workspace_bytes >= (size_t)2 * C * sizeof(double), "ffc_bias_grad: workspace too small");
    ffc_stream_t st = (ffc_stream_t)stream;
    double* sums = (double*)workspace;
    int nparts = 1;
    if (B > 0) {
        ChanReduceParams rp{dy, nullptr, nullptr, nullptr, nullptr, nullptr, sums, B, C, HW, 0, reduce_split(C, (long long)B * HW), 0.f, 0};
        rp.partial = workspace_bytes >= (size_t)2 * C * rp.nsplit * sizeof(double) ? 1 : 0;
        if (rp.partial) nparts = rp.nsplit;
        else FFC_CHECK(ffc_memset_async(sums, 0, (size_t)2 * C * sizeof(double), st));
        FFC_CHECK((ffc_launch<ChanReduceKernel<2>>(C, rp.nsplit, 1, FFC_RED_THREADS, ChanReduceKernel<2>::smem_bytes(), st, rp)));
    } else {
        FFC_CHECK(ffc_memset_async(sums, 0, (size_t)2 * C * sizeof(double), st));
    }
    D2FParams dp{sums, db, C, nparts, 2 * C};
    return ffc_launch<D2FKernel>(ffc_cdiv(C, 256), 1, 1, 256, 0, st, dp);
}

// SE forward.  mode: 0 identity, 1 nearest x2 upsample, 2 avgpool 2x2 (applied to x before the gate).
// Workspace: B*C doubles.  Saves mean [B,C], hidden [B,hid], gate [B,C] for backward.
extern "C" int ffc_se_fwd(const float* x, const float* w1, const float* w2, float* y,
                          float* save_mean, float* save_hidden, float* save_gate,
                          int B, int C, int hid, int Hi, int Wi, int mode,
                          void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(x && y && save_mean && save_gate, "ffc_se_fwd: null pointer");
    FFC_REQUIRE(B >= 0 && C > 0 && hid >= 0 && Hi > 0 && Wi > 0 && mode >= 0 && mode <= 2, "ffc_se_fwd: bad arguments");
    FFC_REQUIRE(hid == 0 || (w1 && w2 && save_hidden), "ffc_se_fwd: weights missing");
    FFC_REQUIRE(mode != 2 || (Hi % 2 == 0 && Wi % 2 == 0), "ffc_se_fwd: avgpool needs even sizes");
    if (B == 0) return FFC_OK;
    FFC_REQUIRE(workspace && workspace_bytes >= (size_t)B * C * sizeof(double), "ffc_se_fwd: workspace too small");
    ffc_stream_t st = (ffc_stream_t)stream;
    double* pooled = (double*)workspace;
    SePoolParams pp{x, pooled, B * C, Hi * Wi};
    FFC_CHECK((ffc_launch<SePoolKernel>(B * C, 1, 1, SePoolKernel::kThreads, SePoolKernel::smem_bytes(), st, pp)));
    SeGateParams gp{pooled, w1, w2, save_mean, save_hidden, save_gate, B, C, hid, 1.0f / (float)(Hi * Wi)};
    FFC_CHECK((ffc_launch<SeGateKernel>(B, 1, 1, 256, 0, st, gp)));
    SeScaleParams sp{x, save_gate, y, B * C, Hi, Wi, mode};
    const long long outn = (long long)B * C * Hi * Wi * (mode == 1 ? 4 : 1) / (mode == 2 ? 4 : 1);
    return ffc_launch<SeScaleKernel>(ew_grid(outn, 256), 1, 1, 256, 0, st, sp);
}

// SE backward.  Workspace: B*(2C + hid) floats... (dgate [B*C], scratch [B*(C+hid)]); dmean uses dx-independent buffer.
extern "C" int ffc_se_bwd(const float* x, const float* dy, const float* w1, const float* w2,
                          const float* save_mean, const float* save_hidden, const float* save_gate,
                          float* dx, float* dw1, float* dw2,
                          int B, int C, int hid, int Hi, int Wi, int mode,
                          void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(x && dy && dx && save_mean && save_gate, "ffc_se_bwd: null pointer");
    FFC_REQUIRE(B >= 0 && C > 0 && hid >= 0 && Hi > 0 && Wi > 0 && mode >= 0 && mode <= 2, "ffc_se_bwd: bad arguments");
    FFC_REQUIRE(hid == 0 || (w1 && w2 && save_hidden && dw1 && dw2), "ffc_se_bwd: weights missing");
    ffc_stream_t st = (ffc_stream_t)stream;
    if (hid > 0) {
        FFC_CHECK(ffc_memset_async(dw1, 0, (size_t)hid * C * sizeof(float), st));
        FFC_CHECK(ffc_memset_async(dw2, 0, (size_t)hid * C * sizeof(float), st));
    }
    if (B == 0) return FFC_OK;
    const size_t need = ((size_t)B * C * 2 + (size_t)B * (C + hid)) * sizeof(float);
    FFC_REQUIRE(workspace && workspace_bytes >= need, "ffc_se_bwd: workspace too small");
    float* dgate = (float*)workspace;
    float* dmean = dgate + (size_t)B * C;
    float* scratch = dmean + (size_t)B * C;
    SeDgateParams dp{x, dy, dgate, B * C, Hi, Wi, mode};
    FFC_CHECK((ffc_launch<SeDgateKernel>(B * C, 1, 1, SeDgateKernel::kThreads, SeDgateKernel::smem_bytes(), st, dp)));
    SeGateBwdParams bp{dgate, save_gate, save_hidden, save_mean, w1, w2, dw1, dw2, dmean, scratch, B, C, hid};
    FFC_CHECK((ffc_launch<SeGateBwdKernel>(B, 1, 1, 256, 0, st, bp)));
    SeDxParams xp{dy, save_gate, dmean, dx, B * C, Hi, Wi, mode};
    return ffc_launch<SeDxKernel>(ew_grid((long long)B * C * Hi * Wi, 256), 1, 1, 256, 0, st, xp);
}
