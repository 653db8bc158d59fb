// Fused FourierUnitSN forward (layers/ffc/fourier_unity.py:32-58): per image tile
//   global x -> shared real planes -> rfft2 (in-register radix-2 FFTs, shared-memory transposes)
//   -> real/imag channel mix (packed FP32x2 FMA, weights broadcast from shared memory)
//   -> BatchNorm + ReLU -> irfft2 -> [+ residual] -> global out,
// so the (B, 2C, H, W/2+1) spectrum never leaves shared memory.
//
// Training-mode BatchNorm needs statistics over the whole batch before anything can be normalised.
// PASS 0 ("stats") runs load -> rfft2 -> mix and reduces sum(y), sum(y^2) per channel (double);
// PASS 1 ("apply") recomputes load -> rfft2 -> mix from x (a second read of x, normally an L2 hit),
// normalises with the finished statistics and runs the inverse.  Eval mode is PASS 1 only.
//
// Tile = IMGS images x all channels; persistent CTAs loop over tiles.  ONE shared-memory region of
// IMGS * CB planes (CB = max(Cin, Cout)) holds the real planes and the spectrum alternately: the row
// transforms run in place because spectrum row u (Wf = W/2+1 complex = W+2 floats) overlays real row u
// (row stride RS = W+4 floats), and each thread overwrites exactly the two rows it has read.
// Supported: H == W in {4, 8, 16, 32}, Cin, Cout <= 32 (CP = 8, 16 or 32 is the register tile of the mix).
#include "ffc_fft2.cuh"

struct FuFwdParams {
    const float* x;          // (B, Cin, N, N)
    const float* w;          // [2*Cout][2*Cin]
    const float* gamma; const float* beta;          // [2*Cout]
    float* running_mean; float* running_var;        // [2*Cout] (updated by PASS 1, block 0, training)
    float* save_mean; float* save_invstd;           // [2*Cout] written by PASS 1
    const float* residual;   // (B, Cout, N, N) or null
    float* out;              // (B, Cout, N, N)
    double* sums;            // [4*Cout]: sum(y) then sum(y^2)
    int B, Cin, Cout, imgs, training;
    float eps, momentum;
};

template <int N, int CP, int PASS>
struct FuFwdKernel {
    typedef FuFwdParams Params;
    typedef Fft2Plan<N, N> PL;
    // CTA shape by tile footprint: in-register FFT-32 threads need ~100 registers, so small tiles run as
    // several narrow CTAs per SM (4 x 160 threads) while the 147 KB tile of 32 channels @ 32x32 gets one wide CTA.
    static constexpr bool kBig = (N == 32 && CP == 32), kMid = (N == 32 && CP == 16);
    static constexpr int kThreads = kBig ? 512 : (kMid ? 256 : 160);
    static constexpr int kMinBlocks = kBig ? 1 : (kMid ? 2 : 4);
    static constexpr int Wf = PL::Wf;
    static constexpr int SPS = PL::RS / 2;          // float2 per spectrum row (in-place layout)
    static constexpr int BINS = N * Wf;
    static constexpr int BPT = CP <= 16 ? 2 : 1;    // bins per thread in the mix
    struct Acc { double v[4]; };

    static size_t smem_floats(int imgs, int cb, int cout, int nt) {
        return (size_t)imgs * cb * PL::REGION + (size_t)nt * 8 + (size_t)cout * CP * 4 + 3 * 2 * cout + 8;
    }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        const int ntiles = (p.B + p.imgs - 1) / p.imgs;
        run_tiles(p, ctx, smem, ctx.bx, ntiles);
    }

    // inverse columns + rows of the tile's Cout planes (in place) and the coalesced, batched store
    static FFC_DEVICE void inverse_and_store(const Params& p, const BlockCtx& ctx, float* spec, int img0, int ni, PlaneMap map_out) {
        const int nt = ctx.nt, Cout = p.Cout;
        FFC_PHASE { fft2_cols_L1<N, N, +1, SPS>(tid, nt, ni * Cout, spec, map_out); } FFC_SYNC;
        FFC_PHASE { fft2_rows_inv_L1<N, N, SPS>(tid, nt, ni * Cout, spec, spec, 1.0f, map_out, map_out); } FFC_SYNC;
        FFC_PHASE {
            constexpr int Q = N / 4;
            constexpr int LDU = 8;
            const int rows = ni * Cout * N;
            const int j = tid % Q, rstep = nt / Q;
            const size_t g0 = (size_t)img0 * Cout * N * N;
            float4* dst = reinterpret_cast<float4*>(p.out + g0) + j;
            const float4* res = p.residual ? reinterpret_cast<const float4*>(p.residual + g0) + j : nullptr;
            for (int r0 = tid / Q; r0 < rows; r0 += rstep * LDU) {
                float4 q[LDU];
                if (res) {
                    FFC_UNROLL
                    for (int u = 0; u < LDU; ++u) {
                        const int r = r0 + u * rstep;
                        if (r < rows) q[u] = FFC_LDG(res + r * Q);
                    }
                }
                FFC_UNROLL
                for (int u = 0; u < LDU; ++u) {
                    const int r = r0 + u * rstep;
                    if (r < rows) {
                        float4 v = *reinterpret_cast<const float4*>(spec + (size_t)map_out(r / N) * PL::REGION + (r % N) * PL::RS + 4 * j);
                        if (res) { v.x += q[u].x; v.y += q[u].y; v.z += q[u].z; v.w += q[u].w; }
                        dst[r * Q] = v;
                    }
                }
            }
        } FFC_SYNC;
    }

    // tiles [t_begin, t_end) with stride gridDim.x starting at t_begin (t_end = t_begin + 1: exactly one tile)
    static FFC_DEVICE void run_tiles(const Params& p, const BlockCtx& ctx, float* smem, int t_begin, int t_end) {
        const int Cin = p.Cin, Cout = p.Cout;
        const int CB = Cin > Cout ? Cin : Cout;
        const int nt = ctx.nt;
        float* spec = smem;
        double* red = reinterpret_cast<double*>(spec + (size_t)p.imgs * CB * PL::REGION);   // [nt][4] doubles
        float* wq = reinterpret_cast<float*>(red) + (size_t)nt * 8;                          // [Cout][CP] float4
        float* bn_mu = wq + (size_t)Cout * CP * 4;                                          // [2Cout]
        float* bn_a = bn_mu + 2 * Cout;
        float* bn_b = bn_a + 2 * Cout;
        const float scale = 1.0f / (float)N;                    // ortho: 1/sqrt(N*N), applied once per direction
        const double count = (double)p.B * BINS;
        const int S = nt / Cout;                                // stats: slices per complex channel
        // identity maps (no runtime integer division) in the usual Cin == Cout case
        PlaneMap map_in; map_in.cn = (Cin == CB) ? 0 : Cin; map_in.cb = CB;
        PlaneMap map_out; map_out.cn = (Cout == CB) ? 0 : Cout; map_out.cb = CB;

        FFC_TLS(Acc, acc);
        // ---- prologue: mix weights as FFMA2 operand pairs (forward ortho scale folded in), BN constants
        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            acc.v[0] = acc.v[1] = acc.v[2] = acc.v[3] = 0.0;
            for (int e = tid; e < Cout * CP; e += nt) {
                const int o2 = e / CP, c = e % CP;
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < Cin) {
                    const float* r0 = p.w + (size_t)(2 * o2) * 2 * Cin + 2 * c;      // W[2o][2c], W[2o][2c+1]
                    const float* r1 = r0 + 2 * Cin;                                  // W[2o+1][2c], W[2o+1][2c+1]
                    q = make_float4(FFC_LDG(r0) * scale, FFC_LDG(r1 + 1) * scale, FFC_LDG(r1) * scale, FFC_LDG(r0 + 1) * scale);
                }
                reinterpret_cast<float4*>(wq)[e] = q;
            }
            if (PASS == 1) {
                for (int o = tid; o < 2 * Cout; o += nt) {
                    float mean, invstd;
                    if (p.training) {
                        const double m = p.sums[o] / count;
                        double var = p.sums[2 * Cout + o] / count - m * m;
                        if (var < 0.0) var = 0.0;
                        mean = (float)m;
                        invstd = 1.0f / sqrtf((float)var + p.eps);
                        if (ctx.bx == 0 && p.running_mean) {
                            const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
                            p.running_mean[o] = (1.f - p.momentum) * p.running_mean[o] + p.momentum * mean;
                            p.running_var[o] = (1.f - p.momentum) * p.running_var[o] + p.momentum * (float)unb;
                        }
                    } else {
                        mean = p.running_mean[o];
                        invstd = 1.0f / sqrtf(p.running_var[o] + p.eps);
                    }
                    if (ctx.bx == 0) { p.save_mean[o] = mean; p.save_invstd[o] = invstd; }
                    // relu(z) * scale == relu(z * scale): the inverse ortho scale rides in the BN constants
                    bn_mu[o] = mean;
                    bn_a[o] = invstd * FFC_LDG(p.gamma + o) * scale;
                    bn_b[o] = FFC_LDG(p.beta + o) * scale;
                }
            }
        }   // no barrier: the first tile's load phase writes a different shared region and ends with one

        for (int tile = t_begin; tile < t_end; tile += ctx.gx) {
            const int img0 = tile * p.imgs;
            const int ni = (p.B - img0) < p.imgs ? (p.B - img0) : p.imgs;
            // ---- coalesced load of the tile's planes into the real layout.  LDU float4 loads are issued
            // back to back before the first shared-memory store so one DRAM latency covers the batch.
            FFC_PHASE {
                // The tile's planes form one tall (rows x N) matrix that is contiguous in global memory.  A thread
                // owns one float4 column j and walks rows with a fixed stride, so addresses advance by constants;
                // LDU loads are in flight before the first shared-memory store.
                constexpr int Q = N / 4;                        // float4 per row
                constexpr int LDU = 8;
                const int rows = ni * Cin * N;
                const int j = tid % Q, rstep = nt / Q;          // nt is a multiple of 32, Q divides 32
                const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)img0 * Cin * N * N) + j;
                for (int r0 = tid / Q; r0 < rows; r0 += rstep * LDU) {
                    float4 v[LDU];
                    FFC_UNROLL
                    for (int u = 0; u < LDU; ++u) {
                        const int r = r0 + u * rstep;
                        if (r < rows) v[u] = FFC_LDG(src + r * Q);
                    }
                    FFC_UNROLL
                    for (int u = 0; u < LDU; ++u) {
                        const int r = r0 + u * rstep;
                        if (r < rows)
                            *reinterpret_cast<float4*>(spec + (size_t)map_in(r / N) * PL::REGION + (r % N) * PL::RS + 4 * j) = v[u];
                    }
                }
            } FFC_SYNC;
            FFC_PHASE { fft2_rows_fwd_L1<N, N, SPS>(tid, nt, ni * Cin, spec, spec, map_in, map_in); } FFC_SYNC;
            FFC_PHASE { fft2_cols_L1<N, N, -1, SPS>(tid, nt, ni * Cin, spec, map_in); } FFC_SYNC;
            // ---- channel mix per bin, in place (one thread owns BPT bins across all channels)
            FFC_PHASE {
                constexpr int BG = (BINS + BPT - 1) / BPT;           // bin groups per image
                for (int it = tid; it < ni * BG; it += nt) {
                    const int im = it / BG, bg = it % BG;
                    float2* base[BPT];
                    bool live[BPT];
                    FFC_UNROLL
                    for (int b = 0; b < BPT; ++b) {
                        int bin = bg + b * BG;
                        live[b] = bin < BINS;
                        if (!live[b]) bin = 0;
                        base[b] = reinterpret_cast<float2*>(spec + (size_t)im * CB * PL::REGION) + (bin / Wf) * SPS + (bin % Wf);
                    }
                    float2 s[BPT][CP];
                    FFC_UNROLL
                    for (int b = 0; b < BPT; ++b) {
                        FFC_UNROLL
                        for (int c = 0; c < CP; ++c) {
                            s[b][c] = make_float2(0.f, 0.f);
                            if (c < Cin) s[b][c] = base[b][(size_t)c * (PL::REGION / 2)];
                        }
                    }
                    constexpr int OU = CP <= 16 ? CP : 4;      // full unroll for small tiles: immediate offsets everywhere
#pragma unroll OU
                    for (int o2 = 0; o2 < CP; ++o2) {
                        if (o2 >= Cout) break;
                        const float4* wrow = reinterpret_cast<const float4*>(wq) + (size_t)o2 * CP;
                        float2 pa[BPT], pb[BPT];
                        FFC_UNROLL
                        for (int b = 0; b < BPT; ++b) { pa[b] = make_float2(0.f, 0.f); pb[b] = make_float2(0.f, 0.f); }
                        FFC_UNROLL
                        for (int c = 0; c < CP; ++c) {
                            const float4 q = wrow[c];
                            const float2 qa = make_float2(q.x, q.y), qb = make_float2(q.z, q.w);
                            FFC_UNROLL
                            for (int b = 0; b < BPT; ++b) {
                                pa[b] = ffc_fma2(qa, s[b][c], pa[b]);     // (W[2o][2c] re, W[2o+1][2c+1] im)
                                pb[b] = ffc_fma2(qb, s[b][c], pb[b]);     // (W[2o+1][2c] re, W[2o][2c+1] im)
                            }
                        }
                        FFC_UNROLL
                        for (int b = 0; b < BPT; ++b) {
                            float yr = pa[b].x + pb[b].y, yi = pb[b].x + pa[b].y;
                            if (PASS == 1) {
                                yr = (yr - bn_mu[2 * o2]) * bn_a[2 * o2] + bn_b[2 * o2];
                                yi = (yi - bn_mu[2 * o2 + 1]) * bn_a[2 * o2 + 1] + bn_b[2 * o2 + 1];
                                yr = yr > 0.f ? yr : 0.f;
                                yi = yi > 0.f ? yi : 0.f;
                            }
                            if (live[b]) base[b][(size_t)o2 * (PL::REGION / 2)] = make_float2(yr, yi);
                        }
                    }
                }
            } FFC_SYNC;
            if (PASS == 0) {
                // ---- statistics: thread = (complex channel c2, slice); partial sums stay in registers across tiles
                FFC_PHASE {
                    FFC_TLS_REF(Acc, acc);
                    const int c2 = tid / S, sl = tid % S;
                    if (c2 < Cout) {
                        for (int it = sl; it < ni * BINS; it += S) {
                            const int im = it / BINS, bin = it % BINS;
                            const float2 y = reinterpret_cast<const float2*>(spec + (size_t)(im * CB + c2) * PL::REGION)[(bin / Wf) * SPS + (bin % Wf)];
                            acc.v[0] += y.x; acc.v[1] += (double)y.x * y.x;
                            acc.v[2] += y.y; acc.v[3] += (double)y.y * y.y;
                        }
                    }
                } FFC_SYNC;
            } else {
                inverse_and_store(p, ctx, spec, img0, ni, map_out);
            }
        }
        if (PASS == 0) {
            FFC_PHASE {
                FFC_TLS_REF(Acc, acc);
                const int c2 = tid / S, sl = tid % S;
                if (c2 < Cout) {
                    FFC_UNROLL
                    for (int j = 0; j < 4; ++j) red[((size_t)c2 * S + sl) * 4 + j] = acc.v[j];
                }
            } FFC_SYNC;
            FFC_PHASE {
                if (tid < Cout * 4) {
                    const int c2 = tid / 4, j = tid % 4;
                    double s = 0.0;
                    for (int sl = 0; sl < S; ++sl) s += red[((size_t)c2 * S + sl) * 4 + j];
                    const int chn = 2 * c2 + (j >> 1);
                    ffc_atomic_add(p.sums + ((j & 1) ? 2 * Cout + chn : chn), s);
                }
            } FFC_SYNC;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Training-mode forward in ONE pass: when every image tile of the batch fits on the chip at once (B200: 148 SMs x
// up to 227 KB of shared memory), the mixed spectrum simply stays in shared memory across a grid-wide barrier while
// the BatchNorm statistics are finished, instead of being recomputed from x by a second pass.
//   part0: load -> rfft2 -> mix -> per-channel sum / sum^2 -> atomics          (spectrum kept in smem)
//   ---- cooperative grid barrier ----
//   part1: BN constants -> normalise + ReLU in place -> irfft2 -> (+residual) store
// One tile per CTA (grid == number of tiles); the host falls back to the two-pass form when the grid does not fit.
// ---------------------------------------------------------------------------------------------
template <int N, int CP>
struct FuFwdCoop {
    typedef FuFwdParams Params;
    typedef FuFwdKernel<N, CP, 0> Base;
    typedef Fft2Plan<N, N> PL;
    static constexpr int kThreads = Base::kThreads;
    static constexpr int kMinBlocks = Base::kMinBlocks;
    static constexpr int Wf = PL::Wf, SPS = Base::SPS, BINS = Base::BINS;

    struct Lay { float* spec; double* red; float* wq; float* bn_mu; float* bn_a; float* bn_b; int CB; };
    static FFC_DEVICE Lay layout(const Params& p, int nt, float* smem) {
        Lay l;
        l.CB = p.Cin > p.Cout ? p.Cin : p.Cout;
        l.spec = smem;
        l.red = reinterpret_cast<double*>(smem + (size_t)p.imgs * l.CB * PL::REGION);
        l.wq = reinterpret_cast<float*>(l.red) + (size_t)nt * 8;
        l.bn_mu = l.wq + (size_t)p.Cout * CP * 4;
        l.bn_a = l.bn_mu + 2 * p.Cout;
        l.bn_b = l.bn_a + 2 * p.Cout;
        return l;
    }

    static FFC_DEVICE void part0(const Params& p, const BlockCtx& ctx, float* smem) {
        // identical to the stats pass over the single tile ctx.bx; the spectrum (mixed, not normalised) stays in smem
        Params q = p;
        FuFwdKernel<N, CP, 0>::run_tiles(q, ctx, smem, ctx.bx, ctx.bx + 1);
    }

    static FFC_DEVICE void part1(const Params& p, const BlockCtx& ctx, float* smem) {
        const int Cin = p.Cin, Cout = p.Cout, nt = ctx.nt;
        const Lay l = layout(p, nt, smem);
        const int CB = l.CB;
        const float scale = 1.0f / (float)N;
        const double count = (double)p.B * BINS;
        const int img0 = ctx.bx * p.imgs;
        const int ni = (p.B - img0) < p.imgs ? (p.B - img0) : p.imgs;
        PlaneMap map_out; map_out.cn = (Cout == CB) ? 0 : Cout; map_out.cb = CB;
        (void)Cin;
        FFC_PHASE {
            for (int o = tid; o < 2 * Cout; o += nt) {
                const double m = p.sums[o] / count;
                double var = p.sums[2 * Cout + o] / count - m * m;
                if (var < 0.0) var = 0.0;
                const float mean = (float)m;
                const float invstd = 1.0f / sqrtf((float)var + p.eps);
                if (ctx.bx == 0) {
                    p.save_mean[o] = mean; p.save_invstd[o] = invstd;
                    if (p.running_mean) {
                        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
                        p.running_mean[o] = (1.f - p.momentum) * p.running_mean[o] + p.momentum * mean;
                        p.running_var[o] = (1.f - p.momentum) * p.running_var[o] + p.momentum * (float)unb;
                    }
                }
                l.bn_mu[o] = mean;
                l.bn_a[o] = invstd * FFC_LDG(p.gamma + o) * scale;
                l.bn_b[o] = FFC_LDG(p.beta + o) * scale;
            }
        } FFC_SYNC;
        FFC_PHASE {      // normalise + ReLU in place over the Cout spectrum planes of the tile
            for (int it = tid; it < ni * Cout * BINS; it += nt) {
                const int bin = it % BINS, c2 = (it / BINS) % Cout, im = it / (BINS * Cout);
                float2* q = reinterpret_cast<float2*>(l.spec + (size_t)(im * CB + c2) * PL::REGION) + (bin / Wf) * SPS + (bin % Wf);
                float2 y = *q;
                y.x = (y.x - l.bn_mu[2 * c2]) * l.bn_a[2 * c2] + l.bn_b[2 * c2];
                y.y = (y.y - l.bn_mu[2 * c2 + 1]) * l.bn_a[2 * c2 + 1] + l.bn_b[2 * c2 + 1];
                *q = make_float2(y.x > 0.f ? y.x : 0.f, y.y > 0.f ? y.y : 0.f);
            }
        } FFC_SYNC;
        FuFwdKernel<N, CP, 1>::inverse_and_store(p, ctx, l.spec, img0, ni, map_out);
    }
};

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct FuPlan { int imgs, nt, grid; size_t smem; bool ok; };
// test hook (ffc_fu2.cu): 1 disables the cooperative single-pass training forward
extern int ffc_fu2_force_two_pass;
#define ffc_fu_force_two_pass ffc_fu2_force_two_pass

template <int N, int CP>
static FuPlan fu_plan(int B, int Cin, int Cout) {
    typedef Fft2Plan<N, N> PL;
    typedef FuFwdKernel<N, CP, 1> K;
    FuPlan pl; pl.ok = false;
    const int CB = Cin > Cout ? Cin : Cout;
    const size_t per_block = (size_t)(224 * 1024) / K::kMinBlocks - 1024;     // shared memory one CTA may take
    const int items_img = CB * PL::Wf;                                         // column FFTs per image (widest FFT phase)
    int imgs = (K::kThreads + items_img / 2) / items_img;                      // ~one full CTA of FFT work
    if (imgs > 8) imgs = 8;
    if (imgs > B) imgs = B;
    if (imgs < 1) imgs = 1;
    for (; imgs >= 1; --imgs) {
        const int total = imgs * items_img;
        const int iters = (total + K::kThreads - 1) / K::kThreads;
        int nt = ((total + iters - 1) / iters + 31) / 32 * 32;
        const int need = (4 * Cout + 31) / 32 * 32;                            // stats combine: Cout*4 threads
        if (nt < need) nt = need;
        if (nt < 64) nt = 64;
        if (nt > K::kThreads) nt = K::kThreads;
        const size_t bytes = K::smem_floats(imgs, CB, Cout, nt) * 4;
        if (bytes <= per_block) {
            pl.imgs = imgs; pl.nt = nt; pl.smem = bytes; pl.ok = true;
            const int ntiles = (B + imgs - 1) / imgs;
            int per_sm = (int)((224 * 1024) / (bytes + 1024));
            const int by_threads = 2048 / nt;
            if (per_sm > by_threads) per_sm = by_threads;
            if (per_sm > 8) per_sm = 8;
            if (per_sm < 1) per_sm = 1;
            pl.grid = ntiles < ffc_sm_count() * per_sm ? ntiles : ffc_sm_count() * per_sm;
            return pl;
        }
    }
    return pl;
}

template <int N, int CP>
static int fu_fwd_launch(const FuFwdParams& p0, ffc_stream_t st) {
    FuPlan pl = fu_plan<N, CP>(p0.B, p0.Cin, p0.Cout);
    if (!pl.ok) { ffc_set_error("ffc_fu_fwd: shape does not fit the fused kernel"); return FFC_ERR_BAD_ARG; }
    FuFwdParams p = p0;
    p.imgs = pl.imgs;
    if (p.training) {
        FFC_CHECK(ffc_memset_async(p.sums, 0, (size_t)4 * p.Cout * sizeof(double), st));
        const int ntiles = (p.B + pl.imgs - 1) / pl.imgs;
        if (!ffc_fu_force_two_pass && ntiles <= ffc_coop_capacity_blocks<FuFwdCoop<N, CP>>(pl.nt, pl.smem))
            return ffc_launch_coop<FuFwdCoop<N, CP>>(ntiles, pl.nt, pl.smem, st, p);      // single pass, spectrum held on chip
        FFC_CHECK((ffc_launch<FuFwdKernel<N, CP, 0>>(pl.grid, 1, 1, pl.nt, pl.smem, st, p)));
    }
    return ffc_launch<FuFwdKernel<N, CP, 1>>(pl.grid, 1, 1, pl.nt, pl.smem, st, p);
}

template <int N>
static int fu_fwd_dispatch_cp(const FuFwdParams& p, ffc_stream_t st) {
    const int cm = p.Cin > p.Cout ? p.Cin : p.Cout;
    if (cm <= 8) return fu_fwd_launch<N, 8>(p, st);
    if (cm <= 16) return fu_fwd_launch<N, 16>(p, st);
    return fu_fwd_launch<N, 32>(p, st);
}

template <int N>
static bool fu_fits(int B, int Cin, int Cout) {
    const int cm = Cin > Cout ? Cin : Cout;
    if (cm <= 8) return fu_plan<N, 8>(B, Cin, Cout).ok;
    if (cm <= 16) return fu_plan<N, 16>(B, Cin, Cout).ok;
    return fu_plan<N, 32>(B, Cin, Cout).ok;
}

// 1 when ffc_fu_fwd supports the shape (otherwise callers use ffc_rfft2 | ffc_conv2d_fwd | ffc_bn_act_fwd | ffc_irfft2)
// First-generation kernel: since ffc_fu2.cu it only serves the 4x4 planes (several images per CTA).
extern "C" int ffc_fu1_supported(int B, int Cin, int Cout, int H, int W) {
    if (H != 4 || W != 4 || B < 1 || Cin < 1 || Cout < 1 || Cin > 32 || Cout > 32) return 0;
    return fu_fits<4>(B, Cin, Cout);
}

// Fused FourierUnitSN forward.  w: conv_layer.weight viewed [2*Cout][2*Cin]; gamma/beta/running_*: bn.* [2*Cout];
// save_mean/save_invstd [2*Cout] are written; out = [residual +] irfft2(relu(bn(mix(rfft2(x))))).
// workspace >= 4*Cout doubles.
extern "C" int ffc_fu1_fwd(const float* x, const float* w, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                          const float* residual, float* out,
                          int B, int Cin, int Cout, int H, int W, int training, float eps, float momentum,
                          void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(x && w && gamma && beta && save_mean && save_invstd && out, "ffc_fu_fwd: null pointer");
    FFC_REQUIRE(training || (running_mean && running_var), "ffc_fu_fwd: eval mode needs running statistics");
    FFC_REQUIRE(B >= 0, "ffc_fu_fwd: negative batch");
    if (B == 0) return FFC_OK;
    FFC_REQUIRE(ffc_fu1_supported(B, Cin, Cout, H, W), "ffc_fu_fwd: unsupported shape B=%d Cin=%d Cout=%d %dx%d", B, Cin, Cout, H, W);
    FFC_REQUIRE((((uintptr_t)x | (uintptr_t)out | (uintptr_t)residual) & 15) == 0, "ffc_fu_fwd: x/out/residual must be 16-byte aligned");
    if (!(workspace && workspace_bytes >= (size_t)4 * Cout * sizeof(double))) { ffc_set_error("ffc_fu_fwd: workspace too small"); return FFC_ERR_WORKSPACE; }
    FuFwdParams p;
    p.x = x; p.w = w; p.gamma = gamma; p.beta = beta; p.running_mean = running_mean; p.running_var = running_var;
    p.save_mean = save_mean; p.save_invstd = save_invstd; p.residual = residual; p.out = out;
    p.sums = (double*)workspace; p.B = B; p.Cin = Cin; p.Cout = Cout; p.imgs = 1; p.training = training;
    p.eps = eps; p.momentum = momentum;
    ffc_stream_t st = (ffc_stream_t)stream;
    return fu_fwd_dispatch_cp<4>(p, st);
}
