// Fused FourierUnitSN forward, second generation (layers/ffc/fourier_unity.py:32-58), planes 8x8 .. 32x32:
//   x -> rfft2 -> real/imag channel mix -> BatchNorm + ReLU -> irfft2 -> [+ residual] -> out
// with the spectrum held in shared memory.  One image per CTA at a time, 160..640 threads per image (see
// ffc_fu2.cuh for the transform decomposition), so that a batch of only 1-2 images per SM still fills the SM.
//
// Three kernels share the phases:
//   Fu2Fwd<N,CP,0>  statistics pass  (load, rfft2, mix, per-channel sum / sum^2 -> double atomics)
//   Fu2Fwd<N,CP,1>  apply pass       (load, rfft2, mix, BN+ReLU with finished statistics, irfft2, store); eval mode
//   Fu2Coop<N,CP>   training forward in ONE cooperative launch when every image has its own resident CTA: the mixed
//                   spectrum waits in shared memory across the grid barrier while the statistics are finished, then
//                   BN+ReLU is applied as the inverse column pass loads it.
// Channel mix: thread = (bin, group of OG output channels); inputs of the bin are read once into registers, the
// (scaled, re-paired) weights are broadcast from shared memory as float4, two packed FP32x2 FMAs per weight.
// When OG < Cout the mix is written to a second plane region (a bin is then shared by several threads).
#include "ffc_fu2.cuh"


template <int N, int CP, int PASS>
struct Fu2Fwd {
    typedef Fu2Params Params;
    typedef Fu2G<N> G;
    typedef Fu2Cfg<N, CP> CFG;
    static constexpr int kThreads = CFG::kThreads;
    static constexpr int kMinBlocks = CFG::kMinBlocks;
    static constexpr int OG = CFG::OG;
    static constexpr int Wf = G::Wf, BINS = G::BINS;
    struct Acc { float v[4]; };

    // shared-memory carve-up (floats)
    struct Lay { float* sreg; float* yreg; float4* wq; float2* tw; float2* bn_a; float2* bn_b; float* red; bool inplace; };
    static FFC_HDM bool one_region(int Cout) { return CFG::kInPlaceAlways || Cout <= OG; }
    static size_t smem_floats(int Cin, int Cout, int nt) {
        const int CB = Cin > Cout ? Cin : Cout;
        const size_t planes = one_region(Cout) ? (size_t)CB : (size_t)(Cin + Cout);
        return planes * G::REGION + (size_t)Cout * CP * 4 + 2 * N + 4 * Cout + (size_t)nt * 8 + 8;   // red: two buffers of nt*4
    }
    static FFC_DEVICE Lay layout(const Params& p, float* smem) {
        Lay l;
        const int CB = p.Cin > p.Cout ? p.Cin : p.Cout;
        l.inplace = one_region(p.Cout);
        l.sreg = smem;
        l.yreg = l.inplace ? smem : smem + (size_t)p.Cin * G::REGION;
        float* q = smem + (l.inplace ? (size_t)CB : (size_t)(p.Cin + p.Cout)) * G::REGION;
        l.wq = reinterpret_cast<float4*>(q); q += (size_t)p.Cout * CP * 4;
        l.tw = reinterpret_cast<float2*>(q); q += 2 * N;
        l.bn_a = reinterpret_cast<float2*>(q); q += 2 * p.Cout;
        l.bn_b = reinterpret_cast<float2*>(q); q += 2 * p.Cout;
        l.red = q;
        return l;
    }

    static FFC_DEVICE void prologue(const Params& p, const BlockCtx& ctx, const Lay& l, int tid) {
        fu2_prologue<N, CP>(p.w, p.Cin, p.Cout, l.wq, l.tw, tid, ctx.nt);
    }

    // BatchNorm constants y -> relu(y*a + b) (inverse ortho scale folded in: relu(z)*s == relu(z*s)); block 0 also
    // publishes the saved statistics and updates the running ones
    static FFC_DEVICE void bn_constants(const Params& p, const BlockCtx& ctx, const Lay& l, int tid) {
        const float scale = 1.0f / (float)N;
        const double count = (double)p.B * BINS;
        for (int o = tid; o < 2 * p.Cout; o += ctx.nt) {
            float mean, invstd;
            if (p.training) {
                const double m = p.sums[o] / count;
                double var = p.sums[2 * p.Cout + o] / count - m * m;
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
            const float a = invstd * FFC_LDG(p.gamma + o) * scale;
            const float b = FFC_LDG(p.beta + o) * scale - mean * a;
            reinterpret_cast<float*>(l.bn_a)[o] = a;
            reinterpret_cast<float*>(l.bn_b)[o] = b;
        }
    }

    // channel mix of one image: sreg (Cin planes) -> yreg (Cout planes); BNF applies BN+ReLU to the result.
    // FULL: Cin == CP and Cout is a multiple of OG (every BASELINE shape) -- no channel guards, immediate offsets.
    // Items run over the bins v < M first (conflict-free lanes, power-of-two index math), the Nyquist bins follow.
    template <bool BNF, bool FULL>
    static FFC_DEVICE void mix_impl(const Params& p, const BlockCtx& ctx, const Lay& l, int tid) {
        const int Cin = p.Cin, Cout = p.Cout;
        const int ngrp = (Cout + OG - 1) / OG;
        constexpr int RF2 = G::REGION / 2;
        for (int it = tid; it < ngrp * BINS; it += ctx.nt) {
            const int bin = it % BINS, grp = it / BINS;
            const int off = fu2_bin_off<N>(bin);
            const float2* sp = reinterpret_cast<const float2*>(l.sreg) + off;
            float2* yp = reinterpret_cast<float2*>(l.yreg) + off + grp * OG * RF2;
            const float4* wg = l.wq + grp * OG * CP;
            float2 s[CP];
            FFC_UNROLL
            for (int c = 0; c < CP; ++c) {
                if (FULL) s[c] = sp[c * RF2];
                else { s[c] = make_float2(0.f, 0.f); if (c < Cin) s[c] = sp[c * RF2]; }
            }
            constexpr int OU = (OG <= 8) ? OG : 4;
#pragma unroll OU
            for (int j = 0; j < OG; ++j) {
                if (!FULL && grp * OG + j >= Cout) break;
                float2 pa = make_float2(0.f, 0.f), pb = make_float2(0.f, 0.f);
                FFC_UNROLL
                for (int c = 0; c < CP; ++c) {
                    const float4 q = wg[j * CP + c];
                    pa = ffc_fma2(make_float2(q.x, q.y), s[c], pa);     // (W[2o][2c] re, W[2o+1][2c+1] im)
                    pb = ffc_fma2(make_float2(q.z, q.w), s[c], pb);     // (W[2o+1][2c] re, W[2o][2c+1] im)
                }
                float2 y = make_float2(pa.x + pb.y, pb.x + pa.y);
                if (BNF) y = fu2_bn_relu(y, l.bn_a[grp * OG + j], l.bn_b[grp * OG + j]);
                yp[j * RF2] = y;
            }
        }
    }
    template <bool BNF>
    static FFC_DEVICE void mix(const Params& p, const BlockCtx& ctx, const Lay& l, int tid) {
        if (p.Cin == CP && p.Cout % OG == 0) mix_impl<BNF, true>(p, ctx, l, tid);
        else mix_impl<BNF, false>(p, ctx, l, tid);
    }

    // per-image statistics of the mixed spectrum: thread = (complex channel, slice) sums whole rows (float4 reads,
    // no per-bin index math); partial sums stay in registers
    static FFC_DEVICE void stats_accumulate(const Params& p, const BlockCtx& ctx, const Lay& l, int tid, Acc& acc) {
        const int S = ctx.nt / p.Cout;
        const int c2 = tid / S, sl = tid % S;
        if (c2 < p.Cout) {
            const float* plane = l.yreg + (size_t)c2 * G::REGION;
            for (int u = sl; u < N; u += S) {
                const float4* row = reinterpret_cast<const float4*>(plane + u * G::RS);
                FFC_UNROLL
                for (int j = 0; j < G::RS / 4; ++j) {
                    const float4 a = row[j];
                    acc.v[0] += a.x; acc.v[1] = fmaf(a.x, a.x, acc.v[1]);
                    acc.v[2] += a.y; acc.v[3] = fmaf(a.y, a.y, acc.v[3]);
                    if (2 * j + 1 < Wf) {                   // the last float4 of a row holds one bin and one pad
                        acc.v[0] += a.z; acc.v[1] = fmaf(a.z, a.z, acc.v[1]);
                        acc.v[2] += a.w; acc.v[3] = fmaf(a.w, a.w, acc.v[3]);
                    }
                }
            }
        }
    }
    static FFC_DEVICE void stats_flush(const Params& p, const BlockCtx& ctx, const Lay& l) { fu2_flush_sums(ctx, l.red, p.Cout, p.sums); }

    // load -> rfft2 -> mix of image `img` (phases with barriers)
    template <bool BNF>
    static FFC_DEVICE void forward_half(const Params& p, const BlockCtx& ctx, const Lay& l, int img) {
        const int Cin = p.Cin;
        FFC_PHASE { fu2_load_rows<N>(tid, ctx.nt, Cin * N, p.x + (size_t)img * Cin * N * N, l.sreg); } FFC_SYNC;
        FFC_PHASE { fu2_rows_fwd<N, false>(tid, ctx.nt, Cin * N, l.sreg); } FFC_SYNC;
        FU2_COLS_FWD(N, Cin, l.sreg, l.tw);
        FFC_PHASE { mix<BNF>(p, ctx, l, tid); } FFC_SYNC;
    }
    // irfft2 of the Cout planes in yreg -> out (+ residual); BNL applies BN+ReLU while the first pass loads
    template <bool BNL>
    static FFC_DEVICE void inverse_half(const Params& p, const BlockCtx& ctx, const Lay& l, int img) {
        const int Cout = p.Cout;
        Fu2Bn bn; bn.a = l.bn_a; bn.b = l.bn_b;
        FU2_COLS_INV(N, BNL, Cout, l.yreg, l.tw, bn);
        FFC_PHASE { fu2_rows_inv<N, false>(tid, ctx.nt, Cout * N, l.yreg, 1.0f); } FFC_SYNC;
        FFC_PHASE {
            const size_t g0 = (size_t)img * Cout * N * N;
            fu2_store_rows<N>(tid, ctx.nt, Cout * N, l.yreg, p.residual ? p.residual + g0 : nullptr, p.out + g0);
        } FFC_SYNC;
    }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        const Lay l = layout(p, smem);
        FFC_TLS(Acc, acc);
        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            acc.v[0] = acc.v[1] = acc.v[2] = acc.v[3] = 0.f;
            prologue(p, ctx, l, tid);
            if (PASS == 1) bn_constants(p, ctx, l, tid);
        }   // the first barrier of forward_half orders these writes before their first use
        for (int img = ctx.bx; img < p.B; img += ctx.gx) {
            if (PASS == 0) {
                forward_half<false>(p, ctx, l, img);
                FFC_PHASE { FFC_TLS_REF(Acc, acc); stats_accumulate(p, ctx, l, tid, acc); } FFC_SYNC;
            } else {
                forward_half<true>(p, ctx, l, img);
                inverse_half<false>(p, ctx, l, img);
            }
        }
        if (PASS == 0) {
            FFC_PHASE {
                FFC_TLS_REF(Acc, acc);
                FFC_UNROLL
                for (int j = 0; j < 4; ++j) l.red[(size_t)tid * 4 + j] = acc.v[j];
            } FFC_SYNC;
            stats_flush(p, ctx, l);
        }
    }
};

template <int N, int CP>
struct Fu2Coop {
    typedef Fu2Params Params;
    typedef Fu2Fwd<N, CP, 0> K;
    static constexpr int kThreads = K::kThreads;
    static constexpr int kMinBlocks = K::kMinBlocks;
    static FFC_DEVICE void part0(const Params& p, const BlockCtx& ctx, float* smem) {
        const typename K::Lay l = K::layout(p, smem);
        FFC_PHASE { K::prologue(p, ctx, l, tid); }
        K::template forward_half<false>(p, ctx, l, ctx.bx);
        FFC_PHASE {
            typename K::Acc acc;
            acc.v[0] = acc.v[1] = acc.v[2] = acc.v[3] = 0.f;
            K::stats_accumulate(p, ctx, l, tid, acc);
            FFC_UNROLL
            for (int j = 0; j < 4; ++j) l.red[(size_t)tid * 4 + j] = acc.v[j];
        } FFC_SYNC;
        K::stats_flush(p, ctx, l);
    }
    static FFC_DEVICE void part1(const Params& p, const BlockCtx& ctx, float* smem) {
        const typename K::Lay l = K::layout(p, smem);
        FFC_PHASE { K::bn_constants(p, ctx, l, tid); } FFC_SYNC;
        K::template inverse_half<true>(p, ctx, l, ctx.bx);
    }
};

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Fu2Plan { int nt; size_t smem; bool ok; };

template <int N, int CP>
static Fu2Plan fu2_plan(int Cin, int Cout) {
    typedef Fu2Fwd<N, CP, 1> K;
    typedef Fu2G<N> G;
    Fu2Plan pl; pl.ok = false;
    const int CB = Cin > Cout ? Cin : Cout;
    // widest phases: column pass A/B items and mix items; round to whole warps, cap at the kernel's bound
    const int col_items = CB * G::Wf * G::N2;            // pass A (the heavy one): N2 threads per column
    const int mix_items = ((Cout + K::OG - 1) / K::OG) * G::BINS;
    int want = col_items > mix_items ? col_items : mix_items;
    if (want < CB * N) want = CB * N;
    int nt = (want + 31) / 32 * 32;
    if (nt > K::kThreads) {                              // several rounds: balance them
        const int rounds = (want + K::kThreads - 1) / K::kThreads;
        nt = ((want + rounds - 1) / rounds + 31) / 32 * 32;
        if (nt > K::kThreads) nt = K::kThreads;
    }
    const int need = (Cout * 4 + 31) / 32 * 32;          // stats reduction: at least Cout * 4 threads
    if (nt < need) nt = need;
    if (nt > K::kThreads) return pl;
    pl.nt = nt;
    pl.smem = K::smem_floats(Cin, Cout, nt) * 4;
    pl.ok = pl.smem <= (size_t)227 * 1024;
    return pl;
}

int ffc_fu2_force_two_pass = 0;          // shared with ffc_fu_fused.cu (4x4 planes)
extern "C" void ffc_debug_fu_two_pass(int on) { ffc_fu2_force_two_pass = on; }

template <int N, int CP>
static int fu2_launch(const Fu2Params& p, ffc_stream_t st) {
    const Fu2Plan pl = fu2_plan<N, CP>(p.Cin, p.Cout);
    if (!pl.ok) { ffc_set_error("ffc_fu_fwd: shape does not fit the fused kernel"); return FFC_ERR_BAD_ARG; }
    int per_sm = (int)(((size_t)227 * 1024) / (pl.smem + 1024));
    if (per_sm > 2048 / pl.nt) per_sm = 2048 / pl.nt;
    if (per_sm > Fu2Cfg<N, CP>::kMinBlocks) per_sm = Fu2Cfg<N, CP>::kMinBlocks;
    if (per_sm < 1) per_sm = 1;
    const int sms = ffc_sm_count();
    const int grid = p.B < sms * per_sm ? p.B : sms * per_sm;
    if (p.training) {
        FFC_CHECK(ffc_memset_async(p.sums, 0, (size_t)4 * p.Cout * sizeof(double), st));
        if (!ffc_fu2_force_two_pass && p.B <= ffc_coop_capacity_blocks<Fu2Coop<N, CP>>(pl.nt, pl.smem))
            return ffc_launch_coop<Fu2Coop<N, CP>>(p.B, pl.nt, pl.smem, st, p);
        FFC_CHECK((ffc_launch<Fu2Fwd<N, CP, 0>>(grid, 1, 1, pl.nt, pl.smem, st, p)));
    }
    return ffc_launch<Fu2Fwd<N, CP, 1>>(grid, 1, 1, pl.nt, pl.smem, st, p);
}

template <int N>
static int fu2_dispatch(const Fu2Params& p, ffc_stream_t st) {
    const int cm = p.Cin > p.Cout ? p.Cin : p.Cout;
    if (cm <= 8) return fu2_launch<N, 8>(p, st);
    if (cm <= 16) return fu2_launch<N, 16>(p, st);
    return fu2_launch<N, 32>(p, st);
}
template <int N>
static bool fu2_fits(int Cin, int Cout) {
    const int cm = Cin > Cout ? Cin : Cout;
    if (cm <= 8) return fu2_plan<N, 8>(Cin, Cout).ok;
    if (cm <= 16) return fu2_plan<N, 16>(Cin, Cout).ok;
    return fu2_plan<N, 32>(Cin, Cout).ok;
}

// 32x32 planes with up to 8 channels: third generation (ffc_fu4.cu; device build only)
bool ffc_fu4_supported(int Cin, int Cout, int H, int W);
int ffc_fu4_launch(const Fu2Params& p, size_t workspace_bytes, ffc_stream_t st);

// the 4x4 planes stay on the first-generation kernel (ffc_fu_fused.cu)
extern "C" int ffc_fu1_supported(int B, int Cin, int Cout, int H, int W);
extern "C" int ffc_fu1_fwd(const float* x, const float* w, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                           const float* residual, float* out,
                           int B, int Cin, int Cout, int H, int W, int training, float eps, float momentum,
                           void* workspace, size_t workspace_bytes, void* stream);

// 1 when ffc_fu_fwd supports the shape (otherwise callers use ffc_rfft2 | ffc_conv2d_fwd | ffc_bn_act_fwd | ffc_irfft2)
extern "C" int ffc_fu_fused_supported(int B, int Cin, int Cout, int H, int W) {
    if (H != W || B < 1 || Cin < 1 || Cout < 1 || Cin > 32 || Cout > 32) return 0;
    switch (H) {
        case 4: return ffc_fu1_supported(B, Cin, Cout, H, W);
        case 8: return fu2_fits<8>(Cin, Cout);
        case 16: return fu2_fits<16>(Cin, Cout);
        case 32: return fu2_fits<32>(Cin, Cout);
        default: return 0;
    }
}

// Workspace ffc_fu_fwd makes the best use of (the minimum stays 4*Cout doubles): room for one float per (sum, image) lets the
// training kernel of ffc_fu4.cu meet its batch statistics without a zeroing launch and without atomics.
extern "C" size_t ffc_fu_workspace_bytes(int B, int Cout) {
    return ((size_t)4 * Cout * sizeof(double) + 255) / 256 * 256 + (size_t)4 * Cout * (B > 0 ? B : 0) * sizeof(float);
}

// Fused FourierUnitSN forward.  w: conv_layer.weight viewed [2*Cout][2*Cin]; gamma/beta/running_*: bn.* [2*Cout];
// save_mean/save_invstd [2*Cout] are written; out = [residual +] irfft2(relu(bn(mix(rfft2(x))))).
// workspace >= 4*Cout doubles.
extern "C" int ffc_fu_fwd(const float* x, const float* w, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                          const float* residual, float* out,
                          int B, int Cin, int Cout, int H, int W, int training, float eps, float momentum,
                          void* workspace, size_t workspace_bytes, void* stream) {
    if (H == 4 && W == 4)
        return ffc_fu1_fwd(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out,
                           B, Cin, Cout, H, W, training, eps, momentum, workspace, workspace_bytes, stream);
    FFC_REQUIRE(x && w && gamma && beta && save_mean && save_invstd && out, "ffc_fu_fwd: null pointer");
    FFC_REQUIRE(training || (running_mean && running_var), "ffc_fu_fwd: eval mode needs running statistics");
    FFC_REQUIRE(B >= 0, "ffc_fu_fwd: negative batch");
    if (B == 0) return FFC_OK;
    FFC_REQUIRE(ffc_fu_fused_supported(B, Cin, Cout, H, W), "ffc_fu_fwd: unsupported shape B=%d Cin=%d Cout=%d %dx%d", B, Cin, Cout, H, W);
    FFC_REQUIRE((((uintptr_t)x | (uintptr_t)out | (uintptr_t)residual) & 15) == 0, "ffc_fu_fwd: x/out/residual must be 16-byte aligned");
    if (!(workspace && workspace_bytes >= (size_t)4 * Cout * sizeof(double))) { ffc_set_error("ffc_fu_fwd: workspace too small"); return FFC_ERR_WORKSPACE; }
    Fu2Params p;
    p.x = x; p.w = w; p.gamma = gamma; p.beta = beta; p.running_mean = running_mean; p.running_var = running_var;
    p.save_mean = save_mean; p.save_invstd = save_invstd; p.residual = residual; p.out = out;
    p.sums = (double*)workspace; p.B = B; p.Cin = Cin; p.Cout = Cout; p.training = training;
    p.eps = eps; p.momentum = momentum;
    ffc_stream_t st = (ffc_stream_t)stream;
    if (ffc_fu4_supported(Cin, Cout, H, W)) return ffc_fu4_launch(p, workspace_bytes, st);       // warp-private transforms (ffc_fu4.cu)
    switch (H) {
        case 8: return fu2_dispatch<8>(p, st);
        case 16: return fu2_dispatch<16>(p, st);
        default: return fu2_dispatch<32>(p, st);
    }
}
