// Fused FourierUnitSN backward (autograd of layers/ffc/fourier_unity.py:32-58), planes 8x8 .. 32x32, as ONE
// cooperative kernel with one image per CTA:
//   part 0   x    -> rfft2                       -> S   (region X, Cin planes)
//            dout -> rfft2, interior bins x2     -> G   (region G, Cout planes; the adjoint of the c2r transform)
//            Y = W S recomputed per (channel, bin), y^ = (Y - mean) * invstd, ReLU mask from y^*gamma + beta,
//            dZ = mask * G  (in place over G),  sum(dZ), sum(dZ * y^) per channel -> double atomics
//   -------- grid barrier (BatchNorm backward needs the two sums over the whole batch) --------
//   part 1   dY = gamma*invstd * (dZ - mean(dZ) - y^ * mean(dZ y^))   (y^ recomputed from S; in place over dZ)
//            dW += dY S^T   (register tiles over bins, shared-memory reduction, one float atomic per entry and CTA)
//            dS = W^T dY    (in place over S)
//            dS -> adjoint of rfft2 (interior bins x1/2, complex inverse columns, c2r rows) -> dx
// The spectra never leave shared memory.  Shared values are unnormalised transforms (N * the ortho ones); the 1/N
// factors ride in the weights, in the reduction epilogues and in the final row scale.
#include "ffc_fu2.cuh"


template <int N, int CP>
struct Fu2Bwd {
    typedef Fu2BwdParams Params;
    typedef Fu2G<N> G;
    typedef Fu2Cfg<N, CP> CFG;
    static constexpr int kThreads = CFG::kThreads;
    static constexpr int kMinBlocks = CFG::kMinBlocks;
    static constexpr int OG = CFG::OG < 8 ? CFG::OG : 8;        // channels per thread in the per-bin phases
    static constexpr int BINS = G::BINS, RF2 = G::REGION / 2;
    struct Acc { float v[4]; };

    // shared memory: X planes | G planes | wq | tw | 4 float2 constant tables | red (nt*8 floats)
    struct Lay { float* xreg; float* greg; float4* wq; float2* tw; float2* k0; float2* k1; float2* k2; float2* k3; float* red; };
    static size_t smem_floats(int Cin, int Cout, int nt) {
        return (size_t)(Cin + Cout) * G::REGION + (size_t)Cout * CP * 4 + 2 * N + 8 * Cout + (size_t)nt * 8 + 8;
    }
    static FFC_DEVICE Lay layout(const Params& p, float* smem) {
        Lay l;
        l.xreg = smem;
        l.greg = smem + (size_t)p.Cin * G::REGION;
        float* q = smem + (size_t)(p.Cin + p.Cout) * G::REGION;
        l.wq = reinterpret_cast<float4*>(q); q += (size_t)p.Cout * CP * 4;
        l.tw = reinterpret_cast<float2*>(q); q += 2 * N;
        l.k0 = reinterpret_cast<float2*>(q); q += 2 * p.Cout;
        l.k1 = reinterpret_cast<float2*>(q); q += 2 * p.Cout;
        l.k2 = reinterpret_cast<float2*>(q); q += 2 * p.Cout;
        l.k3 = reinterpret_cast<float2*>(q); q += 2 * p.Cout;
        l.red = q;
        return l;
    }

    // Y (one complex channel) of one bin from the bin's inputs s[] and the weight row wrow[c]
    static FFC_DEVICE float2 mix_one(const float4* wrow, const float2* s) {
        float2 pa = make_float2(0.f, 0.f), pb = make_float2(0.f, 0.f);
        FFC_UNROLL
        for (int c = 0; c < CP; ++c) {
            const float4 q = wrow[c];
            pa = ffc_fma2(make_float2(q.x, q.y), s[c], pa);
            pb = ffc_fma2(make_float2(q.z, q.w), s[c], pb);
        }
        return make_float2(pa.x + pb.y, pb.x + pa.y);
    }

    static FFC_DEVICE void part0(const Params& p, const BlockCtx& ctx, float* smem) {
        const Lay l = layout(p, smem);
        const int Cin = p.Cin, Cout = p.Cout, img = ctx.bx;
        FFC_PHASE {
            fu2_prologue<N, CP>(p.w, Cin, Cout, l.wq, l.tw, tid, ctx.nt);
            // k0 = invstd, k1 = -mean*invstd (y^ = Y*k0 + k1), k2 = gamma, k3 = beta (z = y^*k2 + k3)
            for (int o = tid; o < 2 * Cout; o += ctx.nt) {
                const float is = FFC_LDG(p.save_invstd + o);
                reinterpret_cast<float*>(l.k0)[o] = is;
                reinterpret_cast<float*>(l.k1)[o] = -FFC_LDG(p.save_mean + o) * is;
                reinterpret_cast<float*>(l.k2)[o] = FFC_LDG(p.gamma + o);
                reinterpret_cast<float*>(l.k3)[o] = FFC_LDG(p.beta + o);
            }
            fu2_load_rows<N>(tid, ctx.nt, Cin * N, p.x + (size_t)img * Cin * N * N, l.xreg);
            fu2_load_rows<N>(tid, ctx.nt, Cout * N, p.dout + (size_t)img * Cout * N * N, l.greg);
        } FFC_SYNC;
        FFC_PHASE {
            fu2_rows_fwd<N, false>(tid, ctx.nt, Cin * N, l.xreg);
            fu2_rows_fwd<N, true>(tid, ctx.nt, Cout * N, l.greg);
        } FFC_SYNC;
        FU2_COLS_FWD(N, Cin + Cout, l.xreg, l.tw);          // the G planes follow the X planes
        // ReLU mask, dZ and the two BatchNorm-backward sums: thread = (complex channel, bin slice)
        FFC_PHASE {
            const int S = ctx.nt / Cout;
            const int o2 = tid / S, sl = tid % S;
            Acc acc;
            acc.v[0] = acc.v[1] = acc.v[2] = acc.v[3] = 0.f;
            if (o2 < Cout) {
                const float4* wrow = l.wq + (size_t)o2 * CP;
                const float2 k0 = l.k0[o2], k1 = l.k1[o2], k2 = l.k2[o2], k3 = l.k3[o2];
                float2* gp = reinterpret_cast<float2*>(l.greg) + (size_t)o2 * RF2;
                for (int bin = sl; bin < BINS; bin += S) {
                    const int off = fu2_bin_off<N>(bin);
                    const float2* sp = reinterpret_cast<const float2*>(l.xreg) + off;
                    float2 s[CP];
                    FFC_UNROLL
                    for (int c = 0; c < CP; ++c) { s[c] = make_float2(0.f, 0.f); if (c < Cin) s[c] = sp[c * RF2]; }
                    const float2 yh = ffc_fma2(mix_one(wrow, s), k0, k1);
                    const float2 z = ffc_fma2(yh, k2, k3);
                    const float2 g = gp[off];
                    const float2 dz = make_float2(z.x > 0.f ? g.x : 0.f, z.y > 0.f ? g.y : 0.f);
                    gp[off] = dz;
                    acc.v[0] += dz.x; acc.v[1] = fmaf(dz.x, yh.x, acc.v[1]);
                    acc.v[2] += dz.y; acc.v[3] = fmaf(dz.y, yh.y, acc.v[3]);
                }
            }
            FFC_UNROLL
            for (int j = 0; j < 4; ++j) l.red[(size_t)tid * 4 + j] = acc.v[j];
        } FFC_SYNC;
        fu2_flush_sums(ctx, l.red, Cout, p.sums);
    }

    static FFC_DEVICE void part1(const Params& p, const BlockCtx& ctx, float* smem) {
        const Lay l = layout(p, smem);
        const int Cin = p.Cin, Cout = p.Cout, img = ctx.bx;
        const float inv_n = 1.0f / (float)N;
        // constants of dY = a*(dZ - c1 - y^*c2): k2 <- a = gamma*invstd, k3 <- c1, and c2 in the spare red slots
        float2* c2t = reinterpret_cast<float2*>(l.red);
        FFC_PHASE {
            const double count = (double)p.B * BINS;
            for (int o = tid; o < 2 * Cout; o += ctx.nt) {
                const double s1 = p.sums[o], s2 = p.sums[2 * Cout + o];          // of the unnormalised dZ (= N * true)
                if (ctx.bx == 0) { p.dbeta[o] = (float)(s1 * inv_n); p.dgamma[o] = (float)(s2 * inv_n); }
                const float a = FFC_LDG(p.gamma + o) * FFC_LDG(p.save_invstd + o);
                reinterpret_cast<float*>(l.k2)[o] = a;
                reinterpret_cast<float*>(l.k3)[o] = p.training ? (float)(s1 / count) : 0.f;
                reinterpret_cast<float*>(c2t)[o] = p.training ? (float)(s2 / count) : 0.f;
            }
        } FFC_SYNC;
        // dY in place over dZ: thread = (bin, group of OG channels); y^ recomputed from S
        FFC_PHASE {
            const int ngrp = (Cout + OG - 1) / OG;
            for (int it = tid; it < ngrp * BINS; it += ctx.nt) {
                const int bin = it % BINS, grp = it / BINS;
                const int off = fu2_bin_off<N>(bin);
                const float2* sp = reinterpret_cast<const float2*>(l.xreg) + off;
                float2* gp = reinterpret_cast<float2*>(l.greg) + off;
                float2 s[CP];
                FFC_UNROLL
                for (int c = 0; c < CP; ++c) { s[c] = make_float2(0.f, 0.f); if (c < Cin) s[c] = sp[c * RF2]; }
#pragma unroll 4
                for (int j = 0; j < OG; ++j) {
                    const int o2 = grp * OG + j;
                    if (o2 >= Cout) break;
                    const float2 yh = ffc_fma2(mix_one(l.wq + (size_t)o2 * CP, s), l.k0[o2], l.k1[o2]);
                    const float2 dz = gp[o2 * RF2];
                    const float2 c1 = l.k3[o2], c2 = c2t[o2], a = l.k2[o2];
                    // a * (dz - c1 - yh*c2)
                    const float2 t = ffc_sub2(ffc_sub2(dz, c1), ffc_mul2(yh, c2));
                    gp[o2 * RF2] = ffc_mul2(a, t);
                }
            }
        } FFC_SYNC;
        // dW partials: thread = (tile of 2 output x 1 input complex channels, bin slice), 8 sums per thread
        const int otiles = (Cout + 1) / 2;
        const int ntile = otiles * Cin;
        FFC_PHASE {
            const int slices = ctx.nt / ntile;
            const int tile = tid / slices, sl = tid % slices;
            float2 a1[2], a2[2];
            a1[0] = a1[1] = a2[0] = a2[1] = make_float2(0.f, 0.f);
            if (tile < ntile) {
                const int c = tile % Cin, oa = 2 * (tile / Cin), ob = (oa + 1 < Cout) ? oa + 1 : oa;
                const float2* sp = reinterpret_cast<const float2*>(l.xreg) + (size_t)c * RF2;
                const float2* ga = reinterpret_cast<const float2*>(l.greg) + (size_t)oa * RF2;
                const float2* gb = reinterpret_cast<const float2*>(l.greg) + (size_t)ob * RF2;
                for (int bin = sl; bin < BINS; bin += slices) {
                    const int off = fu2_bin_off<N>(bin);
                    const float2 s = sp[off], da = ga[off], db = gb[off];
                    const float2 sx = make_float2(s.x, s.x), sy = make_float2(s.y, s.y);
                    a1[0] = ffc_fma2(da, sx, a1[0]); a2[0] = ffc_fma2(da, sy, a2[0]);
                    a1[1] = ffc_fma2(db, sx, a1[1]); a2[1] = ffc_fma2(db, sy, a2[1]);
                }
            }
            float* r = l.red + (size_t)tid * 8;
            r[0] = a1[0].x; r[1] = a1[0].y; r[2] = a2[0].x; r[3] = a2[0].y;
            r[4] = a1[1].x; r[5] = a1[1].y; r[6] = a2[1].x; r[7] = a2[1].y;
        } FFC_SYNC;
        FFC_PHASE {      // slices -> one float atomic per weight entry
            const int slices = ctx.nt / ntile;
            for (int e = tid; e < ntile * 8; e += ctx.nt) {
                const int k = e % 8, tile = e / 8;
                const int c = tile % Cin, o2 = 2 * (tile / Cin) + (k >> 2);
                if (o2 < Cout) {
                    float s = 0.f;
                    for (int sl = 0; sl < slices; ++sl) s += l.red[((size_t)tile * slices + sl) * 8 + k];
                    // k&3: 0 -> dW[2o][2c], 1 -> dW[2o+1][2c], 2 -> dW[2o][2c+1], 3 -> dW[2o+1][2c+1]
                    const int row = 2 * o2 + (k & 1), col = 2 * c + ((k >> 1) & 1);
                    ffc_atomic_add(p.dw + (size_t)row * 2 * Cin + col, s * inv_n * inv_n);
                }
            }
        } FFC_SYNC;
        // dS = W^T dY in place over S: thread = (bin, group of OG input channels)
        FFC_PHASE {
            const int ngrp = (Cin + OG - 1) / OG;
            for (int it = tid; it < ngrp * BINS; it += ctx.nt) {
                const int bin = it % BINS, grp = it / BINS;
                const int off = fu2_bin_off<N>(bin);
                const float2* gp = reinterpret_cast<const float2*>(l.greg) + off;
                float2* sp = reinterpret_cast<float2*>(l.xreg) + off;
                float2 d[CP], dsw[CP];
                FFC_UNROLL
                for (int o = 0; o < CP; ++o) {
                    d[o] = make_float2(0.f, 0.f);
                    if (o < Cout) d[o] = gp[o * RF2];
                    dsw[o] = make_float2(d[o].y, d[o].x);
                }
#pragma unroll 4
                for (int j = 0; j < OG; ++j) {
                    const int c = grp * OG + j;
                    if (c >= Cin) break;
                    float2 pa = make_float2(0.f, 0.f), pb = make_float2(0.f, 0.f);
                    FFC_UNROLL
                    for (int o = 0; o < CP; ++o) {
                        if (o < Cout) {
                            const float4 q = l.wq[o * CP + c];
                            pa = ffc_fma2(make_float2(q.x, q.y), d[o], pa);       // W[2o][2c] dYre, W[2o+1][2c+1] dYim
                            pb = ffc_fma2(make_float2(q.z, q.w), dsw[o], pb);     // W[2o+1][2c] dYim, W[2o][2c+1] dYre
                        }
                    }
                    sp[c * RF2] = make_float2(pa.x + pb.x, pa.y + pb.y);
                }
            }
        } FFC_SYNC;
        Fu2Bn nobn; nobn.a = nullptr; nobn.b = nullptr;
        FU2_COLS_INV(N, false, Cin, l.xreg, l.tw, nobn);
        FFC_PHASE { fu2_rows_inv<N, true>(tid, ctx.nt, Cin * N, l.xreg, inv_n); } FFC_SYNC;
        FFC_PHASE { fu2_store_rows<N>(tid, ctx.nt, Cin * N, l.xreg, nullptr, p.dx + (size_t)img * Cin * N * N); } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Fu2BwdPlan { int nt; size_t smem; bool ok; };

template <int N, int CP>
static Fu2BwdPlan fu2_bwd_plan(int Cin, int Cout) {
    typedef Fu2Bwd<N, CP> K;
    typedef Fu2G<N> G;
    Fu2BwdPlan pl; pl.ok = false;
    const int CB = Cin > Cout ? Cin : Cout;
    int want = CB * G::Wf * G::N2;
    const int mix_items = ((CB + K::OG - 1) / K::OG) * G::BINS;
    if (want < mix_items) want = mix_items;
    if (want < CB * N) want = CB * N;
    int nt = (want + 31) / 32 * 32;
    if (nt > K::kThreads) {
        const int rounds = (want + K::kThreads - 1) / K::kThreads;
        nt = ((want + rounds - 1) / rounds + 31) / 32 * 32;
        if (nt > K::kThreads) nt = K::kThreads;
    }
    const int ntile = ((Cout + 1) / 2) * Cin;
    int need = (Cout * 4 + 31) / 32 * 32;
    if (need < (ntile + 31) / 32 * 32) need = (ntile + 31) / 32 * 32;       // at least one slice per dW tile
    if (nt < need) nt = need;
    if (nt > K::kThreads) return pl;
    pl.nt = nt;
    pl.smem = K::smem_floats(Cin, Cout, nt) * 4;
    pl.ok = pl.smem <= (size_t)227 * 1024;
    return pl;
}

template <int N, int CP>
static int fu2_bwd_launch(const Fu2BwdParams& p, ffc_stream_t st, bool query) {
    const Fu2BwdPlan pl = fu2_bwd_plan<N, CP>(p.Cin, p.Cout);
    if (!pl.ok) return query ? 0 : (ffc_set_error("ffc_fu_bwd: shape does not fit the fused kernel"), FFC_ERR_BAD_ARG);
    const int cap = ffc_coop_capacity_blocks<Fu2Bwd<N, CP>>(pl.nt, pl.smem);
    if (query) return p.B <= cap ? 1 : 0;
    if (p.B > cap) { ffc_set_error("ffc_fu_bwd: batch %d exceeds the %d co-resident images of the fused kernel", p.B, cap); return FFC_ERR_BAD_ARG; }
    FFC_CHECK(ffc_memset_async(p.sums, 0, (size_t)4 * p.Cout * sizeof(double), st));
    FFC_CHECK(ffc_memset_async(p.dw, 0, (size_t)4 * p.Cout * p.Cin * sizeof(float), st));
    return ffc_launch_coop<Fu2Bwd<N, CP>>(p.B, pl.nt, pl.smem, st, p);
}

template <int N>
static int fu2_bwd_dispatch(const Fu2BwdParams& p, ffc_stream_t st, bool query) {
    const int cm = p.Cin > p.Cout ? p.Cin : p.Cout;
    if (cm <= 8) return fu2_bwd_launch<N, 8>(p, st, query);
    if (cm <= 16) return fu2_bwd_launch<N, 16>(p, st, query);
    return fu2_bwd_launch<N, 32>(p, st, query);
}

static int fu2_bwd_entry(const Fu2BwdParams& p, int H, ffc_stream_t st, bool query) {
    switch (H) {
        case 8: return fu2_bwd_dispatch<8>(p, st, query);
        case 16: return fu2_bwd_dispatch<16>(p, st, query);
        case 32: return fu2_bwd_dispatch<32>(p, st, query);
        default: return query ? 0 : (ffc_set_error("ffc_fu_bwd: unsupported plane %dx%d", H, H), FFC_ERR_BAD_ARG);
    }
}

// 32x32 planes with up to 8 channels: the warp-private backward of ffc_fu4.cu (device build, workspace of ffc_fu_bwd_workspace_bytes)
bool ffc_fu4_bwd_supported(int B, int Cin, int Cout, int H, int W, size_t workspace_bytes);
int ffc_fu4_bwd_launch(const Fu2BwdParams& p, ffc_stream_t st);

// 1 when ffc_fu_bwd handles the shape on the current device: H == W in {8,16,32}, Cin, Cout <= 32, and all B images
// co-resident (one CTA each) so that the two BatchNorm sums can cross a grid barrier; otherwise callers use the
// general form (ffc_rfft2 | ffc_conv2d_* | ffc_bn_act_bwd | ffc_irfft2).
extern "C" int ffc_fu_bwd_supported(int B, int Cin, int Cout, int H, int W) {
    if (H != W || B < 1 || Cin < 1 || Cout < 1 || Cin > 32 || Cout > 32) return 0;
    Fu2BwdParams p;
    p.B = B; p.Cin = Cin; p.Cout = Cout;
    return fu2_bwd_entry(p, H, nullptr, true);
}

// Fused FourierUnitSN backward.  x, dout as in the forward; w = conv_layer.weight [2*Cout][2*Cin]; gamma, beta = bn.*;
// save_mean / save_invstd = what ffc_fu_fwd wrote (batch statistics in training mode, running ones in eval mode);
// outputs dx (B,Cin,H,W), dw [2*Cout][2*Cin], dgamma, dbeta [2*Cout].  A residual passes its gradient through
// unchanged (dresidual = dout) and is not an argument.  workspace >= 4*Cout*8 bytes.
extern "C" int ffc_fu_bwd(const float* x, const float* dout, const float* w, const float* gamma, const float* beta,
                          const float* save_mean, const float* save_invstd,
                          float* dx, float* dw, float* dgamma, float* dbeta,
                          int B, int Cin, int Cout, int H, int W, int training,
                          void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(x && dout && w && gamma && beta && save_mean && save_invstd && dx && dw && dgamma && dbeta, "ffc_fu_bwd: null pointer");
    FFC_REQUIRE(B >= 0 && H == W && Cin >= 1 && Cout >= 1 && Cin <= 32 && Cout <= 32, "ffc_fu_bwd: unsupported shape B=%d Cin=%d Cout=%d %dx%d", B, Cin, Cout, H, W);
    FFC_REQUIRE((((uintptr_t)x | (uintptr_t)dout | (uintptr_t)dx) & 15) == 0, "ffc_fu_bwd: x/dout/dx must be 16-byte aligned");
    if (!(workspace && workspace_bytes >= (size_t)4 * Cout * sizeof(double))) { ffc_set_error("ffc_fu_bwd: workspace too small"); return FFC_ERR_WORKSPACE; }
    ffc_stream_t st = (ffc_stream_t)stream;
    Fu2BwdParams p;
    p.x = x; p.dout = dout; p.w = w; p.gamma = gamma; p.beta = beta; p.save_mean = save_mean; p.save_invstd = save_invstd;
    p.dx = dx; p.dw = dw; p.dgamma = dgamma; p.dbeta = dbeta; p.sums = (double*)workspace;
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.training = training;
    if (B == 0) {
        FFC_CHECK(ffc_memset_async(dw, 0, (size_t)4 * Cout * Cin * sizeof(float), st));
        FFC_CHECK(ffc_memset_async(dgamma, 0, (size_t)2 * Cout * sizeof(float), st));
        return ffc_memset_async(dbeta, 0, (size_t)2 * Cout * sizeof(float), st);
    }
    if (ffc_fu4_bwd_supported(B, Cin, Cout, H, W, workspace_bytes)) return ffc_fu4_bwd_launch(p, st);
    return fu2_bwd_entry(p, H, st, false);
}
