// FourierUnitSN forward for spectra that do not fit one CTA's shared memory (64x64 / 128x128 planes, or more than 32
// channels): the L2-staged form, third generation (layers/ffc/fourier_unity.py:32-58).
//
//   x --Fu3Rfft2--> S --Fu3Mix (tensor cores) + BN statistics--> Y --Fu3Irfft2 (BN + ReLU on load)--> [residual +] out
//
// The plane-wise transforms want one H x W plane per CTA, the channel mix wants all 2C channels of a bin: for 32 channels
// at 128 x 128 one image's spectrum is 2.1 MB, so the two tilings meet in L2 instead of shared memory.  The batch is
// walked in CHUNKS of images whose spectrum (<= ~24 MB) stays resident in the 126 MB L2 between the kernel that writes it
// and the kernel that reads it, and the scratch buffer is reused by every chunk; in training mode the mixed spectrum Y of
// the whole batch has to wait for the batch statistics and makes one round trip.  Three kernels per chunk (+ statistics
// finalisation), 4-6 tensor passes where the first-generation general form made 7-8 in five slower kernels.
//
// Scratch layout ("shared-memory image"): plane (b, c) is N rows of RS = N + 4 floats = SPS = N/2 + 2 complex slots, bins
// v = 0..N/2 of spectrum row p followed by one zero pad slot, rows in the permuted u order of the two-level column
// transform (ffc_fu2.cuh).  The transforms copy a plane between shared and global memory with linear 16-byte accesses, and
// the mix sees every plane as a flat list of NB = N * SPS complex "bins" (pads are computed and ignored; statistics skip
// them).  Mix, BatchNorm and ReLU are pointwise in (u, v), so neither the permutation nor the pads are ever undone.
#include "ffc_fu2.cuh"
#include "ffc_fu3.cuh"
#ifndef FFC_EMU
#include "ffc_umma.cuh"      // mbarrier + bulk-copy wrappers (the plane kernels move whole rows / planes with cp.async.bulk)
#endif

template <int N> struct Fu3G {
    typedef Fu2G<N> G;
    static constexpr int NB = N * G::SPS;                 // complex slots per plane (pads included)
    static constexpr int REGION = G::REGION;              // floats per plane = 2 * NB
    // planes per CTA and CTA width: the row transforms keep a whole row in registers (thread per row), so a CTA has N
    // threads per plane; small planes are grouped until a CTA has 128 threads
    static constexpr int P = N >= 128 ? 1 : 128 / N;
    static constexpr int kThreads = P * N;
};

// ------------------------------------------------------------------------------------------------------------------
// column passes, two work items per loop iteration.  Same arithmetic and item enumeration as fu2_cols_strided /
// fu2_cols_contig (ffc_fu2.cuh); the loads of both items are issued before either butterfly, so with only ~3 warps per
// scheduler (three 128-thread CTAs per SM, one 67 KB plane each) the shared-memory latency of one item hides behind the
// arithmetic of the other (ncu r02g: short-scoreboard was the top stall of these passes).
// ------------------------------------------------------------------------------------------------------------------
template <int N, int SIGN, bool BN, bool STRIDED>
FFC_DEVICE void fu3_cols_pass(int tid, int nt, int np, float* planes, const float2* tw, Fu2Bn bn) {
    typedef Fu2G<N> G;
    constexpr int N1 = G::N1, N2 = G::N2, Wf = G::Wf;
    constexpr int R = STRIDED ? N1 : N2;          // points per item
    constexpr int S = STRIDED ? N2 : N1;          // items per column
    const int total = np * Wf * S;
    for (int it0 = tid; it0 < total; it0 += 2 * nt) {
        float2 c[2][R];
        float2* col[2];
        int pls[2], sub[2];
        bool on[2];
        FFC_UNROLL
        for (int u = 0; u < 2; ++u) {
            const int it = it0 + u * nt;
            on[u] = it < total;
            int v = 0, k = 0, pl = 0;
            if (on[u]) {
                if (it < np * G::M * S) { v = it % G::M; k = (it / G::M) % S; pl = it / (G::M * S); }
                else { const int j = it - np * G::M * S; v = G::M; k = j % S; pl = j / S; }
            }
            pls[u] = pl; sub[u] = k;
            col[u] = reinterpret_cast<float2*>(planes + pl * G::REGION) + v + (STRIDED ? k * G::SPS : (N2 * k) * G::SPS);
            if (on[u]) {
                FFC_UNROLL
                for (int i = 0; i < R; ++i) c[u][i] = col[u][(STRIDED ? N2 * i : i) * G::SPS];
            }
        }
        FFC_UNROLL
        for (int u = 0; u < 2; ++u) {
            if (!on[u]) continue;
            if (BN) {
                const float2 a = bn.a[pls[u]], b = bn.b[pls[u]];
                FFC_UNROLL
                for (int i = 0; i < R; ++i) c[u][i] = fu2_bn_relu(c[u][i], a, b);
            }
            ffc_fft_regs<R, SIGN>(c[u]);
        }
        FFC_UNROLL
        for (int u = 0; u < 2; ++u) {
            if (!on[u]) continue;
            FFC_UNROLL
            for (int i = 0; i < R; ++i) {
                float2 o = c[u][i];
                // twiddles between the two levels: after the strided level of the forward transform, after the contiguous
                // level of the inverse one (ffc_fu2.cuh)
                if (STRIDED && SIGN < 0 && N2 > 1 && i > 0) o = ffc_cmul_tw<-1>(o, tw[sub[u] * i]);
                if (!STRIDED && SIGN > 0 && i > 0) o = ffc_cmul_tw<+1>(o, tw[i * sub[u]]);
                col[u][(STRIDED ? N2 * i : i) * G::SPS] = o;
            }
        }
    }
}
#define FU3_COLS_FWD(N, np, planes, tw)                                                                              \
    do {                                                                                                             \
        Fu2Bn nobn_; nobn_.a = nullptr; nobn_.b = nullptr;                                                           \
        FFC_PHASE { fu3_cols_pass<N, -1, false, true>(tid, ctx.nt, np, planes, tw, nobn_); } FFC_SYNC;               \
        if constexpr (Fu2G<N>::N2 > 1) {                                                                             \
            FFC_PHASE { fu3_cols_pass<N, -1, false, false>(tid, ctx.nt, np, planes, tw, nobn_); } FFC_SYNC;          \
        }                                                                                                            \
    } while (0)
#define FU3_COLS_INV(N, BNF, np, planes, tw, bn)                                                                     \
    do {                                                                                                             \
        Fu2Bn nobn_; nobn_.a = nullptr; nobn_.b = nullptr;                                                           \
        if constexpr (Fu2G<N>::N2 > 1) {                                                                             \
            FFC_PHASE { fu3_cols_pass<N, +1, BNF, false>(tid, ctx.nt, np, planes, tw, bn); } FFC_SYNC;               \
            FFC_PHASE { fu3_cols_pass<N, +1, false, true>(tid, ctx.nt, np, planes, tw, nobn_); } FFC_SYNC;           \
        } else {                                                                                                     \
            FFC_PHASE { fu3_cols_pass<N, +1, BNF, true>(tid, ctx.nt, np, planes, tw, bn); } FFC_SYNC;                \
        }                                                                                                            \
    } while (0)

// ------------------------------------------------------------------------------------------------------------------
// plane transforms
// ------------------------------------------------------------------------------------------------------------------
struct Fu3FwdFftParams {
    const float* x;        // (nplanes, N, N)
    float* spec;           // (nplanes, N, RS)
    int nplanes;
    // MASK form (backward, nplanes = B * cout): the transformed plane is multiplied by the ReLU mask of the forward,
    // recomputed from the saved BatchNorm input y with the forward's own folded constants (y * a + b > 0), and the two
    // BatchNorm-backward sums of the plane's channel are accumulated: sums[n] += sum(g), sums[2*cout + n] += sum(g * xhat)
    const float* y;        // (nplanes, N, RS)
    const float* bn_a; const float* bn_b; const float* mean; const float* invstd;     // [2*cout]
    double* sums;          // [4*cout]
    int cout;
};

template <int N, bool ADJ, bool MASK = false>
struct Fu3Rfft2 {
    typedef Fu3FwdFftParams Params;
    typedef Fu2G<N> G;
    static constexpr int P = Fu3G<N>::P;
    static constexpr int kThreads = Fu3G<N>::kThreads;
    static constexpr int kMinBlocks = (N == 128) ? 3 : (N == 64 ? 5 : 4);
    static constexpr int kS2 = (N / 4 < 8) ? N / 4 : 8;      // MASK: second-level fan-in of the per-plane reduction
    static size_t smem_bytes() { return ((size_t)P * G::REGION + 2 * N) * 4 + 16 + (MASK ? (size_t)(kThreads * 4 + P * 4 * kS2) * 4 : 0); }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        const int plane0 = ctx.bx * P;
        const int np = (p.nplanes - plane0) < P ? (p.nplanes - plane0) : P;
        float* planes = smem;
        float2* tw = reinterpret_cast<float2*>(smem + (size_t)P * G::REGION);
        float* red = reinterpret_cast<float*>(tw + N) + 4;      // MASK: [kThreads][4] partials, then [P * 4 * kS2]
#ifndef FFC_EMU
        if (MASK && threadIdx.x == 0)                             // the saved planes are needed last: bring them into L2 meanwhile
            umma::bulk_prefetch_l2(p.y + (size_t)plane0 * G::REGION, (uint32_t)(np * G::REGION * 4));
        // Device: every thread owns one row (kThreads == P * N) and fetches it with ONE bulk copy (cp.async.bulk, N*4 bytes
        // into the padded shared-memory row); all rows of the CTA are in flight at once and complete on one mbarrier, so the
        // load costs one memory latency instead of one per batch of register-staged loads.
        uint64_t* bar = reinterpret_cast<uint64_t*>(tw + N);
        {
            const int tid = (int)threadIdx.x;
            if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_proxy_async_smem(); }   // (the cluster-scope release fence flushes L1: CCTL.IVALL)
            for (int k = tid; k < N; k += ctx.nt) tw[k] = c_tw128[k * (FFC_TW_N / N)];
            __syncthreads();
            if (tid == 0) umma::mbar_arrive_expect_tx(bar, (uint32_t)(np * N * N * 4));
            for (int r = tid; r < np * N; r += ctx.nt)
                umma::bulk_g2s(planes + (size_t)r * G::RS, p.x + ((size_t)plane0 * N + r) * N, (uint32_t)(N * 4), bar);
            umma::mbar_wait(bar, 0);
        }
#else
        FFC_PHASE {
            for (int k = tid; k < N; k += ctx.nt) tw[k] = c_tw128[k * (FFC_TW_N / N)];
            fu2_load_rows<N>(tid, ctx.nt, np * N, p.x + (size_t)plane0 * N * N, planes);
        } FFC_SYNC;
#endif
        FFC_PHASE {
            fu2_rows_fwd<N, ADJ>(tid, ctx.nt, np * N, planes);
            for (int r = tid; r < np * N; r += ctx.nt)           // the pad slot of every row (read by the mix, never used)
                reinterpret_cast<float2*>(planes + (size_t)r * G::RS)[G::M + 1] = make_float2(0.f, 0.f);
        } FFC_SYNC;
        FU3_COLS_FWD(N, np, planes, tw);
        if constexpr (MASK) {
            // thread t works on plane t / N: g = transformed dout * [y * a + b > 0], partial sums of g and g * xhat (re | im)
            FFC_PHASE {
                const int pl = tid / N, j = tid % N;
                float s_re = 0.f, s_im = 0.f, q_re = 0.f, q_im = 0.f;
                if (pl < np) {
                    const int o = (plane0 + pl) % p.cout;
                    const float ar = FFC_LDG(p.bn_a + 2 * o), ai = FFC_LDG(p.bn_a + 2 * o + 1), br = FFC_LDG(p.bn_b + 2 * o), bi = FFC_LDG(p.bn_b + 2 * o + 1);
                    const float mr = FFC_LDG(p.mean + 2 * o), mi = FFC_LDG(p.mean + 2 * o + 1), ir = FFC_LDG(p.invstd + 2 * o), ii = FFC_LDG(p.invstd + 2 * o + 1);
                    const float4* y4 = reinterpret_cast<const float4*>(p.y + (size_t)(plane0 + pl) * G::REGION);
                    float4* g4 = reinterpret_cast<float4*>(planes + (size_t)pl * G::REGION);
                    constexpr int PER = G::REGION / 4, LDU = 4;
                    int i = j;
                    for (; i + (LDU - 1) * N < PER; i += LDU * N) {
                        float4 yv[LDU];
                        FFC_UNROLL
                        for (int u = 0; u < LDU; ++u) yv[u] = FFC_LDG(y4 + i + u * N);
                        FFC_UNROLL
                        for (int u = 0; u < LDU; ++u) {
                            float4 g = g4[i + u * N];
                            g.x = fmaf(yv[u].x, ar, br) > 0.f ? g.x : 0.f; g.y = fmaf(yv[u].y, ai, bi) > 0.f ? g.y : 0.f;
                            g.z = fmaf(yv[u].z, ar, br) > 0.f ? g.z : 0.f; g.w = fmaf(yv[u].w, ai, bi) > 0.f ? g.w : 0.f;
                            g4[i + u * N] = g;
                            s_re += g.x + g.z; s_im += g.y + g.w;
                            q_re = fmaf(g.x, (yv[u].x - mr) * ir, q_re); q_re = fmaf(g.z, (yv[u].z - mr) * ir, q_re);
                            q_im = fmaf(g.y, (yv[u].y - mi) * ii, q_im); q_im = fmaf(g.w, (yv[u].w - mi) * ii, q_im);
                        }
                    }
                    for (; i < PER; i += N) {
                        const float4 yv = FFC_LDG(y4 + i);
                        float4 g = g4[i];
                        g.x = fmaf(yv.x, ar, br) > 0.f ? g.x : 0.f; g.y = fmaf(yv.y, ai, bi) > 0.f ? g.y : 0.f;
                        g.z = fmaf(yv.z, ar, br) > 0.f ? g.z : 0.f; g.w = fmaf(yv.w, ai, bi) > 0.f ? g.w : 0.f;
                        g4[i] = g;
                        s_re += g.x + g.z; s_im += g.y + g.w;
                        q_re = fmaf(g.x, (yv.x - mr) * ir, q_re); q_re = fmaf(g.z, (yv.z - mr) * ir, q_re);
                        q_im = fmaf(g.y, (yv.y - mi) * ii, q_im); q_im = fmaf(g.w, (yv.w - mi) * ii, q_im);
                    }
                }
                red[tid * 4 + 0] = s_re; red[tid * 4 + 1] = s_im; red[tid * 4 + 2] = q_re; red[tid * 4 + 3] = q_im;
            } FFC_SYNC;
            FFC_PHASE {
                if (tid < np * 4 * kS2) {
                    const int part = tid % kS2, jj = (tid / kS2) % 4, pl = tid / (4 * kS2);
                    float a = 0.f;
                    for (int t = part; t < N; t += kS2) a += red[(pl * N + t) * 4 + jj];
                    red[kThreads * 4 + tid] = a;
                }
            } FFC_SYNC;
            FFC_PHASE {
                if (tid < np * 4) {
                    const int jj = tid % 4, pl = tid / 4;
                    double a = 0.0;
                    for (int part = 0; part < kS2; ++part) a += (double)red[kThreads * 4 + (pl * 4 + jj) * kS2 + part];
                    const int o = (plane0 + pl) % p.cout;
                    ffc_atomic_add(p.sums + (jj >> 1) * 2 * p.cout + 2 * o + (jj & 1), a);
                }
            } FFC_SYNC;
        }
#ifndef FFC_EMU
        // the scratch layout IS the shared-memory image: one bulk copy writes the CTA's planes back
        umma::fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            umma::bulk_s2g(p.spec + (size_t)plane0 * G::REGION, planes, (uint32_t)(np * G::REGION * 4));
            umma::bulk_commit_group();
            umma::bulk_wait_group_read0();      // the shared-memory source has been read: the CTA may retire while the writes drain
        }
#else
        FFC_PHASE {
            const float4* s4 = reinterpret_cast<const float4*>(planes);
            float4* d4 = reinterpret_cast<float4*>(p.spec + (size_t)plane0 * G::REGION);
            const int total = np * (G::REGION / 4);
            for (int i = tid; i < total; i += ctx.nt) d4[i] = s4[i];
        } FFC_SYNC;
#endif
    }
};

struct Fu3InvFftParams {
    const float* spec;       // (nplanes, N, RS), nplanes = B * cout
    const float* bn_a;       // [2*cout] folded BatchNorm scale (inverse transform scale included) or null: no BN + ReLU
    const float* bn_b;       // [2*cout]
    const float* residual;   // (nplanes, N, N) or null
    float* out;              // (nplanes, N, N)
    int nplanes, cout;
    float scale;             // multiplies the result (1 when the BN constants already carry it)
};

template <int N, bool ADJ>
struct Fu3Irfft2 {
    typedef Fu3InvFftParams Params;
    typedef Fu2G<N> G;
    static constexpr int P = Fu3G<N>::P;
    static constexpr int kThreads = Fu3G<N>::kThreads;
    static constexpr int kMinBlocks = (N == 128) ? 3 : (N == 64 ? 5 : 4);
    static size_t smem_bytes() { return ((size_t)P * G::REGION + 2 * N + 4 * P) * 4 + 16; }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        const int plane0 = ctx.bx * P;
        const int np = (p.nplanes - plane0) < P ? (p.nplanes - plane0) : P;
        float* planes = smem;
        float2* tw = reinterpret_cast<float2*>(smem + (size_t)P * G::REGION);
        float2* bnp_a = tw + N;                 // per plane of this CTA: (a_re, a_im), (b_re, b_im)
        float2* bnp_b = bnp_a + P;
#ifndef FFC_EMU
        // Device: the CTA's planes are contiguous in the scratch and the scratch layout IS the shared-memory image: ONE bulk
        // copy brings them in (completion on an mbarrier); the residual rows are prefetched into L2 meanwhile.  BatchNorm +
        // ReLU is applied by the first inverse column pass as it loads the spectrum from shared memory.
        uint64_t* bar = reinterpret_cast<uint64_t*>(bnp_b + P);
        {
            const int tid = (int)threadIdx.x;
            if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_proxy_async_smem(); }   // (the cluster-scope release fence flushes L1: CCTL.IVALL)
            __syncthreads();
            if (tid == 0) {
                umma::mbar_arrive_expect_tx(bar, (uint32_t)(np * G::REGION * 4));
                umma::bulk_g2s(planes, p.spec + (size_t)plane0 * G::REGION, (uint32_t)(np * G::REGION * 4), bar);
                if (p.residual) umma::bulk_prefetch_l2(p.residual + (size_t)plane0 * N * N, (uint32_t)(np * N * N * 4));
            }
            for (int k = tid; k < N; k += ctx.nt) tw[k] = c_tw128[k * (FFC_TW_N / N)];
            if (p.bn_a && tid < np) {
                const int o = (plane0 + tid) % p.cout;
                bnp_a[tid] = make_float2(FFC_LDG(p.bn_a + 2 * o), FFC_LDG(p.bn_a + 2 * o + 1));
                bnp_b[tid] = make_float2(FFC_LDG(p.bn_b + 2 * o), FFC_LDG(p.bn_b + 2 * o + 1));
            }
            umma::mbar_wait(bar, 0);
            __syncthreads();
        }
        Fu2Bn bn; bn.a = bnp_a; bn.b = bnp_b;
        if (p.bn_a) { FU3_COLS_INV(N, true, np, planes, tw, bn); }
        else { FU3_COLS_INV(N, false, np, planes, tw, bn); }
#else
        FFC_PHASE {
            for (int k = tid; k < N; k += ctx.nt) tw[k] = c_tw128[k * (FFC_TW_N / N)];
            const float4* s4 = reinterpret_cast<const float4*>(p.spec + (size_t)plane0 * G::REGION);
            float4* d4 = reinterpret_cast<float4*>(planes);
            constexpr int PER = G::REGION / 4;
            const int total = np * PER;
            for (int i = tid; i < total; i += ctx.nt) {
                float4 q = s4[i];
                if (p.bn_a) {              // relu(y * a + b) on (re, im) pairs of output channel o: fourier_unity.py:49
                    const int o = (plane0 + i / PER) % p.cout;
                    const float ar = p.bn_a[2 * o], ai = p.bn_a[2 * o + 1], br = p.bn_b[2 * o], bi = p.bn_b[2 * o + 1];
                    q.x = fmaf(q.x, ar, br); q.y = fmaf(q.y, ai, bi); q.z = fmaf(q.z, ar, br); q.w = fmaf(q.w, ai, bi);
                    q.x = q.x > 0.f ? q.x : 0.f; q.y = q.y > 0.f ? q.y : 0.f; q.z = q.z > 0.f ? q.z : 0.f; q.w = q.w > 0.f ? q.w : 0.f;
                }
                d4[i] = q;
            }
        } FFC_SYNC;
        Fu2Bn nobn; nobn.a = nullptr; nobn.b = nullptr;
        FU3_COLS_INV(N, false, np, planes, tw, nobn);
#endif
        FFC_PHASE { fu2_rows_inv<N, ADJ>(tid, ctx.nt, np * N, planes, p.scale); } FFC_SYNC;
#ifndef FFC_EMU
        if (!p.residual) {
            // no residual: every thread sends its finished row with one bulk copy (shared -> global, N*4 bytes)
            umma::fence_proxy_async_smem();
            __syncthreads();
            const int tid = (int)threadIdx.x;
            for (int r = tid; r < np * N; r += ctx.nt)
                umma::bulk_s2g(p.out + ((size_t)plane0 * N + r) * N, planes + (size_t)r * G::RS, (uint32_t)(N * 4));
            umma::bulk_commit_group();
            umma::bulk_wait_group_read0();      // the shared-memory source has been read: the CTA may retire while the writes drain
            return;
        }
#endif
        FFC_PHASE {
            const size_t g0 = (size_t)plane0 * N * N;
            if (p.residual) fu2_store_rows_impl<N, true, 8>(tid, ctx.nt, np * N, planes, p.residual + g0, p.out + g0);
            else fu2_store_rows_impl<N, false, 8>(tid, ctx.nt, np * N, planes, nullptr, p.out + g0);
        } FFC_SYNC;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// channel mix, plain FP32 form: the host emulation build's mix and the device cross-check of the tensor-core kernel
// (ffc_fu3_mix.cu).  One thread per complex slot.
// ------------------------------------------------------------------------------------------------------------------
struct Fu3MixSimt {
    typedef Fu3MixParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            const long long total = (long long)p.G * p.NB;
            for (long long m = (long long)ctx.bx * ctx.nt + tid; m < total; m += (long long)ctx.gx * ctx.nt) {
                const int b = (int)(m / p.NB), r = (int)(m % p.NB);
                const bool real_bin = (r % p.SPS) != p.SPS - 1;
                const float2* sp = reinterpret_cast<const float2*>(p.s) + (size_t)b * p.Cin * p.NB + r;
                for (int o = 0; o < p.Cout; ++o) {
                    const float* w0 = p.w + (size_t)(2 * o) * 2 * p.Cin;
                    const float* w1 = w0 + 2 * p.Cin;
                    float yr = 0.f, yi = 0.f;
                    for (int c = 0; c < p.Cin; ++c) {
                        const float2 v = sp[(size_t)c * p.NB];
                        yr = fmaf(FFC_LDG(w0 + 2 * c), v.x, yr); yr = fmaf(FFC_LDG(w0 + 2 * c + 1), v.y, yr);
                        yi = fmaf(FFC_LDG(w1 + 2 * c), v.x, yi); yi = fmaf(FFC_LDG(w1 + 2 * c + 1), v.y, yi);
                    }
                    yr *= p.scale; yi *= p.scale;
                    if (p.sums && real_bin) {
                        ffc_atomic_add(p.sums + 2 * o, (double)yr);
                        ffc_atomic_add(p.sums + 2 * o + 1, (double)yi);
                        ffc_atomic_add(p.sums + 2 * p.Cout + 2 * o, (double)yr * (double)yr);
                        ffc_atomic_add(p.sums + 2 * p.Cout + 2 * o + 1, (double)yi * (double)yi);
                    }
                    if (p.bn_a) {
                        yr = fmaf(yr, FFC_LDG(p.bn_a + 2 * o), FFC_LDG(p.bn_b + 2 * o));
                        yi = fmaf(yi, FFC_LDG(p.bn_a + 2 * o + 1), FFC_LDG(p.bn_b + 2 * o + 1));
                        yr = yr > 0.f ? yr : 0.f; yi = yi > 0.f ? yi : 0.f;
                    }
                    if (p.y) reinterpret_cast<float2*>(p.y)[((size_t)b * p.Cout + o) * p.NB + r] = make_float2(yr, yi);
                }
            }
        } FFC_SYNC;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// BatchNorm constants: y -> relu(y * a + b) with the inverse transform's 1/N folded in (relu(z) * s == relu(z * s));
// training mode finishes the batch statistics, publishes them and updates the running ones (fourier_unity.py:49)
// ------------------------------------------------------------------------------------------------------------------
struct Fu3FinalizeParams {
    const double* sums; const float* gamma; const float* beta;
    float* running_mean; float* running_var; float* save_mean; float* save_invstd;
    float* bn_a; float* bn_b;
    int Cout, training;
    double count;
    float eps, momentum, scale;
};
struct Fu3Finalize {
    typedef Fu3FinalizeParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            for (int o = tid; o < 2 * p.Cout; o += ctx.nt) {
                float mean, invstd;
                if (p.training) {
                    const double m = p.sums[o] / p.count;
                    double var = p.sums[2 * p.Cout + o] / p.count - m * m;
                    if (var < 0.0) var = 0.0;
                    mean = (float)m;
                    invstd = 1.0f / sqrtf((float)var + p.eps);
                    if (p.running_mean) {
                        const double unb = p.count > 1.0 ? var * p.count / (p.count - 1.0) : var;
                        p.running_mean[o] = (1.f - p.momentum) * p.running_mean[o] + p.momentum * mean;
                        p.running_var[o] = (1.f - p.momentum) * p.running_var[o] + p.momentum * (float)unb;
                    }
                } else {
                    mean = p.running_mean[o];
                    invstd = 1.0f / sqrtf(p.running_var[o] + p.eps);
                }
                p.save_mean[o] = mean; p.save_invstd[o] = invstd;
                const float a = invstd * FFC_LDG(p.gamma + o) * p.scale;
                p.bn_a[o] = a;
                p.bn_b[o] = FFC_LDG(p.beta + o) * p.scale - mean * a;
            }
        } FFC_SYNC;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
#ifndef FFC_EMU
// tensor-core mix (ffc_fu3_mix.cu)
bool fu3_mix_tc_supported(int Cin, int Cout);
size_t fu3_mix_tc_packed_floats(int Cin, int Cout);
int fu3_mix_tc_pack(const float* w, float* wp, int Cin, int Cout, float scale, int transposed, ffc_stream_t st);
int fu3_mix_tc_run(const Fu3MixParams& p, ffc_stream_t st);
bool fu3_wgrad_tc_supported(int Cin, int Cout);
size_t fu3_wgrad_tc_part_floats(int Cin, int Cout);
int fu3_wgrad_tc_run(const Fu3BwdWgradParams& p, ffc_stream_t st);
#endif
static int g_fu3_simt_mix = 0;
extern "C" void ffc_debug_fu3_simt_mix(int on) { g_fu3_simt_mix = on; }
// Bytes of spectrum a chunk of images may occupy.  Measured (profiles/r02c_fu3_chunk_sweep.txt, FourierUnitSN(32,32) @128x128,
// batch 64): every kernel of a chunk must fill the 148 SMs for several waves (3 planes per SM and wave), so chunks below
// ~50 MB cost more in partial waves and launches than L2 residency of S returns; 160 MB keeps the largest BASELINE unit in one
// chunk and still bounds the scratch for larger batches.
static const size_t kFu3ChunkDefault = (size_t)160 << 20;
static size_t g_fu3_chunk_bytes = kFu3ChunkDefault;
extern "C" void ffc_debug_fu3_chunk_bytes(size_t bytes) { g_fu3_chunk_bytes = bytes ? bytes : kFu3ChunkDefault; }

static bool fu3_mix_supported(int Cin, int Cout) {
#ifdef FFC_EMU
    (void)Cin; (void)Cout;
    return true;
#else
    return g_fu3_simt_mix || fu3_mix_tc_supported(Cin, Cout);
#endif
}

static size_t fu3_align(size_t v) { return (v + 255) & ~(size_t)255; }

struct Fu3Plan {
    int NB, region, chunk;          // slots per plane, floats per plane, images per chunk
    size_t off_sums, off_consts, off_wp, off_s, off_y, total;
};

static Fu3Plan fu3_plan(int B, int Cin, int Cout, int N, int training) {
    Fu3Plan pl;
    const int SPS = N / 2 + 2;
    pl.NB = N * SPS;
    pl.region = 2 * pl.NB;
    // images per chunk: the spectrum of a chunk (and, in eval mode, its mixed spectrum) stays in L2 between two kernels
    const size_t per_image = (size_t)(Cin + (training ? 0 : Cout)) * pl.region * 4;
    size_t g = g_fu3_chunk_bytes / (per_image ? per_image : 1);
    if (g < 1) g = 1;
    if (g > (size_t)B) g = (size_t)B;
    pl.chunk = (int)g;
    size_t off = 0;
    pl.off_sums = off; off = fu3_align(off + (size_t)4 * Cout * sizeof(double));
    pl.off_consts = off; off = fu3_align(off + (size_t)4 * Cout * sizeof(float));
    pl.off_wp = off;
#ifndef FFC_EMU
    off = fu3_align(off + fu3_mix_tc_packed_floats(Cin, Cout) * sizeof(float));
#endif
    pl.off_s = off; off = fu3_align(off + (size_t)pl.chunk * Cin * pl.region * 4);
    pl.off_y = off; off = fu3_align(off + (size_t)(training ? B : pl.chunk) * Cout * pl.region * 4);
    pl.total = off + 256;
    return pl;
}

// 1 when ffc_fu3_fwd supports the shape
extern "C" int ffc_fu3_supported(int B, int Cin, int Cout, int H, int W) {
    if (H != W || B < 1 || Cin < 1 || Cout < 1) return 0;
    if (!(H == 16 || H == 32 || H == 64 || H == 128)) return 0;
    return fu3_mix_supported(Cin, Cout) ? 1 : 0;
}

extern "C" size_t ffc_fu3_workspace_bytes(int B, int Cin, int Cout, int H, int W, int training) {
    if (!ffc_fu3_supported(B, Cin, Cout, H, W)) return 0;
    return fu3_plan(B, Cin, Cout, H, training).total;
}

template <int N>
static int fu3_run(const float* x, const float* w, const float* gamma, const float* beta, float* running_mean, float* running_var,
                   float* save_mean, float* save_invstd, const float* residual, float* out, int B, int Cin, int Cout,
                   int training, float eps, float momentum, float* s_keep, float* y_keep, unsigned char* ws, ffc_stream_t st) {
    typedef Fu3G<N> G3;
    const Fu3Plan pl = fu3_plan(B, Cin, Cout, N, training);
    double* sums = reinterpret_cast<double*>(ws + pl.off_sums);
    float* bn_a = reinterpret_cast<float*>(ws + pl.off_consts);
    float* bn_b = bn_a + 2 * Cout;
    float* wp = reinterpret_cast<float*>(ws + pl.off_wp);
    float* S = reinterpret_cast<float*>(ws + pl.off_s);
    float* Y = y_keep ? y_keep : reinterpret_cast<float*>(ws + pl.off_y);
    const float scale = 1.0f / (float)N;
    // two-stage form: the mix writes the raw mixed spectrum of the whole batch and the inverse transform applies BatchNorm +
    // ReLU as it loads it.  Training mode needs it (batch statistics); a forward that keeps Y for the backward uses it too.
    const bool staged_y = training || y_keep;
    bool tc = false;
#ifndef FFC_EMU
    tc = !g_fu3_simt_mix && fu3_mix_tc_supported(Cin, Cout);
    if (tc) FFC_CHECK(fu3_mix_tc_pack(w, wp, Cin, Cout, scale, 0, st));
#endif
    Fu3FinalizeParams fp;
    fp.sums = sums; fp.gamma = gamma; fp.beta = beta; fp.running_mean = running_mean; fp.running_var = running_var;
    fp.save_mean = save_mean; fp.save_invstd = save_invstd; fp.bn_a = bn_a; fp.bn_b = bn_b; fp.Cout = Cout; fp.training = training;
    fp.count = (double)B * N * (N / 2 + 1); fp.eps = eps; fp.momentum = momentum; fp.scale = scale;
    if (training) FFC_CHECK(ffc_memset_async(sums, 0, (size_t)4 * Cout * sizeof(double), st));
    else FFC_CHECK((ffc_launch<Fu3Finalize>(1, 1, 1, 256, 0, st, fp)));
    for (int b0 = 0; b0 < B; b0 += pl.chunk) {
        const int g = (B - b0) < pl.chunk ? (B - b0) : pl.chunk;
        float* Sc = s_keep ? s_keep + (size_t)b0 * Cin * pl.region : S;
        Fu3FwdFftParams ap = {};
        ap.x = x + (size_t)b0 * Cin * N * N; ap.spec = Sc; ap.nplanes = g * Cin;
        FFC_CHECK((ffc_launch<Fu3Rfft2<N, false>>(ffc_cdiv(ap.nplanes, G3::P), 1, 1, G3::kThreads, Fu3Rfft2<N, false>::smem_bytes(), st, ap)));
        Fu3MixParams mp;
        mp.s = Sc; mp.y = staged_y ? Y + (size_t)b0 * Cout * pl.region : Y; mp.w = w; mp.wp = tc ? wp : nullptr;
        mp.bn_a = staged_y ? nullptr : bn_a; mp.bn_b = staged_y ? nullptr : bn_b; mp.sums = training ? sums : nullptr;
        mp.G = g; mp.Cin = Cin; mp.Cout = Cout; mp.NB = pl.NB; mp.SPS = N / 2 + 2; mp.scale = scale;
#ifndef FFC_EMU
        if (tc) { FFC_CHECK(fu3_mix_tc_run(mp, st)); }
        else
#endif
        {
            long long blocks = ((long long)g * pl.NB + 255) / 256;
            if (blocks > (long long)ffc_sm_count() * 8) blocks = (long long)ffc_sm_count() * 8;
            FFC_CHECK((ffc_launch<Fu3MixSimt>((int)blocks, 1, 1, 256, 0, st, mp)));
        }
        if (!staged_y) {
            Fu3InvFftParams ip; ip.spec = Y; ip.bn_a = nullptr; ip.bn_b = nullptr;
            ip.residual = residual ? residual + (size_t)b0 * Cout * N * N : nullptr; ip.out = out + (size_t)b0 * Cout * N * N;
            ip.nplanes = g * Cout; ip.cout = Cout; ip.scale = 1.0f;
            FFC_CHECK((ffc_launch<Fu3Irfft2<N, false>>(ffc_cdiv(ip.nplanes, G3::P), 1, 1, G3::kThreads, Fu3Irfft2<N, false>::smem_bytes(), st, ip)));
        }
    }
    if (staged_y) {
        if (training) FFC_CHECK((ffc_launch<Fu3Finalize>(1, 1, 1, 256, 0, st, fp)));
        Fu3InvFftParams ip; ip.spec = Y; ip.bn_a = bn_a; ip.bn_b = bn_b; ip.residual = residual; ip.out = out;
        ip.nplanes = B * Cout; ip.cout = Cout; ip.scale = 1.0f;
        FFC_CHECK((ffc_launch<Fu3Irfft2<N, false>>(ffc_cdiv(ip.nplanes, G3::P), 1, 1, G3::kThreads, Fu3Irfft2<N, false>::smem_bytes(), st, ip)));
    }
    return FFC_OK;
}

static int fu3_fwd_impl(const float* x, const float* w, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                        const float* residual, float* out, float* s_keep, float* y_keep,
                        int B, int Cin, int Cout, int H, int W, int training, float eps, float momentum,
                        void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(x && w && gamma && beta && save_mean && save_invstd && out, "ffc_fu3_fwd: null pointer");
    FFC_REQUIRE(training || (running_mean && running_var), "ffc_fu3_fwd: eval mode needs running statistics");
    FFC_REQUIRE(B >= 0, "ffc_fu3_fwd: negative batch");
    if (B == 0) return FFC_OK;
    FFC_REQUIRE(ffc_fu3_supported(B, Cin, Cout, H, W), "ffc_fu3_fwd: unsupported shape B=%d Cin=%d Cout=%d %dx%d", B, Cin, Cout, H, W);
    FFC_REQUIRE((((uintptr_t)x | (uintptr_t)out | (uintptr_t)residual | (uintptr_t)s_keep | (uintptr_t)y_keep) & 15) == 0,
                "ffc_fu3_fwd: x / out / residual / kept spectra must be 16-byte aligned");
    const uintptr_t wsa = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    if (!workspace || wsa + fu3_plan(B, Cin, Cout, H, training).total - 256 > (uintptr_t)workspace + workspace_bytes) {
        ffc_set_error("ffc_fu3_fwd: workspace too small (%zu bytes needed)", fu3_plan(B, Cin, Cout, H, training).total);
        return FFC_ERR_WORKSPACE;
    }
    unsigned char* ws = reinterpret_cast<unsigned char*>(wsa);
    ffc_stream_t st = (ffc_stream_t)stream;
    switch (H) {
        case 16: return fu3_run<16>(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out, B, Cin, Cout, training, eps, momentum, s_keep, y_keep, ws, st);
        case 32: return fu3_run<32>(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out, B, Cin, Cout, training, eps, momentum, s_keep, y_keep, ws, st);
        case 64: return fu3_run<64>(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out, B, Cin, Cout, training, eps, momentum, s_keep, y_keep, ws, st);
        default: return fu3_run<128>(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out, B, Cin, Cout, training, eps, momentum, s_keep, y_keep, ws, st);
    }
}

// L2-staged FourierUnitSN forward; same contract as ffc_fu_fwd (ffc_fu2.cu) with workspace >= ffc_fu3_workspace_bytes(...).
extern "C" int ffc_fu3_fwd(const float* x, const float* w, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                           const float* residual, float* out,
                           int B, int Cin, int Cout, int H, int W, int training, float eps, float momentum,
                           void* workspace, size_t workspace_bytes, void* stream) {
    return fu3_fwd_impl(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out, nullptr, nullptr,
                        B, Cin, Cout, H, W, training, eps, momentum, workspace, workspace_bytes, stream);
}

// The same forward, keeping what ffc_fu3_bwd needs: s_keep (B, Cin, H, W + 4) receives the unnormalised spectrum of x and
// y_keep (B, Cout, H, W + 4) the mixed spectrum that enters BatchNorm, both in the library's plane layout (opaque to callers).
extern "C" int ffc_fu3_fwd_keep(const float* x, const float* w, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                                const float* residual, float* out, float* s_keep, float* y_keep,
                                int B, int Cin, int Cout, int H, int W, int training, float eps, float momentum,
                                void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(s_keep && y_keep, "ffc_fu3_fwd_keep: null pointer");
    return fu3_fwd_impl(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out, s_keep, y_keep,
                        B, Cin, Cout, H, W, training, eps, momentum, workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// backward of the L2-staged form: nothing is recomputed, the two spectra kept by ffc_fu3_fwd_keep are read once each.
//
//   dout --Fu3Rfft2<ADJ, MASK>--> g = adjoint(irfft2)(dout) * [relu mask]  (+ sum g, sum g * xhat per channel)
//        --Fu3BwdFinalize--> dgamma, dbeta, per-channel constants of the BatchNorm backward
//        --Fu3BwdWgrad--> dY = coef * (g - c1 - xhat * c2)  and  dW = dY^T S / N   (FP32, register-tiled)
//        --Fu3Mix with W^T (tensor cores)--> dS = dY W
//        --Fu3Irfft2<ADJ>--> dx = adjoint(rfft2)(dS)
// ------------------------------------------------------------------------------------------------------------------
struct Fu3BwdPrepParams {
    const float* w; const float* gamma; const float* beta; const float* mean; const float* invstd;
    float* bn_a; float* bn_b; float* wt;      // wt [2*Cin][2*Cout] = w^T (plain FP32 mix) or null
    int Cin, Cout;
    float scale;
};
struct Fu3BwdPrep {
    typedef Fu3BwdPrepParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            for (int o = tid; o < 2 * p.Cout; o += ctx.nt) {       // the forward's folded constants, same arithmetic (Fu3Finalize)
                const float a = FFC_LDG(p.invstd + o) * FFC_LDG(p.gamma + o) * p.scale;
                p.bn_a[o] = a;
                p.bn_b[o] = FFC_LDG(p.beta + o) * p.scale - FFC_LDG(p.mean + o) * a;
            }
            if (p.wt)
                for (int e = tid; e < 4 * p.Cin * p.Cout; e += ctx.nt) {
                    const int k = e / (2 * p.Cout), n = e % (2 * p.Cout);
                    p.wt[e] = FFC_LDG(p.w + (size_t)n * 2 * p.Cin + k);
                }
        } FFC_SYNC;
    }
};

struct Fu3BwdFinalizeParams {
    const double* sums; const float* gamma; const float* invstd;
    float* dgamma; float* dbeta; float* coef; float* c1; float* c2;
    int Cout, training;
    double count;
    float scale;
};
struct Fu3BwdFinalize {
    typedef Fu3BwdFinalizeParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        FFC_PHASE {
            for (int o = tid; o < 2 * p.Cout; o += ctx.nt) {
                const double sg = p.sums[o], sgx = p.sums[2 * p.Cout + o];
                p.dbeta[o] = (float)(sg * (double)p.scale);
                p.dgamma[o] = (float)(sgx * (double)p.scale);
                p.coef[o] = FFC_LDG(p.gamma + o) * FFC_LDG(p.invstd + o) * p.scale;
                p.c1[o] = p.training ? (float)(sg / p.count) : 0.f;
                p.c2[o] = p.training ? (float)(sgx / p.count) : 0.f;
            }
        } FFC_SYNC;
    }
};

// dY and the weight gradient.  CTA tile: 64 rows n = 2*o + (re | im) of dY by 64 columns k = 2*c + (re | im) of S, over
// chunks of 64 bins staged bin-major in shared memory; thread = 4 x 4 outputs (two 16-byte loads per 16 FMAs); persistent
// over the chunks, one float atomic per output and CTA at the end.  The k-block 0 CTAs also write dY.
struct Fu3WgAcc { float v[16]; };
struct Fu3BwdWgrad {
    typedef Fu3BwdWgradParams Params;
    static constexpr int kThreads = 256, KB = 64, LD = 68;       // bins per chunk, floats per staged bin row (64 + 4: 16-byte aligned)
    static constexpr int kMinBlocks = 2;
    static size_t smem_bytes() { return (size_t)2 * KB * LD * 4; }
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        float* As = smem;                 // [KB][LD]: dY, row n0 + i at column i
        float* Bs = smem + KB * LD;       // [KB][LD]: S
        const int o0 = ctx.by * 32, c0 = ctx.bz * 32;          // first complex channel of the tile
        const long long Mtot = (long long)p.B * p.NB;
        const int nchunks = (int)((Mtot + KB - 1) / KB);
        FFC_TLS(Fu3WgAcc, acc);
        FFC_PHASE {
            FFC_TLS_REF(Fu3WgAcc, acc);
            (void)tid;
            FFC_UNROLL
            for (int i = 0; i < 16; ++i) acc.v[i] = 0.f;
        } FFC_SYNC;
        for (int chunk = ctx.bx; chunk < nchunks; chunk += ctx.gx) {
            FFC_PHASE {
                // staging: lanes walk the bins of one channel (coalesced 8-byte loads), 32 channels x 64 bins per operand
                for (int idx = tid; idx < 32 * KB; idx += ctx.nt) {
                    const int ch = idx / KB, j = idx % KB;
                    const long long m = (long long)chunk * KB + j;
                    const bool ok = m < Mtot;
                    const int b = ok ? (int)(m / p.NB) : 0, r = ok ? (int)(m % p.NB) : 0;
                    float2 d = make_float2(0.f, 0.f), sv = make_float2(0.f, 0.f);
                    const int o = o0 + ch, c = c0 + ch;
                    if (ok && o < p.Cout) {
                        const size_t off = ((size_t)b * p.Cout + o) * p.NB + r;
                        const float2 gv = FFC_LDG(reinterpret_cast<const float2*>(p.g) + off);
                        const float2 yv = FFC_LDG(reinterpret_cast<const float2*>(p.y) + off);
                        const int n = 2 * o;
                        d.x = FFC_LDG(p.coef + n) * (gv.x - FFC_LDG(p.c1 + n) - (yv.x - FFC_LDG(p.mean + n)) * FFC_LDG(p.invstd + n) * FFC_LDG(p.c2 + n));
                        d.y = FFC_LDG(p.coef + n + 1) * (gv.y - FFC_LDG(p.c1 + n + 1) - (yv.y - FFC_LDG(p.mean + n + 1)) * FFC_LDG(p.invstd + n + 1) * FFC_LDG(p.c2 + n + 1));
                        if (ctx.bz == 0) reinterpret_cast<float2*>(p.dy)[off] = d;
                    }
                    if (ok && c < p.Cin) sv = FFC_LDG(reinterpret_cast<const float2*>(p.s) + ((size_t)b * p.Cin + c) * p.NB + r);
                    *reinterpret_cast<float2*>(As + j * LD + 2 * ch) = d;
                    *reinterpret_cast<float2*>(Bs + j * LD + 2 * ch) = sv;
                }
            } FFC_SYNC;
            FFC_PHASE {
                FFC_TLS_REF(Fu3WgAcc, acc);
                const int tn = tid % 16, tk = tid / 16;
                FFC_UNROLL
                for (int j = 0; j < KB; ++j) {
                    const float4 a = *reinterpret_cast<const float4*>(As + j * LD + 4 * tn);
                    const float4 b = *reinterpret_cast<const float4*>(Bs + j * LD + 4 * tk);
                    acc.v[0] = fmaf(a.x, b.x, acc.v[0]); acc.v[1] = fmaf(a.x, b.y, acc.v[1]); acc.v[2] = fmaf(a.x, b.z, acc.v[2]); acc.v[3] = fmaf(a.x, b.w, acc.v[3]);
                    acc.v[4] = fmaf(a.y, b.x, acc.v[4]); acc.v[5] = fmaf(a.y, b.y, acc.v[5]); acc.v[6] = fmaf(a.y, b.z, acc.v[6]); acc.v[7] = fmaf(a.y, b.w, acc.v[7]);
                    acc.v[8] = fmaf(a.z, b.x, acc.v[8]); acc.v[9] = fmaf(a.z, b.y, acc.v[9]); acc.v[10] = fmaf(a.z, b.z, acc.v[10]); acc.v[11] = fmaf(a.z, b.w, acc.v[11]);
                    acc.v[12] = fmaf(a.w, b.x, acc.v[12]); acc.v[13] = fmaf(a.w, b.y, acc.v[13]); acc.v[14] = fmaf(a.w, b.z, acc.v[14]); acc.v[15] = fmaf(a.w, b.w, acc.v[15]);
                }
            } FFC_SYNC;
        }
        FFC_PHASE {
            FFC_TLS_REF(Fu3WgAcc, acc);
            const int tn = tid % 16, tk = tid / 16;
            FFC_UNROLL
            for (int i = 0; i < 4; ++i) {
                const int n = 2 * o0 + 4 * tn + i;
                FFC_UNROLL
                for (int l = 0; l < 4; ++l) {
                    const int k = 2 * c0 + 4 * tk + l;
                    if (n < 2 * p.Cout && k < 2 * p.Cin) ffc_atomic_add(p.dw + (size_t)n * 2 * p.Cin + k, acc.v[4 * i + l] * p.scale);
                }
            }
        } FFC_SYNC;
    }
};

struct Fu3BwdPlan {
    int NB, region;
    size_t off_sums, off_consts, off_wp, off_part, off_g, off_dy, off_ds, total;
};
static Fu3BwdPlan fu3_bwd_plan(int B, int Cin, int Cout, int N) {
    Fu3BwdPlan pl;
    pl.NB = N * (N / 2 + 2);
    pl.region = 2 * pl.NB;
    size_t off = 0;
    pl.off_sums = off; off = fu3_align(off + (size_t)4 * Cout * sizeof(double));
    pl.off_consts = off; off = fu3_align(off + (size_t)10 * Cout * sizeof(float));      // bn_a | bn_b | coef | c1 | c2, 2*Cout each
    pl.off_wp = off;
    size_t wbytes = (size_t)4 * Cin * Cout * sizeof(float);                             // w^T for the plain FP32 mix
#ifndef FFC_EMU
    const size_t packed = fu3_mix_tc_packed_floats(Cout, Cin) * sizeof(float);
    if (packed > wbytes) wbytes = packed;
#endif
    off = fu3_align(off + wbytes);
    pl.off_part = off;
#ifndef FFC_EMU
    if (fu3_wgrad_tc_supported(Cin, Cout)) off = fu3_align(off + fu3_wgrad_tc_part_floats(Cin, Cout) * sizeof(float));
#endif
    pl.off_g = off; off = fu3_align(off + (size_t)B * Cout * pl.region * 4);
    pl.off_dy = off; off = fu3_align(off + (size_t)B * Cout * pl.region * 4);
    pl.off_ds = off; off = fu3_align(off + (size_t)B * Cin * pl.region * 4);
    pl.total = off + 256;
    return pl;
}

extern "C" size_t ffc_fu3_bwd_workspace_bytes(int B, int Cin, int Cout, int H, int W) {
    if (!ffc_fu3_supported(B, Cin, Cout, H, W)) return 0;
    return fu3_bwd_plan(B, Cin, Cout, H).total;
}

template <int N>
static int fu3_bwd_run(const float* dout, const float* s_keep, const float* y_keep, const float* w, const float* gamma, const float* beta,
                       const float* save_mean, const float* save_invstd, float* dx, float* dw, float* dgamma, float* dbeta,
                       int B, int Cin, int Cout, int training, unsigned char* ws, ffc_stream_t st) {
    typedef Fu3G<N> G3;
    const Fu3BwdPlan pl = fu3_bwd_plan(B, Cin, Cout, N);
    double* sums = reinterpret_cast<double*>(ws + pl.off_sums);
    float* bn_a = reinterpret_cast<float*>(ws + pl.off_consts);
    float* bn_b = bn_a + 2 * Cout; float* coef = bn_b + 2 * Cout; float* c1 = coef + 2 * Cout; float* c2 = c1 + 2 * Cout;
    float* wp = reinterpret_cast<float*>(ws + pl.off_wp);
    float* Gs = reinterpret_cast<float*>(ws + pl.off_g);
    float* dY = reinterpret_cast<float*>(ws + pl.off_dy);
    float* dS = reinterpret_cast<float*>(ws + pl.off_ds);
    const float scale = 1.0f / (float)N;
    bool tc = false;
#ifndef FFC_EMU
    tc = !g_fu3_simt_mix && fu3_mix_tc_supported(Cout, Cin);
#endif
    FFC_CHECK(ffc_memset_async(sums, 0, (size_t)4 * Cout * sizeof(double), st));
    FFC_CHECK(ffc_memset_async(dw, 0, (size_t)4 * Cin * Cout * sizeof(float), st));
    Fu3BwdPrepParams pp;
    pp.w = w; pp.gamma = gamma; pp.beta = beta; pp.mean = save_mean; pp.invstd = save_invstd; pp.bn_a = bn_a; pp.bn_b = bn_b;
    pp.wt = (dx && !tc) ? wp : nullptr; pp.Cin = Cin; pp.Cout = Cout; pp.scale = scale;
    FFC_CHECK((ffc_launch<Fu3BwdPrep>(1, 1, 1, 256, 0, st, pp)));
#ifndef FFC_EMU
    if (dx && tc) FFC_CHECK(fu3_mix_tc_pack(w, wp, Cout, Cin, 1.0f, 1, st));          // W^T: contraction over the 2*Cout channels of dY
#endif
    Fu3FwdFftParams ap = {};
    ap.x = dout; ap.spec = Gs; ap.nplanes = B * Cout; ap.y = y_keep; ap.bn_a = bn_a; ap.bn_b = bn_b; ap.mean = save_mean; ap.invstd = save_invstd;
    ap.sums = sums; ap.cout = Cout;
    FFC_CHECK((ffc_launch<Fu3Rfft2<N, true, true>>(ffc_cdiv(ap.nplanes, G3::P), 1, 1, G3::kThreads, Fu3Rfft2<N, true, true>::smem_bytes(), st, ap)));
    Fu3BwdFinalizeParams fp;
    fp.sums = sums; fp.gamma = gamma; fp.invstd = save_invstd; fp.dgamma = dgamma; fp.dbeta = dbeta; fp.coef = coef; fp.c1 = c1; fp.c2 = c2;
    fp.Cout = Cout; fp.training = training; fp.count = (double)B * N * (N / 2 + 1); fp.scale = scale;
    FFC_CHECK((ffc_launch<Fu3BwdFinalize>(1, 1, 1, 256, 0, st, fp)));
    Fu3BwdWgradParams wg;
    wg.g = Gs; wg.y = y_keep; wg.s = s_keep; wg.dy = dY; wg.coef = coef; wg.c1 = c1; wg.c2 = c2; wg.mean = save_mean; wg.invstd = save_invstd;
    wg.dw = dw; wg.part = reinterpret_cast<float*>(ws + pl.off_part); wg.B = B; wg.Cin = Cin; wg.Cout = Cout; wg.NB = pl.NB; wg.scale = scale;
#ifndef FFC_EMU
    if (!g_fu3_simt_mix && fu3_wgrad_tc_supported(Cin, Cout)) { FFC_CHECK(fu3_wgrad_tc_run(wg, st)); }
    else
#endif
    {
        const int gy = ffc_cdiv(Cout, 32), gz = ffc_cdiv(Cin, 32);
        const long long nchunks = ((long long)B * pl.NB + Fu3BwdWgrad::KB - 1) / Fu3BwdWgrad::KB;
        long long gx = (long long)2 * ffc_sm_count() / (gy * gz);
        if (gx < 1) gx = 1;
        if (gx > nchunks) gx = nchunks;
        FFC_CHECK((ffc_launch<Fu3BwdWgrad>((int)gx, gy, gz, 256, Fu3BwdWgrad::smem_bytes(), st, wg)));
    }
    if (!dx) return FFC_OK;
    Fu3MixParams mp;
    mp.s = dY; mp.y = dS; mp.w = wp; mp.wp = tc ? wp : nullptr; mp.bn_a = nullptr; mp.bn_b = nullptr; mp.sums = nullptr;
    mp.G = B; mp.Cin = Cout; mp.Cout = Cin; mp.NB = pl.NB; mp.SPS = N / 2 + 2; mp.scale = 1.0f;
#ifndef FFC_EMU
    if (tc) { FFC_CHECK(fu3_mix_tc_run(mp, st)); }
    else
#endif
    {
        long long blocks = ((long long)B * pl.NB + 255) / 256;
        if (blocks > (long long)ffc_sm_count() * 8) blocks = (long long)ffc_sm_count() * 8;
        FFC_CHECK((ffc_launch<Fu3MixSimt>((int)blocks, 1, 1, 256, 0, st, mp)));
    }
    Fu3InvFftParams ip; ip.spec = dS; ip.bn_a = nullptr; ip.bn_b = nullptr; ip.residual = nullptr; ip.out = dx;
    ip.nplanes = B * Cin; ip.cout = Cin; ip.scale = scale;
    FFC_CHECK((ffc_launch<Fu3Irfft2<N, true>>(ffc_cdiv(ip.nplanes, G3::P), 1, 1, G3::kThreads, Fu3Irfft2<N, true>::smem_bytes(), st, ip)));
    return FFC_OK;
}

// Backward of ffc_fu3_fwd_keep.  dout (B, Cout, H, W); s_keep / y_keep as written by the forward; save_mean / save_invstd from
// the forward; dx (B, Cin, H, W) or null; dw [2*Cout][2*Cin]; dgamma / dbeta [2*Cout].  The gradient with respect to the
// residual is dout itself.  workspace >= ffc_fu3_bwd_workspace_bytes(...).
extern "C" int ffc_fu3_bwd(const float* dout, const float* s_keep, const float* y_keep, const float* w, const float* gamma, const float* beta,
                           const float* save_mean, const float* save_invstd, float* dx, float* dw, float* dgamma, float* dbeta,
                           int B, int Cin, int Cout, int H, int W, int training, void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(dout && s_keep && y_keep && w && gamma && beta && save_mean && save_invstd && dw && dgamma && dbeta, "ffc_fu3_bwd: null pointer");
    FFC_REQUIRE(B >= 0, "ffc_fu3_bwd: negative batch");
    ffc_stream_t st = (ffc_stream_t)stream;
    if (B == 0) {
        FFC_CHECK(ffc_memset_async(dw, 0, (size_t)4 * Cin * Cout * sizeof(float), st));
        FFC_CHECK(ffc_memset_async(dgamma, 0, (size_t)2 * Cout * sizeof(float), st));
        return ffc_memset_async(dbeta, 0, (size_t)2 * Cout * sizeof(float), st);
    }
    FFC_REQUIRE(ffc_fu3_supported(B, Cin, Cout, H, W), "ffc_fu3_bwd: unsupported shape B=%d Cin=%d Cout=%d %dx%d", B, Cin, Cout, H, W);
    FFC_REQUIRE((((uintptr_t)dout | (uintptr_t)dx | (uintptr_t)s_keep | (uintptr_t)y_keep) & 15) == 0, "ffc_fu3_bwd: dout / dx / kept spectra must be 16-byte aligned");
    const Fu3BwdPlan pl = fu3_bwd_plan(B, Cin, Cout, H);
    const uintptr_t wsa = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    if (!workspace || wsa + pl.total - 256 > (uintptr_t)workspace + workspace_bytes) {
        ffc_set_error("ffc_fu3_bwd: workspace too small (%zu bytes needed)", pl.total);
        return FFC_ERR_WORKSPACE;
    }
    unsigned char* ws = reinterpret_cast<unsigned char*>(wsa);
    switch (H) {
        case 16: return fu3_bwd_run<16>(dout, s_keep, y_keep, w, gamma, beta, save_mean, save_invstd, dx, dw, dgamma, dbeta, B, Cin, Cout, training, ws, st);
        case 32: return fu3_bwd_run<32>(dout, s_keep, y_keep, w, gamma, beta, save_mean, save_invstd, dx, dw, dgamma, dbeta, B, Cin, Cout, training, ws, st);
        case 64: return fu3_bwd_run<64>(dout, s_keep, y_keep, w, gamma, beta, save_mean, save_invstd, dx, dw, dgamma, dbeta, B, Cin, Cout, training, ws, st);
        default: return fu3_bwd_run<128>(dout, s_keep, y_keep, w, gamma, beta, save_mean, save_invstd, dx, dw, dgamma, dbeta, B, Cin, Cout, training, ws, st);
    }
}
