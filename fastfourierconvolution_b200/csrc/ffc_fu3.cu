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
};

template <int N, bool ADJ>
struct Fu3Rfft2 {
    typedef Fu3FwdFftParams Params;
    typedef Fu2G<N> G;
    static constexpr int P = Fu3G<N>::P;
    static constexpr int kThreads = Fu3G<N>::kThreads;
    static constexpr int kMinBlocks = (N == 128) ? 3 : (N == 64 ? 4 : 4);
    static size_t smem_bytes() { return ((size_t)P * G::REGION + 2 * N) * 4 + 16; }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        const int plane0 = ctx.bx * P;
        const int np = (p.nplanes - plane0) < P ? (p.nplanes - plane0) : P;
        float* planes = smem;
        float2* tw = reinterpret_cast<float2*>(smem + (size_t)P * G::REGION);
#ifndef FFC_EMU
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
#ifndef FFC_EMU
        // the scratch layout IS the shared-memory image: one bulk copy writes the CTA's planes back
        umma::fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            umma::bulk_s2g(p.spec + (size_t)plane0 * G::REGION, planes, (uint32_t)(np * G::REGION * 4));
            umma::bulk_commit_group();
            umma::bulk_wait_group0();
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
    static constexpr int kMinBlocks = (N == 128) ? 3 : 4;
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
            umma::bulk_wait_group0();
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
int fu3_mix_tc_pack(const float* w, float* wp, int Cin, int Cout, float scale, ffc_stream_t st);
int fu3_mix_tc_run(const Fu3MixParams& p, ffc_stream_t st);
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
                   int training, float eps, float momentum, unsigned char* ws, ffc_stream_t st) {
    typedef Fu3G<N> G3;
    const Fu3Plan pl = fu3_plan(B, Cin, Cout, N, training);
    double* sums = reinterpret_cast<double*>(ws + pl.off_sums);
    float* bn_a = reinterpret_cast<float*>(ws + pl.off_consts);
    float* bn_b = bn_a + 2 * Cout;
    float* wp = reinterpret_cast<float*>(ws + pl.off_wp);
    float* S = reinterpret_cast<float*>(ws + pl.off_s);
    float* Y = reinterpret_cast<float*>(ws + pl.off_y);
    const float scale = 1.0f / (float)N;
    bool tc = false;
#ifndef FFC_EMU
    tc = !g_fu3_simt_mix && fu3_mix_tc_supported(Cin, Cout);
    if (tc) FFC_CHECK(fu3_mix_tc_pack(w, wp, Cin, Cout, scale, st));
#endif
    Fu3FinalizeParams fp;
    fp.sums = sums; fp.gamma = gamma; fp.beta = beta; fp.running_mean = running_mean; fp.running_var = running_var;
    fp.save_mean = save_mean; fp.save_invstd = save_invstd; fp.bn_a = bn_a; fp.bn_b = bn_b; fp.Cout = Cout; fp.training = training;
    fp.count = (double)B * N * (N / 2 + 1); fp.eps = eps; fp.momentum = momentum; fp.scale = scale;
    if (training) FFC_CHECK(ffc_memset_async(sums, 0, (size_t)4 * Cout * sizeof(double), st));
    else FFC_CHECK((ffc_launch<Fu3Finalize>(1, 1, 1, 256, 0, st, fp)));
    for (int b0 = 0; b0 < B; b0 += pl.chunk) {
        const int g = (B - b0) < pl.chunk ? (B - b0) : pl.chunk;
        Fu3FwdFftParams ap; ap.x = x + (size_t)b0 * Cin * N * N; ap.spec = S; ap.nplanes = g * Cin;
        FFC_CHECK((ffc_launch<Fu3Rfft2<N, false>>(ffc_cdiv(ap.nplanes, G3::P), 1, 1, G3::kThreads, Fu3Rfft2<N, false>::smem_bytes(), st, ap)));
        Fu3MixParams mp;
        mp.s = S; mp.y = training ? Y + (size_t)b0 * Cout * pl.region : Y; mp.w = w; mp.wp = tc ? wp : nullptr;
        mp.bn_a = training ? nullptr : bn_a; mp.bn_b = training ? nullptr : bn_b; mp.sums = training ? sums : nullptr;
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
        if (!training) {
            Fu3InvFftParams ip; ip.spec = Y; ip.bn_a = nullptr; ip.bn_b = nullptr;
            ip.residual = residual ? residual + (size_t)b0 * Cout * N * N : nullptr; ip.out = out + (size_t)b0 * Cout * N * N;
            ip.nplanes = g * Cout; ip.cout = Cout; ip.scale = 1.0f;
            FFC_CHECK((ffc_launch<Fu3Irfft2<N, false>>(ffc_cdiv(ip.nplanes, G3::P), 1, 1, G3::kThreads, Fu3Irfft2<N, false>::smem_bytes(), st, ip)));
        }
    }
    if (training) {
        FFC_CHECK((ffc_launch<Fu3Finalize>(1, 1, 1, 256, 0, st, fp)));
        Fu3InvFftParams ip; ip.spec = Y; ip.bn_a = bn_a; ip.bn_b = bn_b; ip.residual = residual; ip.out = out;
        ip.nplanes = B * Cout; ip.cout = Cout; ip.scale = 1.0f;
        FFC_CHECK((ffc_launch<Fu3Irfft2<N, false>>(ffc_cdiv(ip.nplanes, G3::P), 1, 1, G3::kThreads, Fu3Irfft2<N, false>::smem_bytes(), st, ip)));
    }
    return FFC_OK;
}

// L2-staged FourierUnitSN forward; same contract as ffc_fu_fwd (ffc_fu2.cu) with workspace >= ffc_fu3_workspace_bytes(...).
extern "C" int ffc_fu3_fwd(const float* x, const float* w, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                           const float* residual, float* out,
                           int B, int Cin, int Cout, int H, int W, int training, float eps, float momentum,
                           void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(x && w && gamma && beta && save_mean && save_invstd && out, "ffc_fu3_fwd: null pointer");
    FFC_REQUIRE(training || (running_mean && running_var), "ffc_fu3_fwd: eval mode needs running statistics");
    FFC_REQUIRE(B >= 0, "ffc_fu3_fwd: negative batch");
    if (B == 0) return FFC_OK;
    FFC_REQUIRE(ffc_fu3_supported(B, Cin, Cout, H, W), "ffc_fu3_fwd: unsupported shape B=%d Cin=%d Cout=%d %dx%d", B, Cin, Cout, H, W);
    FFC_REQUIRE((((uintptr_t)x | (uintptr_t)out | (uintptr_t)residual) & 15) == 0, "ffc_fu3_fwd: x/out/residual must be 16-byte aligned");
    const uintptr_t wsa = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    if (!workspace || wsa + fu3_plan(B, Cin, Cout, H, training).total - 256 > (uintptr_t)workspace + workspace_bytes) {
        ffc_set_error("ffc_fu3_fwd: workspace too small (%zu bytes needed)", fu3_plan(B, Cin, Cout, H, training).total);
        return FFC_ERR_WORKSPACE;
    }
    unsigned char* ws = reinterpret_cast<unsigned char*>(wsa);
    ffc_stream_t st = (ffc_stream_t)stream;
    switch (H) {
        case 16: return fu3_run<16>(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out, B, Cin, Cout, training, eps, momentum, ws, st);
        case 32: return fu3_run<32>(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out, B, Cin, Cout, training, eps, momentum, ws, st);
        case 64: return fu3_run<64>(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out, B, Cin, Cout, training, eps, momentum, ws, st);
        default: return fu3_run<128>(x, w, gamma, beta, running_mean, running_var, save_mean, save_invstd, residual, out, B, Cin, Cout, training, eps, momentum, ws, st);
    }
}
