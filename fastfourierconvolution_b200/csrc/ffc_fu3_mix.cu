// Fu3Mix: the real/imag channel mix of FourierUnitSN (layers/ffc/fourier_unity.py:45, a 1x1 convolution over the 2C
// re/im channels of every spectrum bin) on the 5th-generation tensor cores, with the BatchNorm statistics or the
// BatchNorm + ReLU of :49 in its epilogue.  Device build only (the host emulation build uses Fu3MixSimt, ffc_fu3.cu).
//
//   Y[bin][n] = sum_k S[bin][k] * W[n][k],   bin = one complex slot of one image (M = images * NB rows), k = 2c + {re, im}
//
//   GEMM tile: D[128 bins][NT = 2*Cout padded to 16] in tensor memory, K = 2*Cin in chunks of 32, FP32 accuracy by 3xTF32
//   (hi = x truncated to TF32, lo = x - hi; D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, FP32 accumulation in TMEM).
//
//   * PERSISTENT CTAs of NWG independent warpgroups.  Each warpgroup owns 128 TMEM lanes x (64 + NT) columns and walks
//     its own tiles: gather -> MMA -> epilogue.  No dedicated MMA / loader warps: tcgen05.mma is issued by one elected
//     thread, so the warpgroup's first warp issues its own MMAs after a warpgroup-wide named barrier.  While one
//     warpgroup waits for global loads or runs its epilogue the others keep the tensor core and the LSU busy.
//   * A never touches shared memory: thread = bin = TMEM lane.  A chunk is 16 complex channels = 16 coalesced 8-byte
//     loads (consecutive lanes = consecutive bins of a plane), split hi/lo in registers, tcgen05.st into the
//     warpgroup's A stage; the MMA reads A from TMEM (.ts form).  The loads of chunk c+1 are issued before the wait for
//     the MMAs of chunk c (one A stage per warpgroup suffices: the other warpgroups fill the gap).
//   * B: the whole packed weight matrix (hi | lo, K-major no-swizzle UMMA tiles, forward transform scale folded in)
//     is resident in shared memory for the life of the CTA: one cp.async.bulk per K chunk at start.
//   * Epilogue per tile: tcgen05.ld 16 columns (= 8 complex output channels) at a time; either BN + ReLU with folded
//     constants, or the per-channel sum / sum of squares over the real bins (pad slots masked): the tile goes column-major
//     into a shared-memory staging tile of the warpgroup, each thread then adds up 64 bins of one column (a warp-shuffle
//     butterfly cost 40 % of the kernel's instructions, ncu r02d), accumulates in double over all its tiles, and the CTA
//     flushes one double atomic per channel at the end; one coalesced 8-byte store per channel writes Y.
//   * The first chunk of the NEXT tile is gathered before the epilogue of the current one, so its latency is hidden.
#include "ffc_fu3.cuh"

#ifndef FFC_EMU
#include "ffc_umma.cuh"

static constexpr int FM_BK = 32;                 // K per chunk (16 complex channels)

static inline int fm_kchunks(int Cin) { return (2 * Cin + FM_BK - 1) / FM_BK; }

static constexpr int FM_SCOL = 132;               // floats per column of the statistics staging tile (128 bins + 4: conflict-free LDS.128)
static constexpr int FM_SBUF = 64 * FM_SCOL;      // floats per warpgroup

// Tiling of the 2*Cout output columns.  Up to 128 columns the whole packed weight matrix (hi | lo) is resident in one CTA's
// shared memory (one column tile).  Wider mixes (more than 64 output channels: the sweep's 96 / 128 / 192) are cut into column
// tiles of NT <= 64 that fit beside the statistics staging tiles; gridDim.y walks the tiles and every tile re-reads the
// spectrum (these shapes are tensor-bound: 2*Cin MACs x 3 per output value).
struct FmPlan { int NT, ntiles, nwg; size_t smem_weights; bool ok; };
static FmPlan fm_plan(int Cin, int Cout) {
    FmPlan pl; pl.ok = false; pl.NT = 0; pl.ntiles = 0; pl.nwg = 0; pl.smem_weights = 0;
    if (Cin < 1 || Cout < 1) return pl;
    const int KC = fm_kchunks(Cin), full = (2 * Cout + 15) / 16 * 16;
    const size_t budget = (size_t)224 * 1024, fixed = 2 * 128 * 4 + 16 * 8 + 64;
    const int widths[] = {full <= 128 ? full : 0, 64, 48, 32, 16};
    for (int wi = 0; wi < 5; ++wi) {
        const int NT = widths[wi];
        if (NT < 16 || NT > full) continue;
        const size_t wbytes = (size_t)KC * 2 * NT * FM_BK * 4;
        const int nwgs[] = {NT <= 64 ? 4 : 2, 2, 1};
        for (int gi = 0; gi < 3; ++gi) {
            const int nwg = nwgs[gi];
            if (nwg * (64 + NT) > 512) continue;
            if (wbytes + fixed + (size_t)nwg * FM_SBUF * 4 > budget) continue;
            pl.NT = NT; pl.ntiles = (2 * Cout + NT - 1) / NT; pl.nwg = nwg; pl.smem_weights = wbytes; pl.ok = true;
            return pl;
        }
    }
    return pl;
}

bool fu3_mix_tc_supported(int Cin, int Cout) { return fm_plan(Cin, Cout).ok; }
size_t fu3_mix_tc_packed_floats(int Cin, int Cout) {
    const FmPlan pl = fm_plan(Cin, Cout);
    return pl.ok ? (size_t)pl.ntiles * fm_kchunks(Cin) * 2 * pl.NT * FM_BK : 0;
}

// wp[tile][chunk][hi | lo][n/8][kk/4][n%8][kk%4]  (the shared-memory image of a K-major no-swizzle UMMA B tile, per chunk)
// transposed: w is [2*Cin][2*Cout] (the forward's weight seen from the backward mix dS = dY W: contraction over its rows)
__global__ void __launch_bounds__(256) fu3_pack_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cin, int Cout, int NT, int KC, int ntiles, float scale, int transposed) {
    const int total = ntiles * KC * NT * FM_BK;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kk = e % FM_BK, nl = (e / FM_BK) % NT, chunk = (e / (FM_BK * NT)) % KC, tile = e / (FM_BK * NT * KC);
        const int k = chunk * FM_BK + kk, n = tile * NT + nl;
        float v = 0.f;
        if (n < 2 * Cout && k < 2 * Cin) v = __ldg(transposed ? w + (size_t)k * 2 * Cout + n : w + (size_t)n * 2 * Cin + k) * scale;
        const float hi = ffc_tf32_hi(v);
        float* dst = wp + ((size_t)tile * KC + chunk) * 2 * NT * FM_BK + (nl / 8) * 256 + (kk / 4) * 32 + (nl % 8) * 4 + (kk % 4);
        dst[0] = hi;
        dst[(size_t)NT * FM_BK] = v - hi;
    }
}

int fu3_mix_tc_pack(const float* w, float* wp, int Cin, int Cout, float scale, int transposed, ffc_stream_t st) {
    const FmPlan pl = fm_plan(Cin, Cout);
    const int KC = fm_kchunks(Cin);
    int gx = (pl.ntiles * KC * pl.NT * FM_BK + 255) / 256; if (gx > 296) gx = 296;
    fu3_pack_kernel<<<gx, 256, 0, st>>>(w, wp, Cin, Cout, pl.NT, KC, pl.ntiles, scale, transposed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("fu3_pack launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}

__device__ __forceinline__ void fm_named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// N = plane size: NB (complex slots per plane) and SPS (slots per spectrum row) are compile-time, so the per-channel
// address offsets of the gather and of the stores are immediates and the bin -> (image, slot) split is a constant division
template <int NWG, int N>
__global__ void __launch_bounds__(NWG * 128, 1) fu3_mix_kernel(const Fu3MixParams p, const int NT, const int KC, const long long Mtot, const int ntiles) {
    constexpr int SPS = N / 2 + 2, NB = N * SPS;
    extern __shared__ __align__(128) unsigned char fm_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = warp >> 2;
    const uint32_t chunk_bytes = (uint32_t)(2 * NT * FM_BK * 4);
    unsigned char* bsm = fm_smem;                                                       // KC chunks of (hi | lo)
    float* bn_a = reinterpret_cast<float*>(fm_smem + (size_t)KC * chunk_bytes);        // [NT]
    float* bn_b = bn_a + NT;                                                            // [NT]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bn_b + NT);
    uint64_t* w_full = bars;                 // [1]
    uint64_t* a_free = bars + 1;             // [NWG]
    uint64_t* acc_done = a_free + NWG;       // [NWG]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + NWG);
    float* sbuf = reinterpret_cast<float*>(bars + 16) + (size_t)wg * FM_SBUF;           // statistics staging tile of this warpgroup

    if (tid == 0) {
        umma::mbar_init(w_full, 1);
        for (int i = 0; i < NWG; ++i) { umma::mbar_init(&a_free[i], 1); umma::mbar_init(&acc_done[i], 1); }
        umma::fence_barrier_init();
        umma::mbar_arrive_expect_tx(w_full, (uint32_t)KC * chunk_bytes);
        for (int c = 0; c < KC; ++c)
            umma::bulk_g2s(bsm + (size_t)c * chunk_bytes, p.wp + ((size_t)blockIdx.y * KC + c) * 2 * NT * FM_BK, chunk_bytes, w_full);
    }
    for (int i = tid; i < NT; i += blockDim.x) {
        const int n = (int)blockIdx.y * NT + i;            // column tile blockIdx.y covers output columns [nbase, nbase + NT)
        bn_a[i] = (p.bn_a && n < 2 * p.Cout) ? __ldg(p.bn_a + n) : 0.f;
        bn_b[i] = (p.bn_b && n < 2 * p.Cout) ? __ldg(p.bn_b + n) : 0.f;
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t wg_cols = 64u + (uint32_t)NT;                    // A stage (hi 32 | lo 32) + accumulator
    const uint32_t a_col = tbase + (uint32_t)wg * wg_cols, d_col = a_col + 64u;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;   // this warp's TMEM lanes
    const int row = tid & 127;
    const uint32_t idesc = umma::idesc_tf32(128, NT);
    const uint32_t b0 = umma::smem_u32(bsm);
    const bool issuer = (warp & 3) == 0;                             // first warp of the warpgroup issues its MMAs
    const int Cin = p.Cin, Cout = p.Cout, nbase = (int)blockIdx.y * NT;
    const float2* S = reinterpret_cast<const float2*>(p.s);
    float2* Y = reinterpret_cast<float2*>(p.y);
    const bool do_bn = p.bn_a != nullptr, do_stats = p.sums != nullptr;
    const int stride = gridDim.x * NWG;

    // statistics: this thread sums column (64 * half + (row & 63)) over bins [64 * (row >> 6), +64) of every tile
    double st_sum[2] = {0.0, 0.0}, st_sq[2] = {0.0, 0.0};

    auto locate = [&](int tile, bool& ok, int& b, int& r) {
        const long long m = (long long)tile * 128 + row;
        ok = tile < ntiles && m < Mtot;
        b = ok ? (int)(m / NB) : 0;
        r = ok ? (int)(m % NB) : 0;
    };
    // 16 complex channels of one bin: coalesced 8-byte loads, immediate offsets j * NB
    auto gather = [&](float2 (&v)[16], const float2* sp, bool ok, int c) {
        const float2* q = sp + (size_t)c * 16 * NB;
        const int left = Cin - c * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (ok && j < left) ? __ldg(q + (size_t)j * NB) : make_float2(0.f, 0.f);
    };

    uint32_t q = 0;             // chunks this warpgroup has handed to the tensor core so far (a_free phase counter)
    uint32_t t_done = 0;        // tiles finished (acc_done phase counter)
    bool weights_ready = false;
    int tile = blockIdx.x * NWG + wg;
    bool ok; int b, r;
    locate(tile, ok, b, r);
    float2 v[16];
    gather(v, S + (size_t)b * Cin * NB + r, ok, 0);
    while (tile < ntiles) {
        const float2* sp = S + (size_t)b * Cin * NB + r;
        for (int c = 0; c < KC; ++c) {
            if (q > 0) umma::mbar_wait(&a_free[wg], (q - 1) & 1u);          // the MMAs that read the A stage are done
            umma::fence_after_sync();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 x = v[8 * h + j];
                    const float xh = ffc_tf32_hi(x.x);
                    const float yh = ffc_tf32_hi(x.y);
                    hi[2 * j] = __float_as_uint(xh); hi[2 * j + 1] = __float_as_uint(yh);
                    lo[2 * j] = __float_as_uint(x.x - xh); lo[2 * j + 1] = __float_as_uint(x.y - yh);
                }
                umma::tmem_st16(lane_base + a_col + 16 * h, hi);
                umma::tmem_st16(lane_base + a_col + 32 + 16 * h, lo);
            }
            umma::wait_st();
            umma::fence_before_sync();
            fm_named_barrier(1 + wg, 128);                                   // the whole A chunk of this warpgroup is in TMEM
            if (issuer) {
                if (!weights_ready) { umma::mbar_wait(w_full, 0); weights_ready = true; }
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    const uint32_t b_hi = b0 + (uint32_t)c * chunk_bytes, b_lo = b_hi + (uint32_t)(NT * FM_BK * 4);
#pragma unroll
                    for (int ks = 0; ks < FM_BK / 8; ++ks) {
                        const uint64_t dh = umma::smem_desc_kmajor_noswizzle(b_hi + ks * 256, 128, 1024);
                        const uint64_t dl = umma::smem_desc_kmajor_noswizzle(b_lo + ks * 256, 128, 1024);
                        umma::mma_tf32_ts(d_col, a_col + ks * 8, dh, idesc, (c | ks) ? 1u : 0u);          // hi * hi
                        umma::mma_tf32_ts(d_col, a_col + 32 + ks * 8, dh, idesc, 1u);                    // lo * hi
                        umma::mma_tf32_ts(d_col, a_col + ks * 8, dl, idesc, 1u);                         // hi * lo
                    }
                    umma::commit(&a_free[wg]);
                    if (c == KC - 1) umma::commit(&acc_done[wg]);
                }
                __syncwarp();
            }
            ++q;
            if (c + 1 < KC) gather(v, sp, ok, c + 1);          // in flight while the tensor core works on chunk c
        }
        // ---------------- the next tile's first chunk is fetched now: its latency hides behind this tile's epilogue
        const int ntile = tile + stride;
        bool nok; int nb_, nr;
        locate(ntile, nok, nb_, nr);
        gather(v, S + (size_t)nb_ * Cin * NB + nr, nok, 0);
        // ---------------- epilogue of this tile
        umma::mbar_wait(&acc_done[wg], t_done & 1u);
        ++t_done;
        umma::fence_after_sync();
        const bool real_bin = ok && (r % SPS) != SPS - 1;
        float2* yp = Y ? Y + (size_t)b * Cout * NB + r : nullptr;
#pragma unroll
        for (int half = 0; half < 2; ++half) {                  // 64 output columns (32 complex channels) per pass
            if (64 * half >= NT) break;
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
                const int n0 = 64 * half + 16 * gg;
                if (n0 >= NT) break;
                uint32_t rr[16];
                umma::tmem_ld16(lane_base + d_col + (uint32_t)n0, rr);
                umma::wait_ld();
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(rr[j]);
                if (do_stats) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) sbuf[(16 * gg + j) * FM_SCOL + row] = real_bin ? f[j] : 0.f;
                }
                if (do_bn) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float z = fmaf(f[j], bn_a[n0 + j], bn_b[n0 + j]);
                        f[j] = z > 0.f ? z : 0.f;
                    }
                }
                if (yp && ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int o = ((nbase + n0) >> 1) + j;
                        if (o < Cout) yp[(size_t)o * NB] = make_float2(f[2 * j], f[2 * j + 1]);
                    }
                }
            }
            if (do_stats) {
                // transposed reduction through shared memory: the tile's 64 columns x 128 bins were written bin-major per
                // column; now each thread adds up 64 bins of one column (16 conflict-free 16-byte loads)
                fm_named_barrier(1 + wg, 128);
                const float4* src = reinterpret_cast<const float4*>(sbuf + (row & 63) * FM_SCOL + (row >> 6) * 64);
                float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float4 a = src[i], c4 = src[i + 1];
                    s0 += (a.x + a.y) + (a.z + a.w); s1 += (c4.x + c4.y) + (c4.z + c4.w);
                    q0 = fmaf(a.x, a.x, q0); q0 = fmaf(a.y, a.y, q0); q0 = fmaf(a.z, a.z, q0); q0 = fmaf(a.w, a.w, q0);
                    q1 = fmaf(c4.x, c4.x, q1); q1 = fmaf(c4.y, c4.y, q1); q1 = fmaf(c4.z, c4.z, q1); q1 = fmaf(c4.w, c4.w, q1);
                }
                st_sum[half] += (double)(s0 + s1);
                st_sq[half] += (double)(q0 + q1);
                fm_named_barrier(1 + wg, 128);                   // the staging tile may be overwritten
            }
        }
        umma::fence_before_sync();       // the TMEM loads above are ordered before the next tile's MMAs (issued after a barrier)
        tile = ntile; ok = nok; b = nb_; r = nr;
    }
    if (do_stats) {
        // CTA-level reduction before the global atomics: the weight tiles are dead now (every MMA of this CTA has completed),
        // their shared memory holds the partials [warpgroup][bin half][column][sum | sum of squares]
        __syncthreads();
        double* red = reinterpret_cast<double*>(bsm);
        const int part = wg * 2 + (row >> 6);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int n = 64 * half + (row & 63);
            if (n < NT) {
                red[((size_t)part * NT + n) * 2] = st_sum[half];
                red[((size_t)part * NT + n) * 2 + 1] = st_sq[half];
            }
        }
        __syncthreads();
        for (int i = tid; i < 2 * NT; i += blockDim.x) {
            const int n = i >> 1, which = i & 1;
            if (nbase + n < 2 * Cout) {
                double acc = 0.0;
                for (int w = 0; w < 2 * NWG; ++w) acc += red[((size_t)w * NT + n) * 2 + which];
                atomicAdd(p.sums + (which ? 2 * Cout + nbase + n : nbase + n), acc);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

template <int NWG, int N>
static int fu3_mix_launch(const Fu3MixParams& p, int NT, int KC, long long Mtot, int ntiles, int grid, int ncoltiles, size_t smem, ffc_stream_t st) {
    static FfcPerDevice configured_dev = {};
    size_t& configured = *ffc_device_slot(configured_dev);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(fu3_mix_kernel<NWG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ffc_set_error("cudaFuncSetAttribute(fu3_mix, %zu B): %s", smem, cudaGetErrorString(e)); return FFC_ERR_CUDA; }
        configured = smem;
    }
    fu3_mix_kernel<NWG, N><<<dim3(grid, ncoltiles), NWG * 128, smem, st>>>(p, NT, KC, Mtot, ntiles);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("fu3_mix launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}

template <int N>
static int fu3_mix_run_n(const Fu3MixParams& p, ffc_stream_t st) {
    const FmPlan pl = fm_plan(p.Cin, p.Cout);
    if (!pl.ok) { ffc_set_error("fu3_mix: %d -> %d channels do not fit the tensor-core mix", p.Cin, p.Cout); return FFC_ERR_BAD_ARG; }
    const int NT = pl.NT, KC = fm_kchunks(p.Cin), nwg = pl.nwg;
    const long long Mtot = (long long)p.G * p.NB;
    const int ntiles = (int)((Mtot + 127) / 128);
    // weights | BN constants | barriers (16 x 8 bytes) | statistics staging tiles (training mode)
    const size_t smem = (size_t)KC * 2 * NT * FM_BK * 4 + 2 * NT * 4 + 16 * 8 + (p.sums ? (size_t)nwg * FM_SBUF * 4 : 0) + 64;
    int grid = (ntiles + nwg - 1) / nwg;
    int per_col = ffc_sm_count() / pl.ntiles;                  // persistent CTAs per column tile
    if (per_col < 1) per_col = 1;
    if (grid > per_col) grid = per_col;
    if (grid < 1) grid = 1;
    if (nwg == 4) return fu3_mix_launch<4, N>(p, NT, KC, Mtot, ntiles, grid, pl.ntiles, smem, st);
    if (nwg == 2) return fu3_mix_launch<2, N>(p, NT, KC, Mtot, ntiles, grid, pl.ntiles, smem, st);
    return fu3_mix_launch<1, N>(p, NT, KC, Mtot, ntiles, grid, pl.ntiles, smem, st);
}

int fu3_mix_tc_run(const Fu3MixParams& p, ffc_stream_t st) {
    switch (p.SPS) {                          // SPS = N / 2 + 2
        case 10: return fu3_mix_run_n<16>(p, st);
        case 18: return fu3_mix_run_n<32>(p, st);
        case 34: return fu3_mix_run_n<64>(p, st);
        case 66: return fu3_mix_run_n<128>(p, st);
        default: ffc_set_error("fu3_mix: unsupported plane (SPS = %d)", p.SPS); return FFC_ERR_BAD_ARG;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// fu3_wgrad_kernel: BatchNorm backward + the weight gradient of the channel mix on the tensor cores.
//
//   dY[bin][n] = coef[n] * (g[bin][n] - c1[n] - xhat[bin][n] * c2[n])          (written back: the dS mix reads it)
//   dW[n][k]   = scale * sum_bins dY[bin][n] * S[bin][k]                        n = 2*o + {re, im}, k = 2*c + {re, im}
//
//   GEMM: D[128 rows n][NT columns k] in tensor memory, contraction over the BINS in chunks of 32; both operands come from
//   shared memory (SS form), K-major no-swizzle canonical tiles [row / 8][bin / 4][row % 8][bin % 4], hi | lo for 3xTF32.
//   * 8 BUILDER warps: a unit is (complex channel, 4 consecutive bins) = 32 contiguous bytes of a plane; eight lanes read
//     256 contiguous bytes of one channel, compute dY (A side), split hi / lo and store two rows (re, im) of the tile as
//     16-byte stores that are bank-conflict free (lane -> (channel % 4, bin quad)).  All loads of a thread's units are
//     issued before the first use.  fence.proxy.async, then one mbarrier arrival per thread on the stage's `full`.
//   * 1 MMA warp: 12 MMAs per chunk (4 k-steps x {hi*hi, lo*hi, hi*lo}), tcgen05.commit on the stage's `empty`.
//   * persistent CTAs over the chunks (a chunk never straddles two images: NB % 32 == 0); at the end the builder warps
//     read the accumulator and add it to dW with one float atomic per element and CTA.
// ------------------------------------------------------------------------------------------------------------------
static constexpr int WG_BK = 32;                          // bins per chunk
static constexpr int WG_BUILDERS = 256;

__device__ __forceinline__ void wg_cp16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(umma::smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void wg_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wg_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// a = (re0, im0, re1, im1), b = (re2, im2, re3, im3) of complex channel `ch`, bins 4*jq .. 4*jq + 3 -> rows 2*ch, 2*ch + 1 of
// the canonical tile (hi and lo).  Lanes are (channel % 4, bin quad): lanes of an even bin quad store the re row first, lanes
// of an odd one the im row, so the eight lanes of a quarter-warp cover all 32 banks with each 16-byte store.  (Without the
// ordering every store is a 2-way conflict: 13 M extra wavefronts and the L1 / shared-memory pipe at 89 %, ncu r03h -- the
// pipe, shared with the cp.async ring and the tensor core's operand reads, is what bounds this kernel.)
__device__ __forceinline__ void wg_store_rows(float* tile_hi, float* tile_lo, int ch, int jq, const float4 a, const float4 b) {
    const bool odd = jq & 1;
    // first / second row of this lane: (re, im) for even bin quads, (im, re) for odd ones
    const float f[4] = {odd ? a.y : a.x, odd ? a.w : a.z, odd ? b.y : b.x, odd ? b.w : b.z};
    const float g[4] = {odd ? a.x : a.y, odd ? a.z : a.w, odd ? b.x : b.y, odd ? b.z : b.w};
    float fh[4], fl[4], gh[4], gl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        fh[i] = ffc_tf32_hi(f[i]); fl[i] = f[i] - fh[i];
        gh[i] = ffc_tf32_hi(g[i]); gl[i] = g[i] - gh[i];
    }
    const int off = ((2 * ch) >> 3) * 256 + jq * 32 + ((2 * ch) & 7) * 4;
    const int o1 = off + (odd ? 4 : 0), o2 = off + (odd ? 0 : 4);
    *reinterpret_cast<float4*>(tile_hi + o1) = make_float4(fh[0], fh[1], fh[2], fh[3]);
    *reinterpret_cast<float4*>(tile_hi + o2) = make_float4(gh[0], gh[1], gh[2], gh[3]);
    *reinterpret_cast<float4*>(tile_lo + o1) = make_float4(fl[0], fl[1], fl[2], fl[3]);
    *reinterpret_cast<float4*>(tile_lo + o2) = make_float4(gl[0], gl[1], gl[2], gl[3]);
}

// RAW: stages of the cp.async ring (prefetch distance RAW - 1 chunks); NCANON: canonical (tensor-core) stages; U: passes of 32
// channels per operand (1: <= 32 channels).  Warps 0-7 build the A operand (g, y -> dY), warps 8-15 the B operand (S), warp 16
// issues the MMAs: 16 builder warps keep four warps per scheduler busy (with 8, the serial chain wait -> LDS -> split -> STS ->
// fence -> arrive of a warp was exposed: ncu r02u, issue active 21 %).
// STACK (2*Cout <= 64): the hi and the lo half of dY are rows 0..63 and 64..127 of ONE 128-row A tile, and the hi and lo halves of
// S are columns [0, NT) and [NT, 2*NT) of ONE B tile, so a single MMA per 8 bins computes hi*hi, hi*lo, lo*hi (and lo*lo) into
// four blocks of a 128 x 2*NT accumulator that the reduction kernel adds up: 4 MMAs per chunk instead of 12.  Without STACK
// (up to 128 rows) the B halves are still concatenated: 2 MMAs per 8 bins.  (Measured: the kernel ran at ~240 cycles per
// K = 8 MMA whatever the prefetch depth -- the MMA count, not the memory system, set its pace.)
template <int RAW, int NCANON, int U, bool STACK>
__global__ void __launch_bounds__(2 * WG_BUILDERS + 32, 1) fu3_wgrad_kernel(const Fu3BwdWgradParams p, const int NT, const int nchunks, const int tmem_cols) {
    extern __shared__ __align__(128) unsigned char wg_smem[];
    constexpr uint32_t rawA = 4 * U * WG_BUILDERS * 16, rawB = 2 * U * WG_BUILDERS * 16, raw_bytes = rawA + rawB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t a_rows = STACK ? 64 : 128;                                        // rows of one of hi / lo
    const uint32_t a_bytes = a_rows * WG_BK * 4, b_bytes = (uint32_t)NT * WG_BK * 4;     // one of hi / lo
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    unsigned char* raw = wg_smem + (size_t)NCANON * stage_bytes;
    float* consts = reinterpret_cast<float*>(raw + (size_t)RAW * raw_bytes);             // alpha | beta | gamma, 128 each (+ 256 spare)
    uint64_t* full = reinterpret_cast<uint64_t*>(consts + 5 * 128);
    uint64_t* empty = full + 4;
    uint64_t* done = empty + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    // channel tile of this CTA: 64 complex channels of dY (blockIdx.y) by 64 complex channels of S (blockIdx.z)
    const int NB = p.NB, o0 = (int)blockIdx.y * 64, c0 = (int)blockIdx.z * 64;
    const int Cout = (p.Cout - o0) < 64 ? (p.Cout - o0) : 64, Cin = (p.Cin - c0) < 64 ? (p.Cin - c0) : 64;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { umma::mbar_init(&full[i], 2 * WG_BUILDERS / 32); umma::mbar_init(&empty[i], 1); }      // one arrival per builder warp
        umma::mbar_init(done, 1);
        umma::fence_barrier_init();
    }
    // rows beyond 2*Cout / 2*Cin of every canonical stage stay zero for the life of the CTA
    for (uint32_t i = tid; i < (uint32_t)NCANON * stage_bytes / 16; i += blockDim.x) reinterpret_cast<float4*>(wg_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // dY = coef * (g - c1 - (y - mean) * invstd * c2) = alpha * g + beta * y + gamma: two FMAs per value
    for (int i = tid; i < 128; i += blockDim.x) {
        const bool ok = i < 2 * Cout;
        const int n = 2 * o0 + i;
        const float coef = ok ? __ldg(p.coef + n) : 0.f, c1 = ok ? __ldg(p.c1 + n) : 0.f, c2 = ok ? __ldg(p.c2 + n) : 0.f;
        const float mean = ok ? __ldg(p.mean + n) : 0.f, invstd = ok ? __ldg(p.invstd + n) : 0.f;
        const float beta = -coef * invstd * c2;
        consts[i] = coef; consts[128 + i] = beta; consts[256 + i] = -coef * c1 - beta * mean;
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
    umma::fence_proxy_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;

    if (warp < 2 * WG_BUILDERS / 32) {
        // ------------------------------------------------------------------------------------------------ builders
        const bool sideB = warp >= WG_BUILDERS / 32;
        const int bt = tid & (WG_BUILDERS - 1), bw = warp & 7;
        const int jq = lane >> 2, chl = bw * 4 + (lane & 3);     // bin quad, channel within a pass of 32 channels
        const int C = sideB ? Cin : Cout, Ctot = sideB ? p.Cin : p.Cout, cbase = sideB ? c0 : o0;
        const unsigned cpi = (unsigned)(NB / WG_BK);             // chunks per image (NB % 32 == 0)
        if constexpr (RAW == 0) {
            // DIRECT form: the units of the next chunk are loaded straight into registers (one chunk ahead) while this chunk is
            // processed.  The cp.async ring below costs a shared-memory write and a read per loaded byte, and the L1 / shared
            // memory pipe -- shared with the canonical-tile stores and the tensor core's operand reads -- bounds this kernel.
            constexpr int NR = 4 * U;
            float4 cur[NR], nxt[NR];
            auto load = [&](int chunk, float4 (&dst)[NR]) {
                const int b = (int)((unsigned)chunk / cpi), r = (int)((unsigned)chunk % cpi) * WG_BK;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int o = chl + 32 * u;
                    if (o < C) {
                        const size_t off = (((size_t)b * Ctot + cbase + o) * NB + r) * 2 + 8 * jq;
                        if (sideB) {
                            dst[4 * u + 0] = __ldg(reinterpret_cast<const float4*>(p.s + off)); dst[4 * u + 1] = __ldg(reinterpret_cast<const float4*>(p.s + off + 4));
                        } else {
                            dst[4 * u + 0] = __ldg(reinterpret_cast<const float4*>(p.g + off)); dst[4 * u + 1] = __ldg(reinterpret_cast<const float4*>(p.g + off + 4));
                            dst[4 * u + 2] = __ldg(reinterpret_cast<const float4*>(p.y + off)); dst[4 * u + 3] = __ldg(reinterpret_cast<const float4*>(p.y + off + 4));
                        }
                    }
                }
            };
            if ((int)blockIdx.x < nchunks) load(blockIdx.x, cur);
            int it = 0;
            for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
                const int s = it % NCANON, use = it / NCANON;
                const int b = (int)((unsigned)chunk / cpi), r = (int)((unsigned)chunk % cpi) * WG_BK;
                if (chunk + (int)gridDim.x < nchunks) load(chunk + gridDim.x, nxt);
                if (use > 0) umma::mbar_wait(&empty[s], (uint32_t)(use - 1) & 1u);          // the MMAs that read this stage are done
                float* a_hi = reinterpret_cast<float*>(wg_smem + (size_t)s * stage_bytes);
                float* a_lo = a_hi + a_rows * WG_BK;
                float* b_hi = a_lo + a_rows * WG_BK;
                float* b_lo = b_hi + NT * WG_BK;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int o = chl + 32 * u;
                    if (o >= C) continue;
                    if (sideB) {
                        wg_store_rows(b_hi, b_lo, o, jq, cur[4 * u + 0], cur[4 * u + 1]);
                    } else {
                        const float4 g0 = cur[4 * u + 0], g1 = cur[4 * u + 1], y0 = cur[4 * u + 2], y1 = cur[4 * u + 3];
                        const int n = 2 * o;
                        const float ar = consts[n], ai = consts[n + 1], br = consts[128 + n], bi = consts[128 + n + 1];
                        const float cr = consts[256 + n], ci = consts[256 + n + 1];
                        float4 d0, d1;
                        d0.x = fmaf(ar, g0.x, fmaf(br, y0.x, cr)); d0.y = fmaf(ai, g0.y, fmaf(bi, y0.y, ci));
                        d0.z = fmaf(ar, g0.z, fmaf(br, y0.z, cr)); d0.w = fmaf(ai, g0.w, fmaf(bi, y0.w, ci));
                        d1.x = fmaf(ar, g1.x, fmaf(br, y1.x, cr)); d1.y = fmaf(ai, g1.y, fmaf(bi, y1.y, ci));
                        d1.z = fmaf(ar, g1.z, fmaf(br, y1.z, cr)); d1.w = fmaf(ai, g1.w, fmaf(bi, y1.w, ci));
                        if (blockIdx.z == 0) {                      // one column tile writes dY
                            const size_t off = (((size_t)b * p.Cout + o0 + o) * NB + r) * 2 + 8 * jq;
                            *reinterpret_cast<float4*>(p.dy + off) = d0; *reinterpret_cast<float4*>(p.dy + off + 4) = d1;
                        }
                        wg_store_rows(a_hi, a_lo, o, jq, d0, d1);
                    }
                }
                umma::fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&full[s]);           // one arrival per warp: 512 arrivals on one mbarrier serialise
#pragma unroll
                for (int i = 0; i < NR; ++i) cur[i] = nxt[i];
            }
        } else {
        // every thread copies ITS OWN units of chunk i + RAW - 1 into the raw ring with cp.async (16 bytes each) and reads them
        // back after cp.async.wait_group: RAW - 1 chunks of loads per thread stay in flight without holding registers, and no
        // other thread ever touches these bytes, so the ring needs no barrier.
        auto issue = [&](int chunk, int rs) {
            const int b = (int)((unsigned)chunk / cpi), r = (int)((unsigned)chunk % cpi) * WG_BK;     // 32-bit: a 64-bit division is ~100 instructions
            unsigned char* dst = raw + (size_t)rs * raw_bytes + (sideB ? rawA : 0) + (size_t)bt * 16;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int o = chl + 32 * u;
                if (o < C) {
                    const size_t off = (((size_t)b * Ctot + cbase + o) * NB + r) * 2 + 8 * jq;
                    if (sideB) {
                        wg_cp16(dst + (size_t)(2 * u + 0) * WG_BUILDERS * 16, p.s + off); wg_cp16(dst + (size_t)(2 * u + 1) * WG_BUILDERS * 16, p.s + off + 4);
                    } else {
                        wg_cp16(dst + (size_t)(4 * u + 0) * WG_BUILDERS * 16, p.g + off); wg_cp16(dst + (size_t)(4 * u + 1) * WG_BUILDERS * 16, p.g + off + 4);
                        wg_cp16(dst + (size_t)(4 * u + 2) * WG_BUILDERS * 16, p.y + off); wg_cp16(dst + (size_t)(4 * u + 3) * WG_BUILDERS * 16, p.y + off + 4);
                    }
                }
            }
        };
        {
            int c = blockIdx.x;
#pragma unroll
            for (int d = 0; d < RAW - 1; ++d, c += gridDim.x) {
                if (c < nchunks) issue(c, d);
                wg_cp_commit();
            }
        }
        int it = 0;
        for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
            const int s = it % NCANON, use = it / NCANON;
            const int b = (int)((unsigned)chunk / cpi), r = (int)((unsigned)chunk % cpi) * WG_BK;     // 32-bit: a 64-bit division is ~100 instructions
            {
                const long long ahead = (long long)chunk + (long long)(RAW - 1) * gridDim.x;
                if (ahead < nchunks) issue((int)ahead, (it + RAW - 1) % RAW);
                wg_cp_commit();
            }
            wg_cp_wait<(RAW > 0 ? RAW - 1 : 0)>();                               // this thread's copies of chunk `it` have landed
            const float4* mine = reinterpret_cast<const float4*>(raw + (size_t)(it % RAW) * raw_bytes + (sideB ? rawA : 0)) + bt;
            if (use > 0) umma::mbar_wait(&empty[s], (uint32_t)(use - 1) & 1u);          // the MMAs that read this stage are done
            float* a_hi = reinterpret_cast<float*>(wg_smem + (size_t)s * stage_bytes);
            float* a_lo = a_hi + a_rows * WG_BK;
            float* b_hi = a_lo + a_rows * WG_BK;
            float* b_lo = b_hi + NT * WG_BK;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int o = chl + 32 * u;
                if (o >= C) continue;
                if (sideB) {
                    wg_store_rows(b_hi, b_lo, o, jq, mine[(2 * u + 0) * WG_BUILDERS], mine[(2 * u + 1) * WG_BUILDERS]);
                } else {
                    const float4 g0 = mine[(4 * u + 0) * WG_BUILDERS], g1 = mine[(4 * u + 1) * WG_BUILDERS];
                    const float4 y0 = mine[(4 * u + 2) * WG_BUILDERS], y1 = mine[(4 * u + 3) * WG_BUILDERS];
                    const int n = 2 * o;
                    const float ar = consts[n], ai = consts[n + 1], br = consts[128 + n], bi = consts[128 + n + 1];
                    const float cr = consts[256 + n], ci = consts[256 + n + 1];
                    float4 d0, d1;
                    d0.x = fmaf(ar, g0.x, fmaf(br, y0.x, cr)); d0.y = fmaf(ai, g0.y, fmaf(bi, y0.y, ci));
                    d0.z = fmaf(ar, g0.z, fmaf(br, y0.z, cr)); d0.w = fmaf(ai, g0.w, fmaf(bi, y0.w, ci));
                    d1.x = fmaf(ar, g1.x, fmaf(br, y1.x, cr)); d1.y = fmaf(ai, g1.y, fmaf(bi, y1.y, ci));
                    d1.z = fmaf(ar, g1.z, fmaf(br, y1.z, cr)); d1.w = fmaf(ai, g1.w, fmaf(bi, y1.w, ci));
                    if (blockIdx.z == 0) {                      // one column tile writes dY
                        const size_t off = (((size_t)b * p.Cout + o0 + o) * NB + r) * 2 + 8 * jq;
                        *reinterpret_cast<float4*>(p.dy + off) = d0; *reinterpret_cast<float4*>(p.dy + off + 4) = d1;
                    }
                    wg_store_rows(a_hi, a_lo, o, jq, d0, d1);
                }
            }
            umma::fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&full[s]);           // one arrival per warp: 512 arrivals on one mbarrier serialise
        }
        }
        wg_cp_wait<0>();
        // ------------------------------------------------------------------------------------------------ epilogue
        if (warp < 8) {
            umma::mbar_wait(done, 0);
            umma::fence_after_sync();
            const int row = (warp & 3) * 32 + lane;                              // accumulator row = TMEM lane
            const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
            const int col_half = warp >> 2;                                      // warps 0-3: even column groups, 4-7: odd ones
            const int ncols16 = 2 * NT / 16;
            // this CTA's partial tile goes to its own slice of the workspace ([CTA][128][NT], coalesced over the rows' columns is not
            // possible from the TMEM row-per-lane view, so each lane writes 64 contiguous bytes); fu3_wgrad_reduce adds the slices.
            // (One float atomic per element and CTA made 148-way contention on every address: ~40 us for a 128 x 128 tile.)
            float* part = p.part + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 128 * 2 * NT;
            for (int cg = col_half; cg < ncols16; cg += 2) {
                uint32_t rr[16];
                umma::tmem_ld16(lane_base + tbase + (uint32_t)(16 * cg), rr);
                umma::wait_ld();
                float4* dst = reinterpret_cast<float4*>(part + (size_t)row * 2 * NT + 16 * cg);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dst[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
            }
            umma::fence_before_sync();
        }
    } else {
        // ------------------------------------------------------------------------------------------------ MMA warp
        const uint32_t idesc2 = umma::idesc_tf32(128, 2 * NT), idesc1 = umma::idesc_tf32(128, NT);
        const uint32_t base = umma::smem_u32(wg_smem);
        int it = 0;
        for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
            const int s = it % NCANON, use = it / NCANON;
            umma::mbar_wait(&full[s], (uint32_t)use & 1u);
            umma::fence_after_sync();
            if (umma::elect_one()) {
                const uint32_t a_hi = base + (uint32_t)s * stage_bytes, a_lo = a_hi + a_bytes, b_hi = a_lo + a_bytes;
#pragma unroll
                for (int ks = 0; ks < WG_BK / 8; ++ks) {
                    // B descriptor over 2*NT rows: the lo tile follows the hi tile in the canonical layout
                    const uint64_t ah = umma::smem_desc_kmajor_noswizzle(a_hi + ks * 256, 128, 1024);
                    const uint64_t bb = umma::smem_desc_kmajor_noswizzle(b_hi + ks * 256, 128, 1024);
                    umma::mma_tf32_ss(tbase, ah, bb, idesc2, (it | ks) ? 1u : 0u);          // STACK: rows (hi | lo) x columns (hi | lo)
                    if (!STACK) {
                        const uint64_t al = umma::smem_desc_kmajor_noswizzle(a_lo + ks * 256, 128, 1024);
                        umma::mma_tf32_ss(tbase, al, bb, idesc1, 1u);                        // lo * hi into the first NT columns
                    }
                }
                umma::commit(&empty[s]);
            }
            __syncwarp();
        }
        if (umma::elect_one()) umma::commit(done);
        __syncwarp();
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, (uint32_t)tmem_cols);
}

// dw[n][k] = scale * sum over the CTAs' partial tiles [CTA][128][2*NT] of the blocks that hold hi*hi, hi*lo, lo*hi (lo*lo).
// One block per 8 consecutive outputs; 32 groups of threads split the CTAs' tiles, shared-memory sum at the end.
__global__ void __launch_bounds__(256) fu3_wgrad_reduce(const float* __restrict__ part, float* __restrict__ dw, int nparts, int NT, int ld, int row0, int col0,
                                                        int rows, int cols, float scale, int stacked) {
    __shared__ float red[32][9];
    const int el = threadIdx.x & 7, pg = threadIdx.x >> 3;       // 8 consecutive outputs (one 32-byte sector per tile) x 32 groups of tiles
    const int e = blockIdx.x * 8 + el;
    float acc = 0.f;
    if (e < rows * cols) {
        const int n = e / cols, k = e % cols;
        const size_t tile = (size_t)128 * 2 * NT;
        const float* src = part + (size_t)n * 2 * NT + k;
        for (int i = pg; i < nparts; i += 32) {
            const float* q = src + (size_t)i * tile;
            float v = __ldg(q) + __ldg(q + NT);
            if (stacked) v += __ldg(q + (size_t)64 * 2 * NT) + __ldg(q + (size_t)64 * 2 * NT + NT);
            acc += v;
        }
    }
    red[pg][el] = acc;
    __syncthreads();
    if (pg == 0 && e < rows * cols) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) a += red[i][el];
        dw[(size_t)(row0 + e / cols) * ld + col0 + e % cols] = a * scale;
    }
}
// channel tiles of 64 x 64 complex channels, persistent CTAs per tile pair
static void wg_tiling(int Cin, int Cout, int& ty, int& tz, int& NT, int& gx_cap) {
    ty = (Cout + 63) / 64; tz = (Cin + 63) / 64;
    const int cw = Cin < 64 ? Cin : 64;
    NT = (2 * cw + 15) / 16 * 16;
    gx_cap = 296 / (ty * tz);
    if (gx_cap < 1) gx_cap = 1;
}
size_t fu3_wgrad_tc_part_floats(int Cin, int Cout) {
    int ty, tz, NT, cap; wg_tiling(Cin, Cout, ty, tz, NT, cap);
    return (size_t)ty * tz * cap * 128 * 2 * NT;
}

bool fu3_wgrad_tc_supported(int Cin, int Cout) { return Cin >= 1 && Cout >= 1 && ((Cin + 63) / 64) * ((Cout + 63) / 64) <= 64; }

template <int RAW, int NCANON, int U, bool STACK>
static int fu3_wgrad_launch(const Fu3BwdWgradParams& p, int NT, int nchunks, int tmem_cols, size_t smem, ffc_stream_t st) {
    static FfcPerDevice configured_dev = {};
    size_t& configured = *ffc_device_slot(configured_dev);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(fu3_wgrad_kernel<RAW, NCANON, U, STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ffc_set_error("cudaFuncSetAttribute(fu3_wgrad, %zu B): %s", smem, cudaGetErrorString(e)); return FFC_ERR_CUDA; }
        configured = smem;
    }
    int ty, tz, nt_, cap; wg_tiling(p.Cin, p.Cout, ty, tz, nt_, cap);
    int gx = ffc_sm_count() / (ty * tz);
    if (gx < 1) gx = 1;
    if (gx > cap) gx = cap;
    if (gx > nchunks) gx = nchunks;
    fu3_wgrad_kernel<RAW, NCANON, U, STACK><<<dim3(gx, ty, tz), 2 * WG_BUILDERS + 32, smem, st>>>(p, NT, nchunks, tmem_cols);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("fu3_wgrad launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    for (int z = 0; z < tz; ++z)
        for (int y = 0; y < ty; ++y) {
            const int rows = 2 * ((p.Cout - 64 * y) < 64 ? (p.Cout - 64 * y) : 64), cols = 2 * ((p.Cin - 64 * z) < 64 ? (p.Cin - 64 * z) : 64);
            fu3_wgrad_reduce<<<(rows * cols + 7) / 8, 256, 0, st>>>(p.part + (size_t)(z * ty + y) * gx * 128 * 2 * NT, p.dw, gx, NT, 2 * p.Cin,
                                                                    128 * y, 128 * z, rows, cols, p.scale, STACK ? 1 : 0);
            e = cudaGetLastError();
            if (e != cudaSuccess) { ffc_set_error("fu3_wgrad_reduce launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
            ffc_count_launch();
        }
    return FFC_OK;
}

int fu3_wgrad_tc_run(const Fu3BwdWgradParams& p, ffc_stream_t st) {
    int ty, tz, NT, cap; wg_tiling(p.Cin, p.Cout, ty, tz, NT, cap);
    const long long Mtot = (long long)p.B * p.NB;
    if (p.NB % WG_BK != 0) { ffc_set_error("fu3_wgrad: plane slots (%d) must be a multiple of %d", p.NB, WG_BK); return FFC_ERR_BAD_ARG; }
    const int nchunks = (int)(Mtot / WG_BK);
    const int tmem_cols = 2 * NT <= 32 ? 32 : (2 * NT <= 64 ? 64 : (2 * NT <= 128 ? 128 : 256));
    const size_t tail = 5 * 128 * 4 + 10 * 8 + 64, b_tile = (size_t)2 * NT * WG_BK * 4;
    if (p.Cin > 32 || p.Cout > 32) {          // two passes of 32 channels per operand: 2 x (32 KB A + <= 32 KB B) + 2 x 48 KB raw
        const size_t stage = (size_t)2 * 128 * WG_BK * 4 + b_tile, raw = (size_t)12 * WG_BUILDERS * 16;
        (void)raw;
        return fu3_wgrad_launch<0, 3, 2, false>(p, NT, nchunks, tmem_cols, 3 * stage + tail, st);
    }
    // <= 32 channels: stacked A tile (16 KB) + <= 16 KB B, three canonical stages, four raw stages of 24 KB
    const size_t stage = (size_t)2 * 64 * WG_BK * 4 + b_tile, raw = (size_t)6 * WG_BUILDERS * 16;
    (void)raw;
    return fu3_wgrad_launch<0, 4, 1, true>(p, NT, nchunks, tmem_cols, 4 * stage + tail, st);
}
#endif  // !FFC_EMU
