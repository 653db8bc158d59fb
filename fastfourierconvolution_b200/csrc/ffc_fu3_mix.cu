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

static inline int fm_nt(int Cout) { return (2 * Cout + 15) / 16 * 16; }
static inline int fm_kchunks(int Cin) { return (2 * Cin + FM_BK - 1) / FM_BK; }

bool fu3_mix_tc_supported(int Cin, int Cout) {
    const int NT = fm_nt(Cout), KC = fm_kchunks(Cin);
    if (NT > 128 || NT < 16) return false;
    const int nwg = (NT <= 64) ? 4 : 2;                   // + statistics staging tiles of the training-mode epilogue
    return (size_t)KC * 2 * NT * FM_BK * 4 + 4 * NT * 4 + 512 + (size_t)nwg * 64 * 132 * 4 <= (size_t)225 * 1024;
}
size_t fu3_mix_tc_packed_floats(int Cin, int Cout) { return (size_t)fm_kchunks(Cin) * 2 * fm_nt(Cout) * FM_BK; }

// wp[chunk][hi | lo][n/8][kk/4][n%8][kk%4]  (the shared-memory image of a K-major no-swizzle UMMA B tile, per chunk)
// transposed: w is [2*Cin][2*Cout] (the forward's weight seen from the backward mix dS = dY W: contraction over its rows)
__global__ void __launch_bounds__(256) fu3_pack_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cin, int Cout, int NT, int KC, float scale, int transposed) {
    const int total = KC * NT * FM_BK;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kk = e % FM_BK, n = (e / FM_BK) % NT, chunk = e / (FM_BK * NT);
        const int k = chunk * FM_BK + kk;
        float v = 0.f;
        if (n < 2 * Cout && k < 2 * Cin) v = __ldg(transposed ? w + (size_t)k * 2 * Cout + n : w + (size_t)n * 2 * Cin + k) * scale;
        const float hi = ffc_tf32_hi(v);
        float* dst = wp + (size_t)chunk * 2 * NT * FM_BK + (n / 8) * 256 + (kk / 4) * 32 + (n % 8) * 4 + (kk % 4);
        dst[0] = hi;
        dst[(size_t)NT * FM_BK] = v - hi;
    }
}

int fu3_mix_tc_pack(const float* w, float* wp, int Cin, int Cout, float scale, int transposed, ffc_stream_t st) {
    const int NT = fm_nt(Cout), KC = fm_kchunks(Cin);
    int gx = (KC * NT * FM_BK + 255) / 256; if (gx > 64) gx = 64;
    fu3_pack_kernel<<<gx, 256, 0, st>>>(w, wp, Cin, Cout, NT, KC, scale, transposed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("fu3_pack launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}

__device__ __forceinline__ void fm_named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

static constexpr int FM_SCOL = 132;               // floats per column of the statistics staging tile (128 bins + 4: conflict-free LDS.128)
static constexpr int FM_SBUF = 64 * FM_SCOL;      // floats per warpgroup

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
            umma::bulk_g2s(bsm + (size_t)c * chunk_bytes, p.wp + (size_t)c * 2 * NT * FM_BK, chunk_bytes, w_full);
    }
    for (int i = tid; i < NT; i += blockDim.x) {
        bn_a[i] = (p.bn_a && i < 2 * p.Cout) ? __ldg(p.bn_a + i) : 0.f;
        bn_b[i] = (p.bn_b && i < 2 * p.Cout) ? __ldg(p.bn_b + i) : 0.f;
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
    const int Cin = p.Cin, Cout = p.Cout;
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
                        const int o = (n0 >> 1) + j;
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
            if (n < 2 * Cout) {
                double acc = 0.0;
                for (int w = 0; w < 2 * NWG; ++w) acc += red[((size_t)w * NT + n) * 2 + which];
                atomicAdd(p.sums + (which ? 2 * Cout + n : n), acc);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

template <int NWG, int N>
static int fu3_mix_launch(const Fu3MixParams& p, int NT, int KC, long long Mtot, int ntiles, int grid, size_t smem, ffc_stream_t st) {
    static FfcPerDevice configured_dev = {};
    size_t& configured = *ffc_device_slot(configured_dev);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(fu3_mix_kernel<NWG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ffc_set_error("cudaFuncSetAttribute(fu3_mix, %zu B): %s", smem, cudaGetErrorString(e)); return FFC_ERR_CUDA; }
        configured = smem;
    }
    fu3_mix_kernel<NWG, N><<<grid, NWG * 128, smem, st>>>(p, NT, KC, Mtot, ntiles);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("fu3_mix launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}

template <int N>
static int fu3_mix_run_n(const Fu3MixParams& p, ffc_stream_t st) {
    const int NT = fm_nt(p.Cout), KC = fm_kchunks(p.Cin);
    const long long Mtot = (long long)p.G * p.NB;
    const int ntiles = (int)((Mtot + 127) / 128);
    const int nwg = (NT <= 64) ? 4 : 2;
    // weights | BN constants | barriers (16 x 8 bytes) | statistics staging tiles (training mode)
    const size_t smem = (size_t)KC * 2 * NT * FM_BK * 4 + 2 * NT * 4 + 16 * 8 + (p.sums ? (size_t)nwg * FM_SBUF * 4 : 0) + 64;
    int grid = (ntiles + nwg - 1) / nwg;
    if (grid > ffc_sm_count()) grid = ffc_sm_count();
    if (grid < 1) grid = 1;
    if (nwg == 4) return fu3_mix_launch<4, N>(p, NT, KC, Mtot, ntiles, grid, smem, st);
    return fu3_mix_launch<2, N>(p, NT, KC, Mtot, ntiles, grid, smem, st);
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
#endif  // !FFC_EMU
