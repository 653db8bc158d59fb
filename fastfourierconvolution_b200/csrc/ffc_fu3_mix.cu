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
//     constants, or the per-channel sum / sum of squares over the real bins (pad slots masked) by a transposing warp
//     butterfly (16 shuffles per 16 columns), accumulated in double per lane over all tiles of the warpgroup, reduced over
//     the CTA in shared memory and flushed with one double atomic per channel and CTA; then one coalesced 8-byte store
//     per channel.
#include "ffc_fu3.cuh"

#ifndef FFC_EMU
#include "ffc_umma.cuh"

static constexpr int FM_BK = 32;                 // K per chunk (16 complex channels)

static inline int fm_nt(int Cout) { return (2 * Cout + 15) / 16 * 16; }
static inline int fm_kchunks(int Cin) { return (2 * Cin + FM_BK - 1) / FM_BK; }

bool fu3_mix_tc_supported(int Cin, int Cout) {
    const int NT = fm_nt(Cout), KC = fm_kchunks(Cin);
    if (NT > 128 || NT < 16) return false;
    return (size_t)KC * 2 * NT * FM_BK * 4 + 4 * NT * 4 + 512 <= (size_t)200 * 1024;
}
size_t fu3_mix_tc_packed_floats(int Cin, int Cout) { return (size_t)fm_kchunks(Cin) * 2 * fm_nt(Cout) * FM_BK; }

// wp[chunk][hi | lo][n/8][kk/4][n%8][kk%4]  (the shared-memory image of a K-major no-swizzle UMMA B tile, per chunk)
__global__ void __launch_bounds__(256) fu3_pack_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cin, int Cout, int NT, int KC, float scale) {
    const int total = KC * NT * FM_BK;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kk = e % FM_BK, n = (e / FM_BK) % NT, chunk = e / (FM_BK * NT);
        const int k = chunk * FM_BK + kk;
        float v = 0.f;
        if (n < 2 * Cout && k < 2 * Cin) v = __ldg(w + (size_t)n * 2 * Cin + k) * scale;
        const float hi = ffc_tf32_hi(v);
        float* dst = wp + (size_t)chunk * 2 * NT * FM_BK + (n / 8) * 256 + (kk / 4) * 32 + (n % 8) * 4 + (kk % 4);
        dst[0] = hi;
        dst[(size_t)NT * FM_BK] = v - hi;
    }
}

int fu3_mix_tc_pack(const float* w, float* wp, int Cin, int Cout, float scale, ffc_stream_t st) {
    const int NT = fm_nt(Cout), KC = fm_kchunks(Cin);
    int gx = (KC * NT * FM_BK + 255) / 256; if (gx > 64) gx = 64;
    fu3_pack_kernel<<<gx, 256, 0, st>>>(w, wp, Cin, Cout, NT, KC, scale);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("fu3_pack launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}

__device__ __forceinline__ void fm_named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// sum over the 32 lanes of 16 per-lane values by a transposing butterfly: lane l ends with the total of column
// (l >> 1) & 15 ... precisely column col(l) = 8*b4 + 4*b3 + 2*b2 + b1 with b_i = bit i of l (lanes l and l^1 hold the same)
__device__ __forceinline__ float fm_colsum16(const float* v, int lane) {
    float a8[8], a4[4], a2[2];
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float keep = h16 ? v[8 + j] : v[j], send = h16 ? v[j] : v[8 + j];
        a8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = h8 ? a8[4 + j] : a8[j], send = h8 ? a8[j] : a8[4 + j];
        a4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float keep = h4 ? a4[2 + j] : a4[j], send = h4 ? a4[j] : a4[2 + j];
        a2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float keep = h2 ? a2[1] : a2[0], send = h2 ? a2[0] : a2[1];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}
__device__ __forceinline__ int fm_col_of_lane(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

template <int NWG>
__global__ void __launch_bounds__(NWG * 128, 1) fu3_mix_kernel(const Fu3MixParams p, const int NT, const int KC, const long long Mtot, const int ntiles) {
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
    const int NB = p.NB, Cin = p.Cin, Cout = p.Cout;
    const float2* S = reinterpret_cast<const float2*>(p.s);
    float2* Y = reinterpret_cast<float2*>(p.y);
    const bool do_bn = p.bn_a != nullptr, do_stats = p.sums != nullptr;

    // per-lane statistics accumulators: column group g (16 columns) -> (sum, sum of squares) of column 16 g + col(lane)
    double st_sum[8], st_sq[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) { st_sum[g] = 0.0; st_sq[g] = 0.0; }

    uint32_t q = 0;             // chunks this warpgroup has handed to the tensor core so far (a_free phase counter)
    uint32_t t_done = 0;        // tiles finished (acc_done phase counter)
    bool weights_ready = false;
    for (int tile = blockIdx.x * NWG + wg; tile < ntiles; tile += gridDim.x * NWG) {
        const long long m = (long long)tile * 128 + row;
        const bool ok = m < Mtot;
        const int b = ok ? (int)(m / NB) : 0, r = ok ? (int)(m % NB) : 0;
        const float2* sp = S + (size_t)b * Cin * NB + r;
        for (int c = 0; c < KC; ++c) {
            float2 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int ch = c * 16 + j;
                v[j] = (ok && ch < Cin) ? __ldg(sp + (size_t)ch * NB) : make_float2(0.f, 0.f);
            }
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
        }
        // ---------------- epilogue of this tile
        umma::mbar_wait(&acc_done[wg], t_done & 1u);
        ++t_done;
        umma::fence_after_sync();
        const bool real_bin = ok && (r % p.SPS) != p.SPS - 1;
        float2* yp = Y ? Y + (size_t)b * Cout * NB + r : nullptr;
#pragma unroll
        for (int g = 0; g < 8; ++g) {                 // column groups of 16 (fully unrolled: the statistics stay in registers)
            const int n0 = 16 * g;
            if (n0 >= NT) break;
            uint32_t rr[16];
            umma::tmem_ld16(lane_base + d_col + (uint32_t)n0, rr);
            umma::wait_ld();
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(rr[j]);
            if (do_stats) {
                float sq[16], sv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { sv[j] = real_bin ? f[j] : 0.f; sq[j] = sv[j] * sv[j]; }
                st_sum[g] += (double)fm_colsum16(sv, lane);
                st_sq[g] += (double)fm_colsum16(sq, lane);
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
        umma::fence_before_sync();       // the TMEM loads above are ordered before the next tile's MMAs (issued after a barrier)
    }
    if (do_stats) {
        // CTA-level reduction before the global atomics (148 CTAs x 16 warps hammering 4*Cout addresses serialise in L2):
        // the weight tiles are dead now (every MMA of this CTA has completed), their shared memory holds the partials
        __syncthreads();
        double* red = reinterpret_cast<double*>(bsm);                  // [warps][NT][2]
        const int nwarps = NWG * 4;
        if (!(lane & 1)) {
            const int col = fm_col_of_lane(lane);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                if (16 * g < NT) {
                    red[((size_t)warp * NT + 16 * g + col) * 2] = st_sum[g];
                    red[((size_t)warp * NT + 16 * g + col) * 2 + 1] = st_sq[g];
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < 2 * NT; i += blockDim.x) {
            const int n = i >> 1, which = i & 1;
            if (n < 2 * Cout) {
                double acc = 0.0;
                for (int w = 0; w < nwarps; ++w) acc += red[((size_t)w * NT + n) * 2 + which];
                atomicAdd(p.sums + (which ? 2 * Cout + n : n), acc);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

int fu3_mix_tc_run(const Fu3MixParams& p, ffc_stream_t st) {
    const int NT = fm_nt(p.Cout), KC = fm_kchunks(p.Cin);
    const long long Mtot = (long long)p.G * p.NB;
    const int ntiles = (int)((Mtot + 127) / 128);
    const size_t smem = (size_t)KC * 2 * NT * FM_BK * 4 + 2 * NT * 4 + 16 * 8 + 64;
    const int nwg = (NT <= 64) ? 4 : 2;
    int grid = (ntiles + nwg - 1) / nwg;
    if (grid > ffc_sm_count()) grid = ffc_sm_count();
    if (grid < 1) grid = 1;
    static FfcPerDevice configured_dev = {};
    size_t& configured = *ffc_device_slot(configured_dev);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(fu3_mix_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fu3_mix_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ffc_set_error("cudaFuncSetAttribute(fu3_mix, %zu B): %s", smem, cudaGetErrorString(e)); return FFC_ERR_CUDA; }
        configured = smem;
    }
    if (nwg == 4) fu3_mix_kernel<4><<<grid, 512, smem, st>>>(p, NT, KC, Mtot, ntiles);
    else fu3_mix_kernel<2><<<grid, 256, smem, st>>>(p, NT, KC, Mtot, ntiles);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("fu3_mix launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}
#endif  // !FFC_EMU
