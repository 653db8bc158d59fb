// ConvFwdV5: implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 / UMMA) at FP32 accuracy.
// Same contract as ffc_conv2d_fwd_ws (nn.Conv2d / nn.ConvTranspose2d forward and each other's data-gradient, two
// summed input segments, fused bias / addend).  Device build only: the host emulation build keeps ConvFwdV4.
//
//   GEMM per output parity class:  D[M = 128 pixels][N = cout tile <= 256] += A[M][K] * B[N][K]^T,  K = (segment, tap, ci)
//   3xTF32: D_hi += A_hi*B_hi, D_lo += A_lo*B_hi + A_hi*B_lo with hi = x truncated to TF32, lo = x - hi (exact), FP32
//   accumulation in tensor memory (TMEM) over the whole K.  The tensor core truncates when it folds a product group
//   into its accumulator; keeping the 2^-11 times smaller cross terms in their own accumulator means the main sum sees one
//   truncation per K-step instead of three (measured 3x smaller error), and the two are added in the epilogue.
//
//   * A (im2col of the NCHW activations) is never staged in shared memory: each of 128 gather threads owns one output
//     pixel (= one TMEM lane), loads the 32 channels of the current K chunk with coalesced 4-byte loads (consecutive lanes
//     = consecutive pixels), splits hi/lo in registers and writes both straight into TMEM (tcgen05.st); the MMA reads A
//     from TMEM (the .ts form).  Up to four gather warpgroups rotate over as many TMEM stages, so the global-load
//     latency of one chunk hides behind the others.
//   * B (weights) is re-packed once per call by PackV5 into the exact shared-memory image of a K-major no-swizzle UMMA
//     operand tile, hi and lo halves adjacent, zero padded; one 1-D bulk copy (cp.async.bulk) per K chunk and stage,
//     completion counted on an mbarrier.
//   * One thread issues the MMAs (12 per chunk: 4 K-steps x 3 products); tcgen05.commit releases the A stage and the B
//     stage to their producers and finally hands the accumulator to the epilogue.
//   * Epilogue: both warpgroups read their lanes of D from TMEM (tcgen05.ld), add bias / addend and store; for a fixed
//     output channel consecutive lanes are consecutive pixels.
//
// Warp roles (576 threads): four gather warpgroups (warps 0-15) rotate over the K chunks, each with its own TMEM stage
// (two of them when N > 128 leaves room for two stages only); warp 16 issues MMAs and owns the TMEM allocation, warp 17
// streams B.  One CTA per SM (all 512 TMEM columns).
#include "ffc_conv_geom.cuh"

#ifndef FFC_EMU
#include "ffc_umma.cuh"

#define FFC_V5_MAXCLS 4
static constexpr int V5_BK = 32;          // K per chunk (= 4 MMA K-steps of 8)
static constexpr int V5_SB_MAX = 8;       // B stages: as many as fit the shared-memory budget (ConvV5Params::nsb), at most 8
// V5_GW = gather warpgroups = A stages in TMEM is a template parameter of the kernel: 2 (320 threads; two CTAs per SM
// when N <= 64 so that one CTA's prologue / epilogue overlaps the other's main loop; also N > 128, where only two
// stages fit beside the accumulators) or 4 (576 threads, one CTA per SM, 64 < N <= 128).

struct ConvV5Params {
    const float* x[2]; int cin[2]; int cps[2];     // segments: input, channels, 32-channel chunks per tap
    int nseg;
    const float* wp;           // packed weights: [class][n tile][chunk][hi | lo][NT*32]
    long long cls_off[FFC_V5_MAXCLS];     // float offset of each class
    int nt_full;               // N of a full tile (multiple of 16, <= 192)
    int nsb;                   // B stages in shared memory (2 .. V5_SB_MAX)
    int ksplit;                // K chunks of a tile are dealt to ksplit CTAs (gridDim.z = classes * ksplit), which add their
                               // partial sums into a zeroed y with float atomics; 1 = one CTA per tile, plain stores
    const float* bias; const float* addend; float* y;
    float slope;               // epilogue activation x > 0 ? x : slope * x (1 = none, 0.1 = LeakyReLU(0.1), 0 = ReLU)
    float* y1; int cout0;      // block form: output channels [cout0, cout) go to y1 (cout - cout0 channels); else y1 == null, cout0 == cout
    int B, cout, Hi, Wi, Ho, Wo, k, stride, pad, transposed;
};

// N (padded to 16) of tile `t` when cout is cut into tiles of nt_full
__host__ __device__ __forceinline__ int v5_tile_n(int cout, int nt_full, int t) {
    const int rem = cout - t * nt_full;
    const int n = rem < nt_full ? rem : nt_full;
    return (n + 15) / 16 * 16;
}

// ---------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------
struct PackV5Params {
    const float* w[2]; int cin[2]; int cps[2]; int nseg;
    const float* w0b; int cout0;   // block form: segment 0 uses w[0] for co < cout0 and w0b for the rest, segment 1 is zero there
    float* wp; long long cls_off[FFC_V5_MAXCLS];
    int nt_full, ntiles, cout, k, stride, pad, transposed;
    float* zero[2]; long long nzero[2];    // split-K outputs the convolution adds into: zero-filled here (one graph node less)
};

__global__ void __launch_bounds__(256) pack_v5_kernel(const PackV5Params p) {
    const int cls = blockIdx.y;
    const ConvClassGeom g = ffc_conv_class_geom(cls, p.k, p.stride, p.pad, p.transposed);
    const int T = g.Ta * g.Tb, KK = p.k * p.k;
    const int nchunks = T * (p.cps[0] + (p.nseg > 1 ? p.cps[1] : 0));
    // element e of the class: (tile, chunk, half, n, kk) with n-major inside (n/8, kk/4, n%8, kk%4) order
    long long tile_base = 0;
    for (int t = 0; t < p.ntiles; ++t) {
        const int NT = v5_tile_n(p.cout, p.nt_full, t);
        const long long per_chunk = 2LL * NT * V5_BK;
        const long long total = (long long)nchunks * NT * V5_BK;
        // 32-bit index arithmetic (total < 2^31 for every weight this library sees): the three 64-bit divisions per element
        // were most of this kernel's instructions
        for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < (unsigned)total; e += gridDim.x * blockDim.x) {
            const int kk = (int)(e % (unsigned)V5_BK);
            const unsigned q = e / (unsigned)V5_BK;
            const int n = (int)(q % (unsigned)NT);
            int chunk = (int)(q / (unsigned)NT);
            const int chunk_all = chunk;
            int sg = 0;
            if (chunk >= T * p.cps[0]) { sg = 1; chunk -= T * p.cps[0]; }
            const int cps = sg ? p.cps[1] : p.cps[0], cin = sg ? p.cin[1] : p.cin[0];
            const int tap = chunk / cps, ci = (chunk % cps) * V5_BK + kk;
            const int co = t * p.nt_full + n;
            float v = 0.f;
            if (co < p.cout && ci < cin) {
                const int a = tap / g.Tb, b = tap % g.Tb;
                // weight tensor that holds (segment, co) and its own channel count / local channel index
                const float* w = sg ? p.w[1] : p.w[0];
                int wc = p.cout0, cl = co;
                if (co >= p.cout0) { w = sg ? nullptr : p.w0b; wc = p.cout - p.cout0; cl = co - p.cout0; }
                if (w) {
                    if (p.transposed) v = __ldg(w + ((size_t)ci * wc + cl) * KK + (g.ky0 + p.stride * a) * p.k + (g.kx0 + p.stride * b));
                    else v = __ldg(w + ((size_t)cl * cin + ci) * KK + a * p.k + b);
                }
            }
            const float hi = ffc_tf32_hi(v);
            float* dst = p.wp + p.cls_off[cls] + tile_base + (long long)chunk_all * per_chunk
                       + (n / 8) * 256 + (kk / 4) * 32 + (n % 8) * 4 + (kk % 4);
            dst[0] = hi;
            dst[(size_t)NT * V5_BK] = v - hi;
        }
        tile_base += (long long)nchunks * per_chunk;
    }
    // zero fill of the split-K outputs (one graph node less than a separate fill in front of the convolution)
    const long long nthreads = (long long)gridDim.x * gridDim.y * blockDim.x, me = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
#pragma unroll
    for (int z = 0; z < 2; ++z) {
        float* q = p.zero[z];
        const long long n = p.nzero[z];
        if (!q || n <= 0) continue;
        if ((((uintptr_t)q) & 15) == 0) {
            float4* q4 = reinterpret_cast<float4*>(q);
            for (long long i = me; i < n / 4; i += nthreads) q4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (long long i = (n / 4) * 4 + me; i < n; i += nthreads) q[i] = 0.f;
        } else {
            for (long long i = me; i < n; i += nthreads) q[i] = 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
template <int V5_GW>
__global__ void __launch_bounds__(V5_GW * 128 + 64, V5_GW == 2 ? 2 : 1) conv_v5_kernel(const ConvV5Params p) {
    constexpr int V5_MMAWARP = V5_GW * 4;
    extern __shared__ __align__(128) unsigned char v5_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int s = p.transposed ? p.stride : 1;
    const int ncls = s * s;
    const int cls = blockIdx.z % ncls, ksp = blockIdx.z / ncls, py = cls / s, px = cls % s;
    const int Hc = (p.Ho - py + s - 1) / s, Wc = (p.Wo - px + s - 1) / s;
    const int Mc = p.B * Hc * Wc;
    const int m0 = blockIdx.x * 128;
    if (m0 >= Mc) return;                               // whole CTA exits together
    const int ntile = blockIdx.y;
    const int NT = v5_tile_n(p.cout, p.nt_full, ntile);
    const ConvClassGeom g = ffc_conv_class_geom(cls, p.k, p.stride, p.pad, p.transposed);
    const int T = g.Ta * g.Tb;
    const int nchunks_all = T * (p.cps[0] + (p.nseg > 1 ? p.cps[1] : 0));
    // this CTA's share of the K chunks: [cb, cb + nchunks)
    const int per_split = (nchunks_all + p.ksplit - 1) / p.ksplit;
    const int cb = ksp * per_split;
    const int nchunks = (cb + per_split <= nchunks_all) ? per_split : (nchunks_all > cb ? nchunks_all - cb : 0);
    if (p.ksplit > 1 && nchunks == 0) return;           // nothing to add (whole CTA)
    const uint32_t stage_bytes = (uint32_t)(2 * NT * V5_BK * 4);
    // packed weights of this (class, tile): tiles before it have nt_full columns
    const float* wp = p.wp + p.cls_off[cls] + (long long)ntile * nchunks_all * 2LL * p.nt_full * V5_BK
                      + (long long)cb * 2LL * NT * V5_BK;

    const int nsb = p.nsb;
    unsigned char* bstage = v5_smem;                                                     // nsb stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(v5_smem + (size_t)nsb * stage_bytes);
    uint64_t* b_full = bars;                    // [V5_SB_MAX]
    uint64_t* b_free = bars + V5_SB_MAX;        // [V5_SB_MAX]
    uint64_t* a_ready = bars + 2 * V5_SB_MAX;   // [V5_GW]
    uint64_t* a_free = a_ready + V5_GW;         // [V5_GW]
    uint64_t* acc_done = a_free + V5_GW;        // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

    // TMEM columns: D_hi [0, NT) | D_lo [NT, 2NT) | A stage st at a_col0 + 64*st (hi 32 | lo 32), one stage per
    // gather warpgroup
    const uint32_t tmem_cols = (2 * NT + 64 * V5_GW <= 256) ? 256u : 512u;
    constexpr int nstage = V5_GW;
    if (tid == 0) {
        for (int i = 0; i < nsb; ++i) { umma::mbar_init(&b_full[i], 1); umma::mbar_init(&b_free[i], 1); }
        for (int i = 0; i < V5_GW; ++i) { umma::mbar_init(&a_ready[i], 4); umma::mbar_init(&a_free[i], 1); }
        umma::mbar_init(acc_done, 1);
        umma::fence_barrier_init();
    }
    if (warp == V5_MMAWARP) umma::tmem_alloc(tmem_slot, tmem_cols);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t a_col0 = tmem_cols - 64u * V5_GW;

    if (warp < V5_MMAWARP) {
        // ===================== A producers (then epilogue) =====================
        const int wg = warp >> 2;
        const int row = tid & 127;                          // TMEM lane == pixel of the tile
        const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        const int m = m0 + row;
        const bool ok = m < Mc;
        const int xq = m % Wc, yq = (m / Wc) % Hc, b = m / (Wc * Hc);
        const int HWi = p.Hi * p.Wi;
        for (int c = wg; c < nchunks; c += nstage) {
            // cursor of chunk cb + c: (segment, tap, first channel)
            int r = cb + c, sg = 0;
            if (r >= T * p.cps[0]) { sg = 1; r -= T * p.cps[0]; }
            const int cps = sg ? p.cps[1] : p.cps[0], cin = sg ? p.cin[1] : p.cin[0];
            const int tap = r / cps, c0 = (r % cps) * V5_BK;
            const int ta = tap / g.Tb, tb = tap % g.Tb;
            int iy, ix;
            if (p.transposed) { iy = yq + g.qy - ta; ix = xq + g.qx - tb; }
            else { iy = yq * p.stride - p.pad + ta; ix = xq * p.stride - p.pad + tb; }
            const bool okp = ok && iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi;
            const float* xs = sg ? p.x[1] : p.x[0];
            const float* xp = xs + ((size_t)(b * cin + c0) * HWi + iy * p.Wi + ix);
            float v[V5_BK];
#pragma unroll
            for (int j = 0; j < V5_BK; ++j) v[j] = (okp && c0 + j < cin) ? __ldg(xp + (size_t)j * HWi) : 0.f;
            const int it = c / nstage;
            if (it > 0) umma::mbar_wait(&a_free[wg], (uint32_t)((it - 1) & 1));      // the MMAs that read this stage are done
            umma::fence_after_sync();
            const uint32_t acol = lane_base + a_col0 + 64u * (uint32_t)wg;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float xv = v[16 * h + j];
                    const float xh = ffc_tf32_hi(xv);
                    hi[j] = __float_as_uint(xh);
                    lo[j] = __float_as_uint(xv - xh);
                }
                umma::tmem_st16(acol + 16 * h, hi);
                umma::tmem_st16(acol + 32 + 16 * h, lo);
            }
            umma::wait_st();
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&a_ready[wg]);       // one arrival per warp: 128 arrivals on one mbarrier serialise
        }
        // ===================== epilogue =====================
        if (nchunks > 0) { umma::mbar_wait(acc_done, 0); umma::fence_after_sync(); }
        {
            const int oy = yq * s + py, ox = xq * s + px;
            const size_t HWo = (size_t)p.Ho * p.Wo;
            const int co0 = ntile * p.nt_full;
            const size_t pix = (size_t)oy * p.Wo + ox;
            for (int n0 = 16 * wg; n0 < NT; n0 += 16 * V5_GW) {
                uint32_t r[16];
                if (nchunks > 0) {                            // warp-wide: every lane takes part
                    uint32_t q[16];
                    umma::tmem_ld16(lane_base + (uint32_t)n0, r);
                    umma::tmem_ld16(lane_base + (uint32_t)(NT + n0), q);
                    umma::wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(q[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = 0u;
                }
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int co = co0 + n0 + j;
                        if (co < p.cout) {
                            float v2 = __uint_as_float(r[j]);
                            if (p.bias && ksp == 0) v2 += __ldg(p.bias + co);
                            float* dst;
                            if (co < p.cout0) {
                                const size_t o = ((size_t)b * p.cout0 + co) * HWo + pix;
                                if (p.addend && ksp == 0) v2 += __ldg(p.addend + o);
                                dst = p.y + o;
                            } else {
                                dst = p.y1 + ((size_t)b * (p.cout - p.cout0) + (co - p.cout0)) * HWo + pix;
                            }
                            if (p.ksplit > 1) atomicAdd(dst, v2);                 // slope == 1 in this mode (host)
                            else *dst = v2 > 0.f ? v2 : v2 * p.slope;
                        }
                    }
                }
            }
        }
    } else if (warp == V5_MMAWARP) {
        // ===================== MMA issuer =====================
        // The whole warp walks the loop converged (uniform values stay in uniform registers) and one elected lane issues:
        // inside an `if (lane == 0)` region every tcgen05 instruction was wrapped in its own elect / branch sequence and
        // the issue rate, not the tensor pipe, set the pace (~150 cycles per MMA against a 64-96 cycle floor, ncu r01n).
        {
            const uint32_t idesc = umma::idesc_tf32(128, NT);
            const uint32_t b0 = umma::smem_u32(bstage);
            int sb = 0; uint32_t sb_phase = 0;
            for (int c = 0; c < nchunks; ++c) {
                const int sa = c % nstage;
                umma::mbar_wait(&b_full[sb], sb_phase);
                umma::mbar_wait(&a_ready[sa], (uint32_t)((c / nstage) & 1));
                umma::fence_after_sync();
                if (umma::elect_one()) {
                    const uint32_t b_hi = b0 + (uint32_t)sb * stage_bytes, b_lo = b_hi + (uint32_t)(NT * V5_BK * 4);
                    const uint32_t a_hi = tbase + a_col0 + 64u * (uint32_t)sa, a_lo = a_hi + 32u;
                    if (2 * NT <= 256) {
                        // hi and lo halves of the weights are adjacent rows of the stage: ONE MMA of width 2*NT computes
                        // a_hi * b_hi into D_hi and a_hi * b_lo into D_lo, a second one of width NT adds a_lo * b_hi to D_lo --
                        // 8 MMAs per chunk instead of 12 (the issue / operand-fetch latency of a K = 8 MMA, not the tensor
                        // pipe, sets the pace on these narrow tiles)
                        const uint32_t idesc2 = umma::idesc_tf32(128, 2 * NT);
#pragma unroll
                        for (int ks = 0; ks < V5_BK / 8; ++ks) {
                            const uint64_t dh = umma::smem_desc_kmajor_noswizzle(b_hi + ks * 256, 128, 1024);
                            umma::mma_tf32_ts(tbase, a_hi + ks * 8, dh, idesc2, (c | ks) ? 1u : 0u);
                            umma::mma_tf32_ts(tbase + (uint32_t)NT, a_lo + ks * 8, dh, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int ks = 0; ks < V5_BK / 8; ++ks) {        // cross terms first, then the main products: two
                            const uint64_t dh = umma::smem_desc_kmajor_noswizzle(b_hi + ks * 256, 128, 1024);   // accumulator switches per chunk
                            const uint64_t dl = umma::smem_desc_kmajor_noswizzle(b_lo + ks * 256, 128, 1024);
                            umma::mma_tf32_ts(tbase + (uint32_t)NT, a_lo + ks * 8, dh, idesc, (c | ks) ? 1u : 0u);
                            umma::mma_tf32_ts(tbase + (uint32_t)NT, a_hi + ks * 8, dl, idesc, 1u);
                        }
#pragma unroll
                        for (int ks = 0; ks < V5_BK / 8; ++ks) {
                            const uint64_t dh = umma::smem_desc_kmajor_noswizzle(b_hi + ks * 256, 128, 1024);
                            umma::mma_tf32_ts(tbase, a_hi + ks * 8, dh, idesc, (c | ks) ? 1u : 0u);
                        }
                    }
                    umma::commit(&a_free[sa]);
                    umma::commit(&b_free[sb]);
                }
                __syncwarp();
                if (++sb == nsb) { sb = 0; sb_phase ^= 1u; }
            }
            if (umma::elect_one()) umma::commit(acc_done);
            __syncwarp();
        }
    } else {
        // ===================== B loader =====================
        if (lane == 0) {
            int sb = 0; uint32_t free_phase = 1;               // first pass over the ring: nothing to wait for
            for (int c = 0; c < nchunks; ++c) {
                if (c >= nsb) umma::mbar_wait(&b_free[sb], free_phase);
                umma::mbar_arrive_expect_tx(&b_full[sb], stage_bytes);
                umma::bulk_g2s(bstage + (size_t)sb * stage_bytes, wp + (size_t)c * 2 * NT * V5_BK, stage_bytes, &b_full[sb]);
                if (++sb == nsb) { sb = 0; free_phase ^= 1u; }
            }
        }
        __syncwarp();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == V5_MMAWARP) umma::tmem_dealloc(tbase, tmem_cols);
}

// v -> v > 0 ? v : slope * v in place over one or two tensors (the activation of a split-K convolution, applied after the sum)
__global__ void __launch_bounds__(256) leaky_inplace_kernel(float* __restrict__ a, long long na, float* __restrict__ b, long long nb, float slope) {
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int which = 0; which < 2; ++which) {
        float* p = which ? b : a;
        const long long n = which ? nb : na;
        if (!p || n == 0) continue;
        if ((((uintptr_t)p) & 15) == 0) {
            const long long n4 = n / 4;
            for (long long i = t0; i < n4; i += stride) {
                float4 v = reinterpret_cast<float4*>(p)[i];
                v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
                v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
                reinterpret_cast<float4*>(p)[i] = v;
            }
            for (long long i = 4 * n4 + t0; i < n; i += stride) { const float v = p[i]; p[i] = v > 0.f ? v : v * slope; }
        } else {
            for (long long i = t0; i < n; i += stride) { const float v = p[i]; p[i] = v > 0.f ? v : v * slope; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct ConvV5Plan { int nt_full, ntiles, cps[2], ncls; long long cls_off[FFC_V5_MAXCLS]; long long total_floats; };

static ConvV5Plan conv_v5_plan(int cin0, int cin1, int cout, int k, int stride, int pad, int transposed, int nt_max = 192) {
    ConvV5Plan pl;
    const int c16 = (cout + 15) / 16 * 16;
    // widest tile: 2 accumulators of N columns + 128 columns of A stages in the 512 TMEM columns; an even split
    // (e.g. 384 -> 192 + 192, 512 -> 3 x 176) keeps the tiles alike
    const int nsplit = (c16 + nt_max - 1) / nt_max;
    pl.nt_full = ((c16 + nsplit - 1) / nsplit + 15) / 16 * 16;
    pl.ntiles = (cout + pl.nt_full - 1) / pl.nt_full;
    pl.cps[0] = (cin0 + V5_BK - 1) / V5_BK;
    pl.cps[1] = cin1 ? (cin1 + V5_BK - 1) / V5_BK : 0;
    pl.ncls = transposed ? stride * stride : 1;
    long long off = 0, cols = 0;
    for (int t = 0; t < pl.ntiles; ++t) cols += v5_tile_n(cout, pl.nt_full, t);
    for (int c = 0; c < FFC_V5_MAXCLS; ++c) {
        pl.cls_off[c] = off;
        if (c < pl.ncls) {
            const ConvClassGeom g = ffc_conv_class_geom(c, k, stride, pad, transposed);
            off += (long long)g.Ta * g.Tb * (pl.cps[0] + pl.cps[1]) * 2LL * cols * V5_BK;
        }
    }
    pl.total_floats = off;
    return pl;
}

size_t conv_v5_workspace_bytes(int cin0, int cin1, int cout, int k, int stride, int pad, int transposed) {
    const long long a = conv_v5_plan(cin0, cin1, cout, k, stride, pad, transposed).total_floats;
    const long long b = conv_v5_plan(cin0, cin1, cout, k, stride, pad, transposed, 128).total_floats;     // the small-grid tiling (conv_v5_run_block)
    return (size_t)(a > b ? a : b) * sizeof(float) + 256;
}

// arguments validated by ffc_conv2d_fwd_ws
int conv_v5_run_block(const float* x0, const float* w0, const float* w0b, int cin0, const float* x1, const float* w1, int cin1,
                      const float* bias, const float* addend, float* y, float* y1, int cout0, int B, int cout, int Hi, int Wi, int Ho, int Wo,
                      int k, int stride, int pad, int transposed, float slope, void* workspace, size_t workspace_bytes, ffc_stream_t st);

int conv_v5_run(const float* x0, const float* w0, int cin0, const float* x1, const float* w1, int cin1,
                const float* bias, const float* addend, float* y, int B, int cout, int Hi, int Wi, int Ho, int Wo,
                int k, int stride, int pad, int transposed, void* workspace, size_t workspace_bytes, ffc_stream_t st) {
    return conv_v5_run_block(x0, w0, nullptr, cin0, x1, w1, cin1, bias, addend, y, nullptr, cout, B, cout, Hi, Wi, Ho, Wo,
                             k, stride, pad, transposed, 1.f, workspace, workspace_bytes, st);
}

// Block form: y (cout0 channels) = conv(x0, w0) + conv(x1, w1) [+ bias[:cout0]] [+ addend];  y1 (cout - cout0 channels) =
// conv(x0, w0b) [+ bias[cout0:]] -- ONE implicit GEMM over the concatenated output channels, so the operand gathered from x0
// is shared by both outputs (the convl2l | convl2g pair of FFC.forward, ffc.py:91-96).  y1 == null: the plain form.
int conv_v5_run_block(const float* x0, const float* w0, const float* w0b, int cin0, const float* x1, const float* w1, int cin1,
                      const float* bias, const float* addend, float* y, float* y1, int cout0, int B, int cout, int Hi, int Wi, int Ho, int Wo,
                      int k, int stride, int pad, int transposed, float slope, void* workspace, size_t workspace_bytes, ffc_stream_t st) {
    ConvV5Plan pl = conv_v5_plan(cin0, cin1, cout, k, stride, pad, transposed);
    {
        // Wide outputs on a small grid (deep layers on 4x4 / 8x8 planes: 512 output channels as 3 tiles of 176 on 32 pixel tiles =
        // 96 CTAs): when the 128-wide tiling still fits one wave it wins -- every CTA walks the same K loop with a narrower MMA
        // (896 instead of 1248 tensor cycles per chunk), and the extra A gathers of a fourth column tile are L2 hits.
        const int s_ = transposed ? stride : 1;
        const int mt = ffc_cdiv(B * ffc_cdiv(Ho, s_) * ffc_cdiv(Wo, s_), 128) * s_ * s_;
        if (cout > 128 && mt * pl.ntiles < ffc_sm_count()) {
            const ConvV5Plan p128 = conv_v5_plan(cin0, cin1, cout, k, stride, pad, transposed, 128);
            if (p128.ntiles > pl.ntiles && mt * p128.ntiles <= ffc_sm_count()) pl = p128;
        }
    }
    const uintptr_t wsa = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    if (!workspace || wsa + (size_t)pl.total_floats * sizeof(float) > (uintptr_t)workspace + workspace_bytes) {
        ffc_set_error("ffc_conv2d_fwd_ws: workspace too small (%zu bytes needed)", (size_t)pl.total_floats * sizeof(float) + 256);
        return FFC_ERR_WORKSPACE;
    }
    cudaError_t e = cudaSuccess;
    ConvV5Params p;
    p.x[0] = x0; p.x[1] = x1; p.cin[0] = cin0; p.cin[1] = cin1; p.cps[0] = pl.cps[0]; p.cps[1] = pl.cps[1]; p.nseg = x1 ? 2 : 1;
    p.wp = (const float*)wsa; p.nt_full = pl.nt_full;
    for (int c = 0; c < FFC_V5_MAXCLS; ++c) p.cls_off[c] = pl.cls_off[c];
    p.bias = bias; p.addend = addend; p.y = y; p.y1 = y1; p.cout0 = cout0; p.slope = slope; p.B = B; p.cout = cout; p.Hi = Hi; p.Wi = Wi; p.Ho = Ho; p.Wo = Wo;
    p.k = k; p.stride = stride; p.pad = pad; p.transposed = transposed;
    const int s = transposed ? stride : 1;
    const int Mc = B * ffc_cdiv(Ho, s) * ffc_cdiv(Wo, s);
    // B ring: as deep as shared memory allows (the weight stream of a CTA is 2x its activation stream and arrives by
    // bulk copies of one stage each; with 3 stages the MMA warp waited on b_full half of the time, ncu r01n).  The
    // narrow-tile kernel keeps 2 CTAs per SM, and ~60 KB stay with the L1 for the gather loads.
    const bool four_wg = pl.nt_full > 64 && pl.nt_full <= 128;
    const size_t stage_b = (size_t)2 * pl.nt_full * V5_BK * 4;
    const size_t budget = (!four_wg && pl.nt_full <= 64) ? 96 * 1024 : 196 * 1024;
    int nsb = (int)(budget / stage_b);
    if (nsb > V5_SB_MAX) nsb = V5_SB_MAX;
    if (nsb < 2) nsb = 2;
    p.nsb = nsb;
    const size_t smem = (size_t)nsb * stage_b + 256;
    static FfcPerDevice configured_dev = {};
    size_t& configured = *ffc_device_slot(configured_dev);
    if (smem > configured) {
        e = cudaFuncSetAttribute(conv_v5_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_v5_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ffc_set_error("cudaFuncSetAttribute(conv_v5, %zu B): %s", smem, cudaGetErrorString(e)); return FFC_ERR_CUDA; }
        configured = smem;
    }
    // small grids (deep layers on small planes, small batch shards): split K over CTAs until the SMs are covered; the
    // partial sums meet in a zeroed output through float atomics.  Not with a fused activation.
    int ksplit = 1;
    {
        const int ctas = ffc_cdiv(Mc, 128) * pl.ntiles * s * s;
        int min_chunks = 1 << 30;
        for (int c = 0; c < pl.ncls; ++c) {
            const ConvClassGeom g = ffc_conv_class_geom(c, k, stride, pad, transposed);
            const int n = g.Ta * g.Tb * (pl.cps[0] + pl.cps[1]);
            if (n < min_chunks) min_chunks = n;
        }
        // resident CTA slots: the two-warpgroup kernel on narrow tiles keeps two CTAs per SM.  A fused activation no longer blocks
        // the split: the kernel then stores plain sums (slope 1) and one in-place pass applies the activation afterwards -- the
        // deep discriminator layers (4x4 / 8x8 planes, 64-96 CTAs with 72-256 chunks each) ran at 0.3-0.6 of a wave.
        const int slots = ffc_sm_count() * ((!four_wg && pl.nt_full <= 64) ? 2 : 1);
        if (ctas * 2 <= slots) {
            ksplit = slots / ctas;
            if (ksplit > min_chunks / 8) ksplit = min_chunks / 8;
            if (ksplit > 16) ksplit = 16;
            if (ksplit < 1) ksplit = 1;
        }
    }
    p.ksplit = ksplit;
    const bool act_after = ksplit > 1 && slope != 1.f;
    if (act_after) p.slope = 1.f;
    PackV5Params pp;
    pp.w[0] = w0; pp.w[1] = w1; pp.cin[0] = cin0; pp.cin[1] = cin1; pp.cps[0] = pl.cps[0]; pp.cps[1] = pl.cps[1];
    pp.nseg = x1 ? 2 : 1; pp.wp = (float*)wsa; pp.nt_full = pl.nt_full; pp.ntiles = pl.ntiles; pp.cout = cout;
    pp.k = k; pp.stride = stride; pp.pad = pad; pp.transposed = transposed; pp.w0b = w0b; pp.cout0 = cout0;
    for (int c = 0; c < FFC_V5_MAXCLS; ++c) pp.cls_off[c] = pl.cls_off[c];
    pp.zero[0] = pp.zero[1] = nullptr; pp.nzero[0] = pp.nzero[1] = 0;
    if (ksplit > 1) {                 // the partial sums of the K split meet in a zeroed output: the packing launch zero-fills it
        const long long HWo = (long long)Ho * Wo;
        pp.zero[0] = y; pp.nzero[0] = (long long)B * cout0 * HWo;
        if (y1) { pp.zero[1] = y1; pp.nzero[1] = (long long)B * (cout - cout0) * HWo; }
    }
    const long long per_cls = (long long)k * k * (pl.cps[0] + pl.cps[1]) * V5_BK * ((cout + 15) / 16 * 16);
    if (per_cls >= (1LL << 31)) { ffc_set_error("ffc_conv2d_fwd_ws: weight too large for the packing kernel (%lld elements per class)", per_cls); return FFC_ERR_BAD_ARG; }
    int gx = (int)((per_cls + 255) / 256); if (gx > ffc_sm_count() * 4) gx = ffc_sm_count() * 4; if (gx < 1) gx = 1;
    pack_v5_kernel<<<dim3(gx, pl.ncls), 256, 0, st>>>(pp);
    e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("pack_v5 launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();

    const dim3 grid(ffc_cdiv(Mc, 128), pl.ntiles, s * s * ksplit);
    if (four_wg) conv_v5_kernel<4><<<grid, 4 * 128 + 64, smem, st>>>(p);
    else conv_v5_kernel<2><<<grid, 2 * 128 + 64, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("conv_v5 launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    if (act_after) {
        const size_t HWo = (size_t)Ho * Wo;
        const long long n0 = (long long)B * cout0 * HWo, n1 = y1 ? (long long)B * (cout - cout0) * HWo : 0;
        const long long most = n0 > n1 ? n0 : n1;
        int gx = (int)((most / 4 + 255) / 256); if (gx > ffc_sm_count() * 8) gx = ffc_sm_count() * 8; if (gx < 1) gx = 1;
        leaky_inplace_kernel<<<gx, 256, 0, st>>>(y, n0, y1, n1, slope);
        e = cudaGetLastError();
        if (e != cudaSuccess) { ffc_set_error("conv_v5 activation pass launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
        ffc_count_launch();
    }
    return FFC_OK;
}
#endif  // !FFC_EMU
