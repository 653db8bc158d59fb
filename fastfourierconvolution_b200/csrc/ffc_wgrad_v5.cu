// ConvWgradV5: convolution weight gradient on the 5th-generation tensor cores (tcgen05 / UMMA) at FP32 accuracy.
// Same contract as ffc_conv2d_wgrad:
//   dW[sc][lc][ky][kx] = sum_{b,y,x} S[b,sc,y,x] * L[b,lc,y*stride-pad+ky,x*stride-pad+kx]
// GEMM with the reduction over the pixels of the small-side grid:
//   D[m = (lc,ky,kx), 128 per tile][n = sc, <= 192 per tile] += A[m][k = pixel] * B[n][k]^T,  K split over CTAs (atomics)
//   * A = L gathered with the tap shift: thread = GEMM row (one TMEM lane) walks the 32 pixels of the K chunk, bounds
//     checks the shifted position, splits hi/lo and writes its row into TMEM (tcgen05.st); four gather warpgroups rotate
//     over four TMEM stages so that the global-load latency of one chunk hides behind the other three.
//   * B = S, contiguous along the pixels: a third warpgroup loads float4, splits hi/lo and stores the K-major no-swizzle
//     UMMA tile image into shared memory (generic stores + fence.proxy.async), three stages.
//   * one thread issues the MMAs: D_hi += A_hi B_hi, D_lo += A_lo B_hi + A_hi B_lo (3xTF32, see ffc_conv_v5.cu);
//   * epilogue: D_hi + D_lo from TMEM, one float atomic per weight entry (dW is zeroed first; K is split over gridDim.z).
// Device build only.  Requires Hs*Ws % 4 == 0 (float4 groups of pixels never straddle an image).
#include "ffc_common.cuh"

#ifndef FFC_EMU
#include "ffc_umma.cuh"

static constexpr int WG5_BK = 32;
static constexpr int WG5_SB = 4;
static constexpr int WG5_GW = 3;              // gather warpgroups == A stages in TMEM
static constexpr int WG5_BW = 3;              // B-builder warpgroups, alternating over the chunks
static constexpr int WG5_THREADS = WG5_GW * 128 + WG5_BW * 128 + 32;       // gather warps, B-builder warps, 1 MMA warp
static constexpr int WG5_BUILD0 = WG5_GW * 128;                   // first B-builder thread
static constexpr int WG5_MMAWARP = WG5_GW * 4 + WG5_BW * 4;

struct WgradV5Params {
    const float* S; const float* L; float* dW;
    int B, SC, LC, Hs, Ws, Hl, Wl, k, stride, pad;
    int nt_full;          // columns (sc) of a full N tile, multiple of 16, <= 128
    int chunks_per_split; // K chunks of 32 pixels per CTA
};

__global__ void __launch_bounds__(WG5_THREADS, 1) wgrad_v5_kernel(const WgradV5Params p) {
    extern __shared__ __align__(128) unsigned char wg5_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KK = p.k * p.k, Mtot = p.LC * KK;
    const int HWs = p.Hs * p.Ws, Ktot = p.B * HWs;
    const int m0 = blockIdx.x * 128;
    const int n0 = blockIdx.y * p.nt_full;
    int NT = p.SC - n0; if (NT > p.nt_full) NT = p.nt_full; NT = (NT + 15) / 16 * 16;
    const int kbeg = blockIdx.z * p.chunks_per_split * WG5_BK;
    if (kbeg >= Ktot) return;
    int nchunks = (Ktot - kbeg + WG5_BK - 1) / WG5_BK;
    if (nchunks > p.chunks_per_split) nchunks = p.chunks_per_split;
    const uint32_t half_bytes = (uint32_t)(NT * WG5_BK * 4), stage_bytes = 2 * half_bytes;

    unsigned char* bstage = wg5_smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wg5_smem + (size_t)WG5_SB * stage_bytes);
    uint64_t* b_full = bars;                    // [SB] count 4 (one arrival per builder warp)
    uint64_t* b_free = bars + WG5_SB;           // [SB]
    uint64_t* a_ready = bars + 2 * WG5_SB;      // [GW] count 4 (one arrival per gather warp)
    uint64_t* a_free = a_ready + WG5_GW;        // [GW]
    uint64_t* acc_done = a_free + WG5_GW;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);
    const uint32_t tmem_cols = 512u;
    if (tid == 0) {
        for (int i = 0; i < WG5_SB; ++i) { umma::mbar_init(&b_full[i], 4); umma::mbar_init(&b_free[i], 1); }
        for (int i = 0; i < WG5_GW; ++i) { umma::mbar_init(&a_ready[i], 4); umma::mbar_init(&a_free[i], 1); }
        umma::mbar_init(acc_done, 1);
        umma::fence_barrier_init();
    }
    if (warp == WG5_MMAWARP) umma::tmem_alloc(tmem_slot, tmem_cols);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t a_col0 = 256u;      // D_hi [0,NT) | D_lo [NT,2NT) | A stage st at 256 + 64*st (hi 32 | lo 32)

    if (warp < WG5_GW * 4) {
        // ===================== A producers: gathered L rows =====================
        const int wg = warp >> 2, row = tid & 127;
        const uint32_t lane_base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        const int mg = m0 + row;
        const bool okm = mg < Mtot;
        const int lc = okm ? mg / KK : 0, t = okm ? mg % KK : 0;
        const int ky = t / p.k, kx = t % p.k;
        const int HWl = p.Hl * p.Wl;
        const float* Lp = p.L + (size_t)lc * HWl + (kx - p.pad);        // + image, + row, + x*stride
        const size_t imgL = (size_t)p.LC * HWl;
        // pixels of a chunk are walked in runs of SEG consecutive x of one row (SEG = min(Ws, 32) when that divides 32
        // and Ws; otherwise single pixels): one row check / row pointer per run, one unsigned compare per pixel
        const int SEG = (p.Ws >= 32 && p.Ws % 32 == 0) ? 32 : ((p.Ws == 16 || p.Ws == 8 || p.Ws == 4) ? p.Ws : 1);
        for (int c = wg; c < nchunks; c += WG5_GW) {
            const int kk0 = kbeg + c * WG5_BK;
            float v[WG5_BK];
#define WG5_GATHER(SEGC)                                                                                          \
            {                                                                                                    \
                _Pragma("unroll")                                                                                \
                for (int sg = 0; sg < WG5_BK / SEGC; ++sg) {                                                     \
                    const int kk = kk0 + sg * SEGC;                                                              \
                    const int b = kk / HWs, r = kk - b * HWs;                                                    \
                    const int y = r / p.Ws, x0 = r - y * p.Ws;                                                   \
                    const int ly = y * p.stride - p.pad + ky;                                                    \
                    const bool okr = okm && b < p.B && ly >= 0 && ly < p.Hl;                                     \
                    const float* rp = Lp + (size_t)b * imgL + ly * p.Wl + x0 * p.stride;                         \
                    const int lx0 = x0 * p.stride - p.pad + kx;                                                  \
                    _Pragma("unroll")                                                                            \
                    for (int j = 0; j < SEGC; ++j)                                                               \
                        v[sg * SEGC + j] = (okr && (unsigned)(lx0 + j * p.stride) < (unsigned)p.Wl) ? __ldg(rp + j * p.stride) : 0.f; \
                }                                                                                                \
            }
            if (SEG == 32) WG5_GATHER(32) else if (SEG == 16) WG5_GATHER(16) else if (SEG == 8) WG5_GATHER(8)
            else if (SEG == 4) WG5_GATHER(4) else WG5_GATHER(1)
#undef WG5_GATHER
            const int it = c / WG5_GW;
            if (it > 0) umma::mbar_wait(&a_free[wg], (uint32_t)((it - 1) & 1));
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
            if (lane == 0) umma::mbar_arrive(&a_ready[wg]);       // one arrival per warp
        }
        // ===================== epilogue =====================
        umma::mbar_wait(acc_done, 0);
        umma::fence_after_sync();
        const size_t Ntot = (size_t)Mtot;               // dW row length = LC*KK
        for (int c0 = 16 * wg; c0 < NT; c0 += 16 * WG5_GW) {
            uint32_t r[16], q[16];
            umma::tmem_ld16(lane_base + (uint32_t)c0, r);
            umma::tmem_ld16(lane_base + (uint32_t)(NT + c0), q);
            umma::wait_ld();
            if (okm) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int sc = n0 + c0 + j;
                    if (sc < p.SC) atomicAdd(p.dW + (size_t)sc * Ntot + mg, __uint_as_float(r[j]) + __uint_as_float(q[j]));
                }
            }
        }
    } else if (warp < WG5_MMAWARP) {
        // ===================== B builders: S tile -> K-major no-swizzle UMMA image (hi | lo) =====================
        const int bt = (tid - WG5_BUILD0) & 127;        // 0..127 within the builder warpgroup
        const int bw = (tid - WG5_BUILD0) >> 7;         // builder warpgroup: chunks bw, bw + WG5_BW, ...
        const int n8 = bt & 7, qs = (bt >> 3) & 3, grp0 = bt >> 5;       // lane -> (row in 8-group, 16-byte piece), warp -> group
        const int ngroups = (NT / 8) * 2;
        constexpr int MAXG = (128 / 8) * 2 / 4;         // groups per builder warp at the widest tile (NT = 128)
        for (int c = bw; c < nchunks; c += WG5_BW) {
            const int sb = c % WG5_SB;
            const int kk0 = kbeg + c * WG5_BK;
            // all loads of the chunk are in flight before the first is used (and before the stage wait)
            float4 v[MAXG];
#pragma unroll
            for (int i = 0; i < MAXG; ++i) {
                const int g = grp0 + 4 * i;
                const int n = 8 * (g >> 1) + n8, q = 4 * (g & 1) + qs;      // row (sc - n0), 16-byte piece (4 pixels)
                const int kk = kk0 + 4 * q;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g < ngroups && n0 + n < p.SC && kk < Ktot) {
                    const int b = kk / HWs, r = kk - b * HWs;
                    v[i] = __ldg(reinterpret_cast<const float4*>(p.S + ((size_t)b * p.SC + n0 + n) * HWs + r));
                }
            }
            if (c >= WG5_SB) umma::mbar_wait(&b_free[sb], (uint32_t)((c / WG5_SB - 1) & 1));
            unsigned char* st = bstage + (size_t)sb * stage_bytes;
#pragma unroll
            for (int i = 0; i < MAXG; ++i) {
                const int g = grp0 + 4 * i;
                if (g < ngroups) {
                    const int n = 8 * (g >> 1) + n8, q = 4 * (g & 1) + qs;
                    float4 h;
                    h.x = ffc_tf32_hi(v[i].x); h.y = ffc_tf32_hi(v[i].y);
                    h.z = ffc_tf32_hi(v[i].z); h.w = ffc_tf32_hi(v[i].w);
                    const uint32_t off = (uint32_t)((n >> 3) * 1024 + q * 128 + (n & 7) * 16);
                    *reinterpret_cast<float4*>(st + off) = h;
                    *reinterpret_cast<float4*>(st + half_bytes + off) = make_float4(v[i].x - h.x, v[i].y - h.y, v[i].z - h.z, v[i].w - h.w);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&b_full[sb]);
        }
    } else {
        // ===================== MMA issuer =====================
        // whole warp converged, one elected lane issues (see ffc_conv_v5.cu: inside `if (lane == 0)` each tcgen05
        // instruction got its own elect / branch sequence and the issue rate set the pace)
        const uint32_t idesc = umma::idesc_tf32(128, NT);
        const uint32_t b0 = umma::smem_u32(bstage);
        for (int c = 0; c < nchunks; ++c) {
            const int sa = c % WG5_GW, sb = c % WG5_SB;
            umma::mbar_wait(&b_full[sb], (uint32_t)((c / WG5_SB) & 1));
            umma::mbar_wait(&a_ready[sa], (uint32_t)((c / WG5_GW) & 1));
            umma::fence_after_sync();
            if (umma::elect_one()) {
                const uint32_t b_hi = b0 + (uint32_t)sb * stage_bytes, b_lo = b_hi + half_bytes;
                const uint32_t a_hi = tbase + a_col0 + 64u * (uint32_t)sa, a_lo = a_hi + 32u;
                if (2 * NT <= 256) {
                    // the hi and lo halves of the B stage are adjacent rows: one MMA of width 2*NT gives a_hi * b_hi (D_hi) and
                    // a_hi * b_lo (D_lo), one of width NT adds a_lo * b_hi to D_lo: 8 MMAs per chunk instead of 12
                    const uint32_t idesc2 = umma::idesc_tf32(128, 2 * NT);
#pragma unroll
                    for (int ks = 0; ks < WG5_BK / 8; ++ks) {
                        const uint64_t dh = umma::smem_desc_kmajor_noswizzle(b_hi + ks * 256, 128, 1024);
                        umma::mma_tf32_ts(tbase, a_hi + ks * 8, dh, idesc2, (c | ks) ? 1u : 0u);
                        umma::mma_tf32_ts(tbase + (uint32_t)NT, a_lo + ks * 8, dh, idesc, 1u);
                    }
                } else {
#pragma unroll
                    for (int ks = 0; ks < WG5_BK / 8; ++ks) {          // cross terms first, then the main products: two
                        const uint64_t dh = umma::smem_desc_kmajor_noswizzle(b_hi + ks * 256, 128, 1024);   // accumulator switches per chunk
                        const uint64_t dl = umma::smem_desc_kmajor_noswizzle(b_lo + ks * 256, 128, 1024);
                        umma::mma_tf32_ts(tbase + (uint32_t)NT, a_lo + ks * 8, dh, idesc, (c | ks) ? 1u : 0u);
                        umma::mma_tf32_ts(tbase + (uint32_t)NT, a_hi + ks * 8, dl, idesc, 1u);
                    }
#pragma unroll
                    for (int ks = 0; ks < WG5_BK / 8; ++ks) {
                        const uint64_t dh = umma::smem_desc_kmajor_noswizzle(b_hi + ks * 256, 128, 1024);
                        umma::mma_tf32_ts(tbase, a_hi + ks * 8, dh, idesc, (c | ks) ? 1u : 0u);
                    }
                }
                umma::commit(&a_free[sa]);
                umma::commit(&b_free[sb]);
            }
            __syncwarp();
        }
        if (umma::elect_one()) umma::commit(acc_done);
        __syncwarp();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == WG5_MMAWARP) umma::tmem_dealloc(tbase, tmem_cols);
}

// dW must be zeroed by the caller (ffc_conv2d_wgrad does it).  Returns false when the shape is not handled.
bool wgrad_v5_supported(int SC, int LC, int Hs, int Ws, int k) {
    (void)LC; (void)k;
    return (Hs * Ws) % 4 == 0 && SC >= 5;
}

int wgrad_v5_run(const float* S, const float* L, float* dW, int B, int SC, int LC, int Hs, int Ws, int Hl, int Wl,
                 int k, int stride, int pad, ffc_stream_t st) {
    WgradV5Params p;
    p.S = S; p.L = L; p.dW = dW; p.B = B; p.SC = SC; p.LC = LC; p.Hs = Hs; p.Ws = Ws; p.Hl = Hl; p.Wl = Wl;
    p.k = k; p.stride = stride; p.pad = pad;
    const int c16 = (SC + 15) / 16 * 16, nt_max = 128;     // 2*128 accumulator + 4*64 A-stage columns = 512
    const int nsplit_n = (c16 + nt_max - 1) / nt_max;
    p.nt_full = ((c16 + nsplit_n - 1) / nsplit_n + 15) / 16 * 16;
    const int ntiles = (SC + p.nt_full - 1) / p.nt_full;
    const int mtiles = ffc_cdiv(LC * k * k, 128);
    const int Ktot = B * Hs * Ws, kchunks = ffc_cdiv(Ktot, WG5_BK);
    // Split K over CTAs.  One CTA is resident per SM (800 threads, the whole tensor memory), so the launch runs in whole waves
    // of sm_count CTAs and costs ~ waves * (chunks per CTA + a fixed prologue / epilogue of ~6 chunks): pick the split that
    // minimises that, at least 8 chunks per CTA.  (The former rule -- about two CTAs per SM, rounded up -- produced 2.05 to
    // 2.4 waves on the 128- and 256-channel discriminator layers, i.e. a third wave for a few CTAs: 114 -> 89 us and
    // 104 -> 87 us per launch when the grid fits whole waves, tools/bench_wgrad.py.)
    const int base = mtiles * ntiles, sms = ffc_sm_count();
    int max_split = kchunks / 8;
    if (max_split < 1) max_split = 1;
    if (max_split > 4 * sms) max_split = 4 * sms;
    int ksplit = 1;
    long long best = -1;
    for (int ks = 1; ks <= max_split; ++ks) {
        const int cps = ffc_cdiv(kchunks, ks), ks2 = ffc_cdiv(kchunks, cps);
        const long long waves = ffc_cdiv(base * ks2, sms);
        const long long cost = waves * (cps + 6);
        if (best < 0 || cost < best) { best = cost; ksplit = ks2; }
    }
    p.chunks_per_split = ffc_cdiv(kchunks, ksplit);
    ksplit = ffc_cdiv(kchunks, p.chunks_per_split);
    const size_t smem = (size_t)WG5_SB * 2 * p.nt_full * WG5_BK * 4 + 256;
    static FfcPerDevice configured_dev = {};
    size_t& configured = *ffc_device_slot(configured_dev);
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ffc_set_error("cudaFuncSetAttribute(wgrad_v5, %zu B): %s", smem, cudaGetErrorString(e)); return FFC_ERR_CUDA; }
        configured = smem;
    }
    wgrad_v5_kernel<<<dim3(mtiles, ntiles, ksplit), WG5_THREADS, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("wgrad_v5 launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}
#endif  // !FFC_EMU
