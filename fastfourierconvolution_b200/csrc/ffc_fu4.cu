// Fused FourierUnitSN forward, third generation, 32x32 planes with up to 8 channels (the fgan32 / fgan64 / cfg1 unit
// FourierUnitSN(8,8)@32x32, layers/ffc/fourier_unity.py:32-58): WARP-PRIVATE transforms.
//
//   CTA = one image, warp = one plane.  Both 2-D transforms of a plane run inside ONE warp -- lane = row for the row pass,
//   lane = (column v < 16, half h) for the column pass -- with the plane moving through a 4.6 KB warp-private shared-memory
//   tile and __syncwarp only:
//     rows    : lane r holds its row of 32 reals = 16 complex, in-register 16-point FFT + even/odd post-processing;
//     columns : 32 = 2 x 16 decimation in time: lane (v, h) runs the 16-point FFT of the rows of parity h of column v and
//               meets its partner lane (v, 1-h) through 16 complex shuffles; the Nyquist column (one real value per row,
//               still in a register after the row pass) is a 32-point FFT ACROSS the lanes, five shuffle butterflies;
//     so a lane ends up with 17 bins of its plane in registers (16 of column v + one Nyquist bin).
//   The only CTA-wide barriers are the two around the channel mix (the one data exchange between planes): every warp
//   publishes its 17 x 32 bins (lane-contiguous, conflict free), warp o accumulates output plane o for its own 17 bins
//   from all input planes with its 4 * Cin weights in registers (two packed FFMA2 per weight pair), and the result is
//   already in the register layout the inverse column pass needs.  BatchNorm statistics: 17 values per lane, a shuffle
//   reduction, four double atomics per warp; the mixed plane waits IN REGISTERS across the cooperative grid barrier.
//   ffc_fu2.cu's version of the same unit walks 16 __syncthreads phases with 544-thread CTAs (ncu: 9.8 barrier-stall
//   cycles per issued instruction); it stays as the fallback for other shapes and for the host emulation build.
#include "ffc_fu2.cuh"
#ifndef FFC_EMU
#include <cooperative_groups.h>

namespace fu4 {
constexpr int N = 32, M = 16, NW = 8, RSF = 36, PBF = N * RSF, BINS = N * (M + 1);   // tile row = 36 floats (9 float4: odd)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float2 shfl_xor2(float2 a, int m) {
    return make_float2(__shfl_xor_sync(FULL, a.x, m), __shfl_xor_sync(FULL, a.y, m));
}
__device__ __forceinline__ float2 sel2(bool c, float2 a, float2 b) { return make_float2(c ? a.x : b.x, c ? a.y : b.y); }

// 32-point FFT across the lanes of a warp, decimation in frequency: natural order in, lane l ends with frequency bitrev5(l).
// tws[s] = w_(32>>s)^(lane & ((16>>s)-1)) as (cos, sin).
__device__ __forceinline__ float2 lane_fft_fwd(float2 m, const float2* tws, int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int half = 16 >> s;
        const float2 p = shfl_xor2(m, half);
        const bool upper = (lane & half) != 0;
        const float sg = upper ? -1.0f : 1.0f;
        float2 d = ffc_fma2(m, make_float2(sg, sg), p);          // lower: m + p;  upper: p - m
        if (s < 4 && upper) d = ffc_cmul_tw<-1>(d, tws[s]);      // predicated, no divergence
        m = d;
    }
    return m;
}
// the unnormalised inverse: lane l holds frequency bitrev5(l), natural order out (decimation in time)
__device__ __forceinline__ float2 lane_fft_inv(float2 m, const float2* tws, int lane) {
#pragma unroll
    for (int s = 4; s >= 0; --s) {
        const int half = 16 >> s;
        const bool upper = (lane & half) != 0;
        float2 mm = m;
        if (s < 4 && upper) mm = ffc_cmul_tw<+1>(m, tws[s]);
        const float2 p = shfl_xor2(mm, half);
        const float sg = upper ? -1.0f : 1.0f;
        m = ffc_fma2(mm, make_float2(sg, sg), p);                // lower: mm + p;  upper: p - mm
    }
    return m;
}

// rfft2 (unnormalised, times 2) of one 32x32 plane inside one warp.  A lane ends with 17 bins: S[0..15] = bins
// (u = 16*(lane>>4) + k, v = lane & 15), S[16] = bin (u = bitrev5(lane), v = 16); they are published straight into the warp's
// tile pb as the exchange image of the plane -- float4 slots [k/2][lane] for the 16 column bins, float2 slots [16][lane] for
// the Nyquist bin -- so nothing of the plane stays in registers.
// ADJ: the adjoint of the c2r transform (interior bins count twice; used on the incoming gradient by the backward kernel):
// the same pass without the doubling of the DC and Nyquist bins, and without the overall factor 2.
template <bool ADJ>
__device__ __forceinline__ void fwd_plane(const float* __restrict__ src, float* pb, int lane, const float2* tws) {
    {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(s4 + j * 32 + lane);
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(pb + (j * 4 + (lane >> 3)) * RSF + (lane & 7) * 4) = v[j];
    }
    __syncwarp();
    float ny;
    {
        float2 z[M];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(pb + lane * RSF + 4 * q);
            z[2 * q] = make_float2(a.x, a.y);
            z[2 * q + 1] = make_float2(a.z, a.w);
        }
        ffc_fft_regs<M, -1>(z);
        // twice the bins (the 1/2 of the even/odd split is folded into the mix weights), in place over z
        ny = (ADJ ? 1.0f : 2.0f) * (z[0].x - z[0].y);
        z[0] = make_float2((ADJ ? 1.0f : 2.0f) * (z[0].x + z[0].y), 0.f);
#pragma unroll
        for (int k = 1; k <= M / 2; ++k) {
            const float2 a = z[k], b = make_float2(z[M - k].x, -z[M - k].y);
            const float2 e = ffc_cadd(a, b);
            const float2 dd = ffc_csub(a, b);
            const float2 o = make_float2(dd.y, -dd.x);
            const float2 w = fu2_twc<N>(k);
            const float2 t = make_float2(o.x * w.x + o.y * w.y, o.y * w.x - o.x * w.y);
            z[k] = make_float2(e.x + t.x, e.y + t.y);
            if (k != M - k) z[M - k] = make_float2(e.x - t.x, t.y - e.y);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(pb + lane * RSF + 4 * q) = make_float4(z[2 * q].x, z[2 * q].y, z[2 * q + 1].x, z[2 * q + 1].y);
    }
    __syncwarp();
    const int v = lane & 15, h = lane >> 4;
    const float2 sg2 = h ? make_float2(-1.f, -1.f) : make_float2(1.f, 1.f);
    const float2* pb2 = reinterpret_cast<const float2*>(pb);
    float2 a[M];
#pragma unroll
    for (int j = 0; j < M; ++j) a[j] = pb2[(2 * j + h) * (RSF / 2) + v];
    ffc_fft_regs<M, -1>(a);
    __syncwarp();            // every lane has its column: the tile is free for the exchange image
    float4* ex4 = reinterpret_cast<float4*>(pb);
#pragma unroll
    for (int k = 0; k < M; k += 2) {
        float2 m0 = a[k], m1 = a[k + 1];
        if (k != 0 && h) m0 = ffc_cmul_tw<-1>(m0, fu2_twc<N>(k));     // the odd half sends w^k * O[k] (predicated)
        if (h) m1 = ffc_cmul_tw<-1>(m1, fu2_twc<N>(k + 1));
        const float2 r0 = shfl_xor2(m0, 16), r1 = shfl_xor2(m1, 16);
        const float2 s0 = ffc_fma2(m0, sg2, r0), s1 = ffc_fma2(m1, sg2, r1);      // h = 0: E + w^k O;  h = 1: E - w^k O
        ex4[(k / 2) * 32 + lane] = make_float4(s0.x, s0.y, s1.x, s1.y);
    }
    reinterpret_cast<float2*>(pb)[M * 32 + lane] = lane_fft_fwd(make_float2(ny, 0.f), tws, lane);
}

// irfft2 (unnormalised, torch c2r semantics) of the plane whose bins Y[17] sit in the layout of fwd_plane -> dst (+ res).
// ADJ: twice the adjoint of the r2c transform (interior bins count half): the same pass with the DC and Nyquist bins doubled.
template <bool ADJ>
__device__ __forceinline__ void inv_plane(const float2* Y, float* pb, int lane, const float2* tws,
                                          const float* __restrict__ res, float* __restrict__ dst) {
    const int v = lane & 15, h = lane >> 4;
    const float2 sg2 = h ? make_float2(-1.f, -1.f) : make_float2(1.f, 1.f);
    float2* pb2 = reinterpret_cast<float2*>(pb);
    {
        float2 A[M];
#pragma unroll
        for (int k = 0; k < M; ++k) {
            const float2 r = shfl_xor2(Y[k], 16);
            float2 d = ffc_fma2(Y[k], sg2, r);                        // h = 0: X[k] + X[k+16];  h = 1: X[k] - X[k+16]
            if (k != 0 && h) d = ffc_cmul_tw<+1>(d, fu2_twc<N>(k));
            A[k] = d;
        }
        ffc_fft_regs<M, +1>(A);
#pragma unroll
        for (int j = 0; j < M; ++j) pb2[(2 * j + h) * (RSF / 2) + v] = A[j];
    }
    const float xM = lane_fft_inv(Y[M], tws, lane).x;      // Nyquist bin of row `lane` (imaginary part ignored by c2r)
    __syncwarp();
    float2 z[M];
    {
        float2 x[M];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(pb + lane * RSF + 4 * q);
            x[2 * q] = make_float2(a.x, a.y);
            x[2 * q + 1] = make_float2(a.z, a.w);
        }
        z[0] = make_float2((ADJ ? 2.0f : 1.0f) * (x[0].x + xM), (ADJ ? 2.0f : 1.0f) * (x[0].x - xM));
#pragma unroll
        for (int k = 1; k <= M / 2; ++k) {
            const float2 p = x[k], q = make_float2(x[M - k].x, -x[M - k].y);
            const float2 e = make_float2(p.x + q.x, p.y + q.y);
            const float2 dm = make_float2(p.x - q.x, p.y - q.y);
            const float2 w = fu2_twc<N>(k);
            const float2 d = make_float2(dm.x * w.x - dm.y * w.y, dm.x * w.y + dm.y * w.x);
            z[k] = make_float2(e.x - d.y, e.y + d.x);
            if (k != M - k) z[M - k] = make_float2(e.x + d.y, d.x - e.y);
        }
    }
    ffc_fft_regs<M, +1>(z);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(pb + lane * RSF + 4 * q) = make_float4(z[2 * q].x, z[2 * q].y, z[2 * q + 1].x, z[2 * q + 1].y);
    __syncwarp();
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* r4 = reinterpret_cast<const float4*>(res);
    float4 q[8];
    if (res) {
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = __ldg(r4 + j * 32 + lane);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4 o = *reinterpret_cast<const float4*>(pb + (j * 4 + (lane >> 3)) * RSF + (lane & 7) * 4);
        if (res) { o.x += q[j].x; o.y += q[j].y; o.z += q[j].z; o.w += q[j].w; }
        d4[j * 32 + lane] = o;
    }
    __syncwarp();
}

// Channel mix of the NW exchange images at `tiles` (one per warp, PBF floats apart) for this lane's 17 bins, in two halves so
// that only 2 x 9 accumulator pairs are live at a time.  wq = (W[2o][2c], W[2o+1][2c+1], W[2o+1][2c], W[2o][2c+1]) * scale.
// TRANSPOSED = false: Y_o = sum_c W[o][c] S_c with o = warp;  true: dS_c = sum_o W[o][c]^T dY_o with c = warp.
template <bool TRANSPOSED>
__device__ __forceinline__ void mix17(const float* tiles, const float4* wq_s, int warp, int lane, float2* Y) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float2 pa[9], pq[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) pa[i] = pq[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < NW; ++c) {
            const float4* xc4 = reinterpret_cast<const float4*>(tiles + c * PBF) + lane;
            const float4 q = TRANSPOSED ? wq_s[c * NW + warp] : wq_s[warp * NW + c];
            const float2 wa = TRANSPOSED ? make_float2(q.x, q.z) : make_float2(q.x, q.y);
            const float2 wb = TRANSPOSED ? make_float2(q.w, q.y) : make_float2(q.z, q.w);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 s = xc4[(4 * half + i) * 32];
                pa[2 * i] = ffc_fma2(wa, make_float2(s.x, s.y), pa[2 * i]);
                pq[2 * i] = ffc_fma2(wb, make_float2(s.x, s.y), pq[2 * i]);
                pa[2 * i + 1] = ffc_fma2(wa, make_float2(s.z, s.w), pa[2 * i + 1]);
                pq[2 * i + 1] = ffc_fma2(wb, make_float2(s.z, s.w), pq[2 * i + 1]);
            }
            if (half == 1) {
                const float2 sl = reinterpret_cast<const float2*>(tiles + c * PBF)[M * 32 + lane];
                pa[8] = ffc_fma2(wa, sl, pa[8]);
                pq[8] = ffc_fma2(wb, sl, pq[8]);
            }
        }
#pragma unroll
        for (int i = 0; i < (half == 0 ? 8 : 9); ++i)
            Y[8 * half + i] = TRANSPOSED ? make_float2(pa[i].x + pa[i].y, pq[i].x + pq[i].y)
                                         : make_float2(pa[i].x + pq[i].y, pq[i].x + pa[i].y);
    }
}

// a lane's 17 bins <-> the exchange image of its plane (float4 slots [k/2][lane], float2 slots [16][lane])
__device__ __forceinline__ void publish17(float* pb, int lane, const float2* Y) {
    float4* ex4 = reinterpret_cast<float4*>(pb);
#pragma unroll
    for (int i = 0; i < M / 2; ++i) ex4[i * 32 + lane] = make_float4(Y[2 * i].x, Y[2 * i].y, Y[2 * i + 1].x, Y[2 * i + 1].y);
    reinterpret_cast<float2*>(pb)[M * 32 + lane] = Y[M];
}
__device__ __forceinline__ void fetch17(const float* pb, int lane, float2* Y) {
    const float4* ex4 = reinterpret_cast<const float4*>(pb);
#pragma unroll
    for (int i = 0; i < M / 2; ++i) {
        const float4 s = ex4[i * 32 + lane];
        Y[2 * i] = make_float2(s.x, s.y);
        Y[2 * i + 1] = make_float2(s.z, s.w);
    }
    Y[M] = reinterpret_cast<const float2*>(pb)[M * 32 + lane];
}

// BatchNorm constants of output plane o (spectrum channels 2o, 2o+1): y -> relu(y * a + b), the inverse ortho scale folded in;
// block 0 publishes the saved statistics and updates the running ones (same arithmetic as Fu2Fwd::bn_constants)
// sums4: sum y (re, im), sum y^2 (re, im) of output plane o over the batch (training mode only)
__device__ __forceinline__ void bn_constants(const Fu2Params& p, double inv_count, double unbias, int o, int lane, const double* sums4, float2& a2, float2& b2) {
    const float scale = 1.0f / (float)N;
    float a[2], b[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int ch = 2 * o + q;
        float mean, invstd;
        if (p.training) {
            const double m = sums4[q] * inv_count;
            double var = sums4[2 + q] * inv_count - m * m;
            if (var < 0.0) var = 0.0;
            mean = (float)m;
            invstd = 1.0f / sqrtf((float)var + p.eps);
            if (blockIdx.x == 0 && lane == 0 && p.running_mean) {
                const double unb = var * unbias;
                p.running_mean[ch] = (1.f - p.momentum) * p.running_mean[ch] + p.momentum * mean;
                p.running_var[ch] = (1.f - p.momentum) * p.running_var[ch] + p.momentum * (float)unb;
            }
        } else {
            mean = p.running_mean[ch];
            invstd = 1.0f / sqrtf(p.running_var[ch] + p.eps);
        }
        if (blockIdx.x == 0 && lane == 0) { p.save_mean[ch] = mean; p.save_invstd[ch] = invstd; }
        a[q] = invstd * __ldg(p.gamma + ch) * scale;
        b[q] = __ldg(p.beta + ch) * scale - mean * a[q];
    }
    a2 = make_float2(a[0], a[1]);
    b2 = make_float2(b[0], b[1]);
}

// MODE 0: statistics pass;  1: apply pass (eval mode, or the second pass of training);  2: training in one cooperative launch
template <int MODE>
__global__ void __launch_bounds__(NW * 32, 2) fu4_kernel(const Fu2Params p, const double inv_count, const double unbias, float* partial) {
    __shared__ __align__(16) float tiles[NW * PBF];
    __shared__ float2 tw_s[32];
    __shared__ float4 wq_s[NW * NW];
    const int lane = threadIdx.x & 31, warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    float* pb = tiles + warp * PBF;
    if (threadIdx.x < 32) tw_s[threadIdx.x] = c_tw128[threadIdx.x * 4];
    __syncthreads();
    const float2 tws[4] = {tw_s[lane & 15], tw_s[(lane & 7) * 2], tw_s[(lane & 3) * 4], tw_s[(lane & 1) * 8]};
    // mix weights as FFMA2 operand pairs (W[2o][2c], W[2o+1][2c+1] | W[2o+1][2c], W[2o][2c+1]), forward ortho scale folded in
    if (threadIdx.x < NW * NW) {
        const int o = threadIdx.x / NW, c = threadIdx.x % NW;
        const float scale = 0.5f / (float)N;            // forward ortho scale and the 1/2 of the row pass
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o < p.Cout && c < p.Cin) {
            const float* r0 = p.w + (size_t)(2 * o) * 2 * p.Cin + 2 * c;
            const float* r1 = r0 + 2 * p.Cin;
            q = make_float4(__ldg(r0) * scale, __ldg(r1 + 1) * scale, __ldg(r1) * scale, __ldg(r0 + 1) * scale);
        }
        wq_s[threadIdx.x] = q;
    }
    __syncthreads();
    float2 bn_a = make_float2(0.f, 0.f), bn_b = bn_a;
    if (MODE == 1 && warp < p.Cout) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        if (p.training) {
            acc[0] = __ldcg(p.sums + 2 * warp); acc[1] = __ldcg(p.sums + 2 * warp + 1);
            acc[2] = __ldcg(p.sums + 2 * p.Cout + 2 * warp); acc[3] = __ldcg(p.sums + 2 * p.Cout + 2 * warp + 1);
        }
        bn_constants(p, inv_count, unbias, warp, lane, acc, bn_a, bn_b);
    }
    float st[4] = {0.f, 0.f, 0.f, 0.f};            // sum re, sum re^2, sum im, sum im^2 of this warp's plane
    float2* ex = reinterpret_cast<float2*>(pb);
    float4* ex4 = reinterpret_cast<float4*>(pb);
    for (int img = blockIdx.x; img < p.B; img += gridDim.x) {
        float2 Y[M + 1];
        if (warp < p.Cin) {
            fwd_plane<false>(p.x + ((size_t)img * p.Cin + warp) * (N * N), pb, lane, tws);
        } else {
#pragma unroll
            for (int i = 0; i < M / 2; ++i) ex4[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            ex[M * 32 + lane] = make_float2(0.f, 0.f);
        }
        __syncthreads();
        if (warp < p.Cout) {
            mix17<false>(tiles, wq_s, warp, lane, Y);
            if (MODE != 1) {
#pragma unroll
                for (int i = 0; i <= M; ++i) {
                    st[0] += Y[i].x; st[1] = fmaf(Y[i].x, Y[i].x, st[1]);
                    st[2] += Y[i].y; st[3] = fmaf(Y[i].y, Y[i].y, st[3]);
                }
            }
        }
        if (MODE == 0) { __syncthreads(); continue; }       // the exchange images are free again
        if (MODE == 2) {
            // statistics of this image: shuffle reduction, then either one float per (sum, CTA) in the caller's workspace
            // (no zeroing launch, no atomics, deterministic) or four double atomics per warp
            if (warp < p.Cout) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int m = 16; m >= 1; m >>= 1) st[j] += __shfl_xor_sync(FULL, st[j], m);
                }
                if (lane < 4) {
                    const float sv = lane == 0 ? st[0] : (lane == 1 ? st[1] : (lane == 2 ? st[2] : st[3]));
                    const int chn = 2 * warp + (lane >> 1);
                    const int idx = (lane & 1) ? 2 * p.Cout + chn : chn;
                    if (partial) __stcg(partial + (size_t)idx * p.B + blockIdx.x, sv);
                    else atomicAdd(p.sums + idx, (double)sv);
                }
            }
            cooperative_groups::this_grid().sync();
            if (warp < p.Cout) {
                if (partial) {
                    double acc[4] = {0.0, 0.0, 0.0, 0.0};      // sum y, sum y^2 of channels 2*warp and 2*warp + 1 over all CTAs
                    const float* pr = partial + (size_t)(2 * warp) * p.B;
                    for (int b = lane; b < p.B; b += 32) {
                        acc[0] += (double)__ldcg(pr + b);
                        acc[1] += (double)__ldcg(pr + p.B + b);
                        acc[2] += (double)__ldcg(pr + (size_t)2 * p.Cout * p.B + b);
                        acc[3] += (double)__ldcg(pr + (size_t)(2 * p.Cout + 1) * p.B + b);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
#pragma unroll
                        for (int m = 16; m >= 1; m >>= 1) acc[j] += __shfl_xor_sync(FULL, acc[j], m);
                    }
                    bn_constants(p, inv_count, unbias, warp, lane, acc, bn_a, bn_b);
                } else {
                    const double acc[4] = {__ldcg(p.sums + 2 * warp), __ldcg(p.sums + 2 * warp + 1),
                                           __ldcg(p.sums + 2 * p.Cout + 2 * warp), __ldcg(p.sums + 2 * p.Cout + 2 * warp + 1)};
                    bn_constants(p, inv_count, unbias, warp, lane, acc, bn_a, bn_b);
                }
            }
        } else {
            __syncthreads();
        }
        if (warp < p.Cout) {
#pragma unroll
            for (int i = 0; i <= M; ++i) Y[i] = fu2_bn_relu(Y[i], bn_a, bn_b);
            const size_t g0 = ((size_t)img * p.Cout + warp) * (N * N);
            inv_plane<false>(Y, pb, lane, tws, p.residual ? p.residual + g0 : nullptr, p.out + g0);
        }
    }
    if (MODE == 0 && warp < p.Cout) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) st[j] += __shfl_xor_sync(FULL, st[j], m);
        }
        if (lane < 4) {
            const float sv = lane == 0 ? st[0] : (lane == 1 ? st[1] : (lane == 2 ? st[2] : st[3]));
            const int chn = 2 * warp + (lane >> 1);
            atomicAdd(p.sums + ((lane & 1) ? 2 * p.Cout + chn : chn), (double)sv);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of the same unit (autograd of fourier_unity.py:32-58; the math of ffc_fu2_bwd.cu) in the warp-private layout.
// Two tiles per warp.  With S' = 2N * ortho spectrum of x (fwd_plane), G' = N * ortho adjoint-c2r transform of dout
// (fwd_plane<ADJ>) and the weights scaled by 1/(2N):
//   warp w: G'_w -> tile A (private), S'_w -> tile B (published) | barrier | warp o: Y_o = mix (true values), y^ = (Y - mean) *
//   invstd, mask from y^ * gamma + beta, dZ' = mask * G', sum dZ', sum dZ' y^ -> per-CTA partial sums | GRID BARRIER |
//   dY = gamma * invstd / N * (dZ' - mean(dZ') - y^ * mean(dZ' y^)) (true dY), dW partial of this image = dY (x) S' over the
//   lane's bins, 32 sums per warp met by a halving shuffle reduction, written per image; dY -> tile A (published) | barrier |
//   warp c: dS'_c = transposed mix, inverse adjoint plane -> dx (all scale factors cancel) | GRID BARRIER | dW = sum of the
//   per-image partials / (2N) (one warp per weight entry; no atomics, bitwise reproducible).
__global__ void __launch_bounds__(NW * 32, 2) fu4_bwd_kernel(const Fu2BwdParams p, const double inv_count, float* partial, float* dwp) {
    extern __shared__ __align__(16) float smem_bwd[];
    float* tilesA = smem_bwd;
    float* tilesB = smem_bwd + NW * PBF;
    __shared__ float2 tw_s[32];
    __shared__ float4 wq_s[NW * NW];
    const int lane = threadIdx.x & 31, warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
    const int img = blockIdx.x;
    float* pa_tile = tilesA + warp * PBF;
    float* pb_tile = tilesB + warp * PBF;
    if (threadIdx.x < 32) tw_s[threadIdx.x] = c_tw128[threadIdx.x * 4];
    if (threadIdx.x < NW * NW) {
        const int o = threadIdx.x / NW, c = threadIdx.x % NW;
        const float scale = 0.5f / (float)N;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o < p.Cout && c < p.Cin) {
            const float* r0 = p.w + (size_t)(2 * o) * 2 * p.Cin + 2 * c;
            const float* r1 = r0 + 2 * p.Cin;
            q = make_float4(__ldg(r0) * scale, __ldg(r1 + 1) * scale, __ldg(r1) * scale, __ldg(r0 + 1) * scale);
        }
        wq_s[threadIdx.x] = q;
    }
    __syncthreads();
    const float2 tws[4] = {tw_s[lane & 15], tw_s[(lane & 7) * 2], tw_s[(lane & 3) * 4], tw_s[(lane & 1) * 8]};
    if (warp < p.Cout) fwd_plane<true>(p.dout + ((size_t)img * p.Cout + warp) * (N * N), pa_tile, lane, tws);
    if (warp < p.Cin) {
        fwd_plane<false>(p.x + ((size_t)img * p.Cin + warp) * (N * N), pb_tile, lane, tws);
    } else {
        float2 zero[M + 1];
#pragma unroll
        for (int i = 0; i <= M; ++i) zero[i] = make_float2(0.f, 0.f);
        publish17(pb_tile, lane, zero);
    }
    __syncthreads();
    float2 dz[M + 1], yh[M + 1];
    float2 gam = make_float2(0.f, 0.f), istd = gam;
    if (warp < p.Cout) {
        const int ch = 2 * warp;
        istd = make_float2(__ldg(p.save_invstd + ch), __ldg(p.save_invstd + ch + 1));
        const float2 k1 = make_float2(-__ldg(p.save_mean + ch) * istd.x, -__ldg(p.save_mean + ch + 1) * istd.y);
        gam = make_float2(__ldg(p.gamma + ch), __ldg(p.gamma + ch + 1));
        const float2 bet = make_float2(__ldg(p.beta + ch), __ldg(p.beta + ch + 1));
        mix17<false>(tilesB, wq_s, warp, lane, yh);                  // Y of this lane's bins (true values)
        fetch17(pa_tile, lane, dz);                                  // G'
        float st[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i <= M; ++i) {
            yh[i] = ffc_fma2(yh[i], istd, k1);
            const float2 z = ffc_fma2(yh[i], gam, bet);
            dz[i] = make_float2(z.x > 0.f ? dz[i].x : 0.f, z.y > 0.f ? dz[i].y : 0.f);
            st[0] += dz[i].x; st[1] = fmaf(dz[i].x, yh[i].x, st[1]);
            st[2] += dz[i].y; st[3] = fmaf(dz[i].y, yh[i].y, st[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) st[j] += __shfl_xor_sync(FULL, st[j], m);
        }
        if (lane < 4) {          // layout as in the forward: sum(dZ) of channel chn at [chn], sum(dZ y^) at [2*Cout + chn]
            const float sv = lane == 0 ? st[0] : (lane == 1 ? st[1] : (lane == 2 ? st[2] : st[3]));
            const int chn = 2 * warp + (lane >> 1);
            const int idx = (lane & 1) ? 2 * p.Cout + chn : chn;
            __stcg(partial + (size_t)idx * p.B + blockIdx.x, sv);
        }
    }
    cooperative_groups::this_grid().sync();
    if (warp < p.Cout) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const float* pr = partial + (size_t)(2 * warp) * p.B;
        for (int b = lane; b < p.B; b += 32) {
            acc[0] += (double)__ldcg(pr + b);
            acc[1] += (double)__ldcg(pr + p.B + b);
            acc[2] += (double)__ldcg(pr + (size_t)2 * p.Cout * p.B + b);
            acc[3] += (double)__ldcg(pr + (size_t)(2 * p.Cout + 1) * p.B + b);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) acc[j] += __shfl_xor_sync(FULL, acc[j], m);
        }
        const float inv_n = 1.0f / (float)N;
        if (blockIdx.x == 0 && lane == 0) {          // the sums are N * the true ones
            p.dbeta[2 * warp] = (float)(acc[0] * inv_n); p.dbeta[2 * warp + 1] = (float)(acc[1] * inv_n);
            p.dgamma[2 * warp] = (float)(acc[2] * inv_n); p.dgamma[2 * warp + 1] = (float)(acc[3] * inv_n);
        }
        const float2 a = make_float2(gam.x * istd.x * inv_n, gam.y * istd.y * inv_n);
        const float2 c1 = p.training ? make_float2((float)(acc[0] * inv_count), (float)(acc[1] * inv_count)) : make_float2(0.f, 0.f);
        const float2 c2 = p.training ? make_float2((float)(acc[2] * inv_count), (float)(acc[3] * inv_count)) : make_float2(0.f, 0.f);
        // dY = a * (dZ - c1 - y^ * c2), in place over dz
#pragma unroll
        for (int i = 0; i <= M; ++i) {
            const float2 t = ffc_sub2(ffc_sub2(dz[i], c1), ffc_mul2(yh[i], c2));
            dz[i] = ffc_mul2(a, t);
        }
        // dW partial of this image: v[4c + t], t = 0: dW[2o][2c], 1: dW[2o+1][2c], 2: dW[2o][2c+1], 3: dW[2o+1][2c+1]
        float v[4 * NW];
#pragma unroll
        for (int c = 0; c < NW; ++c) {
            float2 a1 = make_float2(0.f, 0.f), a2 = a1;
            const float4* xc4 = reinterpret_cast<const float4*>(tilesB + c * PBF) + lane;
#pragma unroll
            for (int i = 0; i < M / 2; ++i) {
                const float4 s = xc4[i * 32];
                a1 = ffc_fma2(dz[2 * i], make_float2(s.x, s.x), a1);
                a2 = ffc_fma2(dz[2 * i], make_float2(s.y, s.y), a2);
                a1 = ffc_fma2(dz[2 * i + 1], make_float2(s.z, s.z), a1);
                a2 = ffc_fma2(dz[2 * i + 1], make_float2(s.w, s.w), a2);
            }
            const float2 sl = reinterpret_cast<const float2*>(tilesB + c * PBF)[M * 32 + lane];
            a1 = ffc_fma2(dz[M], make_float2(sl.x, sl.x), a1);
            a2 = ffc_fma2(dz[M], make_float2(sl.y, sl.y), a2);
            v[4 * c] = a1.x; v[4 * c + 1] = a1.y; v[4 * c + 2] = a2.x; v[4 * c + 3] = a2.y;
        }
        // 32 sums over the 32 lanes by recursive halving: after the step with mask m a lane keeps the half of the values
        // its bit m selects, so lane j ends with the total of v[j]
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int i = 0; i < m; ++i) {
                const float send = up ? v[i] : v[i + m];
                const float keep = up ? v[i + m] : v[i];
                v[i] = keep + __shfl_xor_sync(FULL, send, m);
            }
        }
        __stcg(dwp + ((size_t)(warp * 32 + lane)) * p.B + blockIdx.x, v[0]);
        publish17(pa_tile, lane, dz);            // dY of plane `warp` (G' is no longer needed)
    } else {
        float2 zero[M + 1];
#pragma unroll
        for (int i = 0; i <= M; ++i) zero[i] = make_float2(0.f, 0.f);
        publish17(pa_tile, lane, zero);
    }
    __syncthreads();
    if (warp < p.Cin) {
        float2 ds[M + 1];
        mix17<true>(tilesA, wq_s, warp, lane, ds);
        inv_plane<true>(ds, pb_tile, lane, tws, nullptr, p.dx + ((size_t)img * p.Cin + warp) * (N * N));
    }
    cooperative_groups::this_grid().sync();
    // dW[row][col] = sum over the images of the partials / (2N): one warp per entry
    for (int e = blockIdx.x * NW + warp; e < 4 * NW * NW; e += gridDim.x * NW) {
        const int o = e >> 5, j = e & 31, c = j >> 2, t = j & 3;
        if (o >= p.Cout || c >= p.Cin) continue;
        double acc = 0.0;
        for (int b = lane; b < p.B; b += 32) acc += (double)__ldcg(dwp + (size_t)e * p.B + b);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(FULL, acc, m);
        if (lane == 0) p.dw[(size_t)(2 * o + (t & 1)) * 2 * p.Cin + 2 * c + (t >> 1)] = (float)(acc * (0.5 / (double)N));
    }
}

static size_t bwd_workspace_bytes(int B, int Cin, int Cout) {
    (void)Cin;
    return ((size_t)4 * Cout * sizeof(double) + 255) / 256 * 256 + ((size_t)4 * Cout * B + (size_t)4 * NW * NW * B) * sizeof(float);
}
static int bwd_capacity() {
    static FfcPerDevice cached = {};
    size_t& c = *ffc_device_slot(cached);
    if (c == 0) {
        const int smem = 2 * NW * PBF * (int)sizeof(float);
        int per_sm = 0, coop = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (!coop || cudaFuncSetAttribute(fu4_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 0;
        cudaFuncSetAttribute(fu4_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fu4_bwd_kernel, NW * 32, smem) != cudaSuccess || per_sm < 1) return 0;
        c = (size_t)per_sm * ffc_sm_count();
    }
    return (int)c;
}

// resident CTAs of one kernel on the current device (prefers the largest shared-memory carve-out; cached per device)
template <int MODE>
static int capacity() {
    static FfcPerDevice cached = {};
    size_t& c = *ffc_device_slot(cached);
    if (c == 0) {
        int per_sm = 0;
        cudaFuncSetAttribute(fu4_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fu4_kernel<MODE>, NW * 32, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        c = (size_t)per_sm * ffc_sm_count();
    }
    return (int)c;
}
template <int MODE>
static int launch(int grid, const Fu2Params& p, double inv_count, double unbias, ffc_stream_t st) {
    fu4_kernel<MODE><<<grid, NW * 32, 0, st>>>(p, inv_count, unbias, nullptr);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("fu4 launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}
}  // namespace fu4

extern int ffc_fu2_force_two_pass;
int ffc_fu4_enabled = 1;
extern "C" void ffc_debug_fu4(int on) { ffc_fu4_enabled = on; }

bool ffc_fu4_supported(int Cin, int Cout, int H, int W) {
    return ffc_fu4_enabled && H == 32 && W == 32 && Cin >= 1 && Cout >= 1 && Cin <= fu4::NW && Cout <= fu4::NW;
}

int ffc_fu4_launch(const Fu2Params& p, size_t workspace_bytes, ffc_stream_t st) {
    using namespace fu4;
    const double count = (double)p.B * BINS;                 // 1/count and the unbiased-variance factor: no FP64 division on the device
    double inv_count = 1.0 / count, unbias = count > 1.0 ? count / (count - 1.0) : 1.0;
    if (p.training) {
        int coop = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        const bool one_launch = coop && !ffc_fu2_force_two_pass && p.B <= capacity<2>();
        // per-CTA partial sums behind the 4*Cout doubles when the caller's workspace has room for them (ffc_fu_workspace_bytes)
        float* partial = nullptr;
        const size_t off = ((size_t)4 * p.Cout * sizeof(double) + 255) / 256 * 256;
        if (one_launch && workspace_bytes >= off + (size_t)4 * p.Cout * p.B * sizeof(float)) partial = reinterpret_cast<float*>(reinterpret_cast<char*>(p.sums) + off);
        if (!partial) FFC_CHECK(ffc_memset_async(p.sums, 0, (size_t)4 * p.Cout * sizeof(double), st));
        if (one_launch) {
            void* args[] = {(void*)&p, (void*)&inv_count, (void*)&unbias, (void*)&partial};
            const cudaError_t e = cudaLaunchCooperativeKernel((const void*)fu4_kernel<2>, dim3(p.B), dim3(NW * 32), args, 0, st);
            if (e != cudaSuccess) { ffc_set_error("fu4 cooperative launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
            ffc_count_launch();
            return FFC_OK;
        }
        const int c0 = capacity<0>();
        FFC_CHECK(launch<0>(p.B < c0 ? p.B : c0, p, inv_count, unbias, st));
    }
    const int c1 = capacity<1>();
    return launch<1>(p.B < c1 ? p.B : c1, p, inv_count, unbias, st);
}

extern "C" size_t ffc_fu_bwd_workspace_bytes(int B, int Cin, int Cout) {
    const size_t small = (size_t)4 * (Cout > 0 ? Cout : 0) * sizeof(double);
    if (B < 1 || Cin < 1 || Cout < 1) return small;
    const size_t big = fu4::bwd_workspace_bytes(B, Cin, Cout);
    return big > small ? big : small;
}
// the warp-private backward handles the shape with this workspace on the current device (all images co-resident)
bool ffc_fu4_bwd_supported(int B, int Cin, int Cout, int H, int W, size_t workspace_bytes) {
    if (!ffc_fu4_supported(Cin, Cout, H, W) || B < 1) return false;
    if (workspace_bytes < fu4::bwd_workspace_bytes(B, Cin, Cout)) return false;
    return B <= fu4::bwd_capacity();
}
int ffc_fu4_bwd_launch(const Fu2BwdParams& p, ffc_stream_t st) {
    using namespace fu4;
    double inv_count = 1.0 / ((double)p.B * BINS);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(p.sums) + ((size_t)4 * p.Cout * sizeof(double) + 255) / 256 * 256);
    float* dwp = partial + (size_t)4 * p.Cout * p.B;
    void* args[] = {(void*)&p, (void*)&inv_count, (void*)&partial, (void*)&dwp};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)fu4_bwd_kernel, dim3(p.B), dim3(NW * 32), args,
                                                      2 * NW * PBF * sizeof(float), st);
    if (e != cudaSuccess) { ffc_set_error("fu4 backward cooperative launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}
#else
bool ffc_fu4_supported(int, int, int, int) { return false; }
extern "C" size_t ffc_fu_bwd_workspace_bytes(int, int, int Cout) { return (size_t)4 * (Cout > 0 ? Cout : 0) * sizeof(double); }
bool ffc_fu4_bwd_supported(int, int, int, int, int, size_t) { return false; }
int ffc_fu4_bwd_launch(const Fu2BwdParams&, ffc_stream_t) { return FFC_ERR_BAD_ARG; }
int ffc_fu4_launch(const Fu2Params&, size_t, ffc_stream_t) { return FFC_ERR_BAD_ARG; }
extern "C" void ffc_debug_fu4(int) {}
#endif
