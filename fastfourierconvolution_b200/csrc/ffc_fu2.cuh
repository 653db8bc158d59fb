// Building blocks of the fused FourierUnit kernels, second generation (ffc_fu2.cu forward, ffc_fu2_bwd.cu backward).
//
// One CTA works on ONE image at a time (all channels), with enough threads per image to fill an SM even when the
// batch only gives one or two images per SM (the BASELINE shapes: 128-256 images of 8-32 channels on 148 SMs):
//   rows    : one thread per real row.  A row of N reals IS N/2 complex numbers z[n] = x[2n] + i x[2n+1]; an
//             in-register N/2-point FFT plus the even/odd post-processing gives the N/2+1 bins of the real FFT,
//             written back over the row (in place; row stride RS = N + 4 floats holds N + 2).
//   columns : N = N1 * N2 (32 = 8 x 4, 16 = 4 x 4, 8 = 8 x 1).  Pass A: N2 threads per column, each an N1-point
//             FFT over the rows N2*n1 + n2, twiddled; pass B: N1 threads (items) per column, each an N2-point
//             FFT over the rows N2*k1 + i.  Frequency k = k1 + N1*k2 ends up at row N2*k1 + k2 ("permuted
//             order"); everything between the forward and the inverse transform is pointwise in (u, v).
// Plane layout: [plane][N rows][RS floats], REGION = N * RS, so row r of plane c is row c*N + r of one tall matrix.
// Bank behaviour: thread-per-row float4 accesses are conflict free because RS/4 is odd; column passes put
// consecutive lanes on consecutive bins v of a row.
#pragma once
#include "ffc_fft.cuh"

// parameters of the fused forward kernels (ffc_fu2.cu, ffc_fu4.cu)
struct Fu2Params {
    const float* x;          // (B, Cin, N, N)
    const float* w;          // [2*Cout][2*Cin]
    const float* gamma; const float* beta;          // [2*Cout]
    float* running_mean; float* running_var;        // [2*Cout]
    float* save_mean; float* save_invstd;           // [2*Cout]
    const float* residual;   // (B, Cout, N, N) or null
    float* out;              // (B, Cout, N, N)
    double* sums;            // [4*Cout]: sum(y) then sum(y^2), channel-major (2*Cout each)
    int B, Cin, Cout, training;
    float eps, momentum;
};

// parameters of the fused backward kernels (ffc_fu2_bwd.cu, ffc_fu4.cu)
struct Fu2BwdParams {
    const float* x;          // (B, Cin, N, N)
    const float* dout;       // (B, Cout, N, N)
    const float* w;          // [2*Cout][2*Cin]
    const float* gamma; const float* beta;              // [2*Cout]
    const float* save_mean; const float* save_invstd;   // [2*Cout]
    float* dx;               // (B, Cin, N, N)
    float* dw;               // [2*Cout][2*Cin], accumulated with atomics (zeroed by the host wrapper)
    float* dgamma; float* dbeta;                         // [2*Cout]
    double* sums;            // [4*Cout]: sum(dZ) then sum(dZ*y^) (zeroed by the host wrapper)
    int B, Cin, Cout, training;
};

template <int N>
struct Fu2G {
    static constexpr int M = N / 2, Wf = M + 1, RS = N + 4, SPS = RS / 2, BINS = N * Wf;
    static constexpr int REGION = N * RS;
    static constexpr int N1 = (N == 16) ? 4 : 8, N2 = N / N1;       // 64 = 8 x 8 and 128 = 8 x 16 serve the L2-staged form (ffc_fu3.cu)
    static_assert(N == 8 || N == 16 || N == 32 || N == 64 || N == 128, "Fu2G: N in {8,16,32,64,128}");
    static_assert((RS / 4) % 2 == 1, "row stride must be an odd number of float4");
};

// w_N^k = (cos, sin)(2 pi k / N) with a compile-time index: a constant-bank operand after unrolling
template <int N> FFC_DEVICE float2 fu2_twc(int k) { return c_tw128[k * (FFC_TW_N / N)]; }

// ---- rows, forward: real row -> N/2+1 bins (unnormalised), in place.  ADJ doubles the interior bins (the adjoint
// of the c2r transform, used on the incoming gradient by the backward kernel).
template <int N, bool ADJ>
FFC_DEVICE void fu2_rows_fwd(int tid, int nt, int nrows, float* planes) {
    typedef Fu2G<N> G;
    constexpr int M = G::M;
    for (int r = tid; r < nrows; r += nt) {
        float* row = planes + (size_t)r * G::RS;
        float2 z[M];
        FFC_UNROLL
        for (int j = 0; j < N / 4; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(row + 4 * j);
            z[2 * j] = make_float2(a.x, a.y);
            z[2 * j + 1] = make_float2(a.z, a.w);
        }
        ffc_fft_regs<M, -1>(z);
        float2* out = reinterpret_cast<float2*>(row);
        constexpr float S = ADJ ? 2.0f : 1.0f;
        out[0] = make_float2(z[0].x + z[0].y, 0.f);
        out[M] = make_float2(z[0].x - z[0].y, 0.f);
        FFC_UNROLL
        for (int k = 1; k <= M / 2; ++k) {
            const float2 a = z[k], b = make_float2(z[M - k].x, -z[M - k].y);       // Z[k], conj(Z[M-k])
            const float2 e = make_float2((0.5f * S) * (a.x + b.x), (0.5f * S) * (a.y + b.y));
            const float2 o = make_float2((0.5f * S) * (a.y - b.y), (-0.5f * S) * (a.x - b.x));   // (a - b) / (2i)
            const float2 w = fu2_twc<N>(k);                                        // w_N^k = (c, -s)
            const float2 t = make_float2(o.x * w.x + o.y * w.y, o.y * w.x - o.x * w.y);
            out[k] = make_float2(e.x + t.x, e.y + t.y);
            if (k != M - k) out[M - k] = make_float2(e.x - t.x, t.y - e.y);       // conj(e - t)
        }
    }
}

// ---- rows, inverse (c2r semantics of torch.fft.irfftn's last dimension: imaginary parts of bins 0 and N/2 are
// ignored), in place, result multiplied by `scale`.  ADJ halves the interior bins (the adjoint of the r2c transform).
template <int N, bool ADJ>
FFC_DEVICE void fu2_rows_inv(int tid, int nt, int nrows, float* planes, float scale) {
    typedef Fu2G<N> G;
    constexpr int M = G::M;
    for (int r = tid; r < nrows; r += nt) {
        float* row = planes + (size_t)r * G::RS;
        float2 x[G::RS / 2];
        FFC_UNROLL
        for (int j = 0; j < G::RS / 4; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(row + 4 * j);
            x[2 * j] = make_float2(a.x, a.y);
            x[2 * j + 1] = make_float2(a.z, a.w);
        }
        float2 z[M];
        z[0] = make_float2(x[0].x + x[M].x, x[0].x - x[M].x);
        constexpr float S = ADJ ? 0.5f : 1.0f;
        FFC_UNROLL
        for (int k = 1; k <= M / 2; ++k) {
            const float2 p = x[k], q = make_float2(x[M - k].x, -x[M - k].y);
            const float2 e = make_float2(S * (p.x + q.x), S * (p.y + q.y));
            const float2 dm = make_float2(S * (p.x - q.x), S * (p.y - q.y));
            const float2 w = fu2_twc<N>(k);                                        // conj(w_N^k) = (c, +s)
            const float2 d = make_float2(dm.x * w.x - dm.y * w.y, dm.x * w.y + dm.y * w.x);
            z[k] = make_float2(e.x - d.y, e.y + d.x);                              // e + i d
            if (k != M - k) z[M - k] = make_float2(e.x + d.y, d.x - e.y);         // conj(e - i d)
        }
        ffc_fft_regs<M, +1>(z);
        FFC_UNROLL
        for (int j = 0; j < N / 4; ++j)
            *reinterpret_cast<float4*>(row + 4 * j) =
                make_float4(z[2 * j].x * scale, z[2 * j].y * scale, z[2 * j + 1].x * scale, z[2 * j + 1].y * scale);
    }
}

// BatchNorm + ReLU folded to one packed FMA and two maxima per complex value: y -> relu(y * a + b)
struct Fu2Bn { const float2* a; const float2* b; };
FFC_DEVICE float2 fu2_bn_relu(float2 y, float2 a, float2 b) {
    const float2 z = ffc_fma2(y, a, b);
    return make_float2(z.x > 0.f ? z.x : 0.f, z.y > 0.f ? z.y : 0.f);
}

// ---- columns, strided level (rows N2*i + n2, i < N1): forward pass A (twiddle after) / inverse pass A' (last)
template <int N, int SIGN, bool BN>
FFC_DEVICE void fu2_cols_strided(int tid, int nt, int np, float* planes, const float2* tw, Fu2Bn bn) {
    typedef Fu2G<N> G;
    constexpr int N1 = G::N1, N2 = G::N2, Wf = G::Wf;
    for (int it = tid; it < np * Wf * N2; it += nt) {
        // bins v < M first: 16 consecutive lanes then sit on 16 consecutive float2 of one row (no bank conflicts,
        // power-of-two index math); the Nyquist bin v = M of every (plane, n2) follows as a short tail
        int v, n2, pl;
        if (it < np * G::M * N2) { v = it % G::M; n2 = (it / G::M) % N2; pl = it / (G::M * N2); }
        else { const int j = it - np * G::M * N2; v = G::M; n2 = j % N2; pl = j / N2; }
        float2* col = reinterpret_cast<float2*>(planes + pl * G::REGION) + v;
        float2 c[N1];
        FFC_UNROLL
        for (int i = 0; i < N1; ++i) c[i] = col[(N2 * i + n2) * G::SPS];
        if (BN) {
            const float2 a = bn.a[pl], b = bn.b[pl];
            FFC_UNROLL
            for (int i = 0; i < N1; ++i) c[i] = fu2_bn_relu(c[i], a, b);
        }
        ffc_fft_regs<N1, SIGN>(c);
        FFC_UNROLL
        for (int i = 0; i < N1; ++i) {
            if (SIGN < 0 && N2 > 1 && i > 0) col[(N2 * i + n2) * G::SPS] = ffc_cmul_tw<-1>(c[i], tw[n2 * i]);
            else col[(N2 * i + n2) * G::SPS] = c[i];
        }
    }
}
// ---- columns, contiguous level (rows N2*k1 + i, i < N2): forward pass B (last) / inverse pass B' (twiddle after)
template <int N, int SIGN, bool BN>
FFC_DEVICE void fu2_cols_contig(int tid, int nt, int np, float* planes, const float2* tw, Fu2Bn bn) {
    typedef Fu2G<N> G;
    constexpr int N1 = G::N1, N2 = G::N2, Wf = G::Wf;
    for (int it = tid; it < np * Wf * N1; it += nt) {
        int v, k1, pl;
        if (it < np * G::M * N1) { v = it % G::M; k1 = (it / G::M) % N1; pl = it / (G::M * N1); }
        else { const int j = it - np * G::M * N1; v = G::M; k1 = j % N1; pl = j / N1; }
        float2* col = reinterpret_cast<float2*>(planes + pl * G::REGION) + v + (N2 * k1) * G::SPS;
        float2 c[N2];
        FFC_UNROLL
        for (int i = 0; i < N2; ++i) c[i] = col[i * G::SPS];
        if (BN) {
            const float2 a = bn.a[pl], b = bn.b[pl];
            FFC_UNROLL
            for (int i = 0; i < N2; ++i) c[i] = fu2_bn_relu(c[i], a, b);
        }
        ffc_fft_regs<N2, SIGN>(c);
        FFC_UNROLL
        for (int i = 0; i < N2; ++i) {
            if (SIGN > 0 && i > 0) col[i * G::SPS] = ffc_cmul_tw<+1>(c[i], tw[i * k1]);
            else col[i * G::SPS] = c[i];
        }
    }
}

// whole column transforms as phase sequences (macros: FFC_PHASE needs `ctx`)
#define FU2_COLS_FWD(N, np, planes, tw)                                                                        \
    do {                                                                                                       \
        Fu2Bn nobn_; nobn_.a = nullptr; nobn_.b = nullptr;                                                     \
        FFC_PHASE { fu2_cols_strided<N, -1, false>(tid, ctx.nt, np, planes, tw, nobn_); } FFC_SYNC;            \
        if constexpr (Fu2G<N>::N2 > 1) {                                                                       \
            FFC_PHASE { fu2_cols_contig<N, -1, false>(tid, ctx.nt, np, planes, tw, nobn_); } FFC_SYNC;         \
        }                                                                                                      \
    } while (0)
// BNF: apply BatchNorm + ReLU (constants `bn`) to every element as it is first loaded
#define FU2_COLS_INV(N, BNF, np, planes, tw, bn)                                                               \
    do {                                                                                                       \
        Fu2Bn nobn_; nobn_.a = nullptr; nobn_.b = nullptr;                                                     \
        if constexpr (Fu2G<N>::N2 > 1) {                                                                       \
            FFC_PHASE { fu2_cols_contig<N, +1, BNF>(tid, ctx.nt, np, planes, tw, bn); } FFC_SYNC;              \
            FFC_PHASE { fu2_cols_strided<N, +1, false>(tid, ctx.nt, np, planes, tw, nobn_); } FFC_SYNC;        \
        } else {                                                                                               \
            FFC_PHASE { fu2_cols_strided<N, +1, BNF>(tid, ctx.nt, np, planes, tw, bn); } FFC_SYNC;             \
        }                                                                                                      \
    } while (0)

// ---- coalesced global <-> shared copies of `nrows` rows of N floats (rows are contiguous in global memory).
// nt is a multiple of the N/4 float4 of a row, so a thread keeps its float4 column and walks rows with constant
// strides in both address spaces (no per-element index math); LDU loads are in flight before the first store.
template <int N>
FFC_DEVICE void fu2_load_rows(int tid, int nt, int nrows, const float* FFC_RESTRICT src, float* planes) {
    typedef Fu2G<N> G;
    constexpr int Q = N / 4, LDU = 4;
    const int rstep = nt / Q;
    const float4* s4 = reinterpret_cast<const float4*>(src) + tid;
    float* d = planes + (tid / Q) * G::RS + 4 * (tid % Q);
    int r = tid / Q;
    for (; r + (LDU - 1) * rstep < nrows; r += LDU * rstep, s4 += LDU * nt, d += LDU * rstep * G::RS) {
        float4 v[LDU];
        FFC_UNROLL
        for (int u = 0; u < LDU; ++u) v[u] = FFC_LDG(s4 + u * nt);
        FFC_UNROLL
        for (int u = 0; u < LDU; ++u) *reinterpret_cast<float4*>(d + u * rstep * G::RS) = v[u];
    }
    for (; r < nrows; r += rstep, s4 += nt, d += rstep * G::RS) *reinterpret_cast<float4*>(d) = FFC_LDG(s4);
}
template <int N, bool RES, int LDU = 4>
FFC_DEVICE void fu2_store_rows_impl(int tid, int nt, int nrows, const float* planes, const float* FFC_RESTRICT res, float* FFC_RESTRICT dst) {
    typedef Fu2G<N> G;
    constexpr int Q = N / 4;
    const int rstep = nt / Q;
    float4* d4 = reinterpret_cast<float4*>(dst) + tid;
    const float4* r4 = reinterpret_cast<const float4*>(res) + tid;
    const float* s = planes + (tid / Q) * G::RS + 4 * (tid % Q);
    int r = tid / Q;
    for (; r + (LDU - 1) * rstep < nrows; r += LDU * rstep, d4 += LDU * nt, r4 += LDU * nt, s += LDU * rstep * G::RS) {
        float4 q[LDU];
        if (RES) {
            FFC_UNROLL
            for (int u = 0; u < LDU; ++u) q[u] = FFC_LDG(r4 + u * nt);
        }
        FFC_UNROLL
        for (int u = 0; u < LDU; ++u) {
            float4 v = *reinterpret_cast<const float4*>(s + u * rstep * G::RS);
            if (RES) { v.x += q[u].x; v.y += q[u].y; v.z += q[u].z; v.w += q[u].w; }
            d4[u * nt] = v;
        }
    }
    for (; r < nrows; r += rstep, d4 += nt, r4 += nt, s += rstep * G::RS) {
        float4 v = *reinterpret_cast<const float4*>(s);
        if (RES) { const float4 q = FFC_LDG(r4); v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
        *d4 = v;
    }
}
template <int N>
FFC_DEVICE void fu2_store_rows(int tid, int nt, int nrows, const float* planes, const float* FFC_RESTRICT res, float* FFC_RESTRICT dst) {
    if (res) fu2_store_rows_impl<N, true>(tid, nt, nrows, planes, res, dst);
    else fu2_store_rows_impl<N, false>(tid, nt, nrows, planes, res, dst);
}

// ---- pieces shared by the forward and the backward kernel ----------------------------------------------------
// per-shape tuning: OG = output (complex) channels per mix thread
template <int N, int CP> struct Fu2Cfg {
    static constexpr int OG = (N == 32) ? (CP < 8 ? CP : (CP == 32 ? 32 : 8)) : (N == 16 ? 4 : 2);
    static constexpr bool kInPlaceAlways = (N == 32 && CP == 32);       // two regions would not fit
    // CTA width = the widest phase of the common channel counts in whole warps (N=32: 8 ch * 17 bins * 4 threads)
    static constexpr int kThreads = (N == 32) ? 544 : (N == 16 ? 576 : 640);
    // two resident CTAs (a batch of 256 images on 148 SMs) wherever the register tile of the mix allows it
    static constexpr int kMinBlocks = (CP == 32 || (N == 32 && CP == 16)) ? 1 : 2;
};

// float2 offset of bin `bin` inside a plane; bins v < M come first (conflict-free lanes), the Nyquist bins follow
template <int N> FFC_DEVICE int fu2_bin_off(int bin) {
    typedef Fu2G<N> G;
    return bin < N * G::M ? (bin / G::M) * G::SPS + (bin % G::M) : (bin - N * G::M) * G::SPS + G::M;
}

// mix weights as FFMA2 operand pairs (W[2o][2c], W[2o+1][2c+1], W[2o+1][2c], W[2o][2c+1]) * 1/N, zero padded to CP
// input channels, and the twiddle table tw[k] = w_N^k
template <int N, int CP>
FFC_DEVICE void fu2_prologue(const float* FFC_RESTRICT w, int Cin, int Cout, float4* wq, float2* tw, int tid, int nt) {
    const float scale = 1.0f / (float)N;
    for (int e = tid; e < Cout * CP; e += nt) {
        const int o2 = e / CP, c = e % CP;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < Cin) {
            const float* r0 = w + (size_t)(2 * o2) * 2 * Cin + 2 * c;      // W[2o][2c], W[2o][2c+1]
            const float* r1 = r0 + 2 * Cin;                                // W[2o+1][2c], W[2o+1][2c+1]
            q = make_float4(FFC_LDG(r0) * scale, FFC_LDG(r1 + 1) * scale, FFC_LDG(r1) * scale, FFC_LDG(r0 + 1) * scale);
        }
        wq[e] = q;
    }
    for (int k = tid; k < N; k += nt) tw[k] = c_tw128[k * (FFC_TW_N / N)];
}

// CTA-level reduction of per-thread partial sums red[tid][4] = (re: s1, s2, im: s1, s2) of the thread layout
// tid = channel * S + slice, S = nt / Cout, through a second buffer red[nt*4 ...], then one double atomic per sum:
// sums[chn] += s1, sums[2*Cout + chn] += s2 with chn = 2*channel + (0 re | 1 im).
static FFC_DEVICE void fu2_flush_sums(const BlockCtx& ctx, float* red, int Cout, double* sums) {
    const int S = ctx.nt / Cout;
    int S2 = ctx.nt / (4 * Cout);              // second-level fan-in: Cout * 4 * S2 threads sum S / S2 slices each
    if (S2 > 8) S2 = 8;
    if (S2 > S) S2 = S;
    FFC_PHASE {
        if (tid < Cout * 4 * S2) {
            const int part = tid % S2, j = (tid / S2) % 4, c2 = tid / (4 * S2);
            float s = 0.f;
            for (int sl = part; sl < S; sl += S2) s += red[((size_t)c2 * S + sl) * 4 + j];
            red[(size_t)ctx.nt * 4 + tid] = s;
        }
    } FFC_SYNC;
    FFC_PHASE {
        if (tid < Cout * 4) {
            const int j = tid % 4, c2 = tid / 4;
            double s = 0.0;
            for (int part = 0; part < S2; ++part) s += (double)red[(size_t)ctx.nt * 4 + (size_t)tid * S2 + part];
            const int chn = 2 * c2 + (j >> 1);
            ffc_atomic_add(sums + ((j & 1) ? 2 * Cout + chn : chn), s);
        }
    } FFC_SYNC;
}
