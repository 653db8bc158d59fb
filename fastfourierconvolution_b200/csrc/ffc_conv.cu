// Local-branch convolutions of FFC / FFCTranspose (layers/ffc/ffc.py:45-68, 91-94;
// layers/ffc/ffc_transpose.py:84-86, 98-104) and the 1x1 convolutions of SpectralTransform /
// FourierUnitSN (spectral_transform.py:52-53, 70-71; fourier_unity.py:23-24), FP32 SIMT
// implicit GEMM.  Two gather forms cover forward and data-gradient of both layer types:
//
//   conv form  (transposed = 0):  y[b,co,oy,ox] = sum_{ci,ky,kx} x[b,ci,oy*s-p+ky,ox*s-p+kx] * w[co][ci][ky][kx]
//       = nn.Conv2d forward,  and  nn.ConvTranspose2d data-gradient (x := dy, w := weight[ci_op][co_op]).
//   convT form (transposed = 1):  y[b,co,oy,ox] = sum_{ci,ky,kx : (oy+p-ky) % s == 0} x[b,ci,(oy+p-ky)/s,...] * w[ci][co][ky][kx]
//       = nn.ConvTranspose2d forward,  and  nn.Conv2d data-gradient (x := dy, w := weight[co_op][ci_op]).
//       The output is processed per parity class (oy % s, ox % s) so only taps that hit are visited.
//
// Up to two (input, weight) segments are summed into one output, which is how
// out_xl = convl2l(x_l) + convg2l(x_g) (ffc.py:91) runs as a single kernel; an optional addend
// tensor implements out_xg = convl2g(x_l) + convg2g(x_g) (ffc.py:94-96) without a separate add.
//
// Weight gradient (both layer types):
//   dW[sc][lc][ky][kx] = sum_{b,y,x} S[b,sc,y,x] * L[b,lc,y*s-p+ky,x*s-p+kx]
//   Conv2d: S = dy, L = x;  ConvTranspose2d: S = x, L = dy.
#include "ffc_common.cuh"

#define FFC_CONV_MAXSEG 2
#define FFC_CONV_BK 16

// N consecutive floats from 16-byte aligned shared memory (N = 4, 2 or 1 elements per access)
template <int N>
FFC_DEVICE void ffc_lds_vec(float* dst, const float* src) {
    if constexpr (N % 4 == 0) {
        FFC_UNROLL
        for (int i = 0; i < N / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
            dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
        }
    } else if constexpr (N % 2 == 0) {
        FFC_UNROLL
        for (int i = 0; i < N / 2; ++i) {
            const float2 v = *reinterpret_cast<const float2*>(src + 2 * i);
            dst[2 * i] = v.x; dst[2 * i + 1] = v.y;
        }
    } else {
        FFC_UNROLL
        for (int i = 0; i < N; ++i) dst[i] = src[i];
    }
}

struct ConvSeg { const float* x; const float* w; int cin; };

struct ConvParams {
    ConvSeg seg[FFC_CONV_MAXSEG];
    int nseg;
    const float* bias;     // [cout] or null
    const float* addend;   // (B, cout, Ho, Wo) or null
    float* y;              // (B, cout, Ho, Wo)
    int B, cout, Hi, Wi, Ho, Wo, k, stride, pad, transposed;
};

template <int BM, int BN, int TM, int TN>
struct ConvFwdKernel {
    typedef ConvParams Params;
    static constexpr int BK = FFC_CONV_BK;
    static constexpr int kThreads = (BM / TM) * (BN / TN);
    static constexpr int AS = BM + 4, BS = BN + 4;          // smem row strides (floats)
    static size_t smem_bytes() { return (size_t)(BK * AS + BK * BS) * 4; }
    struct Acc { float v[TM * TN]; };
    struct Pix { int b, yq, xq; bool ok; };
    static_assert(kThreads % BM == 0, "A-tile loader assumes a fixed pixel per thread");

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        float* As = smem;                  // [BK][AS]
        float* Bs = smem + BK * AS;        // [BK][BS]
        const int s = p.transposed ? p.stride : 1;            // output parity classes per dim
        const int py = ctx.bz / s, px = ctx.bz % s;
        const int Hc = (p.Ho - py + s - 1) / s, Wc = (p.Wo - px + s - 1) / s;   // class grid
        const int Mc = p.B * Hc * Wc;
        const int m0 = ctx.bx * BM, n0 = ctx.by * BN;
        // tap geometry of this class
        int ky0 = 0, kx0 = 0, qy = 0, qx = 0, Ta = p.k, Tb = p.k;
        if (p.transposed) {
            ky0 = (py + p.pad) % s; kx0 = (px + p.pad) % s;
            qy = (py + p.pad - ky0) / s; qx = (px + p.pad - kx0) / s;
            Ta = ky0 < p.k ? (p.k - ky0 + s - 1) / s : 0;
            Tb = kx0 < p.k ? (p.k - kx0 + s - 1) / s : 0;
        }
        const int T = Ta * Tb;
        const int KK = p.k * p.k;
        FFC_TLS(Acc, acc);
        FFC_TLS(Pix, pix);
        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            FFC_TLS_REF(Pix, pix);
            FFC_UNROLL
            for (int i = 0; i < TM * TN; ++i) acc.v[i] = 0.f;
            // the pixel this thread gathers for the A tile is the same in every K step
            const int m = m0 + tid % BM;
            pix.ok = m < Mc;
            pix.xq = m % Wc; pix.yq = (m / Wc) % Hc; pix.b = m / (Wc * Hc);
        }
        if (m0 < Mc && T > 0) {
            for (int sg = 0; sg < p.nseg; ++sg) {
                // (select instead of p.seg[sg]: a dynamic index would force the params into local memory)
                const float* FFC_RESTRICT xs = sg == 0 ? p.seg[0].x : p.seg[1].x;
                const float* FFC_RESTRICT ws = sg == 0 ? p.seg[0].w : p.seg[1].w;
                const int cin = sg == 0 ? p.seg[0].cin : p.seg[1].cin;
                const int Ktot = cin * T;
                for (int k0 = 0; k0 < Ktot; k0 += BK) {
                    FFC_PHASE {
                        // A tile: gathered input pixels, As[kl][ml]
                        FFC_TLS_REF(Pix, pix);
                        for (int e = tid; e < BK * BM; e += kThreads) {
                            const int ml = e % BM, kl = e / BM;
                            const int kk = k0 + kl;
                            float v = 0.f;
                            if (pix.ok && kk < Ktot) {
                                const int xq = pix.xq, yq = pix.yq, b = pix.b;
                                const int ci = kk / T, t = kk % T, a = t / Tb, bb = t % Tb;
                                int iy, ix;
                                if (p.transposed) { iy = yq + qy - a; ix = xq + qx - bb; }
                                else { iy = yq * p.stride - p.pad + a; ix = xq * p.stride - p.pad + bb; }
                                if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi)
                                    v = FFC_LDG(xs + ((size_t)(b * cin + ci) * p.Hi + iy) * p.Wi + ix);
                            }
                            As[kl * AS + ml] = v;
                        }
                        // B tile: weights, Bs[kl][nl]
                        for (int e = tid; e < BK * BN; e += kThreads) {
                            const int nl = e % BN, kl = e / BN;
                            const int co = n0 + nl, kk = k0 + kl;
                            float v = 0.f;
                            if (co < p.cout && kk < Ktot) {
                                const int ci = kk / T, t = kk % T, a = t / Tb, bb = t % Tb;
                                if (p.transposed) {
                                    const int ky = ky0 + s * a, kx = kx0 + s * bb;
                                    v = FFC_LDG(ws + ((size_t)(ci * p.cout + co) * p.k + ky) * p.k + kx);
                                } else {
                                    v = FFC_LDG(ws + (size_t)(co * cin + ci) * KK + a * p.k + bb);
                                }
                            }
                            Bs[kl * BS + nl] = v;
                        }
                    } FFC_SYNC;
                    FFC_PHASE {
                        FFC_TLS_REF(Acc, acc);
                        const int tm = tid % (BM / TM), tn = tid / (BM / TM);
                        FFC_UNROLL
                        for (int kl = 0; kl < BK; ++kl) {
                            float a[TM], b[TN];
                            ffc_lds_vec<TM>(a, As + kl * AS + tm * TM);
                            ffc_lds_vec<TN>(b, Bs + kl * BS + tn * TN);
                            FFC_UNROLL
                            for (int i = 0; i < TM; ++i)
                                FFC_UNROLL
                                for (int j = 0; j < TN; ++j) acc.v[i * TN + j] = fmaf(a[i], b[j], acc.v[i * TN + j]);
                        }
                    } FFC_SYNC;
                }
            }
        }
        // epilogue: bias, addend, store.  Classes with no taps (T == 0) still write bias/addend.
        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            const int tm = tid % (BM / TM), tn = tid / (BM / TM);
            FFC_UNROLL
            for (int i = 0; i < TM; ++i) {
                const int m = m0 + tm * TM + i;
                if (m >= Mc) continue;
                const int xq = m % Wc, yq = (m / Wc) % Hc, b = m / (Wc * Hc);
                const int oy = yq * s + py, ox = xq * s + px;
                FFC_UNROLL
                for (int j = 0; j < TN; ++j) {
                    const int co = n0 + tn * TN + j;
                    if (co >= p.cout) continue;
                    const size_t o = ((size_t)(b * p.cout + co) * p.Ho + oy) * p.Wo + ox;
                    float v = acc.v[i * TN + j];
                    if (p.bias) v += FFC_LDG(p.bias + co);
                    if (p.addend) v += FFC_LDG(p.addend + o);
                    p.y[o] = v;
                }
            }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
// ConvFwdV2: the production forward / data-gradient kernel.  Same math as ConvFwdKernel (kept as the
// simple reference form), restructured for FMA throughput:
//   * K is ordered tap-major (tap, ci): a BK chunk shares one tap, so the gather address of a thread's pixel
//     is computed once per chunk and advances by a constant channel stride;
//   * register prefetch + two shared-memory buffers: the global loads of chunk c+1 are issued before the
//     FMAs of chunk c and stored afterwards, one barrier per chunk;
//   * 128 x BN CTA tile, 8 x TN register tile (TN = BN/16), float4 shared-memory reads that are conflict free.
// ---------------------------------------------------------------------------------------------
template <int BN>
struct ConvFwdV2 {
    typedef ConvParams Params;
    static constexpr int BM = 128, BK = FFC_CONV_BK, TM = 8, TN = BN / 16;
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 2;
    static constexpr int AS = BM + 4, BS = BN + 4;
    static constexpr int A_PER = BK * BM / kThreads;          // 8 gathered inputs per thread per chunk
    static constexpr int B_PER = BK * BN / kThreads;          // 4 / 2 / 1 weights per thread per chunk
    static_assert(BN == 64 || BN == 32 || BN == 16, "BN");
    static size_t smem_bytes() { return (size_t)2 * (BK * AS + BK * BS) * 4; }
    struct Acc { float v[TM * TN]; };
    struct Pix { int b, yq, xq; bool ok; };

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        float* As = smem;                       // [2][BK][AS]
        float* Bs = smem + 2 * BK * AS;         // [2][BK][BS]
        const int s = p.transposed ? p.stride : 1;
        const int py = ctx.bz / s, px = ctx.bz % s;
        const int Hc = (p.Ho - py + s - 1) / s, Wc = (p.Wo - px + s - 1) / s;
        const int Mc = p.B * Hc * Wc;
        const int m0 = ctx.bx * BM, n0 = ctx.by * BN;
        int ky0 = 0, kx0 = 0, qy = 0, qx = 0, Ta = p.k, Tb = p.k;
        if (p.transposed) {
            ky0 = (py + p.pad) % s; kx0 = (px + p.pad) % s;
            qy = (py + p.pad - ky0) / s; qx = (px + p.pad - kx0) / s;
            Ta = ky0 < p.k ? (p.k - ky0 + s - 1) / s : 0;
            Tb = kx0 < p.k ? (p.k - kx0 + s - 1) / s : 0;
        }
        const int T = Ta * Tb;
        const int KK = p.k * p.k;
        const int HWi = p.Hi * p.Wi;
        const int cpk0 = (p.seg[0].cin + BK - 1) / BK;
        const int cpk1 = p.nseg > 1 ? (p.seg[1].cin + BK - 1) / BK : 0;
        const int nch0 = T * cpk0, nchunks = (m0 < Mc) ? T * (cpk0 + cpk1) : 0;

        FFC_TLS(Acc, acc);
        FFC_TLS(Pix, pix);
        // gathers chunk `c` into registers: ra[] (inputs of this thread's pixel) and rb[] (weights of its channel)
#define FFC_CONV_LOAD_CHUNK(c, ra, rb)                                                                      \
        {                                                                                                   \
            const int sg_ = (c) < nch0 ? 0 : 1;                                                             \
            const int r_ = (c) - sg_ * nch0;                                                                \
            const int cpk_ = sg_ ? cpk1 : cpk0;                                                             \
            const int tap_ = r_ / cpk_, c0_ = (r_ % cpk_) * BK;                                             \
            const int a_ = tap_ / Tb, b_ = tap_ % Tb;                                                       \
            const float* FFC_RESTRICT xs_ = sg_ ? p.seg[1].x : p.seg[0].x;                                  \
            const float* FFC_RESTRICT ws_ = sg_ ? p.seg[1].w : p.seg[0].w;                                  \
            const int cin_ = sg_ ? p.seg[1].cin : p.seg[0].cin;                                             \
            int iy_, ix_;                                                                                   \
            if (p.transposed) { iy_ = pix.yq + qy - a_; ix_ = pix.xq + qx - b_; }                           \
            else { iy_ = pix.yq * p.stride - p.pad + a_; ix_ = pix.xq * p.stride - p.pad + b_; }            \
            const bool okp_ = pix.ok && iy_ >= 0 && iy_ < p.Hi && ix_ >= 0 && ix_ < p.Wi;                   \
            const float* xp_ = xs_ + ((size_t)pix.b * cin_ + c0_ + tid / BM) * HWi + iy_ * p.Wi + ix_;      \
            FFC_UNROLL                                                                                      \
            for (int i = 0; i < A_PER; ++i) {                                                               \
                const int ci_ = c0_ + tid / BM + i * (kThreads / BM);                                       \
                ra[i] = (okp_ && ci_ < cin_) ? FFC_LDG(xp_ + (size_t)i * (kThreads / BM) * HWi) : 0.f;      \
            }                                                                                               \
            const int co_ = n0 + tid % BN;                                                                  \
            const int woff_ = p.transposed ? (ky0 + s * a_) * p.k + (kx0 + s * b_) : a_ * p.k + b_;         \
            FFC_UNROLL                                                                                      \
            for (int i = 0; i < B_PER; ++i) {                                                               \
                const int ci_ = c0_ + tid / BN + i * (kThreads / BN);                                       \
                float v_ = 0.f;                                                                             \
                if (co_ < p.cout && ci_ < cin_)                                                             \
                    v_ = p.transposed ? FFC_LDG(ws_ + ((size_t)ci_ * p.cout + co_) * KK + woff_)            \
                                      : FFC_LDG(ws_ + ((size_t)co_ * cin_ + ci_) * KK + woff_);             \
                rb[i] = v_;                                                                                 \
            }                                                                                               \
        }
#define FFC_CONV_STORE_CHUNK(buf, ra, rb)                                                                   \
        {                                                                                                   \
            float* as_ = As + (buf) * BK * AS;                                                              \
            float* bs_ = Bs + (buf) * BK * BS;                                                              \
            FFC_UNROLL                                                                                      \
            for (int i = 0; i < A_PER; ++i) as_[(tid / BM + i * (kThreads / BM)) * AS + tid % BM] = ra[i];  \
            FFC_UNROLL                                                                                      \
            for (int i = 0; i < B_PER; ++i) bs_[(tid / BN + i * (kThreads / BN)) * BS + tid % BN] = rb[i];  \
        }

        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            FFC_TLS_REF(Pix, pix);
            FFC_UNROLL
            for (int i = 0; i < TM * TN; ++i) acc.v[i] = 0.f;
            const int m = m0 + tid % BM;
            pix.ok = m < Mc;
            pix.xq = m % Wc; pix.yq = (m / Wc) % Hc; pix.b = m / (Wc * Hc);
            if (nchunks > 0) {
                float ra[A_PER], rb[B_PER];
                FFC_CONV_LOAD_CHUNK(0, ra, rb);
                FFC_CONV_STORE_CHUNK(0, ra, rb);
            }
        } FFC_SYNC;
        for (int c = 0; c < nchunks; ++c) {
            FFC_PHASE {
                FFC_TLS_REF(Acc, acc);
                FFC_TLS_REF(Pix, pix);
                float ra[A_PER], rb[B_PER];
                const bool more = c + 1 < nchunks;
                if (more) FFC_CONV_LOAD_CHUNK(c + 1, ra, rb);
                const float* as = As + (c & 1) * BK * AS;
                const float* bs = Bs + (c & 1) * BK * BS;
                const int tm = tid % 16, tn = tid / 16;
                FFC_UNROLL
                for (int kl = 0; kl < BK; ++kl) {
                    float a[TM], b[TN];
                    ffc_lds_vec<4>(a, as + kl * AS + tm * 4);
                    ffc_lds_vec<4>(a + 4, as + kl * AS + 64 + tm * 4);
                    ffc_lds_vec<TN>(b, bs + kl * BS + tn * TN);
                    FFC_UNROLL
                    for (int i = 0; i < TM; ++i)
                        FFC_UNROLL
                        for (int j = 0; j < TN; ++j) acc.v[i * TN + j] = fmaf(a[i], b[j], acc.v[i * TN + j]);
                }
                if (more) FFC_CONV_STORE_CHUNK((c + 1) & 1, ra, rb);
            } FFC_SYNC;
        }
#undef FFC_CONV_LOAD_CHUNK
#undef FFC_CONV_STORE_CHUNK
        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            const int tm = tid % 16, tn = tid / 16;
            FFC_UNROLL
            for (int i = 0; i < TM; ++i) {
                const int m = m0 + (i < 4 ? tm * 4 + i : 64 + tm * 4 + (i - 4));
                if (m >= Mc) continue;
                const int xq = m % Wc, yq = (m / Wc) % Hc, b = m / (Wc * Hc);
                const int oy = yq * s + py, ox = xq * s + px;
                FFC_UNROLL
                for (int j = 0; j < TN; ++j) {
                    const int co = n0 + tn * TN + j;
                    if (co >= p.cout) continue;
                    const size_t o = ((size_t)(b * p.cout + co) * p.Ho + oy) * p.Wo + ox;
                    float v = acc.v[i * TN + j];
                    if (p.bias) v += FFC_LDG(p.bias + co);
                    if (p.addend) v += FFC_LDG(p.addend + o);
                    p.y[o] = v;
                }
            }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
struct WgradParams {
    const float* S;     // (B, SC, Hs, Ws)   tensor on the strided ("small") grid
    const float* L;     // (B, LC, Hl, Wl)   tensor that is gathered at y*s - p + ky
    float* dW;          // [SC][LC][k][k], accumulated with atomics (caller zeroes)
    int B, SC, LC, Hs, Ws, Hl, Wl, k, stride, pad;
    int kchunk;         // K elements (of B*Hs*Ws) per blockIdx.z
};

template <int BM, int BN, int TM, int TN>
struct ConvWgradKernel {
    typedef WgradParams Params;
    static constexpr int BK = FFC_CONV_BK;
    static constexpr int kThreads = (BM / TM) * (BN / TN);
    static constexpr int AS = BM + 4, BS = BN + 4;
    static size_t smem_bytes() { return (size_t)(BK * AS + BK * BS) * 4; }
    struct Acc { float v[TM * TN]; };

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        float* As = smem;
        float* Bs = smem + BK * AS;
        const int KK = p.k * p.k;
        const int Ntot = p.LC * KK;
        const int HWs = p.Hs * p.Ws;
        const int Ktot = p.B * HWs;
        const int m0 = ctx.bx * BM, n0 = ctx.by * BN;
        const int kbeg = ctx.bz * p.kchunk;
        const int kend = (kbeg + p.kchunk) < Ktot ? (kbeg + p.kchunk) : Ktot;
        FFC_TLS(Acc, acc);
        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            FFC_UNROLL
            for (int i = 0; i < TM * TN; ++i) acc.v[i] = 0.f;
        }
        for (int k0 = kbeg; k0 < kend; k0 += BK) {
            FFC_PHASE {
                for (int e = tid; e < BK * BM; e += kThreads) {
                    const int kl = e % BK, ml = e / BK;
                    const int kk = k0 + kl, sc = m0 + ml;
                    float v = 0.f;
                    if (kk < kend && sc < p.SC) {
                        const int b = kk / HWs, r = kk % HWs;
                        v = FFC_LDG(p.S + (size_t)(b * p.SC + sc) * HWs + r);
                    }
                    As[kl * AS + ml] = v;
                }
                for (int e = tid; e < BK * BN; e += kThreads) {
                    const int kl = e % BK, nl = e / BK;
                    const int kk = k0 + kl, n = n0 + nl;
                    float v = 0.f;
                    if (kk < kend && n < Ntot) {
                        const int b = kk / HWs, r = kk % HWs, y = r / p.Ws, x = r % p.Ws;
                        const int lc = n / KK, t = n % KK, ky = t / p.k, kx = t % p.k;
                        const int ly = y * p.stride - p.pad + ky, lx = x * p.stride - p.pad + kx;
                        if (ly >= 0 && ly < p.Hl && lx >= 0 && lx < p.Wl)
                            v = FFC_LDG(p.L + ((size_t)(b * p.LC + lc) * p.Hl + ly) * p.Wl + lx);
                    }
                    Bs[kl * BS + nl] = v;
                }
            } FFC_SYNC;
            FFC_PHASE {
                FFC_TLS_REF(Acc, acc);
                const int tn = tid % (BN / TN), tm = tid / (BN / TN);
                FFC_UNROLL
                for (int kl = 0; kl < BK; ++kl) {
                    float a[TM], b[TN];
                    ffc_lds_vec<TM>(a, As + kl * AS + tm * TM);
                    ffc_lds_vec<TN>(b, Bs + kl * BS + tn * TN);
                    FFC_UNROLL
                    for (int i = 0; i < TM; ++i)
                        FFC_UNROLL
                        for (int j = 0; j < TN; ++j) acc.v[i * TN + j] = fmaf(a[i], b[j], acc.v[i * TN + j]);
                }
            } FFC_SYNC;
        }
        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            const int tn = tid % (BN / TN), tm = tid / (BN / TN);
            FFC_UNROLL
            for (int i = 0; i < TM; ++i) {
                const int sc = m0 + tm * TM + i;
                if (sc >= p.SC) continue;
                FFC_UNROLL
                for (int j = 0; j < TN; ++j) {
                    const int n = n0 + tn * TN + j;
                    if (n >= Ntot) continue;
                    ffc_atomic_add(p.dW + (size_t)sc * Ntot + n, acc.v[i * TN + j]);
                }
            }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
// ConvWgradV2: production weight-gradient kernel (same contract as ConvWgradKernel): 64 x 128 tile,
// 4 x 8 register tile, register prefetch + double-buffered shared memory, the (lc, ky, kx) decode of each
// thread's gathered columns hoisted out of the K loop.
// ---------------------------------------------------------------------------------------------
struct ConvWgradV2 {
    typedef WgradParams Params;
    static constexpr int BM = 64, BN = 128, BK = FFC_CONV_BK, TM = 4, TN = 8;
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 2;
    static constexpr int AS = BM + 4, BS = BN + 4;
    static constexpr int A_PER = BK * BM / kThreads;      // 4
    static constexpr int B_PER = BK * BN / kThreads;      // 8
    static size_t smem_bytes() { return (size_t)2 * (BK * AS + BK * BS) * 4; }
    struct Acc { float v[TM * TN]; };
    struct Cols { int off[B_PER]; int kyx[B_PER]; };      // per gathered column: base offset, (ky << 8 | kx) or -1

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        float* As = smem;                       // [2][BK][AS]
        float* Bs = smem + 2 * BK * AS;         // [2][BK][BS]
        const int KK = p.k * p.k;
        const int Ntot = p.LC * KK;
        const int HWs = p.Hs * p.Ws, HWl = p.Hl * p.Wl;
        const int Ktot = p.B * HWs;
        const int m0 = ctx.bx * BM, n0 = ctx.by * BN;
        const int kbeg = ctx.bz * p.kchunk;
        const int kend = (kbeg + p.kchunk) < Ktot ? (kbeg + p.kchunk) : Ktot;
        const int nchunks = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

        FFC_TLS(Acc, acc);
        FFC_TLS(Cols, cols);
#define FFC_WG_LOAD_CHUNK(c, ra, rb)                                                                        \
        {                                                                                                   \
            const int kk_ = kbeg + (c) * BK + tid % BK;                                                     \
            const bool okk_ = kk_ < kend;                                                                   \
            const int b_ = okk_ ? kk_ / HWs : 0, r_ = okk_ ? kk_ % HWs : 0;                                 \
            const int y_ = r_ / p.Ws, x_ = r_ % p.Ws;                                                       \
            const float* sp_ = p.S + ((size_t)b_ * p.SC + m0 + tid / BK) * HWs + r_;                        \
            FFC_UNROLL                                                                                      \
            for (int i = 0; i < A_PER; ++i) {                                                               \
                const int sc_ = m0 + tid / BK + i * (kThreads / BK);                                        \
                ra[i] = (okk_ && sc_ < p.SC) ? FFC_LDG(sp_ + (size_t)i * (kThreads / BK) * HWs) : 0.f;      \
            }                                                                                               \
            const float* lp_ = p.L + (size_t)b_ * p.LC * HWl + (y_ * p.stride) * p.Wl + x_ * p.stride;      \
            FFC_UNROLL                                                                                      \
            for (int i = 0; i < B_PER; ++i) {                                                               \
                float v_ = 0.f;                                                                             \
                if (okk_ && cols.kyx[i] >= 0) {                                                             \
                    const int ly_ = y_ * p.stride - p.pad + (cols.kyx[i] >> 8);                             \
                    const int lx_ = x_ * p.stride - p.pad + (cols.kyx[i] & 255);                            \
                    if (ly_ >= 0 && ly_ < p.Hl && lx_ >= 0 && lx_ < p.Wl) v_ = FFC_LDG(lp_ + cols.off[i]);  \
                }                                                                                           \
                rb[i] = v_;                                                                                 \
            }                                                                                               \
        }
#define FFC_WG_STORE_CHUNK(buf, ra, rb)                                                                     \
        {                                                                                                   \
            float* as_ = As + (buf) * BK * AS + (tid % BK) * AS;                                            \
            float* bs_ = Bs + (buf) * BK * BS + (tid % BK) * BS;                                            \
            FFC_UNROLL                                                                                      \
            for (int i = 0; i < A_PER; ++i) as_[tid / BK + i * (kThreads / BK)] = ra[i];                    \
            FFC_UNROLL                                                                                      \
            for (int i = 0; i < B_PER; ++i) bs_[tid / BK + i * (kThreads / BK)] = rb[i];                    \
        }

        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            FFC_TLS_REF(Cols, cols);
            FFC_UNROLL
            for (int i = 0; i < TM * TN; ++i) acc.v[i] = 0.f;
            FFC_UNROLL
            for (int i = 0; i < B_PER; ++i) {
                const int n = n0 + tid / BK + i * (kThreads / BK);
                if (n < Ntot) {
                    const int lc = n / KK, t = n % KK, ky = t / p.k, kx = t % p.k;
                    cols.off[i] = lc * HWl + (ky - p.pad) * p.Wl + (kx - p.pad);
                    cols.kyx[i] = (ky << 8) | kx;
                } else { cols.off[i] = 0; cols.kyx[i] = -1; }
            }
            if (nchunks > 0) {
                float ra[A_PER], rb[B_PER];
                FFC_WG_LOAD_CHUNK(0, ra, rb);
                FFC_WG_STORE_CHUNK(0, ra, rb);
            }
        } FFC_SYNC;
        for (int c = 0; c < nchunks; ++c) {
            FFC_PHASE {
                FFC_TLS_REF(Acc, acc);
                FFC_TLS_REF(Cols, cols);
                float ra[A_PER], rb[B_PER];
                const bool more = c + 1 < nchunks;
                if (more) FFC_WG_LOAD_CHUNK(c + 1, ra, rb);
                const float* as = As + (c & 1) * BK * AS;
                const float* bs = Bs + (c & 1) * BK * BS;
                const int tm = tid % 16, tn = tid / 16;
                FFC_UNROLL
                for (int kl = 0; kl < BK; ++kl) {
                    float a[TM], b[TN];
                    ffc_lds_vec<4>(a, as + kl * AS + tm * 4);
                    ffc_lds_vec<4>(b, bs + kl * BS + tn * 4);
                    ffc_lds_vec<4>(b + 4, bs + kl * BS + 64 + tn * 4);
                    FFC_UNROLL
                    for (int i = 0; i < TM; ++i)
                        FFC_UNROLL
                        for (int j = 0; j < TN; ++j) acc.v[i * TN + j] = fmaf(a[i], b[j], acc.v[i * TN + j]);
                }
                if (more) FFC_WG_STORE_CHUNK((c + 1) & 1, ra, rb);
            } FFC_SYNC;
        }
#undef FFC_WG_LOAD_CHUNK
#undef FFC_WG_STORE_CHUNK
        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            if (nchunks > 0) {
                const int tm = tid % 16, tn = tid / 16;
                FFC_UNROLL
                for (int i = 0; i < TM; ++i) {
                    const int sc = m0 + tm * 4 + i;
                    if (sc >= p.SC) continue;
                    FFC_UNROLL
                    for (int j = 0; j < TN; ++j) {
                        const int n = n0 + (j < 4 ? tn * 4 + j : 64 + tn * 4 + (j - 4));
                        if (n >= Ntot) continue;
                        ffc_atomic_add(p.dW + (size_t)sc * Ntot + n, acc.v[i * TN + j]);
                    }
                }
            }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
static int conv_fwd_launch(const ConvParams& p, ffc_stream_t st) {
    typedef ConvFwdKernel<BM, BN, TM, TN> K;
    const int s = p.transposed ? p.stride : 1;
    const int Hc = ffc_cdiv(p.Ho, s), Wc = ffc_cdiv(p.Wo, s);       // largest class
    const int Mc = p.B * Hc * Wc;
    return ffc_launch<K>(ffc_cdiv(Mc, BM), ffc_cdiv(p.cout, BN), s * s, K::kThreads, K::smem_bytes(), st, p);
}

template <int BN>
static int conv_fwd_v2_launch(const ConvParams& p, ffc_stream_t st) {
    typedef ConvFwdV2<BN> K;
    const int s = p.transposed ? p.stride : 1;
    const int Hc = ffc_cdiv(p.Ho, s), Wc = ffc_cdiv(p.Wo, s);
    const int Mc = p.B * Hc * Wc;
    return ffc_launch<K>(ffc_cdiv(Mc, K::BM), ffc_cdiv(p.cout, BN), s * s, K::kThreads, K::smem_bytes(), st, p);
}

static int ffc_conv_use_reference_kernel = 0;

// See include/ffc_b200.h for the contract.
extern "C" int ffc_conv2d_fwd(const float* x0, const float* w0, int cin0,
                              const float* x1, const float* w1, int cin1,
                              const float* bias, const float* addend, float* y,
                              int B, int cout, int Hi, int Wi, int Ho, int Wo,
                              int k, int stride, int pad, int transposed, void* stream) {
    FFC_REQUIRE(x0 && w0 && y && cin0 > 0, "ffc_conv2d_fwd: null pointer / empty first segment");
    FFC_REQUIRE((x1 == nullptr) == (cin1 == 0) && (x1 == nullptr) == (w1 == nullptr), "ffc_conv2d_fwd: inconsistent second segment");
    FFC_REQUIRE(k >= 1 && k <= 7 && (stride == 1 || stride == 2) && pad >= 0 && pad < 8, "ffc_conv2d_fwd: unsupported k=%d stride=%d pad=%d", k, stride, pad);
    FFC_REQUIRE(B >= 0 && cout > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "ffc_conv2d_fwd: bad sizes");
    if (!transposed) {
        FFC_REQUIRE(Ho == (Hi + 2 * pad - k) / stride + 1 && Wo == (Wi + 2 * pad - k) / stride + 1,
                    "ffc_conv2d_fwd: output size %dx%d inconsistent with input %dx%d k=%d s=%d p=%d", Ho, Wo, Hi, Wi, k, stride, pad);
    } else {
        const int hmin = (Hi - 1) * stride - 2 * pad + k, wmin = (Wi - 1) * stride - 2 * pad + k;
        FFC_REQUIRE(Ho >= hmin && Ho < hmin + stride && Wo >= wmin && Wo < wmin + stride,
                    "ffc_conv2d_fwd: transposed output size %dx%d inconsistent with input %dx%d k=%d s=%d p=%d", Ho, Wo, Hi, Wi, k, stride, pad);
    }
    if (B == 0) return FFC_OK;
    FFC_REQUIRE((long long)B * Ho * Wo < (1LL << 31) && (long long)B * (cin0 + cin1) * Hi * Wi < (1LL << 31), "ffc_conv2d_fwd: tensor too large for 32-bit pixel indices");
    ConvParams p;
    p.seg[0] = ConvSeg{x0, w0, cin0};
    p.seg[1] = ConvSeg{x1, w1, cin1};
    p.nseg = x1 ? 2 : 1;
    p.bias = bias; p.addend = addend; p.y = y;
    p.B = B; p.cout = cout; p.Hi = Hi; p.Wi = Wi; p.Ho = Ho; p.Wo = Wo;
    p.k = k; p.stride = stride; p.pad = pad; p.transposed = transposed;
    ffc_stream_t st = (ffc_stream_t)stream;
    if (ffc_conv_use_reference_kernel) {          // the simple single-buffered form (tests compare both)
        if (cout <= 8) return conv_fwd_launch<256, 8, 4, 2>(p, st);
        if (cout <= 32) return conv_fwd_launch<128, 32, 4, 4>(p, st);
        return conv_fwd_launch<64, 64, 4, 4>(p, st);
    }
    if (cout <= 16) return conv_fwd_v2_launch<16>(p, st);
    if (cout <= 32 || (cout > 64 && cout % 64 != 0 && cout % 64 <= 32 && cout < 128)) return conv_fwd_v2_launch<32>(p, st);
    return conv_fwd_v2_launch<64>(p, st);
}

// test hook: 1 selects the simple reference-form kernels for ffc_conv2d_fwd
extern "C" void ffc_debug_conv_reference(int on) { ffc_conv_use_reference_kernel = on; }

extern "C" int ffc_conv2d_wgrad(const float* S, const float* L, float* dW,
                                int B, int SC, int LC, int Hs, int Ws, int Hl, int Wl,
                                int k, int stride, int pad, void* stream) {
    FFC_REQUIRE(S && L && dW, "ffc_conv2d_wgrad: null pointer");
    FFC_REQUIRE(k >= 1 && k <= 7 && (stride == 1 || stride == 2) && pad >= 0, "ffc_conv2d_wgrad: unsupported k=%d stride=%d pad=%d", k, stride, pad);
    FFC_REQUIRE(B >= 0 && SC > 0 && LC > 0 && Hs > 0 && Ws > 0 && Hl > 0 && Wl > 0, "ffc_conv2d_wgrad: bad sizes");
    FFC_REQUIRE((long long)B * Hs * Ws < (1LL << 31), "ffc_conv2d_wgrad: reduction too large");
    ffc_stream_t st = (ffc_stream_t)stream;
    const size_t nW = (size_t)SC * LC * k * k;
    FFC_CHECK(ffc_memset_async(dW, 0, nW * sizeof(float), st));
    if (B == 0) return FFC_OK;
    WgradParams p{S, L, dW, B, SC, LC, Hs, Ws, Hl, Wl, k, stride, pad, 0};
    const int Ktot = B * Hs * Ws;
    const bool ref = ffc_conv_use_reference_kernel != 0;
    const int BMt = 64, BNt = ref ? 64 : 128;
    const int gx = ffc_cdiv(SC, BMt), gy = ffc_cdiv(LC * k * k, BNt);
    // split K so that the grid has ~3 CTAs per SM (148 SMs), at least 8 K-steps per CTA
    int nsplit = ffc_cdiv(3 * 148, gx * gy);
    const int max_split = ffc_cdiv(Ktot, 8 * FFC_CONV_BK);
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 65535) nsplit = 65535;
    int kchunk = ffc_cdiv(ffc_cdiv(Ktot, nsplit), FFC_CONV_BK) * FFC_CONV_BK;
    nsplit = ffc_cdiv(Ktot, kchunk);
    p.kchunk = kchunk;
    if (!ref) return ffc_launch<ConvWgradV2>(gx, gy, nsplit, ConvWgradV2::kThreads, ConvWgradV2::smem_bytes(), st, p);
    typedef ConvWgradKernel<64, 64, 4, 4> K;
    return ffc_launch<K>(gx, gy, nsplit, K::kThreads, K::smem_bytes(), st, p);
}
