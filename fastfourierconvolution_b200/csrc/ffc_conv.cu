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

// Chunk gather / scatter shared by ConvFwdV2 and ConvFwdV3 (expanded inside run(); they use its locals).
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
    static constexpr int BM = 128, BK = FFC_CONV_BK, TM = 8, TN = BN / 16, kBN = BN;
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
// ConvFwdV3: ConvFwdV2's pipeline with the FMAs moved to the tensor cores at FP32 accuracy ("3xTF32"):
// every FP32 operand x is split into hi = tf32(x), lo = tf32(x - hi) and a.b is accumulated as
// a_lo.b_hi + a_hi.b_lo + a_hi.b_hi with mma.sync.m16n8k8 (FP32 accumulate); the dropped a_lo.b_lo term is
// ~2^-22 relative.  8 warps = 4 (M) x 2 (N), warp tile 32 x (BN/2), fragments read straight from the FP32
// shared-memory tiles (row strides = 8 mod 32 floats make the fragment loads bank-conflict free).
// The emulation build computes each lane's accumulators directly from shared memory (same C-fragment
// ownership, so the epilogue indexing is exercised; the A/B fragment layout is device-only).
// ---------------------------------------------------------------------------------------------
#ifndef FFC_EMU
FFC_DEVICE unsigned ffc_tf32(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
FFC_DEVICE void ffc_mma_tf32(float* c, const unsigned* a, const unsigned* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif

template <int BN>
struct ConvFwdV3 {
    typedef ConvParams Params;
    static constexpr int BM = 128, BK = FFC_CONV_BK, kBN = BN;
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 2;
    // shared tiles hold (hi, lo) TF32 pairs as float2; row strides = 4 mod 16 float2 keep the 64-bit fragment
    // loads conflict free (lanes g = 0..7 are consecutive, the four t groups land 8 banks apart)
    static constexpr int AS = BM + 4, BS = BN + 4;
    static constexpr int A_PER = BK * BM / kThreads;
    static constexpr int B_PER = BK * BN / kThreads;
    static constexpr int MT = 2, NT = BN / 16;               // m16 / n8 tiles per warp
    static_assert(BN == 64 || BN == 32 || BN == 16, "BN");
    static size_t smem_bytes() { return (size_t)2 * (BK * AS + BK * BS) * 8; }
    struct Acc { float v[MT * NT * 4]; };
    struct State { int b, yq, xq, ok; int seg, tap, ta, tb, c0; };   // pixel of the A loader + chunk cursor

    static FFC_DEVICE float2 split(float x) {
#ifdef FFC_EMU
        return make_float2(x, 0.f);
#else
        // hi = x rounded to the nearest TF32 (IADD + LOP3, ffc_common.cuh), lo = x - hi (exact in FP32, one FADD).  The tensor
        // core reads only the top 19 bits of lo, so a.b keeps ~22 significant bits per product.  (cvt.rna.tf32.f32 expands
        // to a ~7 instruction sequence on sm_100 -- ncu showed it dominating the issue slots.)
        const float hi = ffc_tf32_hi(x);
        return make_float2(hi, x - hi);
#endif
    }

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        float2* As = reinterpret_cast<float2*>(smem);          // [2][BK][AS]
        float2* Bs = As + 2 * BK * AS;                         // [2][BK][BS]
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
        const int nchunks = (m0 < Mc) ? T * (cpk0 + cpk1) : 0;

        FFC_TLS(Acc, acc);
        FFC_TLS(State, st);
        // gather of the chunk under the cursor into registers, then the cursor advances (no integer division)
#define FFC_V3_LOAD(ra, rb)                                                                                 \
        {                                                                                                   \
            const float* FFC_RESTRICT xs_ = st.seg ? p.seg[1].x : p.seg[0].x;                               \
            const float* FFC_RESTRICT ws_ = st.seg ? p.seg[1].w : p.seg[0].w;                               \
            const int cin_ = st.seg ? p.seg[1].cin : p.seg[0].cin;                                          \
            int iy_, ix_;                                                                                   \
            if (p.transposed) { iy_ = st.yq + qy - st.ta; ix_ = st.xq + qx - st.tb; }                       \
            else { iy_ = st.yq * p.stride - p.pad + st.ta; ix_ = st.xq * p.stride - p.pad + st.tb; }        \
            const bool okp_ = st.ok && iy_ >= 0 && iy_ < p.Hi && ix_ >= 0 && ix_ < p.Wi;                    \
            /* 32-bit element offsets (tensor sizes < 2^31 are checked on the host): one IMAD.WIDE per load */ \
            const int ca_ = st.c0 + tid / BM;                                                               \
            const float* xp_ = xs_ + ((st.b * cin_ + ca_) * HWi + iy_ * p.Wi + ix_);                        \
            const int xstep_ = (kThreads / BM) * HWi;                                                       \
            const bool full_ = st.c0 + BK <= cin_;                                                          \
            if (okp_ && full_) {                                                                            \
                FFC_UNROLL                                                                                  \
                for (int i = 0; i < A_PER; ++i) ra[i] = FFC_LDG(xp_ + i * xstep_);                          \
            } else {                                                                                        \
                FFC_UNROLL                                                                                  \
                for (int i = 0; i < A_PER; ++i)                                                             \
                    ra[i] = (okp_ && ca_ + i * (kThreads / BM) < cin_) ? FFC_LDG(xp_ + i * xstep_) : 0.f;   \
            }                                                                                               \
            const int co_ = n0 + tid % BN;                                                                  \
            const int cb_ = st.c0 + tid / BN;                                                               \
            const int woff_ = p.transposed ? (ky0 + s * st.ta) * p.k + (kx0 + s * st.tb) : st.ta * p.k + st.tb; \
            const float* wp_ = ws_ + (p.transposed ? (cb_ * p.cout + co_) * KK + woff_ : (co_ * cin_ + cb_) * KK + woff_); \
            const int wstep_ = (kThreads / BN) * (p.transposed ? p.cout * KK : KK);                         \
            if (co_ < p.cout && full_) {                                                                    \
                FFC_UNROLL                                                                                  \
                for (int i = 0; i < B_PER; ++i) rb[i] = FFC_LDG(wp_ + i * wstep_);                          \
            } else {                                                                                        \
                FFC_UNROLL                                                                                  \
                for (int i = 0; i < B_PER; ++i)                                                             \
                    rb[i] = (co_ < p.cout && cb_ + i * (kThreads / BN) < cin_) ? FFC_LDG(wp_ + i * wstep_) : 0.f; \
            }                                                                                               \
            st.c0 += BK;                                                                                    \
            if (st.c0 >= cin_) {                                                                            \
                st.c0 = 0;                                                                                  \
                if (++st.tb == Tb) { st.tb = 0; ++st.ta; }                                                  \
                if (++st.tap == T) { st.tap = 0; st.ta = 0; st.tb = 0; ++st.seg; }                          \
            }                                                                                               \
        }
#define FFC_V3_STORE(buf, ra, rb)                                                                           \
        {                                                                                                   \
            float2* as_ = As + (buf) * BK * AS;                                                             \
            float2* bs_ = Bs + (buf) * BK * BS;                                                             \
            FFC_UNROLL                                                                                      \
            for (int i = 0; i < A_PER; ++i) as_[(tid / BM + i * (kThreads / BM)) * AS + tid % BM] = split(ra[i]); \
            FFC_UNROLL                                                                                      \
            for (int i = 0; i < B_PER; ++i) bs_[(tid / BN + i * (kThreads / BN)) * BS + tid % BN] = split(rb[i]); \
        }

        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            FFC_TLS_REF(State, st);
            FFC_UNROLL
            for (int i = 0; i < MT * NT * 4; ++i) acc.v[i] = 0.f;
            const int m = m0 + tid % BM;
            st.ok = m < Mc;
            st.xq = m % Wc; st.yq = (m / Wc) % Hc; st.b = m / (Wc * Hc);
            st.seg = 0; st.tap = 0; st.ta = 0; st.tb = 0; st.c0 = 0;
            if (nchunks > 0) {
                float ra[A_PER], rb[B_PER];
                FFC_V3_LOAD(ra, rb);
                FFC_V3_STORE(0, ra, rb);
            }
        } FFC_SYNC;
        for (int c = 0; c < nchunks; ++c) {
            FFC_PHASE {
                FFC_TLS_REF(Acc, acc);
                FFC_TLS_REF(State, st);
                float ra[A_PER], rb[B_PER];
                const bool more = c + 1 < nchunks;
                if (more) FFC_V3_LOAD(ra, rb);
                const float2* as = As + (c & 1) * BK * AS;
                const float2* bs = Bs + (c & 1) * BK * BS;
                const int lane = tid & 31, warp = tid >> 5;
                const int g = lane >> 2, t = lane & 3;
                const int mw = (warp & 3) * 32, nw = (warp >> 2) * (BN / 2);
#ifdef FFC_EMU
                for (int mt = 0; mt < MT; ++mt)
                    for (int nt_ = 0; nt_ < NT; ++nt_)
                        for (int r = 0; r < 4; ++r) {
                            const int row = mw + 16 * mt + g + (r >> 1) * 8, col = nw + 8 * nt_ + 2 * t + (r & 1);
                            float sum = acc.v[(mt * NT + nt_) * 4 + r];
                            for (int kl = 0; kl < BK; ++kl) {
                                const float2 a = as[kl * AS + row], b = bs[kl * BS + col];
                                sum = fmaf(a.x + a.y, b.x + b.y, sum);
                            }
                            acc.v[(mt * NT + nt_) * 4 + r] = sum;
                        }
#else
                FFC_UNROLL
                for (int k8 = 0; k8 < BK; k8 += 8) {
                    unsigned ah[MT][4], al[MT][4], bh[NT][2], bl[NT][2];
                    FFC_UNROLL
                    for (int mt = 0; mt < MT; ++mt) {
                        FFC_UNROLL
                        for (int r = 0; r < 4; ++r) {
                            const float2 x = as[(k8 + t + (r >> 1) * 4) * AS + mw + 16 * mt + g + (r & 1) * 8];
                            ah[mt][r] = __float_as_uint(x.x);
                            al[mt][r] = __float_as_uint(x.y);
                        }
                    }
                    FFC_UNROLL
                    for (int nt_ = 0; nt_ < NT; ++nt_) {
                        FFC_UNROLL
                        for (int r = 0; r < 2; ++r) {
                            const float2 x = bs[(k8 + t + r * 4) * BS + nw + 8 * nt_ + g];
                            bh[nt_][r] = __float_as_uint(x.x);
                            bl[nt_][r] = __float_as_uint(x.y);
                        }
                    }
                    FFC_UNROLL
                    for (int mt = 0; mt < MT; ++mt)
                        FFC_UNROLL
                        for (int nt_ = 0; nt_ < NT; ++nt_) {
                            // The three products of one k8 step are summed in a fresh accumulator and folded into the
                            // running FP32 sum with a round-to-nearest FADD: the tensor core aligns (truncates) its
                            // addend, which over a long K would cost ~1e-5; per-step folding keeps the error ~1e-6.
                            float tmp[4] = {0.f, 0.f, 0.f, 0.f};
                            ffc_mma_tf32(tmp, al[mt], bh[nt_]);
                            ffc_mma_tf32(tmp, ah[mt], bl[nt_]);
                            ffc_mma_tf32(tmp, ah[mt], bh[nt_]);
                            float* cc = acc.v + (mt * NT + nt_) * 4;
                            cc[0] += tmp[0]; cc[1] += tmp[1]; cc[2] += tmp[2]; cc[3] += tmp[3];
                        }
                }
#endif
                if (more) FFC_V3_STORE((c + 1) & 1, ra, rb);
            } FFC_SYNC;
        }
#undef FFC_V3_LOAD
#undef FFC_V3_STORE
        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            const int lane = tid & 31, warp = tid >> 5;
            const int g = lane >> 2, t = lane & 3;
            const int mw = (warp & 3) * 32, nw = (warp >> 2) * (BN / 2);
            FFC_UNROLL
            for (int mt = 0; mt < MT; ++mt) {
                FFC_UNROLL
                for (int h = 0; h < 2; ++h) {
                    const int m = m0 + mw + 16 * mt + g + 8 * h;
                    if (m >= Mc) continue;
                    const int xq = m % Wc, yq = (m / Wc) % Hc, b = m / (Wc * Hc);
                    const int oy = yq * s + py, ox = xq * s + px;
                    FFC_UNROLL
                    for (int nt_ = 0; nt_ < NT; ++nt_) {
                        FFC_UNROLL
                        for (int e = 0; e < 2; ++e) {
                            const int co = n0 + nw + 8 * nt_ + 2 * t + e;
                            if (co >= p.cout) continue;
                            const size_t o = ((size_t)(b * p.cout + co) * p.Ho + oy) * p.Wo + ox;
                            float v = acc.v[(mt * NT + nt_) * 4 + 2 * h + e];
                            if (p.bias) v += FFC_LDG(p.bias + co);
                            if (p.addend) v += FFC_LDG(p.addend + o);
                            p.y[o] = v;
                        }
                    }
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

template <class K>
static int conv_fwd_v2_launch(const ConvParams& p, ffc_stream_t st) {
    const int s = p.transposed ? p.stride : 1;
    const int Hc = ffc_cdiv(p.Ho, s), Wc = ffc_cdiv(p.Wo, s);
    const int Mc = p.B * Hc * Wc;
    return ffc_launch<K>(ffc_cdiv(Mc, K::BM), ffc_cdiv(p.cout, K::kBN), s * s, K::kThreads, K::smem_bytes(), st, p);
}

int ffc_conv_use_reference_kernel = 5;      // shared with ffc_conv_v4.cu; 5 = automatic choice (default) in ffc_conv2d_fwd_ws

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
    FFC_REQUIRE((long long)B * cout * Ho * Wo < (1LL << 31) && (long long)B * (cin0 > cin1 ? cin0 : cin1) * Hi * Wi < (1LL << 31)
                && (long long)cout * (cin0 > cin1 ? cin0 : cin1) * k * k < (1LL << 31),
                "ffc_conv2d_fwd: tensor too large for 32-bit element offsets");
    ConvParams p;
    p.seg[0] = ConvSeg{x0, w0, cin0};
    p.seg[1] = ConvSeg{x1, w1, cin1};
    p.nseg = x1 ? 2 : 1;
    p.bias = bias; p.addend = addend; p.y = y;
    p.B = B; p.cout = cout; p.Hi = Hi; p.Wi = Wi; p.Ho = Ho; p.Wo = Wo;
    p.k = k; p.stride = stride; p.pad = pad; p.transposed = transposed;
    ffc_stream_t st = (ffc_stream_t)stream;
    if (ffc_conv_use_reference_kernel == 1) {     // the simple single-buffered form (tests compare all three)
        if (cout <= 8) return conv_fwd_launch<256, 8, 4, 2>(p, st);
        if (cout <= 32) return conv_fwd_launch<128, 32, 4, 4>(p, st);
        return conv_fwd_launch<64, 64, 4, 4>(p, st);
    }
    const bool simt = ffc_conv_use_reference_kernel == 2;
    if (cout <= 16) return simt ? conv_fwd_v2_launch<ConvFwdV2<16>>(p, st) : conv_fwd_v2_launch<ConvFwdV3<16>>(p, st);
    if (cout <= 32 || (cout > 64 && cout % 64 != 0 && cout % 64 <= 32 && cout < 128))
        return simt ? conv_fwd_v2_launch<ConvFwdV2<32>>(p, st) : conv_fwd_v2_launch<ConvFwdV3<32>>(p, st);
    return simt ? conv_fwd_v2_launch<ConvFwdV2<64>>(p, st) : conv_fwd_v2_launch<ConvFwdV3<64>>(p, st);
}

// test hook: 5 (default) ffc_conv2d_fwd_ws picks ConvFwdV5 (tcgen05) or ConvFwdV4 (mma.sync) by output width and plain
// ffc_conv2d_fwd runs ConvFwdV3; 4 forces ConvFwdV5, 3 forces ConvFwdV4, 0 ConvFwdV3 everywhere, 1 simple reference-form
// kernels, 2 tuned FP32 SIMT kernels
extern "C" void ffc_debug_conv_reference(int on) { ffc_conv_use_reference_kernel = on; }

#ifndef FFC_EMU
bool wgrad_v5_supported(int SC, int LC, int Hs, int Ws, int k);
bool wgrad_small_supported(int SC, int k);
int wgrad_small_run(const float* S, const float* L, float* dW, int B, int SC, int LC, int Hs, int Ws, int Hl, int Wl,
                    int k, int stride, int pad, ffc_stream_t st);
int wgrad_v5_run(const float* S, const float* L, float* dW, int B, int SC, int LC, int Hs, int Ws, int Hl, int Wl,
                 int k, int stride, int pad, ffc_stream_t st);
#endif

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
#ifndef FFC_EMU
    // tcgen05 kernel (ffc_wgrad_v5.cu) for the shapes that fill its 128 x N tile; kernel families 4 and 5 (default)
    if (ffc_conv_use_reference_kernel >= 5 && wgrad_small_supported(SC, k))
        return wgrad_small_run(S, L, dW, B, SC, LC, Hs, Ws, Hl, Wl, k, stride, pad, st);
    if (ffc_conv_use_reference_kernel >= 4 && wgrad_v5_supported(SC, LC, Hs, Ws, k))
        return wgrad_v5_run(S, L, dW, B, SC, LC, Hs, Ws, Hl, Wl, k, stride, pad, st);
#endif
    WgradParams p{S, L, dW, B, SC, LC, Hs, Ws, Hl, Wl, k, stride, pad, 0};
    const int Ktot = B * Hs * Ws;
    const bool ref = ffc_conv_use_reference_kernel == 1;
    const int BMt = 64, BNt = ref ? 64 : 128;
    const int gx = ffc_cdiv(SC, BMt), gy = ffc_cdiv(LC * k * k, BNt);
    // split K so that the grid has ~3 CTAs per SM (148 SMs), at least 8 K-steps per CTA
    int nsplit = ffc_cdiv(3 * ffc_sm_count(), gx * gy);
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
