// ConvFwdV4: tensor-core (3xTF32) implicit-GEMM convolution with an asynchronous, register-free load pipeline.
// Same contract as ffc_conv2d_fwd (ffc_conv.cu) -- nn.Conv2d / nn.ConvTranspose2d forward and each other's
// data-gradient, two summed input segments, fused bias / addend -- restructured around what ncu showed for
// ConvFwdV3 (profiles/r01e..r01g): ~17 issued instructions per HMMA, most of them address arithmetic of the
// register-staged gather and of the strided weight loads.
//
//   * The weights are re-packed once per call (PackWeightsKernel) into the exact B-tile order of the GEMM,
//     already split into (hi, lo) TF32 halves and zero padded:  Wp[class][K_pad][N_pad] float2, K = (segment, tap,
//     ci) tap-major.  A B tile is then BK contiguous rows: two 16-byte cp.async per thread, no index math.
//   * The gathered activations stay raw FP32 in shared memory (4-byte cp.async with zero fill for padding and
//     channel tails); hi = x rounded to the nearest TF32, lo = x - hi: 3 ALU ops.
//   * 3-stage cp.async ring, one __syncthreads per K chunk, no staging registers.
//   * The three products of a chunk are accumulated in a zeroed fragment and folded into the FP32 running sum
//     with a round-to-nearest FADD (the tensor core truncates its addend).
#include "ffc_common.cuh"

#ifndef FFC_EMU
FFC_DEVICE void ffc_cp_async4(float* dst, const float* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n));
}
FFC_DEVICE void ffc_cp_async16(void* dst, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src));
}
FFC_DEVICE void ffc_cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> FFC_DEVICE void ffc_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
FFC_DEVICE void ffc_mma_tf32_v4(float* c, const unsigned* a, const unsigned* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#else
FFC_DEVICE void ffc_cp_async4(float* dst, const float* src, bool valid) { *dst = valid ? *src : 0.f; }
FFC_DEVICE void ffc_cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
FFC_DEVICE void ffc_cp_async_commit() {}
template <int N> FFC_DEVICE void ffc_cp_async_wait() {}
#endif

#define FFC_V4_MAXCLS 4

extern int ffc_conv_use_reference_kernel;       // ffc_conv.cu: kernel family selected by ffc_debug_conv_reference()
extern "C" int ffc_conv2d_fwd(const float* x0, const float* w0, int cin0, const float* x1, const float* w1, int cin1,
                              const float* bias, const float* addend, float* y, int B, int cout, int Hi, int Wi, int Ho, int Wo,
                              int k, int stride, int pad, int transposed, void* stream);

struct ConvV4Params {
    const float* x[2]; int cin[2]; int cpad[2];     // segments: input, channels, channels padded to BK
    int nseg;
    const float2* wp;          // packed weights
    int cls_off[FFC_V4_MAXCLS];   // float2 offset of each class' [K_pad][N_pad] matrix
    int npad;                  // N padded to BN
    const float* bias; const float* addend; float* y;
    int B, cout, Hi, Wi, Ho, Wo, k, stride, pad, transposed;
};

#include "ffc_conv_geom.cuh"

// ---------------------------------------------------------------------------------------------
// weight packing: Wp[cls][(seg, tap, ci)][n] = split(W(seg)[..]) with zero padding
// ---------------------------------------------------------------------------------------------
struct PackWParams {
    const float* w[2]; int cin[2]; int cpad[2]; int nseg;
    float2* wp; int cls_off[FFC_V4_MAXCLS]; int npad, cout, k, stride, pad, transposed, ncls;
};
struct PackWeightsKernel {
    typedef PackWParams Params;
    static constexpr int kThreads = 256;
    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float*) {
        const int cls = ctx.by;
        const ConvClassGeom g = ffc_conv_class_geom(cls, p.k, p.stride, p.pad, p.transposed);
        const int T = g.Ta * g.Tb;
        const int KK = p.k * p.k;
        const int rows = T * (p.cpad[0] + (p.nseg > 1 ? p.cpad[1] : 0));
        const long long total = (long long)rows * p.npad;
        FFC_PHASE {
            for (long long e = (long long)ctx.bx * kThreads + tid; e < total; e += (long long)ctx.gx * kThreads) {
                const int n = (int)(e % p.npad);
                int row = (int)(e / p.npad);
                int sg = 0;
                if (row >= T * p.cpad[0]) { sg = 1; row -= T * p.cpad[0]; }
                const int cpad = sg ? p.cpad[1] : p.cpad[0], cin = sg ? p.cin[1] : p.cin[0];
                const int tap = row / cpad, ci = row % cpad;
                float v = 0.f;
                if (n < p.cout && ci < cin) {
                    const int a = tap / g.Tb, b = tap % g.Tb;
                    const float* w = sg ? p.w[1] : p.w[0];
                    if (p.transposed) v = FFC_LDG(w + ((size_t)ci * p.cout + n) * KK + (g.ky0 + p.stride * a) * p.k + (g.kx0 + p.stride * b));
                    else v = FFC_LDG(w + ((size_t)n * cin + ci) * KK + a * p.k + b);
                }
#ifdef FFC_EMU
                p.wp[p.cls_off[cls] + e] = make_float2(v, 0.f);
#else
                const float hi = ffc_tf32_hi(v);
                p.wp[p.cls_off[cls] + e] = make_float2(hi, v - hi);
#endif
            }
        } FFC_SYNC;
    }
};

// ---------------------------------------------------------------------------------------------
template <int BN, int BK>
struct ConvFwdV4 {
    typedef ConvV4Params Params;
    static constexpr int BM = 128, kBN = BN, STAGES = 3;
    static constexpr int kThreads = 256;
    static constexpr int kMinBlocks = 2;
    static constexpr int AS = BM + 8;            // floats per A row: 8 mod 32 -> conflict-free fragment loads
    static constexpr int BS = BN + 4;            // float2 per B row: 4 mod 16
    static constexpr int A_PER = BK * BM / kThreads;
    static constexpr int B_PER = BK * BN / 2 / kThreads;      // 16-byte pieces per thread
    static constexpr int MT = 2, NT = BN / 16;
    static_assert(BN == 64 || BN == 32, "BN");
    static_assert(B_PER >= 1, "B tile too small for the 16-byte copy mapping");
    static constexpr int STAGE_FLOATS = BK * AS + BK * BS * 2;
    static size_t smem_bytes() { return (size_t)STAGES * STAGE_FLOATS * 4; }
    struct Acc { float v[MT * NT * 4]; };
    // producer cursor: the pixel this thread gathers + the (segment, tap, channel chunk) of the next chunk to issue
    struct State { int b, yq, xq, ok; int seg, tap, ta, tb, c0, issued; };

    static FFC_DEVICE void run(const Params& p, const BlockCtx& ctx, float* smem) {
        const int s = p.transposed ? p.stride : 1;
        const int cls = ctx.bz;
        const int py = cls / s, px = cls % s;
        const int Hc = (p.Ho - py + s - 1) / s, Wc = (p.Wo - px + s - 1) / s;
        const int Mc = p.B * Hc * Wc;
        const int m0 = ctx.bx * BM, n0 = ctx.by * BN;
        const ConvClassGeom g = ffc_conv_class_geom(cls, p.k, p.stride, p.pad, p.transposed);
        const int T = g.Ta * g.Tb;
        const int HWi = p.Hi * p.Wi;
        const int nchunks = (m0 < Mc) ? T * (p.cpad[0] + (p.nseg > 1 ? p.cpad[1] : 0)) / BK : 0;
        const float2* wp = p.wp + p.cls_off[cls] + n0;          // column block of this CTA

        FFC_TLS(Acc, acc);
        FFC_TLS(State, st);
        // issue the cp.async copies of the chunk under the cursor into ring slot `slot`, advance the cursor
#define FFC_V4_ISSUE(slot)                                                                                   \
        {                                                                                                    \
            float* as_ = smem + (slot) * STAGE_FLOATS;                                                       \
            float2* bs_ = reinterpret_cast<float2*>(as_ + BK * AS);                                          \
            const float* FFC_RESTRICT xs_ = st.seg ? p.x[1] : p.x[0];                                        \
            const int cin_ = st.seg ? p.cin[1] : p.cin[0];                                                   \
            int iy_, ix_;                                                                                    \
            if (p.transposed) { iy_ = st.yq + g.qy - st.ta; ix_ = st.xq + g.qx - st.tb; }                    \
            else { iy_ = st.yq * p.stride - p.pad + st.ta; ix_ = st.xq * p.stride - p.pad + st.tb; }         \
            const bool okp_ = st.ok && iy_ >= 0 && iy_ < p.Hi && ix_ >= 0 && ix_ < p.Wi;                     \
            const int ca_ = st.c0 + tid / BM;                                                                \
            const float* xp_ = okp_ ? xs_ + ((st.b * cin_ + ca_) * HWi + iy_ * p.Wi + ix_) : xs_;            \
            const int xstep_ = okp_ ? (kThreads / BM) * HWi : 0;                                             \
            float* ad_ = as_ + (tid / BM) * AS + tid % BM;                                                   \
            FFC_UNROLL                                                                                       \
            for (int i = 0; i < A_PER; ++i)                                                                  \
                ffc_cp_async4(ad_ + i * (kThreads / BM) * AS, xp_ + i * xstep_, okp_ && ca_ + i * (kThreads / BM) < cin_); \
            const float2* wsrc_ = wp + (size_t)st.issued * BK * p.npad;                                      \
            FFC_UNROLL                                                                                       \
            for (int i = 0; i < B_PER; ++i) {                                                                \
                const int pc_ = tid + i * kThreads, row_ = pc_ / (BN / 2), c16_ = pc_ % (BN / 2);            \
                ffc_cp_async16(bs_ + row_ * BS + c16_ * 2, wsrc_ + (size_t)row_ * p.npad + c16_ * 2);        \
            }                                                                                                \
            ++st.issued;                                                                                     \
            st.c0 += BK;                                                                                     \
            if (st.c0 >= (st.seg ? p.cpad[1] : p.cpad[0])) {                                                 \
                st.c0 = 0;                                                                                   \
                if (++st.tb == g.Tb) { st.tb = 0; ++st.ta; }                                                 \
                if (++st.tap == T) { st.tap = 0; st.ta = 0; st.tb = 0; ++st.seg; }                           \
            }                                                                                                \
        }

        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            FFC_TLS_REF(State, st);
            FFC_UNROLL
            for (int i = 0; i < MT * NT * 4; ++i) acc.v[i] = 0.f;
            const int m = m0 + tid % BM;
            st.ok = m < Mc;
            st.xq = m % Wc; st.yq = (m / Wc) % Hc; st.b = m / (Wc * Hc);
            st.seg = 0; st.tap = 0; st.ta = 0; st.tb = 0; st.c0 = 0; st.issued = 0;
            FFC_UNROLL
            for (int pre = 0; pre < STAGES - 1; ++pre) {
                if (pre < nchunks) FFC_V4_ISSUE(pre);
                ffc_cp_async_commit();
            }
        }
        for (int c = 0; c < nchunks; ++c) {
            FFC_PHASE { ffc_cp_async_wait<STAGES - 2>(); } FFC_SYNC;       // chunk c has landed; ring slot (c-1)%3 is free
            FFC_PHASE {
                FFC_TLS_REF(Acc, acc);
                FFC_TLS_REF(State, st);
                if (c + STAGES - 1 < nchunks) FFC_V4_ISSUE((c + STAGES - 1) % STAGES);
                ffc_cp_async_commit();
                const float* as = smem + (c % STAGES) * STAGE_FLOATS;
                const float2* bs = reinterpret_cast<const float2*>(as + BK * AS);
                const int lane = tid & 31, warp = tid >> 5;
                const int gq = lane >> 2, t = lane & 3;
                const int mw = (warp & 3) * 32, nw = (warp >> 2) * (BN / 2);
#ifdef FFC_EMU
                for (int mt = 0; mt < MT; ++mt)
                    for (int nt_ = 0; nt_ < NT; ++nt_)
                        for (int r = 0; r < 4; ++r) {
                            const int row = mw + 16 * mt + gq + (r >> 1) * 8, col = nw + 8 * nt_ + 2 * t + (r & 1);
                            float sum = 0.f;
                            for (int kl = 0; kl < BK; ++kl) {
                                const float2 b = bs[kl * BS + col];
                                sum = fmaf(as[kl * AS + row], b.x + b.y, sum);
                            }
                            acc.v[(mt * NT + nt_) * 4 + r] += sum;
                        }
#else
                float tmp[MT * NT * 4];
                FFC_UNROLL
                for (int i = 0; i < MT * NT * 4; ++i) tmp[i] = 0.f;
                FFC_UNROLL
                for (int k8 = 0; k8 < BK; k8 += 8) {
                    unsigned ah[MT][4], al[MT][4], bh[NT][2], bl[NT][2];
                    FFC_UNROLL
                    for (int mt = 0; mt < MT; ++mt) {
                        FFC_UNROLL
                        for (int r = 0; r < 4; ++r) {
                            const float x = as[(k8 + t + (r >> 1) * 4) * AS + mw + 16 * mt + gq + (r & 1) * 8];
                            const float hi = ffc_tf32_hi(x);
                            ah[mt][r] = __float_as_uint(hi);             // (rounded to nearest: not what the tensor core would truncate x to)
                            al[mt][r] = __float_as_uint(x - hi);
                        }
                    }
                    FFC_UNROLL
                    for (int nt_ = 0; nt_ < NT; ++nt_) {
                        FFC_UNROLL
                        for (int r = 0; r < 2; ++r) {
                            const float2 x = bs[(k8 + t + r * 4) * BS + nw + 8 * nt_ + gq];
                            bh[nt_][r] = __float_as_uint(x.x);
                            bl[nt_][r] = __float_as_uint(x.y);
                        }
                    }
                    FFC_UNROLL
                    for (int mt = 0; mt < MT; ++mt)
                        FFC_UNROLL
                        for (int nt_ = 0; nt_ < NT; ++nt_) {
                            float* cc = tmp + (mt * NT + nt_) * 4;
                            ffc_mma_tf32_v4(cc, al[mt], bh[nt_]);
                            ffc_mma_tf32_v4(cc, ah[mt], bl[nt_]);
                            ffc_mma_tf32_v4(cc, ah[mt], bh[nt_]);
                        }
                }
                FFC_UNROLL
                for (int i = 0; i < MT * NT * 4; ++i) acc.v[i] += tmp[i];
#endif
            }      // no barrier here: the next iteration's wait + barrier orders ring-slot reuse
        }
#undef FFC_V4_ISSUE
        FFC_PHASE { ffc_cp_async_wait<0>(); } FFC_SYNC;
        FFC_PHASE {
            FFC_TLS_REF(Acc, acc);
            const int lane = tid & 31, warp = tid >> 5;
            const int gq = lane >> 2, t = lane & 3;
            const int mw = (warp & 3) * 32, nw = (warp >> 2) * (BN / 2);
            FFC_UNROLL
            for (int mt = 0; mt < MT; ++mt) {
                FFC_UNROLL
                for (int h = 0; h < 2; ++h) {
                    const int m = m0 + mw + 16 * mt + gq + 8 * h;
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
// host side
// ---------------------------------------------------------------------------------------------
static const int kV4BK = 16;
static const int ffc_conv_v4_mode = 3;      // value of ffc_debug_conv_reference() that selects ConvFwdV4
static const int ffc_conv_auto_mode = 5;    // default: V5 or V4 by output width
static const int ffc_conv_v5_mode = 4;      // ... ConvFwdV5 (tcgen05; device build only -- the emulation build runs V4 instead)
#ifndef FFC_EMU
bool conv_small_supported(int cin0, int cin1, int cout, int k);
bool conv1x1_narrow_supported(int cin0, int cin1, int cout, int k, int stride, int pad, int Hi, int Wi, int Ho, int Wo,
                              const void* x, const void* y, const void* addend);
int conv1x1_narrow_run(const float* x, const float* w, int cin, const float* bias, const float* addend, float* y,
                       int B, int cout, int Hi, int Wi, int transposed, ffc_stream_t st);
int conv_small_run(const float* x0, const float* w0, int cin0, const float* x1, const float* w1, int cin1,
                   const float* bias, const float* addend, float* y, int B, int cout, int Hi, int Wi, int Ho, int Wo,
                   int k, int stride, int pad, int transposed, ffc_stream_t st);
size_t conv_v5_workspace_bytes(int cin0, int cin1, int cout, int k, int stride, int pad, int transposed);
int conv_v5_run(const float* x0, const float* w0, int cin0, const float* x1, const float* w1, int cin1,
                const float* bias, const float* addend, float* y, int B, int cout, int Hi, int Wi, int Ho, int Wo,
                int k, int stride, int pad, int transposed, void* workspace, size_t workspace_bytes, ffc_stream_t st);
int conv_v5_run_block(const float* x0, const float* w0, const float* w0b, int cin0, const float* x1, const float* w1, int cin1,
                      const float* bias, const float* addend, float* y, float* y1, int cout0, int B, int cout, int Hi, int Wi, int Ho, int Wo,
                      int k, int stride, int pad, int transposed, float slope, void* workspace, size_t workspace_bytes, ffc_stream_t st);
#endif

static int conv_v4_bn(int cout) { return (cout <= 32 || (cout > 64 && cout % 64 != 0 && cout % 64 <= 32 && cout < 128)) ? 32 : 64; }

struct ConvV4Plan { int bn, npad, cpad[2], ncls, cls_off[FFC_V4_MAXCLS]; size_t wp_float2; };
static ConvV4Plan conv_v4_plan(int cin0, int cin1, int cout, int k, int stride, int pad, int transposed) {
    ConvV4Plan pl;
    pl.bn = conv_v4_bn(cout);
    pl.npad = ffc_cdiv(cout, pl.bn) * pl.bn;
    pl.cpad[0] = ffc_cdiv(cin0, kV4BK) * kV4BK;
    pl.cpad[1] = cin1 ? ffc_cdiv(cin1, kV4BK) * kV4BK : 0;
    pl.ncls = transposed ? stride * stride : 1;
    size_t off = 0;
    for (int c = 0; c < FFC_V4_MAXCLS; ++c) {
        pl.cls_off[c] = (int)off;
        if (c < pl.ncls) {
            const ConvClassGeom g = ffc_conv_class_geom(c, k, stride, pad, transposed);
            off += (size_t)g.Ta * g.Tb * (pl.cpad[0] + pl.cpad[1]) * pl.npad;
        }
    }
    pl.wp_float2 = off;
    return pl;
}

extern "C" size_t ffc_conv2d_workspace_bytes(int cin0, int cin1, int cout, int k, int stride, int pad, int transposed) {
    if (cin0 <= 0 || cout <= 0 || k < 1 || stride < 1) return 0;
    size_t n = conv_v4_plan(cin0, cin1, cout, k, stride, pad, transposed).wp_float2 * sizeof(float2) + 256;
#ifndef FFC_EMU
    const size_t n5 = conv_v5_workspace_bytes(cin0, cin1, cout, k, stride, pad, transposed);
    if (n5 > n) n = n5;
#endif
    return n;
}

template <int BN>
static int conv_v4_launch(const ConvV4Params& p, ffc_stream_t st) {
    typedef ConvFwdV4<BN, kV4BK> K;
    const int s = p.transposed ? p.stride : 1;
    const int Mc = p.B * ffc_cdiv(p.Ho, s) * ffc_cdiv(p.Wo, s);
    return ffc_launch<K>(ffc_cdiv(Mc, K::BM), p.npad / BN, s * s, K::kThreads, K::smem_bytes(), st, p);
}

// Same contract as ffc_conv2d_fwd plus a caller-provided workspace of ffc_conv2d_workspace_bytes(...) bytes that
// receives the packed (hi, lo) weights of this call.
extern "C" int ffc_conv2d_fwd_ws(const float* x0, const float* w0, int cin0,
                                 const float* x1, const float* w1, int cin1,
                                 const float* bias, const float* addend, float* y,
                                 int B, int cout, int Hi, int Wi, int Ho, int Wo,
                                 int k, int stride, int pad, int transposed,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    int mode = ffc_conv_use_reference_kernel;
#ifndef FFC_EMU
    if (mode == ffc_conv_auto_mode && conv_small_supported(cin0, cin1, cout, k)) {
        // <= 4 channels on one side (the RGB output layer and its data gradient): direct FP32 kernel
        FFC_REQUIRE(x0 && w0 && y && cin0 > 0 && B >= 0 && cout > 0, "ffc_conv2d_fwd_ws: null pointer / bad sizes");
        FFC_REQUIRE((x1 == nullptr) == (cin1 == 0) && (x1 == nullptr) == (w1 == nullptr), "ffc_conv2d_fwd_ws: inconsistent second segment");
        FFC_REQUIRE(stride == 1 || stride == 2, "ffc_conv2d_fwd_ws: unsupported stride %d", stride);
        if (B == 0) return FFC_OK;
        return conv_small_run(x0, w0, cin0, x1, w1, cin1, bias, addend, y, B, cout, Hi, Wi, Ho, Wo, k, stride, pad, transposed, (ffc_stream_t)stream);
    }
    if (mode == ffc_conv_auto_mode && x0 && w0 && y && B > 0 &&
        conv1x1_narrow_supported(cin0, cin1, cout, k, stride, pad, Hi, Wi, Ho, Wo, x0, y, addend)) {
        // narrow 1x1 (SpectralTransform conv1 / conv2 and their data gradients): bandwidth-bound direct kernel
        FFC_REQUIRE((long long)B * cout * Ho * Wo < (1LL << 31) && (long long)B * cin0 * Hi * Wi < (1LL << 31), "ffc_conv2d_fwd_ws: tensor too large for 32-bit element offsets");
        return conv1x1_narrow_run(x0, w0, cin0, bias, addend, y, B, cout, Hi, Wi, transposed, (ffc_stream_t)stream);
    }
#endif
    if (mode == ffc_conv_auto_mode) {
        // tcgen05 kernel wherever its 128 x N tile is reasonably filled; the mma.sync kernel for narrow outputs
#ifndef FFC_EMU
        mode = (cout >= 24 || (cout >= 16 && k >= 3)) ? ffc_conv_v5_mode : ffc_conv_v4_mode;
#else
        mode = ffc_conv_v4_mode;
#endif
    }
    if (mode != ffc_conv_v4_mode && mode != ffc_conv_v5_mode)
        return ffc_conv2d_fwd(x0, w0, cin0, x1, w1, cin1, bias, addend, y, B, cout, Hi, Wi, Ho, Wo, k, stride, pad, transposed, stream);   // no workspace needed
    FFC_REQUIRE(x0 && w0 && y && cin0 > 0, "ffc_conv2d_fwd_ws: null pointer / empty first segment");
    FFC_REQUIRE((x1 == nullptr) == (cin1 == 0) && (x1 == nullptr) == (w1 == nullptr), "ffc_conv2d_fwd_ws: inconsistent second segment");
    FFC_REQUIRE(k >= 1 && k <= 7 && (stride == 1 || stride == 2) && pad >= 0 && pad < 8, "ffc_conv2d_fwd_ws: unsupported k=%d stride=%d pad=%d", k, stride, pad);
    FFC_REQUIRE(B >= 0 && cout > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "ffc_conv2d_fwd_ws: bad sizes");
    if (!transposed) {
        FFC_REQUIRE(Ho == (Hi + 2 * pad - k) / stride + 1 && Wo == (Wi + 2 * pad - k) / stride + 1, "ffc_conv2d_fwd_ws: inconsistent output size");
    } else {
        const int hmin = (Hi - 1) * stride - 2 * pad + k, wmin = (Wi - 1) * stride - 2 * pad + k;
        FFC_REQUIRE(Ho >= hmin && Ho < hmin + stride && Wo >= wmin && Wo < wmin + stride, "ffc_conv2d_fwd_ws: inconsistent transposed output size");
    }
    if (B == 0) return FFC_OK;
    const int cmax = cin0 > cin1 ? cin0 : cin1;
    FFC_REQUIRE((long long)B * cout * Ho * Wo < (1LL << 31) && (long long)B * cmax * Hi * Wi < (1LL << 31), "ffc_conv2d_fwd_ws: tensor too large for 32-bit element offsets");
#ifndef FFC_EMU
    if (mode == ffc_conv_v5_mode)
        return conv_v5_run(x0, w0, cin0, x1, w1, cin1, bias, addend, y, B, cout, Hi, Wi, Ho, Wo, k, stride, pad, transposed,
                           workspace, workspace_bytes, (ffc_stream_t)stream);
#endif
    const ConvV4Plan pl = conv_v4_plan(cin0, cin1, cout, k, stride, pad, transposed);
    FFC_REQUIRE(pl.wp_float2 < (1ULL << 31), "ffc_conv2d_fwd_ws: packed weights too large");
    const uintptr_t wsa = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    if (!workspace || wsa + pl.wp_float2 * sizeof(float2) > (uintptr_t)workspace + workspace_bytes) {
        ffc_set_error("ffc_conv2d_fwd_ws: workspace too small (%zu bytes needed)", pl.wp_float2 * sizeof(float2) + 256);
        return FFC_ERR_WORKSPACE;
    }
    ffc_stream_t st = (ffc_stream_t)stream;
    PackWParams pp;
    pp.w[0] = w0; pp.w[1] = w1; pp.cin[0] = cin0; pp.cin[1] = cin1; pp.cpad[0] = pl.cpad[0]; pp.cpad[1] = pl.cpad[1];
    pp.nseg = x1 ? 2 : 1; pp.wp = (float2*)wsa; pp.npad = pl.npad; pp.cout = cout; pp.k = k; pp.stride = stride; pp.pad = pad;
    pp.transposed = transposed; pp.ncls = pl.ncls;
    for (int c = 0; c < FFC_V4_MAXCLS; ++c) pp.cls_off[c] = pl.cls_off[c];
    const long long per_cls = (long long)k * k * (pl.cpad[0] + pl.cpad[1]) * pl.npad;
    int gx = (int)((per_cls + 255) / 256); if (gx > ffc_sm_count() * 8) gx = ffc_sm_count() * 8; if (gx < 1) gx = 1;
    FFC_CHECK((ffc_launch<PackWeightsKernel>(gx, pl.ncls, 1, 256, 0, st, pp)));
    ConvV4Params p;
    p.x[0] = x0; p.x[1] = x1; p.cin[0] = cin0; p.cin[1] = cin1; p.cpad[0] = pl.cpad[0]; p.cpad[1] = pl.cpad[1]; p.nseg = x1 ? 2 : 1;
    p.wp = (const float2*)wsa; p.npad = pl.npad;
    for (int c = 0; c < FFC_V4_MAXCLS; ++c) p.cls_off[c] = pl.cls_off[c];
    p.bias = bias; p.addend = addend; p.y = y; p.B = B; p.cout = cout; p.Hi = Hi; p.Wi = Wi; p.Ho = Ho; p.Wo = Wo;
    p.k = k; p.stride = stride; p.pad = pad; p.transposed = transposed;
    return pl.bn == 32 ? conv_v4_launch<32>(p, st) : conv_v4_launch<64>(p, st);
}

// Block form of the local branches of FFC.forward (ffc.py:91-96; ffc_transpose.py:98-106):
//   y0 (cout0 channels) = conv(x0, w00) + conv(x1, w10) [+ bias[0:cout0]]          convl2l(x_l) + convg2l(x_g)
//   y1 (cout1 channels) = conv(x0, w01)                 [+ bias[cout0:cout0+cout1]]  convl2g(x_l)
// On the tcgen05 path this is ONE implicit GEMM over the concatenated output channels (the operand gathered from x0 is
// shared); otherwise it is two calls of ffc_conv2d_fwd_ws.  x1 / w10 may be null (no second segment).
extern "C" int ffc_conv2d_block_fwd_ws(const float* x0, const float* w00, const float* w01, int cin0,
                                       const float* x1, const float* w10, int cin1, const float* bias,
                                       float* y0, int cout0, float* y1, int cout1,
                                       int B, int Hi, int Wi, int Ho, int Wo, int k, int stride, int pad, int transposed,
                                       void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(x0 && w00 && w01 && y0 && y1 && cin0 > 0 && cout0 > 0 && cout1 > 0, "ffc_conv2d_block_fwd_ws: null pointer / empty block");
    FFC_REQUIRE((x1 == nullptr) == (cin1 == 0) && (x1 == nullptr) == (w10 == nullptr), "ffc_conv2d_block_fwd_ws: inconsistent second segment");
#ifndef FFC_EMU
    const int cout = cout0 + cout1;
    if ((ffc_conv_use_reference_kernel == ffc_conv_auto_mode || ffc_conv_use_reference_kernel == ffc_conv_v5_mode) && cout >= 16) {
        FFC_REQUIRE(k >= 1 && k <= 7 && (stride == 1 || stride == 2) && pad >= 0 && pad < 8, "ffc_conv2d_block_fwd_ws: unsupported k=%d stride=%d pad=%d", k, stride, pad);
        FFC_REQUIRE(B >= 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "ffc_conv2d_block_fwd_ws: bad sizes");
        if (!transposed) {
            FFC_REQUIRE(Ho == (Hi + 2 * pad - k) / stride + 1 && Wo == (Wi + 2 * pad - k) / stride + 1, "ffc_conv2d_block_fwd_ws: inconsistent output size");
        } else {
            const int hmin = (Hi - 1) * stride - 2 * pad + k, wmin = (Wi - 1) * stride - 2 * pad + k;
            FFC_REQUIRE(Ho >= hmin && Ho < hmin + stride && Wo >= wmin && Wo < wmin + stride, "ffc_conv2d_block_fwd_ws: inconsistent transposed output size");
        }
        if (B == 0) return FFC_OK;
        const int cmax = cin0 > cin1 ? cin0 : cin1;
        FFC_REQUIRE((long long)B * cout * Ho * Wo < (1LL << 31) && (long long)B * cmax * Hi * Wi < (1LL << 31), "ffc_conv2d_block_fwd_ws: tensor too large for 32-bit element offsets");
        return conv_v5_run_block(x0, w00, w01, cin0, x1, w10, cin1, bias, nullptr, y0, y1, cout0, B, cout, Hi, Wi, Ho, Wo,
                                 k, stride, pad, transposed, 1.f, workspace, workspace_bytes, (ffc_stream_t)stream);
    }
#endif
    FFC_CHECK(ffc_conv2d_fwd_ws(x0, w00, cin0, x1, w10, cin1, bias, nullptr, y0, B, cout0, Hi, Wi, Ho, Wo, k, stride, pad, transposed,
                                workspace, workspace_bytes, stream));
    return ffc_conv2d_fwd_ws(x0, w01, cin0, nullptr, nullptr, 0, bias ? bias + cout0 : nullptr, nullptr, y1, B, cout1, Hi, Wi, Ho, Wo,
                             k, stride, pad, transposed, workspace, workspace_bytes, stream);
}

extern "C" int ffc_bn_act_fwd(const float* x, float* y, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                              int B, int C, int HW, int norm, int training, float eps, float momentum,
                              int act, float slope, void* workspace, size_t workspace_bytes, void* stream);    // ffc_bnact.cu

// y = act(conv(x, w) + bias), act = LeakyReLU(slope) or ReLU: one stage of the SN conv discriminators of the fgan scripts
// (fgan_complete.py:162-168: ``self.act(self.convN(m))``).  The activation rides in the epilogue of the tcgen05
// kernel; shapes that kernel does not take (<= 4 input channels: the RGB stage) run the convolution and then the
// elementwise kernel in place.  workspace: ffc_conv2d_workspace_bytes(cin, 0, cout, k, stride, pad, 0), at least 2*cout*8 bytes.
extern "C" int ffc_conv2d_act_fwd_ws(const float* x, const float* w, int cin, const float* bias, float* y,
                                     int B, int cout, int Hi, int Wi, int Ho, int Wo, int k, int stride, int pad,
                                     int act, float slope, void* workspace, size_t workspace_bytes, void* stream) {
    FFC_REQUIRE(act == FFC_ACT_LEAKY || act == FFC_ACT_RELU, "ffc_conv2d_act_fwd_ws: LeakyReLU or ReLU only (code %d)", act);
    FFC_REQUIRE(act == FFC_ACT_RELU || slope > 0.f, "ffc_conv2d_act_fwd_ws: LeakyReLU needs slope > 0");
#ifndef FFC_EMU
    if (ffc_conv_use_reference_kernel == ffc_conv_auto_mode && !conv_small_supported(cin, 0, cout, k) && (cout >= 24 || (cout >= 16 && k >= 3))) {
        FFC_REQUIRE(x && w && y && cin > 0, "ffc_conv2d_act_fwd_ws: null pointer / empty input");
        FFC_REQUIRE(k >= 1 && k <= 7 && (stride == 1 || stride == 2) && pad >= 0 && pad < 8, "ffc_conv2d_act_fwd_ws: unsupported k=%d stride=%d pad=%d", k, stride, pad);
        FFC_REQUIRE(B >= 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "ffc_conv2d_act_fwd_ws: bad sizes");
        FFC_REQUIRE(Ho == (Hi + 2 * pad - k) / stride + 1 && Wo == (Wi + 2 * pad - k) / stride + 1, "ffc_conv2d_act_fwd_ws: inconsistent output size");
        if (B == 0) return FFC_OK;
        FFC_REQUIRE((long long)B * cout * Ho * Wo < (1LL << 31) && (long long)B * cin * Hi * Wi < (1LL << 31), "ffc_conv2d_act_fwd_ws: tensor too large for 32-bit element offsets");
        return conv_v5_run_block(x, w, nullptr, cin, nullptr, nullptr, 0, bias, nullptr, y, nullptr, cout, B, cout, Hi, Wi, Ho, Wo,
                                 k, stride, pad, 0, act == FFC_ACT_RELU ? 0.f : slope, workspace, workspace_bytes, (ffc_stream_t)stream);
    }
#endif
    FFC_CHECK(ffc_conv2d_fwd_ws(x, w, cin, nullptr, nullptr, 0, bias, nullptr, y, B, cout, Hi, Wi, Ho, Wo, k, stride, pad, 0,
                                workspace, workspace_bytes, stream));
    return ffc_bn_act_fwd(y, y, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, B, cout, Ho * Wo, 0, 0, 0.f, 0.f, act, slope,
                          workspace, workspace_bytes, stream);
}
