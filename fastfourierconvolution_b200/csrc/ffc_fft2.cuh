// Shared-memory 2-D real FFT building blocks over tiles of P planes (H x W, powers of two 4..128).
//
// Region layout (floats), per plane pl of the tile:
//   "real layout":  rows of RS = W + 4 floats            (float4 reads by a thread-per-row are
//                                                          bank-conflict free: (W/4 + 1) is odd)
//   "spec layout":  [H][Wf] float2, Wf = W/2 + 1 (odd -> thread-per-row / thread-per-column
//                                                          float2 accesses are conflict free)
//   "pair scratch": [H/2][W] float2 (only when W >= 64)
// Every region is REGION = H * RS floats per plane, which bounds all three layouts.
//
// Row transforms pack real rows r and r + H/2 into one complex FFT.  Frequencies along a
// two-level dimension (64, 128) are kept in the permuted order of FftSplit<N>::pos(); the
// channel mix / BatchNorm that sit between the forward and the inverse transform are
// pointwise in (u, v), so the permutation along u never has to be undone.
#pragma once
#include "ffc_fft.cuh"

// Where plane `pl` of a work list lives inside a spectrum region: identity by default; for the fused
// Fourier unit the region holds images of `cb` planes each of which only the first `cn` are transformed.
struct PlaneMap {
    int cn, cb;
    FFC_HDM int operator()(int pl) const { return cn == 0 ? pl : (pl / cn) * cb + (pl % cn); }
};
FFC_HD PlaneMap ffc_identity_map() { PlaneMap m; m.cn = 0; m.cb = 0; return m; }

template <int H, int W>
struct Fft2Plan {
    static constexpr int Wf = W / 2 + 1;
    static constexpr int RS = W + 4;
    static constexpr int REGION = H * RS;              // floats per plane per region
    static constexpr int W1 = FftSplit<W>::N1, W2 = FftSplit<W>::N2;
    static constexpr int H1 = FftSplit<H>::N1, H2 = FftSplit<H>::N2;
    static constexpr bool kRowsTwoLevel = (W2 > 1);
    // work items per plane of the widest phase (used by the host to size the CTA)
    static constexpr int kRowItems = (H / 2) * W2;
    static constexpr int kColItems = Wf * (H2 > 1 ? H1 : 1);
    static constexpr int kMaxItems = kRowItems > kColItems ? kRowItems : kColItems;
    static_assert(H * Wf * 2 <= REGION, "spec layout must fit a region");
    static_assert((H / 2) * W * 2 <= REGION, "pair scratch must fit a region");
};

// ---------------------------------------------------------------------------------------------
// forward rows: real layout (src) -> spec layout.  One-level: src -> dst directly.
// Two-level: src -> scratch(dst region) -> in place -> spec layout written back into src region.
// Returns (via the plan) where the spectrum lives: dst if one-level, src if two-level.
// The functions below are single phases; callers put FFC_SYNC between them.
// ---------------------------------------------------------------------------------------------
// SPS = float2 per spectrum row: Wf (dense spec layout) or RS/2 (spectrum row u overlays real row u, which
// lets the row transforms run IN PLACE: a thread overwrites exactly the two rows it has read).
template <int H, int W, int SPS = W / 2 + 1>
FFC_DEVICE void fft2_rows_fwd_L1(int tid, int nt, int np, const float* src, float* dst, PlaneMap dmap = ffc_identity_map(),
                                 PlaneMap smap = ffc_identity_map()) {
    typedef Fft2Plan<H, W> PL;
    for (int it = tid; it < np * (H / 2); it += nt) {
        const int pl = it / (H / 2), r = it % (H / 2);
        const float* ra = src + smap(pl) * PL::REGION + r * PL::RS;
        const float* rb = ra + (H / 2) * PL::RS;
        float2 z[W];
        FFC_UNROLL
        for (int j = 0; j < W / 4; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(ra + 4 * j);
            const float4 b = *reinterpret_cast<const float4*>(rb + 4 * j);
            z[4 * j + 0] = make_float2(a.x, b.x);
            z[4 * j + 1] = make_float2(a.y, b.y);
            z[4 * j + 2] = make_float2(a.z, b.z);
            z[4 * j + 3] = make_float2(a.w, b.w);
        }
        ffc_fft_regs<W, -1>(z);
        float2* sa = reinterpret_cast<float2*>(dst + dmap(pl) * PL::REGION) + r * SPS;
        float2* sb = sa + (H / 2) * SPS;
        FFC_UNROLL
        for (int v = 0; v < PL::Wf; ++v) {
            const float2 zv = z[v], zc = z[(W - v) % W];
            sa[v] = make_float2(0.5f * (zv.x + zc.x), 0.5f * (zv.y - zc.y));
            sb[v] = make_float2(0.5f * (zv.y + zc.y), 0.5f * (zc.x - zv.x));
        }
    }
}

// two-level rows, phase A: real rows (src) -> pair scratch (dst region), first level + twiddle
template <int H, int W>
FFC_DEVICE void fft2_rows_fwd_L2a(int tid, int nt, int np, const float* src, float* dst, const float2* tw) {
    typedef Fft2Plan<H, W> PL;
    constexpr int W1 = PL::W1, W2 = PL::W2;
    for (int it = tid; it < np * (H / 2) * W2; it += nt) {
        const int n2 = it % W2, r = (it / W2) % (H / 2), pl = it / (W2 * (H / 2));
        const float* ra = src + pl * PL::REGION + r * PL::RS;
        const float* rb = ra + (H / 2) * PL::RS;
        float2 z[W1];
        FFC_UNROLL
        for (int n1 = 0; n1 < W1; ++n1) z[n1] = make_float2(ra[W2 * n1 + n2], rb[W2 * n1 + n2]);
        ffc_fft_regs<W1, -1>(z);
        float2* sc = reinterpret_cast<float2*>(dst + pl * PL::REGION) + r * W;
        FFC_UNROLL
        for (int k1 = 0; k1 < W1; ++k1) {
            const float2 w = tw[(n2 * k1) * (FFC_TW_N / W)];
            sc[W2 * k1 + n2] = ffc_cmul_tw<-1>(z[k1], w);
        }
    }
}
// two-level, phase B (forward or inverse second/first level over the W2 contiguous entries)
template <int H, int W, int SIGN, bool TWIDDLE_AFTER>
FFC_DEVICE void fft2_rows_L2b(int tid, int nt, int np, float* scr, const float2* tw) {
    typedef Fft2Plan<H, W> PL;
    constexpr int W1 = PL::W1, W2 = PL::W2;
    for (int it = tid; it < np * (H / 2) * W1; it += nt) {
        const int k1 = it % W1, r = (it / W1) % (H / 2), pl = it / (W1 * (H / 2));
        float2* sc = reinterpret_cast<float2*>(scr + pl * PL::REGION) + r * W + W2 * k1;
        float2 z[W2];
        FFC_UNROLL
        for (int i = 0; i < W2; ++i) z[i] = sc[i];
        ffc_fft_regs<W2, SIGN>(z);
        FFC_UNROLL
        for (int i = 0; i < W2; ++i) {
            if (TWIDDLE_AFTER) sc[i] = ffc_cmul_tw<SIGN>(z[i], tw[(i * k1) * (FFC_TW_N / W)]);
            else sc[i] = z[i];
        }
    }
}
// two-level, separation: pair scratch -> spec layout (dst region must not alias scr)
template <int H, int W>
FFC_DEVICE void fft2_rows_fwd_L2sep(int tid, int nt, int np, const float* scr, float* dst) {
    typedef Fft2Plan<H, W> PL;
    for (int it = tid; it < np * (H / 2) * PL::Wf; it += nt) {
        const int v = it % PL::Wf, r = (it / PL::Wf) % (H / 2), pl = it / (PL::Wf * (H / 2));
        const float2* sc = reinterpret_cast<const float2*>(scr + pl * PL::REGION) + r * W;
        const float2 zv = sc[FftSplit<W>::pos(v)], zc = sc[FftSplit<W>::pos((W - v) % W)];
        float2* sa = reinterpret_cast<float2*>(dst + pl * PL::REGION) + r * PL::Wf;
        float2* sb = sa + (H / 2) * PL::Wf;
        sa[v] = make_float2(0.5f * (zv.x + zc.x), 0.5f * (zv.y - zc.y));
        sb[v] = make_float2(0.5f * (zv.y + zc.y), 0.5f * (zc.x - zv.x));
    }
}

// ---------------------------------------------------------------------------------------------
// columns (complex, in place in a spec-layout region), element stride Wf
// ---------------------------------------------------------------------------------------------
template <int H, int W, int SIGN, int SPS = W / 2 + 1>
FFC_DEVICE void fft2_cols_L1(int tid, int nt, int np, float* spec, PlaneMap map = ffc_identity_map()) {
    typedef Fft2Plan<H, W> PL;
    for (int it = tid; it < np * PL::Wf; it += nt) {
        const int v = it % PL::Wf, pl = it / PL::Wf;
        float2* col = reinterpret_cast<float2*>(spec + map(pl) * PL::REGION) + v;
        float2 c[H];
        FFC_UNROLL
        for (int u = 0; u < H; ++u) c[u] = col[u * SPS];
        ffc_fft_regs<H, SIGN>(c);
        FFC_UNROLL
        for (int u = 0; u < H; ++u) col[u * SPS] = c[u];
    }
}
// strided level: gathers positions H2*i + n2 (i < H1).  Forward: first level (twiddle after).
// Inverse: second level (no twiddle), output natural time order.
template <int H, int W, int SIGN, bool TWIDDLE_AFTER>
FFC_DEVICE void fft2_cols_L2_strided(int tid, int nt, int np, float* spec, const float2* tw) {
    typedef Fft2Plan<H, W> PL;
    constexpr int H1 = PL::H1, H2 = PL::H2;
    for (int it = tid; it < np * PL::Wf * H2; it += nt) {
        const int v = it % PL::Wf, n2 = (it / PL::Wf) % H2, pl = it / (PL::Wf * H2);
        float2* col = reinterpret_cast<float2*>(spec + pl * PL::REGION) + v;
        float2 c[H1];
        FFC_UNROLL
        for (int i = 0; i < H1; ++i) c[i] = col[(H2 * i + n2) * PL::Wf];
        ffc_fft_regs<H1, SIGN>(c);
        FFC_UNROLL
        for (int i = 0; i < H1; ++i) {
            if (TWIDDLE_AFTER) col[(H2 * i + n2) * PL::Wf] = ffc_cmul_tw<SIGN>(c[i], tw[(n2 * i) * (FFC_TW_N / H)]);
            else col[(H2 * i + n2) * PL::Wf] = c[i];
        }
    }
}
// contiguous level: positions H2*k1 + i (i < H2).  Forward: second level (no twiddle).
// Inverse: first level (twiddle after).
template <int H, int W, int SIGN, bool TWIDDLE_AFTER>
FFC_DEVICE void fft2_cols_L2_contig(int tid, int nt, int np, float* spec, const float2* tw) {
    typedef Fft2Plan<H, W> PL;
    constexpr int H1 = PL::H1, H2 = PL::H2;
    for (int it = tid; it < np * PL::Wf * H1; it += nt) {
        const int v = it % PL::Wf, k1 = (it / PL::Wf) % H1, pl = it / (PL::Wf * H1);
        float2* col = reinterpret_cast<float2*>(spec + pl * PL::REGION) + v;
        float2 c[H2];
        FFC_UNROLL
        for (int i = 0; i < H2; ++i) c[i] = col[(H2 * k1 + i) * PL::Wf];
        ffc_fft_regs<H2, SIGN>(c);
        FFC_UNROLL
        for (int i = 0; i < H2; ++i) {
            if (TWIDDLE_AFTER) col[(H2 * k1 + i) * PL::Wf] = ffc_cmul_tw<SIGN>(c[i], tw[(i * k1) * (FFC_TW_N / H)]);
            else col[(H2 * k1 + i) * PL::Wf] = c[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// inverse rows: spec layout -> real layout, c2r semantics of torch.fft.irfftn's last dimension
// (imaginary parts of bins 0 and W/2 are ignored, interior bins count twice).
// ---------------------------------------------------------------------------------------------
template <int H, int W, int SPS = W / 2 + 1>
FFC_DEVICE void fft2_rows_inv_L1(int tid, int nt, int np, const float* spec, float* dst, const float scale, PlaneMap smap = ffc_identity_map(),
                                 PlaneMap dmap = ffc_identity_map()) {
    typedef Fft2Plan<H, W> PL;
    for (int it = tid; it < np * (H / 2); it += nt) {
        const int pl = it / (H / 2), r = it % (H / 2);
        const float2* sa = reinterpret_cast<const float2*>(spec + smap(pl) * PL::REGION) + r * SPS;
        const float2* sb = sa + (H / 2) * SPS;
        float2 z[W];
        FFC_UNROLL
        for (int v = 0; v <= W / 2; ++v) {
            const float2 a = sa[v], b = sb[v];
            if (v == 0 || v == W / 2) {
                z[v] = make_float2(a.x, b.x);
            } else {
                z[v] = make_float2(a.x - b.y, a.y + b.x);
                z[W - v] = make_float2(a.x + b.y, b.x - a.y);
            }
        }
        ffc_fft_regs<W, +1>(z);
        float* ra = dst + dmap(pl) * PL::REGION + r * PL::RS;
        float* rb = ra + (H / 2) * PL::RS;
        FFC_UNROLL
        for (int j = 0; j < W / 4; ++j) {
            *reinterpret_cast<float4*>(ra + 4 * j) = make_float4(z[4 * j].x * scale, z[4 * j + 1].x * scale, z[4 * j + 2].x * scale, z[4 * j + 3].x * scale);
            *reinterpret_cast<float4*>(rb + 4 * j) = make_float4(z[4 * j].y * scale, z[4 * j + 1].y * scale, z[4 * j + 2].y * scale, z[4 * j + 3].y * scale);
        }
    }
}
// two-level inverse rows, build: spec layout -> pair scratch in permuted frequency order
template <int H, int W>
FFC_DEVICE void fft2_rows_inv_L2build(int tid, int nt, int np, const float* spec, float* scr) {
    typedef Fft2Plan<H, W> PL;
    for (int it = tid; it < np * (H / 2) * PL::Wf; it += nt) {
        const int v = it % PL::Wf, r = (it / PL::Wf) % (H / 2), pl = it / (PL::Wf * (H / 2));
        const float2* sa = reinterpret_cast<const float2*>(spec + pl * PL::REGION) + r * PL::Wf;
        const float2* sb = sa + (H / 2) * PL::Wf;
        float2* sc = reinterpret_cast<float2*>(scr + pl * PL::REGION) + r * W;
        const float2 a = sa[v], b = sb[v];
        if (v == 0 || v == W / 2) {
            sc[FftSplit<W>::pos(v)] = make_float2(a.x, b.x);
        } else {
            sc[FftSplit<W>::pos(v)] = make_float2(a.x - b.y, a.y + b.x);
            sc[FftSplit<W>::pos(W - v)] = make_float2(a.x + b.y, b.x - a.y);
        }
    }
}
// two-level inverse rows, last level: pair scratch (strided gather) -> real rows
template <int H, int W>
FFC_DEVICE void fft2_rows_inv_L2a(int tid, int nt, int np, const float* scr, float* dst, float scale) {
    typedef Fft2Plan<H, W> PL;
    constexpr int W1 = PL::W1, W2 = PL::W2;
    for (int it = tid; it < np * (H / 2) * W2; it += nt) {
        const int n2 = it % W2, r = (it / W2) % (H / 2), pl = it / (W2 * (H / 2));
        const float2* sc = reinterpret_cast<const float2*>(scr + pl * PL::REGION) + r * W;
        float2 z[W1];
        FFC_UNROLL
        for (int k1 = 0; k1 < W1; ++k1) z[k1] = sc[W2 * k1 + n2];
        ffc_fft_regs<W1, +1>(z);
        float* ra = dst + pl * PL::REGION + r * PL::RS;
        float* rb = ra + (H / 2) * PL::RS;
        FFC_UNROLL
        for (int n1 = 0; n1 < W1; ++n1) {
            ra[W2 * n1 + n2] = z[n1].x * scale;
            rb[W2 * n1 + n2] = z[n1].y * scale;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// composite helpers: whole forward / inverse 2-D transform of a tile, barriers included.
// Forward: real layout in region A  -> spectrum in `*spec_region` (B if one-level rows, A else).
// Inverse: spectrum in region S (spec layout), other region O -> real layout; the result region
//          is O for one-level rows and S for two-level rows (returned in *real_region).
// These are macros rather than functions because FFC_PHASE needs `ctx` and expands to a loop.
// ---------------------------------------------------------------------------------------------
#define FFC_FFT2_FORWARD(H, W, np, A, B, tw, spec_out)                                            \
    do {                                                                                          \
        typedef Fft2Plan<H, W> PL_;                                                               \
        if constexpr (!PL_::kRowsTwoLevel) {                                                                \
            FFC_PHASE { fft2_rows_fwd_L1<H, W>(tid, ctx.nt, np, A, B); } FFC_SYNC;                \
            spec_out = B;                                                                         \
        } else {                                                                                  \
            FFC_PHASE { fft2_rows_fwd_L2a<H, W>(tid, ctx.nt, np, A, B, tw); } FFC_SYNC;           \
            FFC_PHASE { fft2_rows_L2b<H, W, -1, false>(tid, ctx.nt, np, B, tw); } FFC_SYNC;       \
            FFC_PHASE { fft2_rows_fwd_L2sep<H, W>(tid, ctx.nt, np, B, A); } FFC_SYNC;             \
            spec_out = A;                                                                         \
        }                                                                                         \
        if constexpr (PL_::H2 == 1) {                                                                       \
            FFC_PHASE { fft2_cols_L1<H, W, -1>(tid, ctx.nt, np, spec_out); } FFC_SYNC;            \
        } else {                                                                                  \
            FFC_PHASE { fft2_cols_L2_strided<H, W, -1, true>(tid, ctx.nt, np, spec_out, tw); } FFC_SYNC;  \
            FFC_PHASE { fft2_cols_L2_contig<H, W, -1, false>(tid, ctx.nt, np, spec_out, tw); } FFC_SYNC;  \
        }                                                                                         \
    } while (0)

#define FFC_FFT2_INVERSE(H, W, np, S, O, tw, scale, real_out)                                     \
    do {                                                                                          \
        typedef Fft2Plan<H, W> PL_;                                                               \
        if constexpr (PL_::H2 == 1) {                                                                       \
            FFC_PHASE { fft2_cols_L1<H, W, +1>(tid, ctx.nt, np, S); } FFC_SYNC;                   \
        } else {                                                                                  \
            FFC_PHASE { fft2_cols_L2_contig<H, W, +1, true>(tid, ctx.nt, np, S, tw); } FFC_SYNC;  \
            FFC_PHASE { fft2_cols_L2_strided<H, W, +1, false>(tid, ctx.nt, np, S, tw); } FFC_SYNC;\
        }                                                                                         \
        if constexpr (!PL_::kRowsTwoLevel) {                                                                \
            FFC_PHASE { fft2_rows_inv_L1<H, W>(tid, ctx.nt, np, S, O, scale); } FFC_SYNC;         \
            real_out = O;                                                                         \
        } else {                                                                                  \
            FFC_PHASE { fft2_rows_inv_L2build<H, W>(tid, ctx.nt, np, S, O); } FFC_SYNC;           \
            FFC_PHASE { fft2_rows_L2b<H, W, +1, true>(tid, ctx.nt, np, O, tw); } FFC_SYNC;        \
            FFC_PHASE { fft2_rows_inv_L2a<H, W>(tid, ctx.nt, np, O, S, scale); } FFC_SYNC;        \
            real_out = S;                                                                         \
        }                                                                                         \
    } while (0)
