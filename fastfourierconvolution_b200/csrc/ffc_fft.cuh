// In-register radix-2 FFTs with compile-time twiddles (fully unrolled; N <= 32 per thread).
// Larger transforms (64, 128) are composed from two in-register levels with one shared-memory
// exchange by the callers (see ffc_fft2.cu), using the "permuted frequency" convention:
//   N = N1 * N2, frequency k = k1 + N1*k2 is stored at position pos(k) = N2*k1 + k2.
// Forward transforms map natural time order -> permuted frequency order, inverse transforms map
// permuted frequency order -> natural time order; for N <= 32, N2 == 1 and pos(k) == k.
#pragma once
#include "ffc_common.cuh"
#include "ffc_twiddles.h"

// Packed FP32x2 arithmetic: sm_100 issues add/mul/fma on a register pair as ONE instruction
// (FADD2 / FMUL2 / FFMA2), which halves the instruction count of the complex butterflies and of the
// channel mix.  The emulation build uses plain scalar math (same rounding: each half is an IEEE op).
#if !defined(FFC_EMU)
FFC_DEVICE float2 ffc_add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
FFC_DEVICE float2 ffc_sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
FFC_DEVICE float2 ffc_mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
FFC_DEVICE float2 ffc_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
FFC_DEVICE float2 ffc_add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FFC_DEVICE float2 ffc_sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FFC_DEVICE float2 ffc_mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
FFC_DEVICE float2 ffc_fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
FFC_DEVICE float2 ffc_cadd(float2 a, float2 b) { return ffc_add2(a, b); }
FFC_DEVICE float2 ffc_csub(float2 a, float2 b) { return ffc_sub2(a, b); }
// a * (c + i*s*sign)
template <int SIGN>
FFC_DEVICE float2 ffc_cmul_tw(float2 a, float2 w) {
    return SIGN > 0 ? make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x)
                    : make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
}

FFC_HD constexpr int ffc_clog2(int n) { return n <= 1 ? 0 : 1 + ffc_clog2(n >> 1); }
FFC_HD constexpr int ffc_bitrev(int i, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    return r;
}

// v[0..N) natural order in, natural order out.  SIGN = -1: X[k] = sum x[n] e^{-2 pi i nk/N};
// SIGN = +1: the unnormalised inverse.  All indices are compile-time after unrolling, so v[]
// lives in registers and the twiddles become immediate / constant-bank operands.
template <int N, int SIGN>
FFC_DEVICE void ffc_fft_regs(float2* v) {
    constexpr int L = ffc_clog2(N);
    FFC_UNROLL
    for (int s = 0; s < L; ++s) {
        const int len = N >> s, half = len >> 1;
        FFC_UNROLL
        for (int blk = 0; blk < N; blk += len) {
            FFC_UNROLL
            for (int j = 0; j < half; ++j) {
                const float2 a = v[blk + j], b = v[blk + j + half];
                v[blk + j] = ffc_cadd(a, b);
                const float2 d = ffc_csub(a, b);
                if (j == 0) {
                    v[blk + j + half] = d;
                } else if (4 * j == len) {          // w = -+ i
                    v[blk + j + half] = SIGN > 0 ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
                } else {
                    v[blk + j + half] = ffc_cmul_tw<SIGN>(d, c_tw128[j * (FFC_TW_N / len)]);
                }
            }
        }
    }
    // bit-reversal (register renaming after unrolling)
    FFC_UNROLL
    for (int i = 0; i < N; ++i) {
        const int r = ffc_bitrev(i, L);
        if (r > i) { const float2 t = v[i]; v[i] = v[r]; v[r] = t; }
    }
}

// decomposition used by the smem-level callers
template <int N> struct FftSplit {
    static constexpr int N1 = (N <= 32) ? N : N / 8;   // first-level in-register size
    static constexpr int N2 = N / N1;                  // 1 or 8
    static FFC_DEVICE int pos(int k) { return N2 * (k % N1) + k / N1; }
};
