// Shared between ffc_fu3.cu (plane transforms, host orchestration, plain FP32 mix) and ffc_fu3_mix.cu (tensor-core mix).
#pragma once
#include "ffc_common.cuh"

struct Fu3MixParams {
    const float* s;          // (G, Cin, NB) complex: spectrum of a chunk of images, "shared-memory image" layout
    float* y;                // (G, Cout, NB) complex, or null (statistics only)
    const float* w;          // [2*Cout][2*Cin] (conv_layer.weight)
    const float* wp;         // packed tensor-core image of w (ffc_fu3_mix.cu) or null
    const float* bn_a;       // [2*Cout] or null: apply relu(y * a + b) before storing
    const float* bn_b;
    double* sums;            // [4*Cout] or null: sum(y) | sum(y^2) over the REAL bins (pads skipped), channel-major
    int G, Cin, Cout, NB, SPS;
    float scale;             // forward transform scale folded into the mix (1/N)
};

// dY = coef * (g - c1 - xhat * c2) and dW = scale * dY^T S (backward of the mix + BatchNorm; ffc_fu3.cu FP32 form, ffc_fu3_mix.cu tensor cores)
struct Fu3BwdWgradParams {
    const float* g;        // (B, Cout, NB) complex
    const float* y;        // (B, Cout, NB)
    const float* s;        // (B, Cin, NB)
    float* dy;             // (B, Cout, NB)
    const float* coef; const float* c1; const float* c2; const float* mean; const float* invstd;     // [2*Cout]
    float* dw;             // [2*Cout][2*Cin], zeroed by the host wrapper
    float* part;           // tensor-core form: per-CTA partial tiles [CTA][128][NT] (workspace)
    int B, Cin, Cout, NB;
    float scale;
};
