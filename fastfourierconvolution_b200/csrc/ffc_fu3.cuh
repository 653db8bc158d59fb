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
