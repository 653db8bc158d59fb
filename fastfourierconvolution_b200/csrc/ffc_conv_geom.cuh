// Output-parity-class geometry shared by the tensor-core convolution kernels (ffc_conv_v4.cu, ffc_conv_v5.cu).
// A transposed convolution of stride s is s*s ordinary (stride-1 gather) convolutions, one per output parity class
// (oy % s, ox % s), each with its own subset of the k*k taps; a plain convolution is the single class 0.
#pragma once
#include "ffc_common.cuh"

// tap geometry of an output parity class (shared by the pack kernel, the main kernel and the host)
struct ConvClassGeom { int ky0, kx0, qy, qx, Ta, Tb; };
FFC_HD ConvClassGeom ffc_conv_class_geom(int cls, int k, int stride, int pad, int transposed) {
    ConvClassGeom g;
    g.ky0 = 0; g.kx0 = 0; g.qy = 0; g.qx = 0; g.Ta = k; g.Tb = k;
    if (transposed) {
        const int s = stride, py = cls / s, px = cls % s;
        g.ky0 = (py + pad) % s; g.kx0 = (px + pad) % s;
        g.qy = (py + pad - g.ky0) / s; g.qx = (px + pad - g.kx0) / s;
        g.Ta = g.ky0 < k ? (k - g.ky0 + s - 1) / s : 0;
        g.Tb = g.kx0 < k ? (k - g.kx0 + s - 1) / s : 0;
    }
    return g;
}

