// Translation unit 1 of libffc_b200.so: API, plane FFT kernels, convolutions, BatchNorm/activation/SE.
// (ffc_unit_fu.cu holds the fused Fourier-unit kernels; the units are compiled in parallel.)
#include "ffc_api.cu"
#include "ffc_fft2.cu"
#include "ffc_conv.cu"
#include "ffc_bnact.cu"
