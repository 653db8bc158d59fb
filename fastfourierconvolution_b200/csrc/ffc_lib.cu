// Single translation unit of libffc_b200.so (keeps the __constant__ tables and templates in one
// place and lets nvcc see every kernel at once).
#include "ffc_api.cu"
#include "ffc_fft2.cu"
#include "ffc_conv.cu"
#include "ffc_bnact.cu"
#include "ffc_fu_fused.cu"
