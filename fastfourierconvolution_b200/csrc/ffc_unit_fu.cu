// Translation unit 2 of libffc_b200.so: the fused FourierUnit kernels (24 template instantiations).
#include "ffc_fu_fused.cu"
