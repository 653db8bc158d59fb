// Micro-benchmark: issue rate of legacy mma.sync (m16n8k8 TF32, m16n8k16 BF16) and FP32 FFMA / FFMA2 on sm_100a.
// Decides whether a 3xTF32 mma.sync convolution can beat a tuned FP32 SIMT kernel on B200.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_tf32(float* out, int iters) {
    float c[8][4] = {};
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16(float* out, int iters) {
    float c[8][4] = {};
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters, float x) {
    float c[16];
    for (int j = 0; j < 16; ++j) c[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) c[j] = fmaf(c[j], x, 1.0f);
    }
    float s = 0; for (int j = 0; j < 16; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, int iters, float x) {
    float2 c[16];
    for (int j = 0; j < 16; ++j) c[j] = make_float2(threadIdx.x + j, j);
    const float2 xx = make_float2(x, x), one = make_float2(1.f, 1.f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) c[j] = __ffma2_rn(c[j], xx, one);
    }
    float s = 0; for (int j = 0; j < 16; ++j) s += c[j].x + c[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    const int iters = 20000, grid = 148 * 4, block = 256;
    double warps = (double)grid * block / 32;
    float ms = timeit([&] { k_tf32<<<grid, block>>>(out, iters); });
    printf("mma.sync m16n8k8 tf32 : %.1f TFLOP/s\n", warps * iters * 8 * (2.0 * 16 * 8 * 8) / ms / 1e9);
    ms = timeit([&] { k_bf16<<<grid, block>>>(out, iters); });
    printf("mma.sync m16n8k16 bf16: %.1f TFLOP/s\n", warps * iters * 8 * (2.0 * 16 * 8 * 16) / ms / 1e9);
    ms = timeit([&] { k_ffma<<<grid, block>>>(out, iters, 0.999f); });
    printf("FFMA  fp32            : %.1f TFLOP/s\n", (double)grid * block * iters * 16 * 2 / ms / 1e9);
    ms = timeit([&] { k_ffma2<<<grid, block>>>(out, iters, 0.999f); });
    printf("FFMA2 fp32x2          : %.1f TFLOP/s\n", (double)grid * block * iters * 16 * 4 / ms / 1e9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
