// Standalone check of the tcgen05 building blocks in csrc/ffc_umma.cuh: D[128 x N] = A[128 x K] * B[N x K]^T at FP32
// accuracy (3xTF32), A written to tensor memory by the producer threads, B from pre-packed no-swizzle K-major tiles.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_gemm_test umma_gemm_test.cu -I../../fastfourierconvolution_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "ffc_umma.cuh"

constexpr int BK = 32;

__global__ void __launch_bounds__(160, 1) gemm_test(const float* A, const float* Bp, float* D, int N, int K) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* bsm = reinterpret_cast<float*>(smem);                                  // [hi | lo] tiles of N x 32
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)2 * N * BK * 4);  // 0: B full, 1: A ready, 2: mma done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int tid = threadIdx.x, warp = tid / 32;
    if (tid == 0) {
        umma::mbar_init(&bars[0], 1); umma::mbar_init(&bars[1], 128); umma::mbar_init(&bars[2], 1);
        umma::fence_barrier_init();
    }
    if (warp == 4) umma::tmem_alloc(tmem_slot, 256);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    const uint32_t d_col = 0, a_col = 128;          // D: columns [0, N), A hi: [128, 160), A lo: [160, 192)
    const int nchunks = K / BK;
    const uint32_t idesc = umma::idesc_tf32(128, N);
    for (int c = 0; c < nchunks; ++c) {
        const uint32_t ph = c & 1;
        if (warp < 4) {
            // producer: row m = tid
            uint32_t hi[BK], lo[BK];
            for (int j = 0; j < BK; ++j) {
                const float x = A[(size_t)tid * K + c * BK + j];
                const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
                hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(x - h);
            }
            const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
            umma::tmem_st16(lane_addr + a_col, hi); umma::tmem_st16(lane_addr + a_col + 16, hi + 16);
            umma::tmem_st16(lane_addr + a_col + 32, lo); umma::tmem_st16(lane_addr + a_col + 48, lo + 16);
            umma::wait_st();
            umma::fence_before_sync();
            umma::mbar_arrive(&bars[1]);
            umma::mbar_wait(&bars[2], ph);           // stage reusable
        } else if (tid == 128) {
            const uint32_t bytes = (uint32_t)(2 * N * BK * 4);
            umma::mbar_arrive_expect_tx(&bars[0], bytes);
            umma::bulk_g2s(bsm, Bp + (size_t)c * 2 * N * BK, bytes, &bars[0]);
            umma::mbar_wait(&bars[0], ph);
            umma::mbar_wait(&bars[1], ph);
            umma::fence_after_sync();
            const uint32_t b_hi = umma::smem_u32(bsm), b_lo = b_hi + N * BK * 4;
            for (int ks = 0; ks < BK / 8; ++ks) {
                const uint64_t dh = umma::smem_desc_kmajor_noswizzle(b_hi + ks * 256, 128, 1024);
                const uint64_t dl = umma::smem_desc_kmajor_noswizzle(b_lo + ks * 256, 128, 1024);
                umma::mma_tf32_ts(tbase + d_col, tbase + a_col + 32 + ks * 8, dh, idesc, (c | ks) ? 1u : 0u);   // lo * hi
                umma::mma_tf32_ts(tbase + d_col, tbase + a_col + ks * 8, dl, idesc, 1u);                      // hi * lo
                umma::mma_tf32_ts(tbase + d_col, tbase + a_col + ks * 8, dh, idesc, 1u);                      // hi * hi
            }
            umma::commit(&bars[2]);
            umma::mbar_wait(&bars[2], ph);
        }
        __syncthreads();
    }
    umma::fence_after_sync();
    if (warp < 4) {
        const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
        for (int n0 = 0; n0 < N; n0 += 16) {
            uint32_t r[16];
            umma::tmem_ld16(lane_addr + d_col + n0, r);
            umma::wait_ld();
            for (int j = 0; j < 16; ++j) D[(size_t)tid * N + n0 + j] = __uint_as_float(r[j]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 4) umma::tmem_dealloc(tbase, 256);
}

int main() {
    const int M = 128, N = 96, K = 64;
    std::vector<float> A(M * K), B(N * K), Bp((size_t)(K / BK) * 2 * N * BK), D(M * N);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX - 0.5f;
    for (auto& v : B) v = (float)rand() / RAND_MAX - 0.5f;
    for (int c = 0; c < K / BK; ++c)
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < BK; ++k) {
                const float x = B[n * K + c * BK + k];
                unsigned u; memcpy(&u, &x, 4); u &= 0xffffe000u; float h; memcpy(&h, &u, 4);
                const size_t off = (size_t)(n / 8) * 1024 / 4 + (k / 4) * 32 + (n % 8) * 4 + (k % 4);
                Bp[(size_t)c * 2 * N * BK + off] = h;
                Bp[(size_t)c * 2 * N * BK + N * BK + off] = x - h;
            }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bp.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bp.data(), Bp.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const size_t smem = (size_t)2 * N * BK * 4 + 64;
    cudaFuncSetAttribute(gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gemm_test<<<1, 160, smem>>>(dA, dB, dD, N, K);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
            maxerr = fmax(maxerr, fabs(s - D[m * N + n])); maxref = fmax(maxref, fabs(s));
        }
    printf("max abs err %.3e, max ref %.3e, rel %.3e  D[0][0..3] = %f %f %f %f\n", maxerr, maxref, maxerr / maxref, D[0], D[1], D[2], D[3]);
    return 0;
}
