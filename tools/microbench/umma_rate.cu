// Issue-rate micro-benchmark of tcgen05.mma kind::tf32 (M = 128, K = 8 per instruction) as the library's kernels use it:
// A from tensor memory (.ts) or from shared memory (.ss), B from no-swizzle K-major shared-memory tiles, N in {64, 128, 256},
// all MMAs accumulating into ONE accumulator or alternating between two.  One CTA per SM, one elected thread issues; a
// tcgen05.commit + mbarrier wait every 32 instructions bounds the queue.  Prints cycles per MMA and the TF32 rate of the chip.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu -I../../fastfourierconvolution_b200/csrc
#include <cstdio>
#include <cstdlib>
#include "ffc_umma.cuh"

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int ss, int two_acc, int iters, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid / 32;
    for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (float)(i % 97);
    if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_proxy_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = *tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma::idesc_tf32(128, N);
        const uint32_t a_smem = umma::smem_u32(smem), b_smem = a_smem + 32 * 1024;     // A tile 128 x 32 floats, B tile up to 256 x 32 floats
        const uint32_t a_tmem = tbase + 448;                                            // 64 columns of A at the top of tensor memory
        uint32_t phase = 0;
        long long t0 = 0;
        for (int rep = 0; rep < 2; ++rep) {               // first repetition warms up
            __syncwarp();
            if (rep == 1) t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                if (umma::elect_one()) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int ks = j & 3;
                        const uint32_t d = tbase + ((two_acc && (j & 4)) ? (uint32_t)N : 0u);
                        const uint64_t bd = umma::smem_desc_kmajor_noswizzle(b_smem + ks * 256, 128, 1024);
                        if (ss) {
                            const uint64_t ad = umma::smem_desc_kmajor_noswizzle(a_smem + ks * 256, 128, 1024);
                            umma::mma_tf32_ss(d, ad, bd, idesc, 1u);
                        } else {
                            umma::mma_tf32_ts(d, a_tmem + ks * 8, bd, idesc, 1u);
                        }
                    }
                    umma::commit(bar);
                }
                __syncwarp();
                umma::mbar_wait(bar, phase);
                phase ^= 1u;
            }
        }
        const long long t1 = clock64();
        if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    long long* d_cycles;
    cudaMalloc(&d_cycles, sizeof(long long) * sms);
    const size_t smem = 96 * 1024 + 64;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 400;
    printf("tcgen05.mma kind::tf32, M = 128, K = 8, %d SMs, 32 MMAs between commits (sm clock attribute %d kHz)\n", sms, khz);
    for (int ss = 0; ss < 2; ++ss)
        for (int two = 0; two < 2; ++two)
            for (int N = 64; N <= 256; N *= 2) {
                if (two && 2 * N > 448) continue;
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                rate_kernel<<<sms, 128, smem>>>(N, ss, two, iters, d_cycles);
                cudaEventRecord(e1);
                cudaError_t err = cudaDeviceSynchronize();
                if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); return 1; }
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                long long c0 = 0;
                cudaMemcpy(&c0, d_cycles, sizeof(long long), cudaMemcpyDeviceToHost);
                const double per = (double)c0 / ((double)iters * 32.0);
                const double flops = 2.0 * 128 * N * 8 * 32.0 * iters * sms;          // measured repetition only
                printf("A %s, %s, N = %3d: %7.1f cycles per MMA, %7.1f TFLOP/s (kernel incl. warm-up %.3f ms)\n", ss ? "smem" : "TMEM",
                       two ? "two accumulators" : "one accumulator ", N, per, flops / ((double)c0 / ((double)khz * 1e3)) / 1e12, ms);
            }
    return 0;
}
