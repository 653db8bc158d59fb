// Common definitions for the ffc_b200 CUDA library (sm_100a).
//
// Every kernel in this library is written as a "phase program": a struct K with
//   using Params = ...;                        (POD, passed by value)
//   static constexpr int kThreads = ...;       (upper bound for __launch_bounds__)
//   static FFC_DEVICE void run(const Params&, const BlockCtx&, float* smem);
// whose body is a sequence of FFC_PHASE { ... } FFC_SYNC; blocks.  On the device a phase is
// executed by every thread of the CTA once and FFC_SYNC is __syncthreads().  When the same
// source is compiled with -DFFC_EMU by a host C++ compiler, a phase is a sequential loop over
// the thread ids and FFC_SYNC is a no-op, which yields a bit-faithful (up to libm / fma
// contraction) host execution of the kernel logic.  The emulation build exists ONLY so that
// tests/ can check index arithmetic on a machine without a GPU; the product never loads it.
//
// Rules that make a kernel emulable: no warp intrinsics, no state kept in plain locals across
// phases (use FFC_TLS for per-thread state that must survive a barrier), atomics through
// ffc_atomic_add, and within one phase no thread reads a location another thread writes.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef FFC_EMU
// ------------------------------------------------------------------ host emulation
#include <vector>
#include <cstring>
#include <cstdlib>
#include <string>
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
typedef void* ffc_stream_t;
#define FFC_HD static inline
#define FFC_HDM inline
#define FFC_DEVICE inline
#define FFC_CONST static const
#define FFC_RESTRICT __restrict__
#define FFC_UNROLL
#define FFC_LDG(p) (*(p))
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
template <class T> static inline void ffc_atomic_add(T* p, T v) { *p += v; }
#else
// ------------------------------------------------------------------ device build
#include <cuda_runtime.h>
typedef cudaStream_t ffc_stream_t;
#define FFC_HD __host__ __device__ __forceinline__
#define FFC_HDM __host__ __device__ __forceinline__
#define FFC_DEVICE __device__ __forceinline__
#define FFC_CONST static __constant__
#define FFC_RESTRICT __restrict__
#define FFC_UNROLL _Pragma("unroll")
#define FFC_LDG(p) __ldg(p)
template <class T> __device__ __forceinline__ void ffc_atomic_add(T* p, T v) { atomicAdd(p, v); }
#endif

struct BlockCtx {
    int bx, by, bz;   // block index
    int gx, gy, gz;   // grid size
    int nt;           // threads per block
};

#ifdef FFC_EMU
#define FFC_PHASE for (int tid = 0; tid < ctx.nt; ++tid)
#define FFC_SYNC ((void)0)
// per-thread state that survives a barrier
#define FFC_TLS(T, name) std::vector<T> name##_tls((size_t)ctx.nt)
#define FFC_TLS_REF(T, name) T& name = name##_tls[(size_t)tid]
#else
#define FFC_PHASE for (int tid = (int)threadIdx.x, _ffc_once = 1; _ffc_once; _ffc_once = 0)
#define FFC_SYNC __syncthreads()
#define FFC_TLS(T, name) T name
#define FFC_TLS_REF(T, name) ((void)0)
#endif

// ------------------------------------------------------------------ error reporting
// 0 = OK.  Non-zero codes are returned by every C-ABI entry point; ffc_last_error() gives text.
enum {
    FFC_OK = 0,
    FFC_ERR_BAD_ARG = 1,       // unsupported shape / mode / null pointer
    FFC_ERR_WORKSPACE = 2,     // caller-provided workspace too small
    FFC_ERR_CUDA = 3           // a CUDA runtime call failed (launch error etc.)
};

void ffc_set_error(const char* fmt, ...);
void ffc_count_launch();

#ifdef FFC_EMU
// ------------------------------------------------------------------ launch (emulated)
template <class K>
static int ffc_launch(int gx, int gy, int gz, int nt, size_t smem_bytes, ffc_stream_t, const typename K::Params& p) {
    std::vector<float4> smem((smem_bytes + 15) / 16 + 1);
    for (int bz = 0; bz < gz; ++bz)
        for (int by = 0; by < gy; ++by)
            for (int bx = 0; bx < gx; ++bx) {
                BlockCtx ctx{bx, by, bz, gx, gy, gz, nt};
                K::run(p, ctx, reinterpret_cast<float*>(smem.data()));
            }
    ffc_count_launch();
    return FFC_OK;
}
// cooperative two-part kernels (a grid-wide barrier between K::part0 and K::part1): every block keeps its own
// shared memory alive across the barrier, exactly like co-resident CTAs do on the device
template <class K>
static int ffc_launch_coop(int gx, int nt, size_t smem_bytes, ffc_stream_t, const typename K::Params& p) {
    std::vector<std::vector<float4>> smem((size_t)gx, std::vector<float4>((smem_bytes + 15) / 16 + 1));
    for (int part = 0; part < 2; ++part)
        for (int bx = 0; bx < gx; ++bx) {
            BlockCtx ctx{bx, 0, 0, gx, 1, 1, nt};
            if (part == 0) K::part0(p, ctx, reinterpret_cast<float*>(smem[bx].data()));
            else K::part1(p, ctx, reinterpret_cast<float*>(smem[bx].data()));
        }
    ffc_count_launch();
    return FFC_OK;
}
template <class K> static int ffc_coop_capacity_blocks(int, size_t) { return 1 << 30; }
template <class K> static int ffc_resident_blocks(int, size_t) { return 1 << 30; }
static inline int ffc_sm_count() { return 148; }
static inline int ffc_memset_async(void* p, int v, size_t n, ffc_stream_t) { memset(p, v, n); return FFC_OK; }
#else
// ------------------------------------------------------------------ launch (device)
#include <type_traits>
// Per-device launch state.  cudaFuncSetAttribute applies to the CURRENT device only, so the "already raised to N bytes"
// high-water marks are kept per device (one process may drive several GPUs, e.g. nn.DataParallel, train_cond.py:67-68).
#define FFC_MAX_DEVICES 64
struct FfcPerDevice { size_t v[FFC_MAX_DEVICES]; };
static inline size_t* ffc_device_slot(FfcPerDevice& s) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return &s.v[dev % FFC_MAX_DEVICES];
}
// multiprocessors of the current device (grids are sized in multiples of it)
static inline int ffc_sm_count() {
    static int cached[FFC_MAX_DEVICES] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    int& c = cached[dev % FFC_MAX_DEVICES];
    if (c <= 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        c = n;
    }
    return c;
}
// optional K::kMinBlocks (resident CTAs per SM the register allocation must allow); default 1
template <class K, class = void> struct FfcMinBlocks { static constexpr int v = 1; };
template <class K> struct FfcMinBlocks<K, std::void_t<decltype(K::kMinBlocks)>> { static constexpr int v = K::kMinBlocks; };

template <class K>
__global__ void __launch_bounds__(K::kThreads, FfcMinBlocks<K>::v) ffc_kernel(const typename K::Params p) {
    extern __shared__ float4 ffc_smem4[];
    BlockCtx ctx{(int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z,
                 (int)gridDim.x, (int)gridDim.y, (int)gridDim.z, (int)blockDim.x};
    K::run(p, ctx, reinterpret_cast<float*>(ffc_smem4));
}

template <class K>
static int ffc_launch(int gx, int gy, int gz, int nt, size_t smem_bytes, ffc_stream_t stream, const typename K::Params& p) {
    if (gx <= 0 || gy <= 0 || gz <= 0) return FFC_OK;   // empty problem
    if (nt > K::kThreads || gy > 65535 || gz > 65535) {
        ffc_set_error("launch config out of range (nt=%d, grid=%d,%d,%d)", nt, gx, gy, gz);
        return FFC_ERR_BAD_ARG;
    }
    // Kernels that stage tiles in shared memory: raise the dynamic limit and ask for the largest shared
    // carve-out so that occupancy is not capped by the default L1/shared split (set once per high-water mark;
    // per device; a benign race if two host threads launch the same kernel for the first time).
    static FfcPerDevice configured = {};
    size_t& configured_smem = *ffc_device_slot(configured);
    if (smem_bytes > 16 * 1024 && smem_bytes > configured_smem) {
        cudaError_t e = cudaSuccess;
        if (smem_bytes > 48 * 1024)
            e = cudaFuncSetAttribute(ffc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(ffc_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { ffc_set_error("cudaFuncSetAttribute(%zu B smem): %s", smem_bytes, cudaGetErrorString(e)); return FFC_ERR_CUDA; }
        configured_smem = smem_bytes;
    }
    ffc_kernel<K><<<dim3(gx, gy, gz), dim3(nt), smem_bytes, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("kernel launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}
// resident CTAs of ffc_kernel<K> on the whole device for this launch shape (persistent kernels size their grid with it)
template <class K>
static int ffc_resident_blocks(int nt, size_t smem_bytes) {
    if (smem_bytes > 48 * 1024) cudaFuncSetAttribute(ffc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (smem_bytes > 16 * 1024) cudaFuncSetAttribute(ffc_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ffc_kernel<K>, nt, smem_bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * ffc_sm_count();
}
// ---- cooperative two-part kernels: part0, grid-wide barrier, part1 (all CTAs co-resident)
#include <cooperative_groups.h>
template <class K>
__global__ void __launch_bounds__(K::kThreads, FfcMinBlocks<K>::v) ffc_kernel_coop(const typename K::Params p) {
    extern __shared__ float4 ffc_smem4[];
    BlockCtx ctx{(int)blockIdx.x, 0, 0, (int)gridDim.x, 1, 1, (int)blockDim.x};
    K::part0(p, ctx, reinterpret_cast<float*>(ffc_smem4));
    cooperative_groups::this_grid().sync();
    K::part1(p, ctx, reinterpret_cast<float*>(ffc_smem4));
}
// how many CTAs of this kernel can be resident at once on the current device (0 on error)
template <class K>
static int ffc_coop_capacity_blocks(int nt, size_t smem_bytes) {
    static FfcPerDevice configured = {};
    size_t& configured_smem = *ffc_device_slot(configured);
    if (smem_bytes > configured_smem) {
        if (smem_bytes > 48 * 1024 &&
            cudaFuncSetAttribute(ffc_kernel_coop<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return 0;
        if (cudaFuncSetAttribute(ffc_kernel_coop<K>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) return 0;
        configured_smem = smem_bytes;
    }
    int dev = 0, sms = 0, per_sm = 0, coop = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ffc_kernel_coop<K>, nt, smem_bytes) != cudaSuccess) return 0;
    return per_sm * sms;
}
template <class K>
static int ffc_launch_coop(int gx, int nt, size_t smem_bytes, ffc_stream_t stream, const typename K::Params& p) {
    void* args[] = {(void*)&p};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)ffc_kernel_coop<K>, dim3(gx), dim3(nt), args, smem_bytes, stream);
    if (e != cudaSuccess) { ffc_set_error("cooperative launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}
// Zero fill as a KERNEL node.  Inside a captured CUDA graph a cudaMemsetAsync node in front of a kernel cost ~5 us of
// dependency latency on B200 (measured on the fused Fourier unit: memset + cooperative kernel vs the kernel alone,
// profiles/r04e_fu4_ab.jsonl), a kernel -> kernel edge ~1 us; the library zeroes ~100 small buffers per training step.
static __global__ void ffc_zero_kernel(uint4* p16, size_t n16, unsigned char* tail, size_t ntail) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p16[i] = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ntail; i += stride) tail[i] = 0;
}
static inline int ffc_memset_async(void* p, int v, size_t n, ffc_stream_t s) {
    if (n == 0) return FFC_OK;
    if (v != 0) {
        cudaError_t e = cudaMemsetAsync(p, v, n, s);
        if (e != cudaSuccess) { ffc_set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
        return FFC_OK;
    }
    unsigned char* b = static_cast<unsigned char*>(p);
    size_t head = (16 - ((uintptr_t)b & 15)) & 15;          // bytes up to the first 16-byte boundary
    if (head > n) head = n;
    const size_t n16 = (n - head) / 16, ntail = (n - head) % 16;
    if (head) {                                             // unaligned start: rare, zero it through the tail path of a first launch
        ffc_zero_kernel<<<1, 32, 0, s>>>(nullptr, 0, b, head);
        ffc_count_launch();
    }
    size_t blocks = (n16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    const size_t cap = (size_t)8 * ffc_sm_count();
    if (blocks > cap) blocks = cap;
    ffc_zero_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<uint4*>(b + head), n16, b + head + n16 * 16, ntail);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("zero-fill launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}
#endif

#define FFC_CHECK(call) do { int _e = (call); if (_e != FFC_OK) return _e; } while (0)
#define FFC_REQUIRE(cond, ...) do { if (!(cond)) { ffc_set_error(__VA_ARGS__); return FFC_ERR_BAD_ARG; } } while (0)

// ------------------------------------------------------------------ 3xTF32 operand split
// hi = x rounded to NEAREST TF32 (10 explicit mantissa bits), lo = x - hi (exact in FP32, |lo| <= 2^-11 |x|).  The tensor core
// truncates its FP32 inputs to TF32: hi passes unchanged, lo loses <= 2^-10 |lo|.  Rounding instead of truncating the split
// halves |lo| and with it every error term of a*b ~ a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (measured on the fgan128 fixture:
// tools/diag_fixture_accuracy.py).  The +0x1000 carries into the exponent correctly; inf / nan do not occur here.
#ifndef FFC_EMU
__device__ __forceinline__ float ffc_tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
#else
static inline float ffc_tf32_hi(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }
#endif

// ------------------------------------------------------------------ small helpers
FFC_HD int ffc_cdiv(int a, int b) { return (a + b - 1) / b; }
FFC_HD int ffc_ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
FFC_HD bool ffc_is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// activation codes shared by the host API and the kernels (include/ffc_b200.h mirrors them)
enum { FFC_ACT_IDENTITY = 0, FFC_ACT_RELU = 1, FFC_ACT_LEAKY = 2, FFC_ACT_GELU = 3, FFC_ACT_TANH = 4, FFC_ACT_SIGMOID = 5 };

FFC_HD float ffc_act_fwd(float z, int act, float slope) {
    switch (act) {
        case FFC_ACT_RELU: return z > 0.f ? z : 0.f;
        case FFC_ACT_LEAKY: return z > 0.f ? z : slope * z;
        case FFC_ACT_GELU: return 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));
        case FFC_ACT_TANH: return tanhf(z);
        case FFC_ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
        default: return z;
    }
}
// derivative of the activation with respect to its input z
FFC_HD float ffc_act_bwd(float z, int act, float slope) {
    switch (act) {
        case FFC_ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case FFC_ACT_LEAKY: return z > 0.f ? 1.f : slope;
        case FFC_ACT_GELU: {
            float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
            float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
            return cdf + z * pdf;
        }
        case FFC_ACT_TANH: { float t = tanhf(z); return 1.0f - t * t; }
        case FFC_ACT_SIGMOID: { float s = 1.0f / (1.0f + expf(-z)); return s * (1.0f - s); }
        default: return 1.f;
    }
}
