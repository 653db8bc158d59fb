// Direct (SIMT, FP32 FMA) convolution kernels for layers with at most 4 channels on one side -- the RGB output
// convolution of every generator (fgan_complete.py:110-113: FFC_BN_ACT(ngf, 3, ...)) and its gradients.  A 128 x N
// tensor-core tile is >90 % padding there and the work is bandwidth bound, so one thread per output pixel with the few
// channels in registers is the right shape:
//   ConvSmallCout : cout <= 4, any cin (two summed segments), acc[4] per pixel, weights broadcast from shared memory
//   ConvSmallCin  : cin  <= 4 (one segment), the k*k*cin gathered inputs of a pixel stay in registers across all cout
//   WgradSmallSc  : SC   <= 4 small-side channels, thread = pixel walker with SC*k*k partial sums in registers,
//                   warp-shuffle + shared-memory reduction, one atomic per weight entry and CTA
// Same contracts as ffc_conv2d_fwd / ffc_conv2d_wgrad (gather form, transposed by tap validity).  Device build only.
#include "ffc_common.cuh"

#ifndef FFC_EMU

struct SmallConvParams {
    const float* x[2]; const float* w[2]; int cin[2]; int nseg;
    const float* bias; const float* addend; float* y;
    int B, cout, Hi, Wi, Ho, Wo, k, stride, pad, transposed;
};

// input offset (iy*Wi + ix) of tap (ky,kx) for output pixel (oy,ox), or -1 when the tap does not contribute
__device__ __forceinline__ int small_tap_off(const SmallConvParams& p, int oy, int ox, int ky, int kx) {
    int iy, ix;
    if (!p.transposed) { iy = oy * p.stride - p.pad + ky; ix = ox * p.stride - p.pad + kx; }
    else {
        const int ty = oy + p.pad - ky, tx = ox + p.pad - kx;
        if (ty < 0 || tx < 0 || ty % p.stride || tx % p.stride) return -1;
        iy = ty / p.stride; ix = tx / p.stride;
    }
    return (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) ? iy * p.Wi + ix : -1;
}
// weight element [co][ci][ky][kx] (conv) or [ci][co][ky][kx] (transposed)
__device__ __forceinline__ float small_w(const SmallConvParams& p, const float* w, int cin, int co, int ci, int t) {
    const int KK = p.k * p.k;
    return p.transposed ? __ldg(w + ((size_t)ci * p.cout + co) * KK + t) : __ldg(w + ((size_t)co * cin + ci) * KK + t);
}

template <int K>
__global__ void __launch_bounds__(256) conv_small_cout_kernel(const SmallConvParams p) {
    extern __shared__ float4 scw[];                 // [seg ci][tap] -> (w of co 0..3)
    constexpr int KK = K * K;
    const int ctot = p.cin[0] + (p.nseg > 1 ? p.cin[1] : 0);
    for (int e = threadIdx.x; e < ctot * KK; e += blockDim.x) {
        const int c = e / KK, t = e % KK;
        const int sg = c >= p.cin[0], ci = sg ? c - p.cin[0] : c;
        float q[4] = {0.f, 0.f, 0.f, 0.f};
        for (int co = 0; co < p.cout; ++co) q[co] = small_w(p, p.w[sg], p.cin[sg], co, ci, t);
        scw[e] = make_float4(q[0], q[1], q[2], q[3]);
    }
    __syncthreads();
    const int HWo = p.Ho * p.Wo, HWi = p.Hi * p.Wi;
    const long long total = (long long)p.B * HWo;
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < total; m += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(m / HWo), r = (int)(m % HWo), oy = r / p.Wo, ox = r % p.Wo;
        int off[KK];
#pragma unroll
        for (int t = 0; t < KK; ++t) off[t] = small_tap_off(p, oy, ox, t / K, t % K);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int cbase = 0;
        for (int sg = 0; sg < p.nseg; ++sg) {
            const float* xp = p.x[sg] + (size_t)b * p.cin[sg] * HWi;
#pragma unroll 4
            for (int ci = 0; ci < p.cin[sg]; ++ci, xp += HWi) {            // 4 channels x k*k loads in flight per thread
                const float4* wq = scw + (size_t)(cbase + ci) * KK;
#pragma unroll
                for (int t = 0; t < KK; ++t) {
                    if (off[t] >= 0) {
                        const float v = __ldg(xp + off[t]);
                        const float4 q = wq[t];
                        acc[0] = fmaf(v, q.x, acc[0]); acc[1] = fmaf(v, q.y, acc[1]);
                        acc[2] = fmaf(v, q.z, acc[2]); acc[3] = fmaf(v, q.w, acc[3]);
                    }
                }
            }
            cbase += p.cin[sg];
        }
        for (int co = 0; co < p.cout; ++co) {
            const size_t o = ((size_t)b * p.cout + co) * HWo + r;
            float v = acc[co];
            if (p.bias) v += __ldg(p.bias + co);
            if (p.addend) v += __ldg(p.addend + o);
            p.y[o] = v;
        }
    }
}

// The RGB layer itself (k3, stride 1, pad 1, width a multiple of 4): a thread computes 4 consecutive output pixels of a
// row, so a (channel, row) costs one aligned float4 plus two edge loads for 4 x 3 taps instead of 12 loads.  The input
// channels are dealt round-robin to 4 thread groups of the CTA (64 strips x 4 groups) and summed through shared memory:
// 4x the threads of a pixel-only mapping, which this latency-bound loop needs to fill the SMs.
__global__ void __launch_bounds__(256) conv_small_cout_k3s1_kernel(const SmallConvParams p) {
    extern __shared__ float4 scw[];                 // [seg ci][tap] -> (w of co 0..3), then 64 x 4 x 4 float4 partials
    const int ctot = p.cin[0] + (p.nseg > 1 ? p.cin[1] : 0);
    float4* part = scw + (size_t)ctot * 9;
    for (int e = threadIdx.x; e < ctot * 9; e += blockDim.x) {
        const int c = e / 9, t = e % 9;
        const int sg = c >= p.cin[0], ci = sg ? c - p.cin[0] : c;
        float q[4] = {0.f, 0.f, 0.f, 0.f};
        // a transposed k3 s1 p1 convolution is the plain one with the taps mirrored (ky -> 2 - ky, kx -> 2 - kx)
        for (int co = 0; co < p.cout; ++co) q[co] = small_w(p, p.w[sg], p.cin[sg], co, ci, p.transposed ? 8 - t : t);
        scw[e] = make_float4(q[0], q[1], q[2], q[3]);
    }
    __syncthreads();
    const int W = p.Wi, H = p.Hi, HW = H * W, W4 = W / 4;
    const long long total = (long long)p.B * H * W4;
    const int strip = threadIdx.x & 63, grp = threadIdx.x >> 6;
    for (long long m0 = (long long)blockIdx.x * 64; m0 < total; m0 += (long long)gridDim.x * 64) {
        const long long m = m0 + strip;
        const bool ok = m < total;
        const int b = ok ? (int)(m / (H * W4)) : 0, r = ok ? (int)(m % (H * W4)) : 0, oy = r / W4, ox0 = 4 * (r % W4);
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
        int cbase = 0;
        for (int sg = 0; sg < p.nseg; ++sg) {
            const float* xp = p.x[sg] + ((size_t)b * p.cin[sg] + grp) * HW + ox0;
#pragma unroll 2
            for (int ci = grp; ci < p.cin[sg] && ok; ci += 4, xp += 4 * (size_t)HW) {
                const float4* wq = scw + (size_t)(cbase + ci) * 9;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const int iy = oy - 1 + dy;
                    if (iy < 0 || iy >= H) continue;
                    const float* row = xp + iy * W;
                    const float4 c4 = __ldg(reinterpret_cast<const float4*>(row));
                    const float v[6] = {ox0 > 0 ? __ldg(row - 1) : 0.f, c4.x, c4.y, c4.z, c4.w, ox0 + 4 < W ? __ldg(row + 4) : 0.f};
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float4 q = wq[dy * 3 + kx];
#pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            acc[px][0] = fmaf(v[px + kx], q.x, acc[px][0]); acc[px][1] = fmaf(v[px + kx], q.y, acc[px][1]);
                            acc[px][2] = fmaf(v[px + kx], q.z, acc[px][2]); acc[px][3] = fmaf(v[px + kx], q.w, acc[px][3]);
                        }
                    }
                }
            }
            cbase += p.cin[sg];
        }
        // sum the 4 channel groups: partial [group][pixel of strip][strip] as float4 over co
        __syncthreads();
#pragma unroll
        for (int px = 0; px < 4; ++px) part[(grp * 4 + px) * 64 + strip] = make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
        __syncthreads();
        if (ok && grp < p.cout) {                   // group g finishes output channel g
            const int co = grp;
            float o4[4];
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                float sum = 0.f;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 q = part[(g * 4 + px) * 64 + strip];
                    sum += co == 0 ? q.x : (co == 1 ? q.y : (co == 2 ? q.z : q.w));
                }
                o4[px] = sum;
            }
            const size_t o = ((size_t)b * p.cout + co) * HW + (size_t)oy * W + ox0;
            float4 v = make_float4(o4[0], o4[1], o4[2], o4[3]);
            if (p.bias) { const float bb = __ldg(p.bias + co); v.x += bb; v.y += bb; v.z += bb; v.w += bb; }
            if (p.addend) { const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.addend + o)); v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w; }
            *reinterpret_cast<float4*>(p.y + o) = v;
        }
    }
}

template <int K>
__global__ void __launch_bounds__(256) conv_small_cin_kernel(const SmallConvParams p) {
    // weights as float4 over 4 consecutive output channels: [co / 4][ci][tap] -> one broadcast LDS.128 feeds 4 FMAs (with
    // scalar weights the loop issued one shared-memory load per FMA and ran at the LDS rate)
    extern __shared__ float4 scv4[];
    constexpr int KK = K * K;
    const int cin = p.cin[0];
    const int co4n = (p.cout + 3) / 4;
    for (int e = threadIdx.x; e < co4n * cin * KK; e += blockDim.x) {
        const int t = e % KK, ci = (e / KK) % cin, c4 = e / (KK * cin);
        float q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = (4 * c4 + j < p.cout) ? small_w(p, p.w[0], cin, 4 * c4 + j, ci, t) : 0.f;
        scv4[e] = make_float4(q[0], q[1], q[2], q[3]);
    }
    __syncthreads();
    const int HWo = p.Ho * p.Wo, HWi = p.Hi * p.Wi;
    const long long total = (long long)p.B * HWo;
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < total; m += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(m / HWo), r = (int)(m % HWo), oy = r / p.Wo, ox = r % p.Wo;
        float v[4][KK];
        const float* xp = p.x[0] + (size_t)b * cin * HWi;
#pragma unroll
        for (int t = 0; t < KK; ++t) {
            const int off = small_tap_off(p, oy, ox, t / K, t % K);
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) v[ci][t] = (off >= 0 && ci < cin) ? __ldg(xp + (size_t)ci * HWi + off) : 0.f;
        }
        for (int c4 = 0; c4 < co4n; ++c4) {
            const float4* wr = scv4 + (size_t)c4 * cin * KK;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
                if (ci < cin) {
#pragma unroll
                    for (int t = 0; t < KK; ++t) {
                        const float4 q = wr[ci * KK + t];
                        a0 = fmaf(v[ci][t], q.x, a0); a1 = fmaf(v[ci][t], q.y, a1);
                        a2 = fmaf(v[ci][t], q.z, a2); a3 = fmaf(v[ci][t], q.w, a3);
                    }
                }
            }
            const float acc[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int co = 4 * c4 + j;
                if (co < p.cout) {
                    const size_t o = ((size_t)b * p.cout + co) * HWo + r;
                    float a = acc[j];
                    if (p.bias) a += __ldg(p.bias + co);
                    if (p.addend) a += __ldg(p.addend + o);
                    p.y[o] = a;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 1x1 convolution with a narrow output (5 <= cout <= 24): conv1 / conv2 of SpectralTransform (spectral_transform.py:52-53,
// 70-71: in_cg -> out_cg / 2 -> out_cg with 8-48 channels) and their data gradients.  Pure bandwidth: a thread owns 4
// consecutive pixels, reads one float4 per input channel (coalesced) and keeps 4 x cout accumulators in registers; the
// weights sit in shared memory as float4 over 4 output channels (one broadcast LDS.128 per 16 FMAs).
// ---------------------------------------------------------------------------------------------
template <int C4>           // groups of 4 output channels
__global__ void __launch_bounds__(256) conv1x1_narrow_kernel(const SmallConvParams p) {
    extern __shared__ float4 w4[];                  // [ci][C4]
    const int cin = p.cin[0];
    for (int e = threadIdx.x; e < cin * C4; e += blockDim.x) {
        const int ci = e / C4, g = e % C4;
        float q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = (4 * g + j < p.cout) ? small_w(p, p.w[0], cin, 4 * g + j, ci, 0) : 0.f;
        w4[e] = make_float4(q[0], q[1], q[2], q[3]);
    }
    __syncthreads();
    const int HW = p.Hi * p.Wi, HW4 = HW / 4;
    const long long total = (long long)p.B * HW4;
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < total; m += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(m / HW4), r4 = (int)(m % HW4);
        const float4* xp = reinterpret_cast<const float4*>(p.x[0] + (size_t)b * cin * HW) + r4;
        float4 acc[4 * C4];
#pragma unroll
        for (int i = 0; i < 4 * C4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int ci = 0; ci < cin; ++ci) {
            const float4 v = __ldg(xp + (size_t)ci * HW4);
#pragma unroll
            for (int g = 0; g < C4; ++g) {
                const float4 q = w4[ci * C4 + g];
                float4& a0 = acc[4 * g], &a1 = acc[4 * g + 1], &a2 = acc[4 * g + 2], &a3 = acc[4 * g + 3];
                a0.x = fmaf(v.x, q.x, a0.x); a0.y = fmaf(v.y, q.x, a0.y); a0.z = fmaf(v.z, q.x, a0.z); a0.w = fmaf(v.w, q.x, a0.w);
                a1.x = fmaf(v.x, q.y, a1.x); a1.y = fmaf(v.y, q.y, a1.y); a1.z = fmaf(v.z, q.y, a1.z); a1.w = fmaf(v.w, q.y, a1.w);
                a2.x = fmaf(v.x, q.z, a2.x); a2.y = fmaf(v.y, q.z, a2.y); a2.z = fmaf(v.z, q.z, a2.z); a2.w = fmaf(v.w, q.z, a2.w);
                a3.x = fmaf(v.x, q.w, a3.x); a3.y = fmaf(v.y, q.w, a3.y); a3.z = fmaf(v.z, q.w, a3.z); a3.w = fmaf(v.w, q.w, a3.w);
            }
        }
#pragma unroll
        for (int co = 0; co < 4 * C4; ++co) {
            if (co < p.cout) {
                float4 a = acc[co];
                const size_t o4 = ((size_t)b * p.cout + co) * HW4 + r4;
                if (p.bias) { const float bb = __ldg(p.bias + co); a.x += bb; a.y += bb; a.z += bb; a.w += bb; }
                if (p.addend) { const float4 d = __ldg(reinterpret_cast<const float4*>(p.addend) + o4); a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w; }
                reinterpret_cast<float4*>(p.y)[o4] = a;
            }
        }
    }
}

bool conv1x1_narrow_supported(int cin0, int cin1, int cout, int k, int stride, int pad, int Hi, int Wi, int Ho, int Wo,
                              const void* x, const void* y, const void* addend) {
    return k == 1 && stride == 1 && pad == 0 && cin1 == 0 && cout >= 5 && cout <= 24 && cin0 >= 1 && cin0 <= 512 && Ho == Hi && Wo == Wi &&
           (Hi * Wi) % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)addend)) & 15) == 0;
}

int conv1x1_narrow_run(const float* x, const float* w, int cin, const float* bias, const float* addend, float* y,
                       int B, int cout, int Hi, int Wi, int transposed, ffc_stream_t st) {
    SmallConvParams p;
    p.x[0] = x; p.x[1] = nullptr; p.w[0] = w; p.w[1] = nullptr; p.cin[0] = cin; p.cin[1] = 0; p.nseg = 1;
    p.bias = bias; p.addend = addend; p.y = y; p.B = B; p.cout = cout; p.Hi = Hi; p.Wi = Wi; p.Ho = Hi; p.Wo = Wi;
    p.k = 1; p.stride = 1; p.pad = 0; p.transposed = transposed;
    const long long total = (long long)B * (Hi * Wi / 4);
    int grid = (int)((total + 255) / 256); if (grid > ffc_sm_count() * 8) grid = ffc_sm_count() * 8; if (grid < 1) grid = 1;
    const int c4 = (cout + 3) / 4;
    const size_t smem = (size_t)cin * c4 * 16;
    switch (c4) {
        case 2: conv1x1_narrow_kernel<2><<<grid, 256, smem, st>>>(p); break;
        case 3: conv1x1_narrow_kernel<3><<<grid, 256, smem, st>>>(p); break;
        case 4: conv1x1_narrow_kernel<4><<<grid, 256, smem, st>>>(p); break;
        case 5: conv1x1_narrow_kernel<5><<<grid, 256, smem, st>>>(p); break;
        default: conv1x1_narrow_kernel<6><<<grid, 256, smem, st>>>(p); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("conv1x1_narrow launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}

bool conv_small_supported(int cin0, int cin1, int cout, int k) {
    if (k != 1 && k != 3 && k != 4) return false;
    if (cout <= 4) return (size_t)(cin0 + cin1) * k * k * 16 <= 96 * 1024;
    return cin1 == 0 && cin0 <= 4 && (size_t)((cout + 3) / 4) * cin0 * k * k * 16 <= 96 * 1024;
}

template <int K>
static int conv_small_launch(const SmallConvParams& p, ffc_stream_t st) {
    const long long total = (long long)p.B * p.Ho * p.Wo;
    int grid = (int)((total + 255) / 256); if (grid > ffc_sm_count() * 16) grid = ffc_sm_count() * 16; if (grid < 1) grid = 1;
    cudaError_t e;
    const uintptr_t al = (uintptr_t)p.x[0] | (uintptr_t)p.x[1] | (uintptr_t)p.y | (uintptr_t)p.addend;
    if (p.cout <= 4 && K == 3 && p.stride == 1 && p.pad == 1 && p.Ho == p.Hi && p.Wo == p.Wi && p.Wi % 4 == 0 && (al & 15) == 0) {
        const size_t smem = (size_t)(p.cin[0] + (p.nseg > 1 ? p.cin[1] : 0)) * 9 * 16 + 64 * 16 * 16;
        if (smem > 48 * 1024) cudaFuncSetAttribute(conv_small_cout_k3s1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int g4 = (int)((total / 4 + 63) / 64); if (g4 > ffc_sm_count() * 16) g4 = ffc_sm_count() * 16; if (g4 < 1) g4 = 1;
        conv_small_cout_k3s1_kernel<<<g4, 256, smem, st>>>(p);
    } else if (p.cout <= 4) {
        const size_t smem = (size_t)(p.cin[0] + (p.nseg > 1 ? p.cin[1] : 0)) * K * K * 16;
        if (smem > 48 * 1024) cudaFuncSetAttribute(conv_small_cout_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv_small_cout_kernel<K><<<grid, 256, smem, st>>>(p);
    } else {
        const size_t smem = (size_t)((p.cout + 3) / 4) * p.cin[0] * K * K * 16;
        if (smem > 48 * 1024) cudaFuncSetAttribute(conv_small_cin_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        conv_small_cin_kernel<K><<<grid, 256, smem, st>>>(p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("conv_small launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}

int conv_small_run(const float* x0, const float* w0, int cin0, const float* x1, const float* w1, int cin1,
                   const float* bias, const float* addend, float* y, int B, int cout, int Hi, int Wi, int Ho, int Wo,
                   int k, int stride, int pad, int transposed, ffc_stream_t st) {
    SmallConvParams p;
    p.x[0] = x0; p.x[1] = x1; p.w[0] = w0; p.w[1] = w1; p.cin[0] = cin0; p.cin[1] = cin1; p.nseg = x1 ? 2 : 1;
    p.bias = bias; p.addend = addend; p.y = y; p.B = B; p.cout = cout; p.Hi = Hi; p.Wi = Wi; p.Ho = Ho; p.Wo = Wo;
    p.k = k; p.stride = stride; p.pad = pad; p.transposed = transposed;
    if (k == 1) return conv_small_launch<1>(p, st);
    if (k == 3) return conv_small_launch<3>(p, st);
    return conv_small_launch<4>(p, st);
}

// ---------------------------------------------------------------------------------------------
// weight gradient with SC <= 4:  dW[sc][lc][ky][kx] = sum_{b,y,x} S[b,sc,y,x] * L[b,lc,y*s-pad+ky,x*s-pad+kx]
// grid = (LC, image groups); a thread walks pixels of its images with SC*K*K partial sums in registers
// ---------------------------------------------------------------------------------------------
struct SmallWgradParams { const float* S; const float* L; float* dW; int B, SC, LC, Hs, Ws, Hl, Wl, stride, pad, imgs; };

template <int K>
__global__ void __launch_bounds__(256) wgrad_small_sc_kernel(const SmallWgradParams p) {
    constexpr int KK = K * K;
    __shared__ float red[8][4 * KK];
    const int lc = blockIdx.x;
    const int HWs = p.Hs * p.Ws, HWl = p.Hl * p.Wl;
    const int b0 = blockIdx.y * p.imgs;
    const int b1 = (b0 + p.imgs) < p.B ? (b0 + p.imgs) : p.B;
    float acc[4][KK];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int t = 0; t < KK; ++t) acc[s][t] = 0.f;
    for (int r = threadIdx.x; r < HWs; r += blockDim.x) {
        const int y = r / p.Ws, x = r % p.Ws;
        int off[KK];
#pragma unroll
        for (int t = 0; t < KK; ++t) {
            const int ly = y * p.stride - p.pad + t / K, lx = x * p.stride - p.pad + t % K;
            off[t] = (ly >= 0 && ly < p.Hl && lx >= 0 && lx < p.Wl) ? ly * p.Wl + lx : -1;
        }
        for (int b = b0; b < b1; ++b) {
            const float* lp = p.L + ((size_t)b * p.LC + lc) * HWl;
            const float* sp = p.S + (size_t)b * p.SC * HWs + r;
            float sv[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) sv[s] = s < p.SC ? __ldg(sp + (size_t)s * HWs) : 0.f;
#pragma unroll
            for (int t = 0; t < KK; ++t) {
                if (off[t] >= 0) {
                    const float v = __ldg(lp + off[t]);
#pragma unroll
                    for (int s = 0; s < 4; ++s) acc[s][t] = fmaf(sv[s], v, acc[s][t]);
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int t = 0; t < KK; ++t) {
            float v = acc[s][t];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) red[warp][s * KK + t] = v;
        }
    __syncthreads();
    for (int e = threadIdx.x; e < p.SC * KK; e += blockDim.x) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][e];
        const int s = e / KK, t = e % KK;
        atomicAdd(p.dW + ((size_t)s * p.LC + lc) * KK + t, v);
    }
}

// k3, stride 1, pad 1, width a multiple of 4 (the RGB layer): 4 consecutive pixels per thread, see conv_small_cout_k3s1_kernel
__global__ void __launch_bounds__(256) wgrad_small_sc_k3s1_kernel(const SmallWgradParams p) {
    __shared__ float red[8][4 * 9];
    const int lc = blockIdx.x;
    const int W = p.Ws, H = p.Hs, HW = H * W, W4 = W / 4;
    const int b0 = blockIdx.y * p.imgs;
    const int b1 = (b0 + p.imgs) < p.B ? (b0 + p.imgs) : p.B;
    float acc[4][9];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[s][t] = 0.f;
    for (int r = threadIdx.x; r < H * W4; r += blockDim.x) {
        const int y = r / W4, x0 = 4 * (r % W4);
        for (int b = b0; b < b1; ++b) {
            const float* lp = p.L + ((size_t)b * p.LC + lc) * HW + x0;
            const float* sp = p.S + (size_t)b * p.SC * HW + (size_t)y * W + x0;
            float4 sv[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) sv[s] = s < p.SC ? __ldg(reinterpret_cast<const float4*>(sp + (size_t)s * HW)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int ly = y - 1 + dy;
                if (ly < 0 || ly >= H) continue;
                const float* row = lp + ly * W;
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(row));
                const float v[6] = {x0 > 0 ? __ldg(row - 1) : 0.f, c4.x, c4.y, c4.z, c4.w, x0 + 4 < W ? __ldg(row + 4) : 0.f};
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int s = 0; s < 4; ++s)
                        acc[s][dy * 3 + kx] += sv[s].x * v[kx] + sv[s].y * v[kx + 1] + sv[s].z * v[kx + 2] + sv[s].w * v[kx + 3];
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            float v = acc[s][t];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) red[warp][s * 9 + t] = v;
        }
    __syncthreads();
    for (int e = threadIdx.x; e < p.SC * 9; e += blockDim.x) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][e];
        atomicAdd(p.dW + ((size_t)(e / 9) * p.LC + lc) * 9 + e % 9, v);
    }
}

bool wgrad_small_supported(int SC, int k) { return SC <= 4 && (k == 1 || k == 3 || k == 4); }

// dW zeroed by the caller
int wgrad_small_run(const float* S, const float* L, float* dW, int B, int SC, int LC, int Hs, int Ws, int Hl, int Wl,
                    int k, int stride, int pad, ffc_stream_t st) {
    SmallWgradParams p{S, L, dW, B, SC, LC, Hs, Ws, Hl, Wl, stride, pad, 1};
    // ~4 CTAs per SM; each CTA reduces `imgs` images of one large-side channel
    int groups = ffc_cdiv(4 * ffc_sm_count(), LC); if (groups > B) groups = B; if (groups < 1) groups = 1;
    p.imgs = ffc_cdiv(B, groups);
    groups = ffc_cdiv(B, p.imgs);
    const dim3 grid(LC, groups);
    if (k == 3 && stride == 1 && pad == 1 && Hs == Hl && Ws == Wl && Ws % 4 == 0 && (((uintptr_t)S | (uintptr_t)L) & 15) == 0)
        wgrad_small_sc_k3s1_kernel<<<grid, 256, 0, st>>>(p);
    else if (k == 1) wgrad_small_sc_kernel<1><<<grid, 256, 0, st>>>(p);
    else if (k == 3) wgrad_small_sc_kernel<3><<<grid, 256, 0, st>>>(p);
    else wgrad_small_sc_kernel<4><<<grid, 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ffc_set_error("wgrad_small launch failed: %s", cudaGetErrorString(e)); return FFC_ERR_CUDA; }
    ffc_count_launch();
    return FFC_OK;
}
#endif  // !FFC_EMU
