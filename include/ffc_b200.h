/* ffc_b200.h -- C ABI of libffc_b200.so (hand-written sm_100a CUDA for the FFC hot path).
 *
 * This is the drop-in boundary for the Fast-Fourier-Convolution layer stack of
 * phbgomes22/FastFourierConvolution (layers/ffc/*.py, layers/snffc/*.py).  The reference is pure
 * PyTorch, so "the reference's FFI for this path" is the set of ATen library calls its modules make;
 * every entry point below names the reference call site(s) it replaces (paths relative to the
 * reference checkout).  The host side (fastfourierconvolution_b200/layers/*) mirrors the reference's
 * nn.Module API on top of these functions through ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - All tensors are contiguous FP32 NCHW device buffers owned by the caller (PyTorch's caching
 *    allocator); the library never allocates, frees or retains pointers.
 *  - `stream` is a cudaStream_t; work is only enqueued, never synchronised.
 *  - `workspace` is caller-provided scratch; the required size is documented per function and
 *    ffc_workspace_bytes() returns a bound that is sufficient for every function.
 *  - Return value: 0 on success, non-zero on error (1 bad argument / unsupported configuration,
 *    2 workspace too small, 3 CUDA error); ffc_last_error() returns the message (thread local).
 *  - Re-entrant, no global mutable state besides the thread-local error string.
 */
#ifndef FFC_B200_H_
#define FFC_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* activation codes (ffc_bn_act.py:63-67: nn.Identity, nn.ReLU, nn.LeakyReLU(0.1), nn.GELU (erf), nn.Tanh, nn.Sigmoid) */
#define FFC_B200_ACT_IDENTITY 0
#define FFC_B200_ACT_RELU     1
#define FFC_B200_ACT_LEAKY    2
#define FFC_B200_ACT_GELU     3
#define FFC_B200_ACT_TANH     4
#define FFC_B200_ACT_SIGMOID  5

/* resampling modes of SpectralTransform.downsample (spectral_transform.py:42-47) */
#define FFC_B200_RESAMPLE_NONE     0
#define FFC_B200_RESAMPLE_UP2      1   /* nn.Upsample(scale_factor=2, mode='nearest') */
#define FFC_B200_RESAMPLE_AVGPOOL2 2   /* nn.AvgPool2d(kernel_size=(2,2), stride=2)   */

int ffc_version(void);                 /* major*1000 + minor*10 + patch */
const char* ffc_last_error(void);
int ffc_is_emulation(void);            /* 1 only for the host emulation build used by tests/ */
unsigned long long ffc_launch_count(void);   /* kernels launched by this library since load */

/* Upper bound of the scratch needed by any entry point for a problem with `batch` images and
 * `max_channels` channels (bytes). */
size_t ffc_workspace_bytes(int batch, int max_channels);

/* ---- Fourier unit, general form (spectrum staged through L2) ------------------------------------
 * ffc_rfft2: replaces torch.fft.rfftn(x, dim=(-2,-1), norm="ortho") + stack/permute/contiguous/view
 *            (layers/ffc/fourier_unity.py:38-42).
 *   x    (nplanes, H, W)          nplanes = B*C
 *   spec (nplanes, 2, H, W/2+1)   == (B, 2C, H, Wf) with channel 2c = Re, 2c+1 = Im.
 *   Along u the bins of H in {64,128} are stored in the library's internal permuted order; the
 *   consumers between ffc_rfft2 and ffc_irfft2 (1x1 conv, BatchNorm, ReLU) are pointwise in (u,v).
 *   colscale = 1 multiplies columns 0 < v < W/2 by 2: this is the adjoint of ffc_irfft2
 *   (autograd of torch.fft.irfftn, fourier_unity.py:56).
 * ffc_irfft2: replaces view/permute/contiguous/torch.complex + torch.fft.irfftn(s=(H,W), norm="ortho")
 *            (fourier_unity.py:51-56); imaginary parts of columns 0 and W/2 are ignored exactly as
 *            cuFFT/pocketfft C2R do.  out = residual + irfft2(spec) when residual != NULL
 *            (the x + fu(x) of spectral_transform.py:108).
 *   colscale = 1 pre-multiplies columns 0 < v < W/2 by 1/2: the adjoint of ffc_rfft2.
 * Supported planes: H == W in {4,8,16,32,64,128} on the tuned kernels (16-byte aligned x / out / residual); every other
 * H, W in 1..128 (odd, non-square, 48x48 of the mg = 6 scripts, fgan_cond_complete.py:325 -- torch.fft accepts any size)
 * on direct-DFT plane kernels (csrc/ffc_dft2.cu), natural order along u, "interior" columns = those a C2R counts twice
 * (0 < v, 2v != W).  ffc_fft2_supported: 2 = tuned, 1 = direct DFT, 0 = unsupported. */
int ffc_fft2_supported(int H, int W);
int ffc_rfft2(const float* x, float* spec, int nplanes, int H, int W, int colscale, void* stream);
int ffc_irfft2(const float* spec, const float* residual, float* out, int nplanes, int H, int W,
               int colscale, void* stream);

/* ---- Fourier unit, fused form (spectrum never leaves shared memory) ------------------------------
 * ffc_fu_fwd replaces the whole of FourierUnitSN.forward (layers/ffc/fourier_unity.py:32-58):
 *   out = [residual +] irfft2( relu( BatchNorm2d( conv_layer( rfft2(x) ) ) ) )
 * in one kernel per tile (eval mode) or two passes over x (training mode: batch statistics first).
 *   x (B,Cin,H,W); w = conv_layer.weight viewed [2*Cout][2*Cin]; gamma/beta/running_* = bn.* [2*Cout];
 *   save_mean / save_invstd [2*Cout] are written (batch statistics in training, running statistics in eval);
 *   residual (B,Cout,H,W) or NULL (the `x + fu(x)` of spectral_transform.py:108); out (B,Cout,H,W).
 *   training = 1 also updates running_mean / running_var in place (momentum, unbiased variance).
 *   workspace >= 4*Cout*8 bytes.
 * ffc_fu_fused_supported returns 1 for shapes this entry point handles (H == W in {4,8,16,32},
 * Cin, Cout <= 32 and the tile fits in shared memory); other shapes use the general form above. */
int ffc_fu_fused_supported(int B, int Cin, int Cout, int H, int W);
/* Workspace ffc_fu_fwd makes the best use of (the minimum stays 4*Cout doubles): with room for one float per (sum, image)
 * the training kernel of the 32x32 / <= 8 channel unit (csrc/ffc_fu4.cu) meets its batch statistics through per-CTA partial
 * sums -- no zeroing launch, no atomics, bitwise reproducible. */
size_t ffc_fu_workspace_bytes(int B, int Cout);
int ffc_fu_fwd(const float* x, const float* w, const float* gamma, const float* beta,
               float* running_mean, float* running_var, float* save_mean, float* save_invstd,
               const float* residual, float* out,
               int B, int Cin, int Cout, int H, int W, int training, float eps, float momentum,
               void* workspace, size_t workspace_bytes, void* stream);

/* ---- Fourier unit, L2-staged form (planes / channel counts beyond one CTA's shared memory) -------------
 * ffc_fu3_fwd: the same contract as ffc_fu_fwd (replaces FourierUnitSN.forward, layers/ffc/fourier_unity.py:32-58) for
 * H == W in {16,32,64,128} and any Cin / Cout whose packed mix weights fit shared memory (2*Cout <= 128 on the tensor-core
 * path): three kernels per chunk of images -- plane rfft2 | tcgen05 channel mix with the BatchNorm statistics (training) or
 * BatchNorm + ReLU (eval) in its epilogue | plane irfft2 (BatchNorm + ReLU applied on load in training mode) -- with the
 * spectrum of a chunk staged in a caller-provided scratch that is sized to stay resident in the 126 MB L2 and reused
 * by every chunk; in training mode the mixed spectrum of the whole batch makes one round trip (it waits for the batch
 * statistics).  workspace >= ffc_fu3_workspace_bytes(B, Cin, Cout, H, W, training) bytes.
 * ffc_debug_fu3_simt_mix(1) routes the channel mix through the plain FP32 kernel (cross-check of the tensor-core one). */
int ffc_fu3_supported(int B, int Cin, int Cout, int H, int W);
size_t ffc_fu3_workspace_bytes(int B, int Cin, int Cout, int H, int W, int training);
int ffc_fu3_fwd(const float* x, const float* w, const float* gamma, const float* beta,
                float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                const float* residual, float* out,
                int B, int Cin, int Cout, int H, int W, int training, float eps, float momentum,
                void* workspace, size_t workspace_bytes, void* stream);
/* ffc_fu3_fwd_keep / ffc_fu3_bwd: FourierUnitSN forward that keeps its two spectra, and the backward that autograd derives for
 * layers/ffc/fourier_unity.py:32-58 (cuFFT c2r/r2c adjoints, cudnn BatchNorm backward, ReLU mask, 1x1 conv dgrad / wgrad) as
 * five kernels that recompute nothing.  s_keep (B, Cin, H, W + 4) and y_keep (B, Cout, H, W + 4) floats are opaque (the
 * library's plane layout).  dx may be null; the gradient with respect to the residual is dout itself.
 * workspace >= ffc_fu3_bwd_workspace_bytes(B, Cin, Cout, H, W). */
int ffc_fu3_fwd_keep(const float* x, const float* w, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                     const float* residual, float* out, float* s_keep, float* y_keep,
                     int B, int Cin, int Cout, int H, int W, int training, float eps, float momentum,
                     void* workspace, size_t workspace_bytes, void* stream);
size_t ffc_fu3_bwd_workspace_bytes(int B, int Cin, int Cout, int H, int W);
int ffc_fu3_bwd(const float* dout, const float* s_keep, const float* y_keep, const float* w, const float* gamma, const float* beta,
                const float* save_mean, const float* save_invstd, float* dx, float* dw, float* dgamma, float* dbeta,
                int B, int Cin, int Cout, int H, int W, int training, void* workspace, size_t workspace_bytes, void* stream);
/* ---- Glue between the FFC layers of the reference's generators, one bandwidth-bound kernel each ----
 * ffc_noise_add_fwd: NoiseInjection.forward (layers/noise_injection.py:20-32; called on both branches after every
 *   upsampling stage, fgan_complete.py:122-131): out = x + weight[c] * noise[b, 0, h, w]; x / out (B, C, HW), noise (B, HW).
 * ffc_noise_add_bwd_w: its weight gradient dweight[c] = sum_{b,hw} dy * noise (dx = dy passes through unchanged).
 * ffc_to_uint8: the eval-mode image epilogue (fgan_complete.py:136-138): out = uint8(255 * (clamp(x, lo, hi) * 0.5 + 0.5));
 *   lo > hi disables the clamp (fgan64_complete.py:150-153 clamps to the tensor's own min / max, an identity). */
int ffc_noise_add_fwd(const float* x, const float* w, const float* noise, float* out, int B, int C, int HW, void* stream);
int ffc_noise_add_bwd_w(const float* dy, const float* noise, float* dw, int B, int C, int HW, void* stream);
int ffc_to_uint8(const float* x, unsigned char* out, long long n, float lo, float hi, void* stream);

/* ---- Linear layers and the optimiser of the reference's training scripts ----
 * ffc_gemm_f32: C (M x N) = A (M x K) * B (K x N) [+ bias[n]] in FP32 with arbitrary element strides: the Linear stem of the
 *   generators (fgan_complete.py:92-95, 117-119: x W^T + b), the SN Linear head of the discriminators (:160-170) and their
 *   gradients dy^T x and dy W are the same kernel.  Skinny products split K (C must then be dense row-major).
 * ffc_colsum_f32: out[n] = sum_m x[m][n] (bias gradient).
 * ffc_adam_step: optim.AdamW / optim.Adam (fgan_complete.py:315-319, sngan_complete.py:247-248) over FLAT parameter, gradient
 *   and moment buffers: one kernel per step for any number of tensors; lr and the step count live on the device (CUDA-graph
 *   replay); grad_scale multiplies the gradient first (1 / world size after a SUM all-reduce); decoupled = 1: AdamW. */
int ffc_gemm_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                 long long sam, long long sak, long long sbk, long long sbn, long long scm, long long scn, void* stream);
int ffc_colsum_f32(const float* x, float* out, int M, int N, void* stream);
int ffc_adam_step(float* p, const float* g, float* m, float* v, long long n, const float* lr, float* step,
                  float beta1, float beta2, float eps, float weight_decay, float grad_scale, int decoupled, void* stream);
/* ffc_adam_step_table: the same step with the gradients left where autograd put them: tensor t's gradient is read through
 *   gptr[t] (device array of device pointers; null = no gradient this step, tensor skipped), its parameters / moments are
 *   offs[t] .. offs[t] + sizes[t] of the flat buffers; block i handles 4096-element piece blk_piece[i] of tensor blk_tensor[i];
 *   step is an array of ntensors counts (torch keeps state["step"] per parameter: a skipped tensor does not age).
 * ffc_gather_table: with the same tables, copies the gradients into one flat buffer (the packing pass before an all-reduce). */
int ffc_adam_step_table(float* p, float* m, float* v, const float* const* gptr, const long long* offs, const long long* sizes,
                        const int* blk_tensor, const int* blk_piece, int ntensors, int nblocks, const float* lr, float* step,
                        float beta1, float beta2, float eps, float weight_decay, float grad_scale, int decoupled, void* stream);
int ffc_gather_table(float* dst, const float* const* gptr, const long long* offs, const long long* sizes,
                     const int* blk_tensor, const int* blk_piece, int nblocks, void* stream);

void ffc_debug_fu3_simt_mix(int on);
void ffc_debug_fu3_chunk_bytes(size_t bytes);      /* spectrum bytes per chunk of images (0 = default 160 MB); tuning / tests */

/* ffc_fu_bwd: the autograd backward of FourierUnitSN.forward (fourier_unity.py:32-58; derived from ATen's
 * fft_r2c / fft_c2r / batch_norm / relu / conv backward formulas) as ONE cooperative kernel, one image per CTA:
 * rfft2(x) and the adjoint-c2r transform of dout stay in shared memory, Y = W S is recomputed, ReLU mask and the two
 * BatchNorm-backward sums cross a grid barrier, then dW (atomics), dS = W^T dY and the adjoint-r2c inverse give dx.
 *   x, dout as in the forward; save_mean / save_invstd = what ffc_fu_fwd wrote; training selects batch-statistics
 *   BatchNorm backward; outputs dx (B,Cin,H,W), dw [2*Cout][2*Cin], dgamma, dbeta [2*Cout] (all overwritten).
 *   The residual's gradient is dout itself.  workspace >= 4*Cout*8 bytes.
 * ffc_fu_bwd_supported: 1 when the shape is handled on the current device (H == W in {8,16,32}, Cin, Cout <= 32,
 * all B images co-resident); otherwise callers compose the general-form entry points. */
int ffc_fu_bwd_supported(int B, int Cin, int Cout, int H, int W);
int ffc_fu_bwd(const float* x, const float* dout, const float* w, const float* gamma, const float* beta,
               const float* save_mean, const float* save_invstd,
               float* dx, float* dw, float* dgamma, float* dbeta,
               int B, int Cin, int Cout, int H, int W, int training,
               void* workspace, size_t workspace_bytes, void* stream);
/* Workspace ffc_fu_bwd makes the best use of (the minimum stays 4*Cout doubles): with room for the per-image partial sums
 * and weight-gradient tiles the 32x32 / <= 8 channel unit runs the warp-private backward of csrc/ffc_fu4.cu (no zeroing
 * launches, no atomics, bitwise reproducible). */
size_t ffc_fu_bwd_workspace_bytes(int B, int Cin, int Cout);

/* ---- Convolutions -----------------------------------------------------------------------------
 * ffc_conv2d_fwd, transposed = 0: nn.Conv2d forward (layers/ffc/ffc.py:45-68 convl2l/convl2g/convg2l,
 *   spectral_transform.py:52-53,70-71 conv1/conv2, fourier_unity.py:23-24 conv_layer) and the
 *   data-gradient of nn.ConvTranspose2d.  Weight layout [cout][cin][k][k].
 * transposed = 1: nn.ConvTranspose2d forward (layers/ffc/ffc_transpose.py:84-86) and the
 *   data-gradient of nn.Conv2d.  Weight layout [cin][cout][k][k].
 * y = conv(x0, w0) [+ conv(x1, w1)] [+ bias] [+ addend]; the second segment fuses
 *   convl2l(x_l) + convg2l(x_g) (ffc.py:91, ffc_transpose.py:98-100); `addend` fuses the sum with the
 *   global branch (ffc.py:94-96).  groups = 1, dilation = 1, square kernel, stride in {1,2}. */
int ffc_conv2d_fwd(const float* x0, const float* w0, int cin0,
                   const float* x1, const float* w1, int cin1,
                   const float* bias, const float* addend, float* y,
                   int B, int cout, int Hi, int Wi, int Ho, int Wo,
                   int k, int stride, int pad, int transposed, void* stream);

/* ffc_conv2d_fwd_ws: same contract as ffc_conv2d_fwd plus caller-provided scratch of
 * ffc_conv2d_workspace_bytes(cin0, cin1, cout, k, stride, pad, transposed) bytes.  The scratch receives the
 * weights re-packed for this call (GEMM tile order, split into TF32 hi/lo halves) and lets the kernel stream both
 * operands with asynchronous copies; the host modules call this form. */
size_t ffc_conv2d_workspace_bytes(int cin0, int cin1, int cout, int k, int stride, int pad, int transposed);
int ffc_conv2d_fwd_ws(const float* x0, const float* w0, int cin0,
                      const float* x1, const float* w1, int cin1,
                      const float* bias, const float* addend, float* y,
                      int B, int cout, int Hi, int Wi, int Ho, int Wo,
                      int k, int stride, int pad, int transposed,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ffc_conv2d_block_fwd_ws: the local branches of FFC.forward (ffc.py:91-96, ffc_transpose.py:98-106) in one call:
 *   y0 (cout0 channels) = conv(x0, w00) + conv(x1, w10) [+ bias[0:cout0]]            convl2l(x_l) + convg2l(x_g)
 *   y1 (cout1 channels) = conv(x0, w01)                 [+ bias[cout0:cout0+cout1]]   convl2g(x_l)
 * On the tcgen05 path both outputs come from ONE implicit GEMM over the concatenated output channels, so the operand
 * gathered from x0 is shared.  x1 / w10 may be NULL.  workspace: ffc_conv2d_workspace_bytes(cin0, cin1, cout0 + cout1, ...). */
int ffc_conv2d_block_fwd_ws(const float* x0, const float* w00, const float* w01, int cin0,
                            const float* x1, const float* w10, int cin1, const float* bias,
                            float* y0, int cout0, float* y1, int cout1,
                            int B, int Hi, int Wi, int Ho, int Wo, int k, int stride, int pad, int transposed,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ffc_conv2d_act_fwd_ws: y = act(conv(x, w) + bias), act = FFC_ACT_LEAKY (slope > 0) or FFC_ACT_RELU -- one stage of the
 * SN conv discriminators the FFC generators are trained against (fgan_complete.py:162-168, ``act(convN(m))``;
 * SURVEY.md 8(f) rank 1).  nn.Conv2d semantics (not transposed, one segment).  The activation rides in the epilogue of the
 * tcgen05 kernel; the <= 4-input-channel RGB stage runs the convolution and then the elementwise kernel in place.
 * Both activations keep the sign of their argument, so the backward mask is taken from y (ffc_bn_act_bwd with x := y).
 * workspace: max(ffc_conv2d_workspace_bytes(cin, 0, cout, k, stride, pad, 0), 2*cout*8) bytes. */
int ffc_conv2d_act_fwd_ws(const float* x, const float* w, int cin, const float* bias, float* y,
                          int B, int cout, int Hi, int Wi, int Ho, int Wo, int k, int stride, int pad,
                          int act, float slope, void* workspace, size_t workspace_bytes, void* stream);

/* dW[sc][lc][ky][kx] = sum_{b,y,x} S[b,sc,y,x] * L[b,lc,y*stride-pad+ky,x*stride-pad+kx]
 * nn.Conv2d:          S = dy (cout, Ho x Wo), L = x  (cin,  Hi x Wi)  -> dW [cout][cin][k][k]
 * nn.ConvTranspose2d: S = x  (cin,  Hi x Wi), L = dy (cout, Ho x Wo)  -> dW [cin][cout][k][k]
 * dW is overwritten. */
int ffc_conv2d_wgrad(const float* S, const float* L, float* dW,
                     int B, int SC, int LC, int Hs, int Ws, int Hl, int Wl,
                     int k, int stride, int pad, void* stream);

/* db[c] = sum_{b,hw} dy[b,c,hw].  workspace >= 2*C*8 bytes. */
int ffc_bias_grad(const float* dy, float* db, int B, int C, int HW,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- BatchNorm2d + activation -----------------------------------------------------------------
 * Replaces nn.BatchNorm2d followed by the activation at layers/ffc/ffc_bn_act.py:73-81,
 * spectral_transform.py:89 and fourier_unity.py:49.
 *   norm = 0: y = act(x) (norm_layer = nn.Identity);  norm = 1: BatchNorm2d(eps, momentum, affine).
 *   training = 1: batch statistics (biased variance for normalisation, unbiased for running_var,
 *   running stats updated in place); training = 0: running statistics.
 *   save_mean / save_invstd [C] are written (norm = 1) and must be handed to ffc_bn_act_bwd.
 *   workspace >= 2*C*8 bytes. */
int ffc_bn_act_fwd(const float* x, float* y, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                   int B, int C, int HW, int norm, int training, float eps, float momentum,
                   int act, float slope, void* workspace, size_t workspace_bytes, void* stream);
int ffc_bn_act_bwd(const float* x, const float* dy, float* dx, const float* gamma, const float* beta,
                   const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta,
                   int B, int C, int HW, int norm, int training, int act, float slope,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ffc_bn_stats: the statistics half of ffc_bn_act_fwd (batch statistics + running-stat update in training mode, running
 * statistics in eval mode) -> save_mean / save_invstd [C].  workspace >= 2*C*8 bytes.
 * ffc_irfft2_bn_relu: out = irfft2(relu(batchnorm(spec))) [+ residual] -- fourier_unity.py:49 folded into :51-56, the
 * normalised spectrum never exists in memory.  spec (nplanes = B*cout, 2, H, W/2+1); mean / invstd / gamma / beta over the
 * 2*cout spectrum channels.  The backward is ffc_rfft2(colscale = 1) followed by ffc_bn_act_bwd on the saved spec. */
int ffc_bn_stats(const float* x, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                 int B, int C, int HW, int training, float eps, float momentum,
                 void* workspace, size_t workspace_bytes, void* stream);
int ffc_irfft2_bn_relu(const float* spec, const float* residual, float* out, int nplanes, int cout, int H, int W,
                       const float* mean, const float* invstd, const float* gamma, const float* beta, void* stream);

/* ---- spectral normalisation ----------------------------------------------------------------------
 * torch.nn.utils.spectral_norm's pre-forward computation (layers/snffc/snffc.py:8, 23-33; fgan_complete.py:147-156) on
 * W = weight_orig viewed as (h = out channels, w = the rest), one power iteration.  kk == 0: the weight is that matrix,
 * row-major (SpectralNorm.dim == 0: nn.Conv2d, nn.Linear); kk = k*k > 0: the weight is stored (w / kk, h, kk)
 * (dim == 1: nn.ConvTranspose2d, matrix view weight.permute(1, 0, 2, 3).reshape(out, -1)):
 *     power_iteration != 0:  v = normalize(W^T u);  u = normalize(W v)     (u, v updated in place, eps as in torch)
 *     sigma = u . (W v);     w_eff = W / sigma
 * u_save (h) / v_save (w) (nullable) receive the vectors sigma was computed with and sigma[0] its value: the caller keeps
 * them for the backward  dW = g / sigma - (sum(g * W) / sigma^2) * u v^T.
 * workspace: ffc_spectral_norm_workspace_bytes(h, w). */
size_t ffc_spectral_norm_workspace_bytes(int h, int w);
int ffc_spectral_norm_fwd(const float* w_orig, float* u, float* v, float* u_save, float* v_save,
                          float* w_eff, float* sigma, int h, int w, int kk, int power_iteration, float eps,
                          void* workspace, size_t workspace_bytes, void* stream);
/* backward through weight = W / sigma (sigma = u^T W v with the u, v the forward saved):
 *   dW = g / sigma - (sum(g * W) / sigma^2) * u v^T   -- torch.nn.utils.spectral_norm's autograd in two kernels.
 * g / w_orig / dw in the weight's own storage order (kk as in the forward); workspace >= 16 bytes. */
int ffc_spectral_norm_bwd(const float* g, const float* w_orig, const float* u, const float* v, const float* sigma,
                          float* dw, int h, int w, int kk, void* workspace, size_t workspace_bytes, void* stream);

/* ---- SE gate + resampling ---------------------------------------------------------------------
 * y = r(x) * sigmoid(W2 relu(W1 mean_hw(r(x)))): SpectralTransform.downsample followed by SELayer
 * (spectral_transform.py:79, 87 -> :12-28).  w1 [hid][C], w2 [C][hid] (hid = C // 16, may be 0).
 * x (B,C,Hi,Wi); y (B,C,Ho,Wo) with Ho = Hi, 2*Hi or Hi/2 for mode 0/1/2.
 * save_mean [B*C], save_hidden [B*hid], save_gate [B*C] are written by fwd and read by bwd.
 * workspace: fwd >= B*C*8 bytes; bwd >= (3*B*C + B*hid)*4 bytes.  dw1/dw2 are overwritten. */
int ffc_se_fwd(const float* x, const float* w1, const float* w2, float* y,
               float* save_mean, float* save_hidden, float* save_gate,
               int B, int C, int hid, int Hi, int Wi, int mode,
               void* workspace, size_t workspace_bytes, void* stream);
int ffc_se_bwd(const float* x, const float* dy, const float* w1, const float* w2,
               const float* save_mean, const float* save_hidden, const float* save_gate,
               float* dx, float* dw1, float* dw2,
               int B, int C, int hid, int Hi, int Wi, int mode,
               void* workspace, size_t workspace_bytes, void* stream);

/* ---- test hook ---------------------------------------------------------------------------------
 * Kernel family used by the convolution entry points.  5 (default): ffc_conv2d_fwd_ws picks, by output width, the
 * tcgen05 / tensor-memory implicit-GEMM kernel (ConvFwdV5) or the packed-weight cp.async-pipelined mma.sync kernel
 * (ConvFwdV4), both 3xTF32 at FP32 accuracy, and plain ffc_conv2d_fwd runs the register-prefetch mma.sync kernel;
 * 4 / 3 force ConvFwdV5 / ConvFwdV4; 0 the register-prefetch mma.sync kernel everywhere; 1 the simple single-buffered
 * FP32 forms of the same math; 2 the tuned FP32 SIMT kernels.  Tests compare all of them. */
void ffc_debug_conv_reference(int mode);
/* Training-mode ffc_fu_fwd runs as ONE cooperative kernel (spectrum held in shared memory across a grid barrier)
 * when all image tiles are co-resident, else as two passes over x; on = 1 forces the two-pass form. */
void ffc_debug_fu_two_pass(int on);

#ifdef __cplusplus
}
#endif
#endif /* FFC_B200_H_ */
