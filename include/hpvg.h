/*
 * hpvg.h — C-ABI of libhpvg.so: the B200 (sm_100a) kernels behind the HP-VAE-GAN per-scale training/generation
 * hot path.  Every entry point takes raw device pointers, extents and a cudaStream_t (as void*); the library never
 * allocates or frees device memory, never synchronises, and launches only on the stream it is given.
 * Return value: 0 on success, negative on error (hpvg_last_error() returns a thread-local message).
 *
 * Tensor formats (the `*_fmt` arguments):
 *   HPVG_FMT_NCDHW_F32  : float32, [N][C][D][H][W] contiguous (what the reference's nn.Module boundary sees:
 *                         3-channel videos/images, 1-channel critic maps, the [N,128,T,H,W] latent).
 *   HPVG_FMT_NDHWC_BF16 : bfloat16, [N][D][H][W][C] contiguous (every wide activation between layers).
 * 2-D networks (modules/networks_2d.py) use D == 1 and KD == 1.
 *
 * Each function names the reference call it replaces (paths relative to the reference repository root).
 */
#ifndef HPVG_H_
#define HPVG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPVG_FMT_NCDHW_F32 0
#define HPVG_FMT_NDHWC_BF16 1

#define HPVG_ACT_NONE 0
#define HPVG_ACT_LRELU 1

/* conv backend selection: 0 = auto (tcgen05 where the shape allows, CUDA-core kernel otherwise),
 * 1 = force the CUDA-core kernels (debug/parity aid), 2 = require tcgen05 (error if the shape does not fit). */
#define HPVG_BACKEND_AUTO 0
#define HPVG_BACKEND_DIRECT 1
#define HPVG_BACKEND_TCGEN05 2

const char* hpvg_last_error(void);
int hpvg_version(void);
int hpvg_set_conv_backend(int backend);
int hpvg_get_conv_backend(void);
/* number of kernels this library has launched since load (all threads) — bench.py's `gpu_launches` */
long long hpvg_launch_count(void);

/* Per-launch device timing of the convolution kernels with CUDA events on the launching stream (bench.py's roofline
 * measurement; adds two event records per launch, so it is off by default and must not be used under stream capture).
 * hpvg_profile_dump waits for the recorded events, aggregates them by (kind, work) and clears the log; it writes up to
 * max_rows rows of 4 doubles {kind, work per launch, launches, total milliseconds} and returns the row count.
 * kind: HPVG_PROF_*; work: algorithmic FLOPs of the launch for the convolution kernels. */
#define HPVG_PROF_CONV_TC 0
#define HPVG_PROF_WGRAD_TC 1
#define HPVG_PROF_CONV_DIRECT 2
#define HPVG_PROF_WGRAD_DIRECT 3
#define HPVG_PROF_CONV_EXPAND 4
#define HPVG_PROF_WGRAD_NARROW 5
#define HPVG_PROF_CONV_BN_FUSED 6
#define HPVG_PROF_CONV_THIN 7
/* development aid: when set (device pointer to >= 8 * grid int64), the tcgen05 kernels write per-CTA phase clocks */
int hpvg_debug_set_clock_buffer(long long* device_buffer);
int hpvg_profile_enable(int on);
/* Programmatic dependent launch of the library's kernels (default OFF — measured slower inside the multi-stream recorded iteration; environment HPVG_PDL=1 or this call turns it on): each kernel is
 * launched with cudaLaunchAttributeProgrammaticStreamSerialization and blocks in griddepcontrol.wait before it touches
 * global memory, so consecutive launches of a stream overlap launch latency and set-up with the predecessor's tail.
 * Returns the previous setting. */
int hpvg_set_pdl(int on);
/* Which tcgen05 kernel runs the 3-D 64-input-channel wide layers: -1 = chosen per layer (default): the brick kernel
 * conv_tc.cu when its 4-slice x 128-voxel units fill at least 80 % of the unit slots of its rounds over the SMs, else the
 * column-streaming kernel conv_col.cu (single-tile work units, no quantisation loss); 0 = brick kernel always; 1 = column
 * kernel whenever it supports the layer.  Initial value from HPVG_TC_COL.  Returns the previous mode. */
int hpvg_set_conv_col_mode(int mode);
/* Weight-gradient kernel of the 3-D wide layers: 1 = kd-stacked form, one N = 192 MMA per (kh,kw) position serves the three
 * kd taps (default; measured on B200: 29.7 vs 37.9 us per 64 -> 64 call at 16 x 64 x 64); 0 = one kd tap plane per CTA, N = 64
 * MMAs; 2 = mode 0 with its fp32 partials written through a swizzled shared-memory tile (coalesced stores).  All three produce
 * the same sums up to fp32 summation order (experiments/check_wgrad_stack.py).  Initial value from HPVG_WGRAD_STACK.
 * Returns the previous mode. */
int hpvg_set_wgrad_mode(int mode);
int hpvg_profile_dump(double* rows, int max_rows);
/* Which kernel hpvg_conv_forward picks for a layer with the AUTO backend, a packed weight image available, plain epilogue
 * (host logic only, no launch; negative = invalid geometry).  COLUMN vs BRICK follows hpvg_set_conv_col_mode's rule. */
#define HPVG_KERNEL_DIRECT 0
#define HPVG_KERNEL_EXPAND 1
#define HPVG_KERNEL_TC_BRICK 2
#define HPVG_KERNEL_TC_COLUMN 3
int hpvg_conv_kernel_choice(int N, int Cin, int Cout, int D, int H, int W, int KD, int pad, int x_fmt, int y_fmt);

/* Do the narrow network ends with `c_thin` channels on their thin side (the 3 -> 64 head convolution, the data gradient of the
 * 64 -> 3 / 64 -> 1 tails, both narrow weight gradients; modules/networks_3d.py:51,63,175,341,362) run on the tcgen05 kernels of
 * narrow_tc.cu?  1 = yes, 0 = the mma.sync / CUDA-core kernels of narrow.cu (c_thin not in {1, 3}, or HPVG_EXPAND_TC=0 /
 * HPVG_NARROW_WGRAD_TC=0).  Host logic only. */
int hpvg_narrow_kernel_choice(int c_thin, int KD);

/* ---------------------------------------------------------------------------------------------------------------
 * Convolution, 3x3x3 (KD == 3) or 3x3 (KD == 1), stride 1, zero padding `pad` in {0,1,2} on every filtered axis.
 * Replaces nn.Conv3d / nn.Conv2d forward inside ConvBlock3D / ConvBlock3DSN / the tail convs
 * (modules/networks_3d.py:51,63,175,290,341,362 ; modules/networks_2d.py:56,68,179,204,225), and — with
 * transposed = 1 — the data-gradient of the same convolution (aten::convolution_backward, grad_input).
 *
 *   transposed == 0 :  y[n,co,o] = act( bias[co] + sum_{ci,k} x[n,ci,o+k-pad] * w[co][ci][k] ),   w is [Cout][Cin][taps]
 *   transposed == 1 :  y[n,co,o] = act( bias[co] + sum_{ci,k} x[n,ci,o+k-pad] * w[ci][co][taps-1-k] ), w is [Cin][Cout][taps]
 *                      (the data gradient of a forward conv with padding p is this call with pad = 2 - p)
 * Output extent per filtered axis = input extent + 2*pad - 2.  D is not filtered when KD == 1.
 * `w_f32` is the float32 master weight in PyTorch layout; `w_packed` (may be NULL) is the bf16 image produced by
 * hpvg_pack_weights for the same `transposed` flag, required for the tcgen05 path (NDHWC_BF16 input with Cin in
 * {64, 128}; NDHWC_BF16 output with Cout a multiple of 64, or NCDHW_F32 output with Cout <= 16 and Cin == 64, bias only).
 * transposed bit 1 (value 2, with or without bit 0): `w_packed` was packed with exactly Cout rows per tap
 * (hpvg_pack_weights(rows = Cout), image [taps][Cout][64]) for a 64 -> <= 4 channel layer with NCDHW_F32 output: the layer then
 * runs as one GEMM per input slab + a shift-add gather (thin_gs.cu) instead of the 16-rows-per-tap tcgen05 thin kernel.
 * `bias` may be NULL.  `stats` (may be NULL) is a float32 [2*Cout] accumulator that
 * receives += per-channel sum and sum of squares of the *stored* output (BatchNorm batch statistics,
 * aten::native_batch_norm's reduction, fused into the conv epilogue); the caller zeroes it.
 * `mask_src` (may be NULL; NDHWC_BF16, same extents as y) multiplies the result by the LeakyReLU derivative of that
 * tensor (1 where > 0, else lrelu_slope): the fused "dgrad then leaky_relu_backward" step of the critic backward.
 * ------------------------------------------------------------------------------------------------------------- */
int hpvg_conv_forward(const void* x, int x_fmt, const float* w_f32, const void* w_packed, const float* bias,
                      void* y, int y_fmt, int N, int Cin, int Cout, int D, int H, int W, int KD, int pad,
                      int transposed, int act, float lrelu_slope, float* stats, const void* mask_src, void* stream);

/* The same call with BatchNorm sums kept PER SAMPLE: stats is [N][2*Cout] when stats_per_sample != 0 (tcgen05 and
 * thin -> wide kernels only).  With hpvg_bn_apply_lrelu_per_sample a batched generation forward computes exactly what N
 * batch-1 forwards compute (the reference draws every sample with batch size 1, train_video.py:226-235). */
int hpvg_conv_forward_ex(const void* x, int x_fmt, const float* w_f32, const void* w_packed, const float* bias,
                         void* y, int y_fmt, int N, int Cin, int Cout, int D, int H, int W, int KD, int pad,
                         int transposed, int act, float lrelu_slope, float* stats, int stats_per_sample,
                         const void* mask_src, void* stream);

/* Weight gradient of the convolution above (aten::convolution_backward grad_weight, and the
 * "wgrad-as-conv" node of the WGAN-GP double backward, modules/utils.py:14-18):
 *   dw[co][ci][k] = sum_{n,o} gy[n,co,o] * x[n,ci,o+k-pad]        (float32, PyTorch layout, overwritten)
 *   dbias[co]     = sum_{n,o} gy[n,co,o]                           (optional)
 * x has extents (D,H,W); gy has extents + 2*pad - 2.  `workspace` must hold hpvg_conv_wgrad_workspace() bytes. */
size_t hpvg_conv_wgrad_workspace(int N, int Cin, int Cout, int D, int H, int W, int KD, int pad, int x_fmt, int gy_fmt);
int hpvg_conv_wgrad(const void* x, int x_fmt, const void* gy, int gy_fmt, float* dw, float* dbias,
                    int N, int Cin, int Cout, int D, int H, int W, int KD, int pad,
                    void* workspace, size_t workspace_bytes, void* stream);

/* float32 [Cout][Cin][taps] (transposed == 0) or [Cin][Cout][taps] (transposed == 1) -> bf16 [taps][rows_per_tap][Cin]
 * K-major tiles for the tcgen05 kernels (rows >= Cout are zero).  rows_per_tap = Cout for wide outputs, 16 for the
 * thin-output kernel (Cout <= 16).  `inv_scale_of` (may be NULL) points to a device float whose reciprocal multiplies
 * every weight (spectral normalisation: W / sigma). */
int hpvg_pack_weights(const float* w_f32, void* w_packed, int Cout, int Cin, int taps, int transposed,
                      const float* inv_scale_of, int rows_per_tap, void* stream);

/* Both images (transposed == 0 into packed_fwd[l], transposed == 1 into packed_tr[l]; either may be NULL) of n <=
 * HPVG_SN_MAX_LAYERS layers with rows_per_tap = Cout, no scaling, in one launch: the critic's spectral-norm weights change in
 * every pass, so every pass repacks all of them.  Arrays of n device pointers / shapes live in host memory. */
int hpvg_pack_weights_pair_batched(int n, const float* const* w_f32, void* const* packed_fwd, void* const* packed_tr,
                                   const int* Cout, const int* Cin, const int* taps, void* stream);

/* float32 [taps][Cin][64] image of the filter of a thin -> wide layer (Cin <= 4, 64 output channels; `transposed` as in
 * hpvg_conv_forward), passed as `w_packed` to hpvg_conv_forward for NCDHW_F32 -> NDHWC_BF16 calls: optional, but lets every
 * block of the kernel copy the filter instead of gathering it from the PyTorch layout. */
int hpvg_pack_weights_expand(const float* w_f32, float* w_tco, int Cin, int taps, int transposed, void* stream);

/* per-channel sum of a tensor: out[c] = sum_{n,o} t[n,c,o]   (bias gradient) */
int hpvg_channel_sum(const void* t, int fmt, float* out, int N, int C, long long spatial, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * BatchNorm (training mode) + LeakyReLU, replaces nn.BatchNorm3d/2d + nn.LeakyReLU(0.2) of ConvBlock3D/2D
 * (modules/networks_3d.py:54-56, modules/networks_2d.py:59-61).
 * bn_finalize: from the conv-epilogue sums -> scale_shift[0..C) = gamma*invstd, [C..2C) = beta - mean*scale ;
 *   mean_invstd[0..C) = mean, [C..2C) = invstd ; running_mean/var updated with `momentum` (unbiased variance),
 *   num_batches_tracked += 1 (int64, may be NULL).
 * bn_apply_lrelu: out = lrelu(y*scale + shift).   bn_finalize_apply_lrelu: both of the above in one launch.
 * bn_lrelu_bwd_reduce: `sums` is float32 [3*C], zeroed here; sums[0..C) = sum dz, sums[C..2C) = sum dz*xhat with
 *   dz = gout * lrelu'(y*scale+shift).
 * bn_lrelu_bwd_apply: gy = scale * (dz - sums0/M - xhat*sums1/M) ; dgamma = sums1, dbeta = sums0 (written once); with
 *   want_chsum, sums[2C..3C) += per-channel sum of the stored gy = the bias gradient of the preceding convolution.
 *   mask_bits (both backward kernels; may be NULL): uint8 [nvox][C/8] written by hpvg_conv_bn_lrelu_fused — bit b of byte
 *   [v][c/8] says that the fp32 pre-activation of channel 8*(c/8)+b at voxel v was positive; when given, lrelu' is read from
 *   it instead of from the sign of the recomputed y*scale+shift (y is stored in bf16, the forward saw fp32).
 * hpvg_conv_bn_lrelu_fused: ConvBlock3D (modules/networks_3d.py:48-56) in ONE launch for 64 -> 64 3x3x3 layers whose work
 *   units fit the SMs (hpvg_conv_bn_lrelu_fused_supported): the tcgen05 convolution keeps its accumulators in TMEM across a
 *   grid-wide barrier on sum(y), sum(y^2) — taken from the fp32 accumulators — then normalises, applies the affine map and
 *   LeakyReLU in fp32 and stores `out` (NDHWC_BF16), optionally `y` (the bf16 conv output the backward needs; NULL: not
 *   stored) and `mask_bits` (NULL: not stored).  `stats` is float32 [2*Cout + 32], ZEROED by the caller: [0, 2*Cout) the sums,
 *   element 2*Cout the grid barrier's arrival counter.  scale_shift / mean_invstd / running statistics as bn_finalize.
 *   Launched cooperatively: never run two of these concurrently on one device from different streams unless both grids fit
 *   the SMs together.
 * ------------------------------------------------------------------------------------------------------------- */
int hpvg_bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, long long* num_batches_tracked, float momentum, float eps,
                     long long count, float* scale_shift, float* mean_invstd, int C, void* stream);
int hpvg_bn_apply_lrelu(const void* y, const float* scale_shift, void* out, long long nvox, int C, float slope,
                        void* stream);
int hpvg_bn_lrelu_bwd_reduce(const void* y, const void* gout, const float* scale_shift, const float* mean_invstd,
                             float* sums, long long nvox, int C, float slope, const void* mask_bits, void* stream);
int hpvg_bn_lrelu_bwd_apply(const void* y, const void* gout, const float* scale_shift, const float* mean_invstd,
                            float* sums, void* gy, float* dgamma, float* dbeta, long long nvox, int C,
                            float slope, int want_chsum, const void* mask_bits, void* stream);
/* Deferred running-statistics updates: entry i applies running <- (1 - momentum) running + momentum {mean, unbiased var} and
 * num_batches_tracked += 1 to its layer from the mean_invstd vector the layer's forward saved, entries in call order (several
 * entries may name the same layer).  Lets passes that share BatchNorm layers run concurrently without racing on the buffers. */
#define HPVG_BN_LOG_MAX 48
int hpvg_bn_running_update_batched(int n, float* const* running_mean, float* const* running_var,
                                   long long* const* num_batches_tracked, const float* const* mean_invstd, const long long* count,
                                   const int* C, const float* momentum, const float* eps, void* stream);
/* bn_lrelu_bwd_reduce + bn_lrelu_bwd_apply in ONE launch: y and gout are read once into shared memory, the partial sums cross a
 * grid-wide barrier, gy is written from shared memory.  `sums`: float32 [3C + 32] ZEROED by the caller ([3C] is the barrier's
 * arrival counter).  Eligible (.._supported) when nvox * C fits the SMs' shared memory; the two-launch pair covers the rest. */
int hpvg_bn_lrelu_bwd_fused_supported(long long nvox, int C);
int hpvg_bn_lrelu_bwd_fused(const void* y, const void* gout, const float* scale_shift, const float* mean_invstd, float* sums,
                            void* gy, float* dgamma, float* dbeta, long long nvox, int C, float slope, int want_chsum,
                            const void* mask_bits, void* stream);
int hpvg_conv_bn_lrelu_fused_supported(int N, int Cin, int Cout, int D, int H, int W, int KD, int pad);
int hpvg_conv_bn_lrelu_fused(const void* x, const void* w_packed, const float* bias, void* y, void* out, int N, int Cin, int Cout,
                             int D, int H, int W, int KD, int pad, float slope, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                             float* stats, float* scale_shift, float* mean_invstd, void* mask_bits, void* stream);
/* inference-only BatchNorm(batch statistics of each sample) + LeakyReLU: y, out NDHWC_BF16 [N][nvox_per_sample][C],
 * stats [N][2C] from hpvg_conv_forward_ex(stats_per_sample = 1); no running statistics, nothing saved for a backward */
int hpvg_bn_apply_lrelu_per_sample(const void* y, const float* stats, const float* gamma, const float* beta, float eps,
                                   void* out, int N, long long nvox_per_sample, int C, float slope, void* stream);
int hpvg_bn_finalize_apply_lrelu(const void* y, const float* stats, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                                 float eps, float* scale_shift, float* mean_invstd, void* out, long long nvox, int C,
                                 float slope, void* stream);

/* gz = gout * (out > 0 ? 1 : slope)   — aten::leaky_relu_backward on the saved in-place output (bf16 NDHWC tensors).
 * chsum (may be NULL; float32 [C], overwritten) receives the per-channel sum of gz: the bias gradient of the
 * spectral-norm convolution in front of this activation, fused into the same pass. */
int hpvg_lrelu_bwd(const void* gout, const void* out_saved, void* gz, long long numel, float slope, int C, float* chsum,
                   void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Trilinear / bilinear resize with align_corners=True on NCDHW float32 tensors, replaces utils.upscale /
 * utils.interpolate_3D / utils.upscale_2d (utils/images.py:9-26,83-105):
 *   out = resize(x) [+ noise_amp * noise]          noise may be NULL (networks_3d.py:395-402)
 * and its adjoint (aten::upsample_trilinear3d_backward) as a gather over the input grid.
 * ------------------------------------------------------------------------------------------------------------- */
int hpvg_upsample_linear_fwd(const float* x, float* out, const float* noise, float noise_amp, int NC,
                             int Di, int Hi, int Wi, int Do, int Ho, int Wo, void* stream);
int hpvg_upsample_linear_bwd(const float* gout, float* gx, int NC, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                             void* stream);

/* The same resize on NDHWC_BF16 tensors (C a multiple of 8) for GeneratorCSG, which resizes its nfc-channel feature maps
 * between stages (networks_3d.py:252-261): out = resize(x) [+ noise_amp * noise], `noise` float32 NCDHW [N][C][Do][Ho][Wo]
 * as the reference draws it (may be NULL); and the adjoint in gather form. */
int hpvg_upsample_linear_wide_fwd(const void* x, void* out, const float* noise, float noise_amp, int N, int C, int Di,
                                  int Hi, int Wi, int Do, int Ho, int Wo, void* stream);
int hpvg_upsample_linear_wide_bwd(const void* gout, void* gx, int N, int C, int Di, int Hi, int Wi, int Do, int Ho,
                                  int Wo, void* stream);
/* F.pad(x, (pad,)*6) with zeros on an NDHWC_BF16 tensor (networks_3d.py:205,248,264); pad < 0 crops (its adjoint):
 * x [N][D][H][W][C] -> out [N][D+2pad][H+2pad][W+2pad][C] */
int hpvg_pad_wide(const void* x, void* out, int N, int C, int D, int H, int W, int pad, void* stream);
/* out = a + b on NDHWC_BF16 tensors: the stage residual x_prev + x_prev_out_up of GeneratorCSG (networks_3d.py:265) */
int hpvg_add_wide(const void* a, const void* b, void* out, long long numel, void* stream);

/* out = tanh(a + b)  (b may be NULL) — torch.tanh(block(x) + x_up), torch.tanh(decoder(z))
 * (networks_3d.py:377,404).  bwd: g = gout * (1 - out^2). */
int hpvg_tanh_add_fwd(const float* a, const float* b, float* out, long long numel, void* stream);
int hpvg_tanh_bwd(const float* gout, const float* out, float* g, long long numel, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * VAE head: reparameterize (networks_3d.py:29-35) on NDHWC_BF16 mu/logvar with NCDHW float32 eps:
 *   z = eps * exp(0.5*logvar) + mu  (NDHWC_BF16) ;  bwd: gmu = gz, glogvar = gz * eps * 0.5 * exp(0.5*logvar)
 * kl_criterion (modules/losses.py:7-9) on NCDHW float32: out[0] = mean(-0.5*(1 + logvar - mu^2 - exp(logvar)))
 * ------------------------------------------------------------------------------------------------------------- */
int hpvg_reparam_fwd(const void* mu, const void* logvar, const float* eps, void* z, int N, int C, long long spatial,
                     void* stream);
int hpvg_reparam_bwd(const void* gz, const void* logvar, const float* eps, void* gmu, void* glogvar, int N, int C,
                     long long spatial, void* stream);
int hpvg_kl_fwd(const float* mu, const float* logvar, float* out, long long numel, void* stream);
int hpvg_kl_bwd(const float* gout, const float* mu, const float* logvar, float* gmu, float* glogvar, long long numel,
                void* stream);

/* WGAN-GP penalty (modules/utils.py:18): out[0] = lambda * mean_{n,voxel} (||g[n,:,voxel]||_2 - 1)^2 over the
 * channel axis of an NCDHW float32 gradient; bwd: gg = gout * lambda * 2(||g||-1)/(N*S) * g/||g||. */
int hpvg_gp_penalty_fwd(const float* g, float* out, int N, int C, long long spatial, float lambda, void* stream);
int hpvg_gp_penalty_bwd(const float* gout, const float* g, float* gg, int N, int C, long long spatial, float lambda,
                        void* stream);

/* layout/dtype conversion between the two formats (differentiable by its own inverse) */
int hpvg_convert_format(const void* src, int src_fmt, void* dst, int dst_fmt, int N, int C, long long spatial,
                        void* stream);

/* out = alpha*a + (1-alpha)*b on float32 (the GP interpolates, modules/utils.py:9); alpha is read from device memory
 * so that a captured CUDA graph sees a fresh value at every replay */
int hpvg_lerp(const float* a, const float* b, float* out, const float* alpha, long long numel, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Data formats either side of the path (bit-exact restatements of host code of the reference).
 * clip_from_frames: what SingleVideoDataset.__getitem__ builds per iteration (datasets/video.py:44-92) from the frames of
 *   the current pyramid level, kept resident as uint8 [num_frames][H][W][3] (RGB):
 *     clip[c][t][h][w] = ((frames[first + t*every][h][w'][c] / 255) - 0.5) / 0.5 ,  w' = hflip ? W-1-w : w     (float32 [3][T][H][W])
 * frames_to_uint8: utils/saver.py:16-18  out[t][h][w][c] = uint8((video[c][t][h][w] + 1) * 127.5)   (uint8 [T][H][W][3])
 * ------------------------------------------------------------------------------------------------------------- */
int hpvg_clip_from_frames(const uint8_t* frames, float* clip, int num_frames, int first, int every, int T, int H, int W,
                          int hflip, void* stream);
int hpvg_frames_to_uint8(const float* video, uint8_t* out, int T, int H, int W, void* stream);
/* the same for a batch of videos: float32 [N][3][T][H][W] -> uint8 [N][T][H][W][3] in one launch */
int hpvg_frames_to_uint8_batched(const float* video, uint8_t* out, int N, int T, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Spectral normalisation, one power iteration (nn.utils.spectral_norm as used by ConvBlock3DSN/2DSN,
 * networks_3d.py:63): w_mat = w_orig viewed [Cout][K].  In place: v <- normalize(W^T u), u <- normalize(W v);
 * sigma[0] = u^T W v.  `scratch` holds K + Cout + 4 floats.  Then w_sn = w_orig / sigma.
 * sn_backward: gw_orig = (gw_sn - (sum(gw_sn * w_sn)) * u v^T) / sigma    (u, v constants, as in torch); its `scratch`
 * holds HPVG_SN_DOT_PARTS floats.  All reductions run in a fixed order (no atomics): sigma is bit-reproducible.
 * ------------------------------------------------------------------------------------------------------------- */
int hpvg_sn_power_iter(const float* w_orig, float* u, float* v, float* sigma, float* w_sn, float* scratch, int Cout,
                       int K, int update_uv, float eps, void* stream);
int hpvg_sn_backward(const float* gw_sn, const float* w_sn, const float* u, const float* v, const float* sigma,
                     float* gw_orig, float* scratch, int Cout, int K, void* stream);
/* the same for all spectral-norm layers of one network at once (n <= HPVG_SN_MAX_LAYERS): arrays of n device pointers
 * (the arrays themselves are host memory) and n shapes; scratch[l] holds K[l] + Cout[l] + 4 floats (forward) / HPVG_SN_DOT_PARTS (backward) */
#define HPVG_SN_MAX_LAYERS 8
#define HPVG_SN_DOT_PARTS 32
int hpvg_sn_power_iter_batched(int n, const float* const* w_orig, float* const* u, float* const* v, float* const* sigma,
                               float* const* w_sn, float* const* scratch, const int* Cout, const int* K, int update_uv,
                               float eps, void* stream);
/* .._ex: with update_uv, also writes the updated u / v of layer l to u_saved[l] / v_saved[l] (device pointers, arrays or entries
 * may be NULL): the copies the backward pass needs, without 2 copy launches per layer */
int hpvg_sn_power_iter_batched_ex(int n, const float* const* w_orig, float* const* u, float* const* v, float* const* sigma,
                                  float* const* w_sn, float* const* scratch, const int* cout, const int* k, int update_uv,
                                  float eps, float* const* u_saved, float* const* v_saved, void* stream);
int hpvg_sn_backward_batched(int n, const float* const* gw_sn, const float* const* w_sn, const float* const* u,
                             const float* const* v, const float* const* sigma, float* const* gw_orig, float* const* scratch,
                             const int* Cout, const int* K, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * End of an iteration: gradient clipping + Adam over all tensors of an optimizer.  Replaces
 * torch.nn.utils.clip_grad_norm_(G_curr.parameters(), opt.grad_clip) (train_video.py:201, train_image.py:216) and
 * optim.Adam(...).step() (train_video.py:55,88,183,202; train_video_baselines.py:51,70: betas (beta1, 0.999), eps 1e-8,
 * no weight decay, no amsgrad).  All tensors are contiguous float32; the pointer / size / lr arrays are HOST arrays of n
 * entries (n <= HPVG_OPT_MAX_TENSORS per call; more tensors = more calls).
 * `state`: HPVG_OPT_STATE_FLOATS device floats owned by the optimizer, zero-initialised once:
 *   [0] step count t   [1] clip coefficient   [2] total gradient norm   [4],[5] block tickets (self-resetting)
 * grad_clip_coef: partials[slot_base + i*HPVG_OPT_BLOCKS + b] = per-block sums of squares of grads[i]; the call with
 *   finalize != 0 (the last one of a sequence covering total_slots slots) reduces all partials in index order and sets
 *   state[1] = min(1, max_norm / (sqrt(sum) + 1e-6)), state[2] = sqrt(sum).  No atomically accumulated floats.
 * adam_step: g <- g*state[1] (when use_clip; written back, as clip_grad_norm_ scales .grad in place), m <- lerp(m, g, 1-beta1),
 *   v <- beta2 v + (1-beta2) g^2, p <- p - lr[i]/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps) with t = state[0] + 1.
 *   exp_avg[i] == NULL: tensor i is not owned by the optimizer, only its gradient is scaled.  advance_step != 0 (the last
 *   call of a step): state[0] <- t when the grid has finished.  The step count lives on the device so that a recorded CUDA
 *   graph advances it at every replay.
 * ------------------------------------------------------------------------------------------------------------- */
#define HPVG_OPT_MAX_TENSORS 32
#define HPVG_OPT_BLOCKS 64
#define HPVG_OPT_STATE_FLOATS 8
int hpvg_grad_clip_coef(int n, const float* const* grads, const long long* numel, float* partials, int slot_base,
                        int total_slots, int finalize, float max_norm, float* state, void* stream);
int hpvg_adam_step(int n, float* const* params, float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const long long* numel, const float* lr, double beta1, double beta2, double eps, int use_clip,
                   int advance_step, float* state, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Gradient averaging over NVLink peer memory (multi-GPU mode, one process per GPU): replaces the reduction of the replicas'
 * gradients that nn.DataParallel's backward performs (train_video.py:91-94, :182, :200) — here one clip per rank and ONE kernel per
 * backward instead of a library collective.
 *
 * Every rank owns one bucket (numel floats, numel a multiple of 4 * world) and one signal pad (HPVG_PEER_SIGNAL_BYTES, zero
 * filled), both allocated with hpvg_peer_alloc (an allocation of its own: an IPC handle covers a whole allocation), exported
 * with hpvg_peer_export (64-byte handle, sent to the other ranks by any host channel) and mapped there with hpvg_peer_import.
 * hpvg_peer_allreduce_avg(bufs, signals, rank, world, numel, stream): bufs[q] / signals[q] are rank q's bucket / pad as mapped in
 * THIS process (q == rank: the local allocations).  The kernel (two-shot):
 *   1. flag barrier with the same-numbered CTA of every peer (st.release.sys into the peer's pad, ld.acquire.sys on the own pad):
 *      every rank's bucket has been filled by the kernels that precede the call on its stream;
 *   2. rank r reduces slice r: pulls it from every bucket in rank order (bit-identical result whatever the arrival order),
 *      multiplies by 1 / world and pushes the mean into slice r of EVERY bucket;
 *   3. second flag barrier: all pushes have landed, nobody reads this rank's bucket any more — the next kernel of the stream
 *      finds the averaged gradients in the local bucket, and the bucket may be refilled.
 * Flags count calls (the count lives in the pad: a recorded CUDA graph advances it at every replay), so nothing is reset between
 * calls.  Every wait is bounded (60 s) and traps.  All ranks must issue the same sequence of calls.
 * ------------------------------------------------------------------------------------------------------------- */
#define HPVG_PEER_MAX_RANKS 8
#define HPVG_PEER_HANDLE_BYTES 64
#define HPVG_PEER_SIGNAL_BYTES 8192
int hpvg_peer_alloc(size_t bytes, void** ptr);
int hpvg_peer_free(void* ptr);
int hpvg_peer_export(const void* ptr, void* handle);
int hpvg_peer_import(const void* handle, void** ptr);
int hpvg_peer_close(void* ptr);
int hpvg_peer_can_access(int device, int peer_device);
int hpvg_peer_allreduce_avg(void* const* bufs, void* const* signals, int rank, int world, long long numel, void* stream);
/* The same exchange with the bucket filled from, and emptied into, n gradient tensors (16-byte aligned, numel[i] floats) by the kernel
 * itself: no pack / unpack launches around it.  Tensor i occupies the float4 slots behind tensor i - 1 (its last slot zero-padded);
 * hpvg_peer_bucket_numel(n, numel, world) is the bucket size that layout needs (host only).  grads[i] <- mean over ranks, in place. */
#define HPVG_PEER_MAX_TENSORS 64
long long hpvg_peer_bucket_numel(int n, const long long* numel, int world);
int hpvg_peer_allreduce_avg_tensors(void* const* bufs, void* const* signals, int rank, int world, long long bucket_numel, int n,
                                    float* const* grads, const long long* numel, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HPVG_H_ */
