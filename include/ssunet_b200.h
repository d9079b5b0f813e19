/* ssunet_b200.h — C ABI of libssunet_b200.so (hand-written sm_100a CUDA for the ssUnet-GAN
 * data-parallel training-step hot path).
 *
 * The reference (ideafisher/ssUnet-GAN) has no FFI of its own: every device operation on the
 * path is an ATen/cuDNN call issued from Python.  Each entry point below
 * replaces one such call site; the comment above it cites the reference file:line
 * (relative to scripts/).  The Python host modules in ssunet-gan_b200/ bind these with ctypes.
 *
 * Conventions
 *  - plain pointers + sizes; no torch / C++ types.  All pointers are DEVICE pointers unless the
 *    name ends in _host.  The caller owns every buffer (inputs, outputs, workspaces).
 *  - activations are NHWC ("channels last") contiguous: element (n,h,w,c) at ((n*H+h)*W+w)*C+c.
 *    `dtype` selects their storage type: SSG_F32 or SSG_BF16.  Arithmetic is always fp32.
 *  - every launch is asynchronous on `stream` (a cudaStream_t passed as void*); no hidden syncs.
 *  - return value: 0 on success, negative ssg_status on error; ssg_last_error() describes it.
 *    Nothing throws across the ABI.
 */
#ifndef SSUNET_B200_H
#define SSUNET_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ssg_stream_t; /* cudaStream_t */

enum ssg_status { SSG_OK = 0, SSG_ERR_INVALID_ARG = -1, SSG_ERR_CUDA = -2, SSG_ERR_UNSUPPORTED = -3 };
enum ssg_dtype { SSG_F32 = 0, SSG_BF16 = 1 };
enum ssg_act { SSG_ACT_NONE = 0, SSG_ACT_RELU = 1, SSG_ACT_LEAKY = 2 };
/* packed-weight layouts produced by ssg_pack_conv_weight */
enum ssg_wlayout {
    SSG_W_RSCK = 0, /* [kh][kw][cin][cout]  : forward operand of the SIMT kernels,
                                              dgrad (K-major in cout) operand of the tcgen05 kernels */
    SSG_W_RSKC = 1, /* [kh][kw][cout][cin]  : dgrad operand of the SIMT kernels,
                                              forward (K-major) operand of the tcgen05 kernels */
    SSG_W_RSCK_FLIP = 2 /* [kh-1-r][kw-1-s][cin][cout]: flipped-tap variant (kept for callers that
                                              express dgrad as a plain forward convolution)     */
};

int ssg_version(void);
const char* ssg_last_error(void);
/* SM count / compute capability of the current device (host query). */
int ssg_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- layout ------------------------------------------------------------------------------ */
/* module boundary: reference tensors are NCHW fp32 (dataset.py:126-142, archs.py:623). */
int ssg_nchw_to_nhwc(const float* src, void* dst, int dtype, int n, int c, int h, int w, ssg_stream_t s);
int ssg_nhwc_to_nchw(const void* src, int dtype, float* dst, int n, int c, int h, int w, ssg_stream_t s);
/* Same with channel-padded NHWC storage (c_dst / c_src >= c; padding channels are zero / ignored).  The tensor-core
 * kernels need the NHWC pixel pitch to be a multiple of 16 bytes (TMA), so 3-channel images and SPADE's 3 / h-channel
 * maps (normalization.py:93-98) are stored with their channel count rounded up to a multiple of 8. */
int ssg_nchw_to_nhwc_pad(const float* src, void* dst, int dtype, int n, int c, int c_dst, int h, int w, ssg_stream_t s);
int ssg_nhwc_to_nchw_pad(const void* src, int dtype, float* dst, int n, int c, int c_src, int h, int w, ssg_stream_t s);
/* dtype cast of a flat buffer (fp32 <-> bf16), n elements */
int ssg_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, ssg_stream_t s);
/* torch.cat([a, b], 1) (archs.py:651,656,661,664,667) and its adjoint; rows = N*H*W */
int ssg_concat2(const void* a, int ca, const void* b, int cb, void* out, int dtype, long long rows, ssg_stream_t s);
int ssg_split2(const void* in, void* a, int ca, void* b, int cb, int dtype, long long rows, ssg_stream_t s);
/* OIHW fp32 parameter -> packed operand (see ssg_wlayout); `scale_dev` (optional, device float)
 * multiplies every element (spectral norm's 1/sigma, spectral_norm.py:86-87). */
int ssg_pack_conv_weight(const float* w_oihw, void* dst, int dtype, int layout, int cout, int cin, int kh, int kw,
                         const float* inv_scale_dev, ssg_stream_t s);
/* Same, zero-padded to [.. cout_p ..][.. cin_p ..] channel extents (operands of channel-padded activations). */
int ssg_pack_conv_weight_pad(const float* w_oihw, void* dst, int dtype, int layout, int cout, int cin, int kh, int kw, int cout_p,
                             int cin_p, const float* inv_scale_dev, ssg_stream_t s);
/* Every packed operand of a network refreshed by ONE launch (after the optimiser step: srgan_utils.py:192-195 + Adam changed all
 * weights).  descs_dev: DEVICE array of n_descs records of SSG_PACK_DESC_BYTES bytes, little-endian, no padding:
 *   u64 src (fp32 OIHW), u64 dst, i32 layout, i32 cout, i32 cin, i32 ksize, i32 cout_p, i32 cin_p, i64 first_block
 * One block converts one SSG_PACK_TILE x SSG_PACK_TILE (output x input channel) tile with all ksize^2 taps: `first_block` is the
 * running sum of ceil(cout_p / SSG_PACK_TILE) * ceil(cin_p / SSG_PACK_TILE) over the preceding records and total_blocks the sum
 * over all of them.  ksize must be 1 or 3 (other kernels: ssg_pack_conv_weight_pad). */
#define SSG_PACK_DESC_BYTES 48
#define SSG_PACK_TILE 32
int ssg_pack_conv_weights_multi(const void* descs_dev, int n_descs, long long total_blocks, int dtype, ssg_stream_t s);

/* ---- convolution, CUDA-core implicit GEMM (any shape; the only path for tiny channel counts) */
/* nn.Conv2d forward (archs.py:210,212,218; normalization.py:93-98; models_seg_gan.py:38-39).
 * w: SSG_W_RSCK in `dtype`; bias fp32 or NULL; y = act(conv(x) + bias). */
int ssg_conv2d_fwd_simt(const void* x, const void* w, const float* bias, void* y, int dtype, int n, int h, int w_,
                        int cin, int cout, int kh, int kw, int stride, int pad, int act, float slope, ssg_stream_t s);
/* dL/dx of the same convolution; w: SSG_W_RSKC; (h,w_) are the INPUT spatial dims, dy is [n,oh,ow,cout]. */
int ssg_conv2d_dgrad_simt(const void* dy, const void* w, void* dx, int dtype, int n, int h, int w_, int cin, int cout,
                          int kh, int kw, int stride, int pad, ssg_stream_t s);
/* dL/dW written as OIHW fp32 (overwrites dw); dy is [n,oh,ow,cout]. */
int ssg_conv2d_wgrad_simt(const void* x, const void* dy, float* dw_oihw, int dtype, int n, int h, int w_, int cin,
                          int cout, int kh, int kw, int stride, int pad, ssg_stream_t s);

/* ---- convolution, tcgen05 / TMEM / TMA implicit GEMM (bf16 NHWC, fp32 accumulate) ------------------- */
/* Channel counts passed to these three calls are the STORED (possibly zero-padded) channel counts of the NHWC
 * tensors; they must be multiples of 8 (16-byte pixel pitch for TMA).  K-blocks are 64 channels wide; a ragged last
 * block (e.g. the 8-channel stems and SPADE's 8..48-channel maps) is zero-filled by the TMA unit.
 *
 * Convolution (1x1 or 3x3, stride 1 or 2, zero padding `pad`) of the channel concatenation [x0 | x1]
 * (x1 may be NULL with c1 == 0; torch.cat is never materialised, archs.py:651-667; c0 % 64 == 0 when x1 is given).
 * (h, w) are the INPUT spatial dims.  w_packed: bf16 [taps][cout][c0+c1] (SSG_W_RSKC, zero rows/columns for
 * padding channels); bias: bias_n fp32 entries or NULL; y = act(conv + bias), bf16 [n,oh,ow,cout].  The stride-2
 * gather is done by the TMA unit (element-strided tensor-map traversal).  Replaces the cuDNN implicit-GEMM calls
 * behind archs.py:210,212,218,593-601,615, normalization.py:93-98 and models_seg_gan.py:38-39.
 * stats (optional, NULL otherwise): fp64 [2][cout], overwritten with the per-channel sum and sum of squares of the
 * stored output -- the BatchNorm statistics of batchnorm.py:59-64 as a by-product of the epilogue (only for the
 * shapes ssg_conv2d_fwd_tc_has_stats() accepts: stride 1, same-size). */
int ssg_conv2d_fwd_tc(const void* x0, int c0, const void* x1, int c1, const void* w_packed, const float* bias, int bias_n, void* y,
                      int n, int h, int w, int cout, int ksize, int stride, int pad, int act, float slope, double* stats,
                      ssg_stream_t s);
/* 1 if ssg_conv2d_fwd_tc can produce `stats` for this geometry (host query, returns 0/1, never fails) */
int ssg_conv2d_fwd_tc_has_stats(int ksize, int stride, int pad);

/* Data gradient of the same convolution: dx bf16 [n,h,w,cin] from dy bf16 [n,oh,ow,cout].
 * w_packed: bf16 [taps][cin][cout] (SSG_W_RSCK).  Stride 2 is computed as four output-parity classes (each a
 * small stride-1 convolution over dy with 1, 2, 2 and 4 taps) written interleaved into dx. */
int ssg_conv2d_dgrad_tc(const void* dy, const void* w_packed, void* dx, int n, int h, int w, int cin, int cout, int ksize,
                        int stride, int pad, ssg_stream_t s);
/* Same for the stride-2 3x3 convolution, with the PRODUCER's activation backward folded into the epilogue: `producer_out` is the
 * post-activation output (shaped like dx) of the layer whose output this convolution consumed -- i.e. this convolution's saved
 * input x -- and dx is multiplied by act'(producer_out) before it is stored.  The discriminator's block 0 (conv + LeakyReLU, no
 * BN: models_seg_gan.py:262-266) then needs no separate LeakyReLU-backward pass over its 16 x 512 x 512 x 64 gradient.  Valid
 * only when x has no other consumer.  ssg_conv2d_dgrad_tc_mask_supported: geometries that offer it. */
int ssg_conv2d_dgrad_tc_mask_supported(int ksize, int stride, int pad);
int ssg_conv2d_dgrad_tc_mask(const void* dy, const void* w_packed, void* dx, const void* producer_out, int producer_act, float producer_slope,
                             int n, int h, int w, int cin, int cout, int ksize, int stride, int pad, ssg_stream_t s);
/* Same, but dx += gradient (TMA reduce-add in the epilogue, bf16 addition at the L2): the second consumer of an activation adds
 * its contribution into the buffer the first one wrote -- replaces autograd's separate addition pass for BasicBlock's
 * conv1 + shortcut (archs.py:229-234) and SPADE's x2map + modulation (normalization.py:112-120).  Only for the geometries
 * ssg_conv2d_dgrad_tc_can_acc() accepts (same-size stride-1 1x1 / 3x3); otherwise SSG_ERR_UNSUPPORTED. */
int ssg_conv2d_dgrad_tc_acc(const void* dy, const void* w_packed, void* dx, int n, int h, int w, int cin, int cout, int ksize,
                            int stride, int pad, ssg_stream_t s);
int ssg_conv2d_dgrad_tc_can_acc(int ksize, int stride, int pad);
/* Data gradient of a convolution over the virtual concatenation [x0 | x1] (ssg_conv2d_fwd_tc with x1 != NULL): gradient channels
 * [0, c0) go to dx0 [n,h,w,c0], channels [c0, c0+c1) to dx1 [n,h,w,c1] (c0 % 64 == 0, c1 % 8 == 0), written or -- accumulate != 0 --
 * added.  Replaces torch.cat's backward (archs.py:651-667).  Same geometries as ssg_conv2d_dgrad_tc_acc. */
int ssg_conv2d_dgrad_tc_split(const void* dy, const void* w_packed, void* dx0, int c0, void* dx1, int c1, int n, int h, int w, int cout,
                              int ksize, int stride, int pad, int accumulate, ssg_stream_t s);

/* Weight gradient of the same convolution on tensor cores (both operands MN-major straight from the NHWC
 * tensors, split-K over pixel tiles, fp32 atomics into dw): dw_oihw fp32 [cout_real][cin_real][k][k] is
 * overwritten (padding channels of x / dy are ignored).  (h, w) are the INPUT spatial dims. */
int ssg_conv2d_wgrad_tc(const void* x0, int c0, const void* x1, int c1, const void* dy, int cout, float* dw_oihw, int cout_real,
                        int cin_real, int n, int h, int w, int ksize, int stride, int pad, ssg_stream_t s);

/* Same, but dw_oihw is ACCUMULATED into (dw += ...) instead of overwritten: the caller passes the parameter's slot of the
 * flat gradient arena (zeroed by zero_grad), which removes the per-layer memset, temporary and gradient-accumulation add. */
int ssg_conv2d_wgrad_tc_acc(const void* x0, int c0, const void* x1, int c1, const void* dy, int cout, float* dw_oihw, int cout_real,
                            int cin_real, int n, int h, int w, int ksize, int stride, int pad, ssg_stream_t s);

/* ---- 3x3 / stride 1 / pad 1 convolutions between thin activations (both stored with 8 channels, <= 8 real channels) ---------
 * SPADE's mlp_shared (label_nc -> nhidden, normalization.py:93-95) at the two finest levels and its adjoints.  CUDA-core kernels,
 * one pixel per thread, weights read from the fp32 OIHW parameter directly (no packed operand): a 128-row tensor-core tile for
 * ~100 multiply-adds per pixel is all issue / drain overhead.  bf16 NHWC storage only. */
int ssg_conv3x3_tiny_supported(int cin_stored, int cout_stored, int cin, int cout);
int ssg_conv3x3_tiny_fwd(const void* x, const float* w_oihw, const float* bias, void* y, int n, int h, int w, int cin, int cout, int act,
                         float slope, ssg_stream_t s);
int ssg_conv3x3_tiny_dgrad(const void* dy, const float* w_oihw, void* dx, int n, int h, int w, int cin, int cout, ssg_stream_t s);
/* dw (fp32 OIHW) and db (fp32, may be NULL) are ACCUMULATED into: pass zeroed buffers or gradient-arena slots. */
int ssg_conv3x3_tiny_wgrad(const void* x, const void* dy, float* dw_oihw, float* db, int n, int h, int w, int cin, int cout, ssg_stream_t s);

/* ---- per-channel statistics / batch norm -------------------------------------------------- */
/* batchnorm.py:59-64 (_sum_ft of x and x**2): sums[0:C] = sum x, sums[C:2C] = sum x^2 (fp64,
 * overwritten).  With with_sq == 0 only sums[0:C] is produced (bias gradients). */
int ssg_channel_stats(const void* x, int dtype, long long rows, int c, double* sums, int with_sq, ssg_stream_t s);
/* batchnorm.py:115-127 (_compute_mean_std) when sync_quirk != 0: inv_std = clamp(var, eps)^-1/2;
 * otherwise F.batch_norm semantics (var + eps), batchnorm.py:52-55.  Updates running stats in
 * place (momentum; unbiased variance) when running_mean != NULL. */
int ssg_bn_finalize(const double* sums, double count, int c, float eps, float momentum, int sync_quirk,
                    float* running_mean, float* running_var, float* mean, float* inv_std, ssg_stream_t s);
/* y = act((x - mean) * inv_std * gamma + beta + residual)   (archs.py:231-234, batchnorm.py:73-77,
 * models_seg_gan.py:43-49).  gamma/beta/residual may be NULL. */
int ssg_bn_apply(const void* x, const void* residual, void* y, int dtype, long long rows, int c, const float* mean,
                 const float* inv_std, const float* gamma, const float* beta, int act, float slope, ssg_stream_t s);
/* eval-mode BN (running stats): same as ssg_bn_apply with inv_std = rsqrt(var + eps) computed inline */
int ssg_bn_eval_prepare(const float* running_mean, const float* running_var, int c, float eps, float* mean,
                        float* inv_std, ssg_stream_t s);
/* backward, pass 1: dz = dy * act'(y); sums[0:C] = sum dz, sums[C:2C] = sum dz * xhat (fp64, overwritten) */
int ssg_bn_bwd_reduce(const void* dy, const void* y, const void* x, int dtype, long long rows, int c,
                      const float* mean, const float* inv_std, int act, float slope, double* sums, ssg_stream_t s);
/* backward, pass 2: dx = gamma*inv_std*(dz - sums0/count - xhat*sums1/count); dres = dz (optional).
 * With training == 0 (eval mode): dx = gamma*inv_std*dz. */
int ssg_bn_bwd_apply(const void* dy, const void* y, const void* x, void* dx, void* dres, int dtype, long long rows,
                     int c, const float* mean, const float* inv_std, const float* gamma, const double* sums,
                     double count, int act, float slope, int training, ssg_stream_t s);
/* dgamma = sums[C:2C], dbeta = sums[0:C] as fp32 */
int ssg_bn_param_grads(const double* sums, int c, float* dgamma, float* dbeta, ssg_stream_t s);
/* Same, ADDED into dgamma / dbeta: the caller passes the parameters' slots of the optimiser's flat gradient arena (zeroed by
 * zero_grad; the discriminator runs two backward passes per step), replacing autograd's AccumulateGrad addition. */
int ssg_bn_param_grads_acc(const double* sums, int c, float* dgamma, float* dbeta, ssg_stream_t s);
/* dst[i] += (float)src[i]: a bias gradient (fp64 column sums of dy) added into its gradient-arena slot. */
int ssg_accum_f64_f32(const double* src, float* dst, int n, ssg_stream_t s);
/* ssg_bn_finalize that also advances nn.BatchNorm2d's `num_batches_tracked` (int64 device scalar, may be NULL): the
 * F.batch_norm path of batchnorm.py:52-55 / torch's BatchNorm2d.forward increments it once per training forward. */
int ssg_bn_finalize_count(const double* sums, double count, int c, float eps, float momentum, int sync_quirk, float* running_mean,
                          float* running_var, float* mean, float* inv_std, long long* num_batches_tracked, const float* gamma,
                          const float* beta, float* sc_out, float* sh_out, ssg_stream_t s);
/* BN backward for layers WITHOUT a residual operand that does not re-read the forward output: the activation mask is
 * recomputed from x with the forward's own expression act'(fma(x, sc, sh)), sc / sh as ssg_bn_finalize_count stored them at
 * forward time (a weight clamp between forward and backward, train.py:111-112, must not move the mask) -- one full-size
 * operand less per pass (archs.py:231, models_seg_gan.py:43-49).  bf16: C % 8 == 0 and C <= 2048; fp32: C % 4 == 0, C <= 1024.
 * `gamma` is the LIVE parameter (dx = gamma * inv_std * (...), as torch's backward reads it). */
int ssg_bn_bwd_reduce_rc(const void* dy, const void* x, int dtype, long long rows, int c, const float* mean, const float* inv_std,
                         const float* fwd_sc, const float* fwd_sh, int act, float slope, double* sums, ssg_stream_t s);
int ssg_bn_bwd_apply_rc(const void* dy, const void* x, void* dx, int dtype, long long rows, int c, const float* mean,
                        const float* inv_std, const float* gamma, const float* fwd_sc, const float* fwd_sh, const double* sums,
                        double count, int act, float slope, int training, ssg_stream_t s);

/* ---- pooling / resampling ------------------------------------------------------------------ */
/* nn.MaxPool2d(2,2,return_indices=True) (archs.py:571): y [n,h/2,w/2,c]; code in {0..3} = 2*dy+dx of the
 * first maximum in row-major window order (ATen tie rule). */
int ssg_maxpool2x2_fwd(const void* x, void* y, uint8_t* code, int dtype, int n, int h, int w, int c, ssg_stream_t s);
/* nn.MaxUnpool2d(2,2) forward (archs.py:572,648,654,659) == max-pool backward: dst [n,2h,2w,c] */
int ssg_scatter2x2(const void* src, const uint8_t* code, void* dst, int dtype, int n, int h, int w, int c, ssg_stream_t s);
/* max-unpool backward == gather by code: src [n,2h,2w,c] -> dst [n,h,w,c] */
int ssg_gather2x2(const void* src, const uint8_t* code, void* dst, int dtype, int n, int h, int w, int c, ssg_stream_t s);
/* nn.Upsample(scale_factor=2, bilinear, align_corners=True) (archs.py:573,664,667) and adjoint */
int ssg_upsample2x_fwd(const void* x, void* y, int dtype, int n, int h, int w, int c, ssg_stream_t s);
int ssg_upsample2x_bwd(const void* dy, void* dx, int dtype, int n, int h, int w, int c, ssg_stream_t s);
/* nn.Upsample(scale_factor=2) (default mode 'nearest'; up_conv, archs.py:848-861) and adjoint (sum of the 2x2 block);
 * x / dx [n,h,w,c], y / dy [n,2h,2w,c] */
int ssg_upsample_nearest2x_fwd(const void* x, void* y, int dtype, int n, int h, int w, int c, ssg_stream_t s);
int ssg_upsample_nearest2x_bwd(const void* dy, void* dx, int dtype, int n, int h, int w, int c, ssg_stream_t s);
/* Attention gate (Attention_block.forward, archs.py:136-142): y[r,:] = x[r,:] * sigmoid(z[r]); z is the [rows] output of
 * psi's BatchNorm2d(1).  Backward: dx = dy * s, dz[r] = s (1 - s) * sum_c dy[r,c] x[r,c]. */
int ssg_pixel_gate_fwd(const void* x, const void* z, void* y, int dtype, long long rows, int c, ssg_stream_t s);
int ssg_pixel_gate_bwd(const void* dy, const void* x, const void* z, void* dx, void* dz, int dtype, long long rows, int c,
                       ssg_stream_t s);
/* nn.AdaptiveAvgPool2d((oh,ow)) + view(batch,-1) (models_seg_gan.py:277,295-296): y is [n, c*oh*ow] in the
 * reference's NCHW flatten order (channel-major). */
int ssg_adaptive_avgpool_flat_fwd(const void* x, void* y, int dtype, int n, int h, int w, int c, int oh, int ow, ssg_stream_t s);
int ssg_adaptive_avgpool_flat_bwd(const void* dy, void* dx, int dtype, int n, int h, int w, int c, int oh, int ow, ssg_stream_t s);

/* ---- SPADE modulation (normalization.py:120) ------------------------------------------------ */
/* gb is [rows, 2C]: gamma = gb[:, :C], beta = gb[:, C:]; y = x * (1 + gamma) + beta */
int ssg_spade_modulate_fwd(const void* x, const void* gb, void* y, int dtype, long long rows, int c, ssg_stream_t s);
/* dx_part = dy*(1+gamma); dgb = [dy*x, dy] */
int ssg_spade_modulate_bwd(const void* dy, const void* x, const void* gb, void* dx, void* dgb, int dtype, long long rows,
                           int c, ssg_stream_t s);

/* Same, and colsum[0:2C] (fp64, overwritten) = per-channel sums of the stored dgb = the bias gradients of the gamma | beta
 * convolutions (normalization.py:97-98): saves the separate reduction pass over dgb.  C % 8 == 0 (bf16) / C % 4 == 0 (fp32). */
int ssg_spade_modulate_bwd_sums(const void* dy, const void* x, const void* gb, void* dx, void* dgb, int dtype, long long rows,
                                int c, double* colsum, ssg_stream_t s);

/* Whole self-conditioned SPADE forward in one kernel (normalization.py:106-122 with segmap = x): x is staged once per 16 x 16
 * tile with a 3-pixel halo, the three thin convolutions run as warp-level mma.sync out of shared memory and the modulation is
 * applied to the accumulators (csrc/spade_fused.cu).  OPT-IN (not the default path yet).  x [n,h,w,c] bf16 with c in {64, 128};
 * w1 bf16 [9][8][c] (x2map, tap-major, rows >= label_nc zero); w2 bf16 [8][80] (mlp_shared, k = tap * 8 + ci, zero padded);
 * w3 bf16 [2c][80] (gamma rows then beta rows); b1 / b2 fp32 [8], b3 fp32 [2c]; outputs seg / actv bf16 [n,h,w,8],
 * gb bf16 [n,h,w,2c] (may be NULL: inference) and y bf16 [n,h,w,c]. */
int ssg_spade_fused_supported(int c, int label_nc, int hidden);
int ssg_spade_fused_fwd(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const void* w3, const float* b3,
                        void* seg, void* actv, void* gb, void* y, int n, int h, int w, int c, ssg_stream_t s);
/* Second version (persistent CTAs, weights-stationary x2map, register-blocked gamma|beta stage; c == 64, label_nc <= 3): w1m is
 * bf16 [32][c] with row tap * label_nc + class (zero rows beyond 9 * label_nc).  Written after v1's measurement, not yet run. */
int ssg_spade_fused_fwd_v2(const void* x, const void* w1m, const float* b1, const void* w2, const float* b2, const void* w3, const float* b3,
                           void* seg, void* actv, void* gb, void* y, int n, int h, int w, int c, int label_nc, ssg_stream_t s);

/* ---- elementwise ---------------------------------------------------------------------------- */
int ssg_act_fwd(const void* x, void* y, int dtype, long long n, int act, float slope, ssg_stream_t s);
int ssg_act_bwd(const void* dy, const void* y, void* dx, int dtype, long long n, int act, float slope, ssg_stream_t s);
/* out = a + b */
int ssg_add(const void* a, const void* b, void* out, int dtype, long long n, ssg_stream_t s);
/* generator_output[isnan] = 0 (train_seg_gan.py:190,269); fp32; bwd zeroes the gradient there */
int ssg_nan_scrub_fwd(const float* x, float* y, long long n, ssg_stream_t s);
int ssg_nan_scrub_bwd(const float* dy, const float* x, float* dx, long long n, ssg_stream_t s);

/* ---- linear layers (models_seg_gan.py:279-283,296-298); w fp32 [nout, k] ---------------------- */
int ssg_linear_fwd(const void* x, const float* w, const float* bias, void* y, int dtype, int m, int k, int nout,
                   int act, float slope, const float* inv_scale_dev, ssg_stream_t s);
/* dx32 is ALWAYS fp32 [m][k] (accumulated with atomics over output-feature splits); dy is `dtype` */
int ssg_linear_dgrad(const void* dy, const float* w, float* dx32, int dtype, int m, int k, int nout,
                     const float* inv_scale_dev, ssg_stream_t s);
int ssg_linear_wgrad(const void* x, const void* dy, float* dw, float* dbias, int dtype, int m, int k, int nout, ssg_stream_t s);

/* ---- losses (losses.py:130-136,274-302; train_seg_gan.py:194-195,204-205,221-222) ------------- */
/* logits/target: fp32 [batch, per_sample] (NCHW flatten).  sums: fp64 [batch][5] =
 * {sum bce_elem, sum sigmoid*t, sum sigmoid, sum t, sum (x-t)^2}, overwritten. */
int ssg_seg_loss_sums(const float* logits, const float* target, int batch, long long per_sample, double* sums, ssg_stream_t s);
/* out[0] = BCEDiceLoss (with the NaN/Inf -> 2*dice branch), out[1] = bce, out[2] = dice term,
 * out[3] = MSELoss, out[4] = 1.0 if the NaN/Inf branch was taken */
int ssg_seg_loss_finalize(const double* sums, int batch, long long per_sample, float* out, ssg_stream_t s);
/* dlogits = g[0] * dBCEDice/dx + g[1] * dMSE/dx + g[2] * dStableBCE/dx ; g_loss is a device float[3] */
int ssg_seg_loss_bwd(const float* logits, const float* target, const double* sums, const float* out, const float* g_loss,
                     int batch, long long per_sample, float* dlogits, ssg_stream_t s);
/* nn.BCEWithLogitsLoss(x, full_like(x, target_value)) mean over n (n small); out: device float */
int ssg_bce_logits_fwd(const float* x, float target_value, int n, float* out, ssg_stream_t s);
int ssg_bce_logits_bwd(const float* x, float target_value, int n, const float* g, float* dx, ssg_stream_t s);

/* ---- metrics (metrics.py:6-35) ---------------------------------------------------------------- */
/* counts[0] = sum((sigmoid(x) > .5) & (t > .5)), counts[1] = sum(|) ; int64, overwritten */
int ssg_iou_counts(const float* logits, const float* target, long long n, long long* counts, ssg_stream_t s);
/* NumPy float32 pairwise-summation tree (PW_BLOCKSIZE 128): host helpers + leaf kernel.
 * ssg_pairwise_leaves_host returns the number of leaves and fills offsets[0..leaves] (offsets[leaves] = n). */
long long ssg_pairwise_leaves_host(long long n, long long* offsets_host, long long capacity);
/* leaf_sums: float [3][n_leaves] = per-leaf sums of sigmoid(x)*t, sigmoid(x), t in NumPy's order.
 * If probs_out != NULL the fp32 probabilities are also written (for oracle comparison). */
int ssg_dice_leaf_sums(const float* logits, const float* target, const long long* offsets_dev, long long n_leaves,
                       float* leaf_sums, float* probs_out, ssg_stream_t s);
/* combines leaf sums (host memory) up the same tree; returns the float32 total */
float ssg_pairwise_combine_host(const float* leaf_sums_host, long long n);

/* ---- data feed (dataset.py:95-144; Normalize / Flip of the pipelines at train_seg_gan.py:366-382) -------------------- */
/* img: uint8 [n,h,w,c] (cv2.imread layout, c <= 8).  out = (float(img) - sub[c]) * mul[c] with sub = mean * max_pixel_value and
 * mul = 1 / (std * max_pixel_value) precomputed in float32 (albumentations' Normalize).  nchw != 0: out is the reference's
 * NCHW tensor [n,c,h,w] in `dtype`; nchw == 0: out is NHWC [n,h,w,c_store] with zero padding channels (the layout the first
 * convolution reads).  post_div != 0: the result is then divided by post_div in float32 (`get_patched_input` divides the
 * normalised patch by 255 once more, aerial_image_segmentation_api.py:367).  flip_codes (optional, int per sample): bit 0
 * reverses x, bit 1 reverses y (albumentations Flip). */
int ssg_feed_image_u8(const unsigned char* img, void* out, int dtype, int nchw, int n, int h, int w, int c, int c_store, const float* sub,
                      const float* mul, float post_div, const int* flip_codes, ssg_stream_t s);
/* mask: uint8 [n,h,w,classes] (the per-class PNGs of dataset.py:126-131 stacked): out fp32 [n,classes,h,w] =
 * (uint8)(float32(mask) / 255.0), i.e. 1.0 only where the PNG holds 255; same flip codes. */
int ssg_feed_mask_u8(const unsigned char* mask, float* out_nchw, int n, int h, int w, int classes, const int* flip_codes, ssg_stream_t s);

/* ---- optimiser (srgan_utils.py:186-195 + torch.optim.Adam, train_seg_gan.py:452,468) ---------- */
int ssg_clamp_(float* g, long long n, float clip, ssg_stream_t s);
/* g <- clamp(g*grad_scale, +-clip) (clip <= 0: no clamp); Adam update of p, m, v in place */
int ssg_clamp_adam(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                   float bias_corr1, float bias_corr2, float clip, float grad_scale, ssg_stream_t s);

/* Same update with the step count t kept on the device (float, incremented by the call before use; bias corrections
 * 1 - beta^t are computed in the kernel): the launch carries no per-step host scalar, so a whole training step can be
 * captured in a CUDA graph and replayed. */
int ssg_clamp_adam_dev(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                       float* step_dev, float clip, float grad_scale, ssg_stream_t s);
/* Both with torch.optim.Adam's L2 `weight_decay` (train.py:290, config "weight_decay"): the moments see g + wd * p, the stored
 * gradient stays the clamped one. */
int ssg_clamp_adam_wd(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                      float bias_corr1, float bias_corr2, float clip, float grad_scale, float weight_decay, ssg_stream_t s);
int ssg_clamp_adam_wd_dev(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                          float* step_dev, float clip, float grad_scale, float weight_decay, ssg_stream_t s);
/* Same with the hyper-parameters on the device too: hp_dev = float[8] {lr, beta1, beta2, eps, clip (0 = none), grad_scale,
 * weight_decay, unused}.  A launch captured in a CUDA graph then follows an LR scheduler or a changed clip: the host rewrites
 * hp_dev between replays. */
int ssg_clamp_adam_hp_dev(float* p, float* g, float* m, float* v, long long n, const float* hp_dev, float* step_dev, ssg_stream_t s);

/* ---- spectral norm (spectral_norm.py:38-88) ---------------------------------------------------- */
/* One power iteration on W [rows, cols] fp32: v = normalize(W^T u); u = normalize(W v); sigma = u.(W v).
 * u, v updated in place when do_power_iteration != 0.  inv_sigma: device float[2] = {1/sigma, sigma}.
 * workspace: (rows + cols) floats. */
int ssg_spectral_sigma(const float* w, float* u, float* v, int rows, int cols, float eps, int do_power_iteration,
                       float* inv_sigma, float* workspace, ssg_stream_t s);
/* dW_orig = (dW_sn - <dW_sn, W_orig> * inv_sigma * u v^T) * inv_sigma ; dot_ws: device double[1] */
int ssg_spectral_weight_bwd(const float* dw_sn, const float* w_orig, const float* u, const float* v, int rows, int cols,
                            const float* inv_sigma, double* dot_ws, float* dw_orig, ssg_stream_t s);
/* out = w * inv_sigma[0] */
int ssg_scale_by_dev(const float* w, const float* inv_sigma, float* out, long long n, ssg_stream_t s);

/* ---- EfficientNet encoder / xResidualBlock: depthwise, squeeze-excite, swish (HBM-bound) ------------------------- */
/* Depthwise convolution (groups == C) of efficientnet_pytorch/model.py:52-55 (Conv2dStaticSamePadding, utils.py:118-141)
 * and xresidualblock.py:17.  w: the parameter itself, fp32 [C][1][k][k]; bias fp32 [C] or NULL.  TF-style "same" padding
 * is asymmetric: (pad_t, pad_l) zero rows/columns lead the input, the trailing padding is whatever the given output
 * size (oh, ow) implies.  C must be a multiple of 8 (bf16) / 4 (fp32); k <= 11; stride 1 or 2. */
int ssg_dwconv2d_fwd(const void* x, const float* w, const float* bias, void* y, int dtype, int n, int h, int w_, int c, int k, int stride,
                     int pad_t, int pad_l, int oh, int ow, ssg_stream_t s);
/* dx [n,h,w_,c] from dy [n,oh,ow,c] */
int ssg_dwconv2d_dgrad(const void* dy, const float* w, void* dx, int dtype, int n, int h, int w_, int c, int k, int stride, int pad_t,
                       int pad_l, int oh, int ow, ssg_stream_t s);
/* dw fp32 [C][1][k][k], overwritten */
int ssg_dwconv2d_wgrad(const void* x, const void* dy, float* dw, int dtype, int n, int h, int w_, int c, int k, int stride, int pad_t,
                       int pad_l, int oh, int ow, ssg_stream_t s);
/* Per-sample, per-channel plane sums: sums[n][c] = sum_hw a[n,:,:,c] * (b ? b[n,:,:,c] : 1), fp32, overwritten.
 * b == NULL: F.adaptive_avg_pool2d(x, 1) * HW (model.py:80); with b: the gate gradient sum_hw dy * x. */
int ssg_plane_sums(const void* a, const void* b, float* sums, int dtype, int n, int hw, int c, ssg_stream_t s);
/* Squeeze-excite gate MLP on the pooled vector (model.py:80-82): pooled = pooled_sum / hw;
 * s_pre = W1 pooled + b1 (w1 fp32 [sq][c]); gate = sigmoid(W2 swish(s_pre) + b2) (w2 fp32 [c][sq]).
 * Outputs pooled [n][c], s_pre [n][sq], gate [n][c] (fp32; kept for the backward). */
int ssg_se_gate_fwd(const float* pooled_sum, int n, int hw, int c, int sq, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* pooled, float* s_pre, float* gate, ssg_stream_t s);
/* dgate [n][c] -> dpooled [n][c] (dL/dx contribution per pixel, i.e. already divided by hw) and the MLP parameter
 * gradients dw1/db1/dw2/db2 (overwritten, summed over samples). */
int ssg_se_gate_bwd(const float* dgate, const float* gate, const float* s_pre, const float* pooled, const float* w1, const float* w2,
                    int n, int hw, int c, int sq, float* dpooled, float* dw1, float* db1, float* dw2, float* db2, ssg_stream_t s);
/* y[n,p,c] = a[n,p,c] * mul[n][c] + (add ? add[n][c] : 0): torch.sigmoid(x_squeezed) * x (model.py:82), drop_connect's
 * per-sample scale (utils.py:83-93), and the SE input gradient dy * gate + dpooled. */
int ssg_plane_scale(const void* a, const float* mul, const float* add, void* y, int dtype, int n, int hw, int c, ssg_stream_t s);
/* MemoryEfficientSwish (utils.py:36-53): y = x * sigmoid(x); dx = dy * (sig * (1 + x * (1 - sig))) */
int ssg_swish_fwd(const void* x, void* y, int dtype, long long n, ssg_stream_t s);
int ssg_swish_bwd(const void* dy, const void* x, void* dx, int dtype, long long n, ssg_stream_t s);
/* xresidualblock.py:4-6,19-23: y = x1 * exp(-z*z) and its gradients */
int ssg_gauss_gate_fwd(const void* x1, const void* z, void* y, int dtype, long long n, ssg_stream_t s);
int ssg_gauss_gate_bwd(const void* dy, const void* x1, const void* z, void* dx1, void* dz, int dtype, long long n, ssg_stream_t s);
/* nn.ZeroPad2d (utils.py:133-136): y[n,oy,ox,:] = x[n,oy-pad_t,ox-pad_l,:] or 0; negative pads crop (the adjoint). */
int ssg_pad2d(const void* x, void* y, int dtype, int n, int h, int w, int c, int pad_t, int pad_l, int oh, int ow, ssg_stream_t s);
/* F.interpolate(size=(oh,ow), mode='bilinear', align_corners=False) (archs.py:459) on NHWC storage; the adjoint
 * accumulates into an fp32 buffer dx32 [n,h,w,c] (overwritten). */
int ssg_resize_bilinear_fwd(const void* x, void* y, int dtype, int n, int h, int w, int c, int oh, int ow, ssg_stream_t s);
int ssg_resize_bilinear_bwd(const void* dy, float* dx32, int dtype, int n, int h, int w, int c, int oh, int ow, ssg_stream_t s);

/* ---- SyncBN statistics exchange over NVLink peer memory (batchnorm.py:95-112, comm.py:74-136) ---------------------- */
/* Size in bytes of the symmetric receive buffer every rank must allocate (zero-filled before first use) and map into all
 * ranks' address spaces: data [2][world][slot_len] fp64 + flags [2][world] uint32. */
long long ssg_p2p_buffer_bytes(int world, int slot_len);
/* In-place sum of `data` (n <= slot_len fp64 values: [sum x | sum x^2] or [sum dy | sum dy*xhat]) over all ranks.
 * peer_bufs_dev: DEVICE array of `world` pointers, entry p = rank p's receive buffer as mapped in THIS process;
 * epoch_dev: device uint32 call counter owned by this rank (starts at 0, advanced by the kernel).  One single-CTA launch:
 * remote stores + release/acquire flags, reduction in rank order (bit-identical on all ranks).  Every rank must issue
 * the same sequence of calls. */
int ssg_p2p_allreduce_f64(double* data, int n, void* const* peer_bufs_dev, int rank, int world, int slot_len, unsigned* epoch_dev,
                          ssg_stream_t s);
/* Same with a wall-clock bound on the wait for the peers: after timeout_ns nanoseconds (%globaltimer) the kernel stops
 * waiting, stores 1 + the index of the first missing source rank into *status_dev (device uint32, 0 = healthy) and returns;
 * the context stays usable and the host reports the failure when it next looks at the flag.  timeout_ns <= 0 waits forever
 * (what NCCL would do). */
int ssg_p2p_allreduce_f64_to(double* data, int n, void* const* peer_bufs_dev, int rank, int world, int slot_len, unsigned* epoch_dev,
                             long long timeout_ns, unsigned* status_dev, ssg_stream_t s);

/* ---- tiled inference merge (aerial_image_segmentation_api.py:129-217) ------------------------------ */
/* values: fp32 [patches][classes][patch_size][patch_size] (NCHW logits with apply_sigmoid != 0, else probabilities in
 * [0, 1]); windows: int32 [patches][2] = (top, left) of each patch in the h x w raster.  Adds, per class and pixel, one
 * vote when (uint8)(p * 255) > 127 into pos_votes int32 [classes][h][w], and 1 into patch_count int32 [h][w] (both must be
 * zeroed by the caller before the first batch; batches accumulate). */
int ssg_mask_vote(const float* values, const int* windows, int patches, int classes, int patch_size, int h, int w, int apply_sigmoid,
                  int* pos_votes, int* patch_count, ssg_stream_t s);
/* masks uint8 [classes][h][w] = post_process((uint8)(votes / max(count, 1) * 255)) in {0, 255}, fp64 arithmetic as numpy */
int ssg_mask_finalize(const int* pos_votes, const int* patch_count, int classes, int h, int w, unsigned char* masks, ssg_stream_t s);
/* The reference's resize bridge (patch_size != network size, e.g. config_v1.json: 1024 vs 512).
 * ssg_resize_u8_linear: cv2.resize(uint8 NHWC [n,h,w,c] -> [n,oh,ow,c]) with the default INTER_LINEAR, bit-exact (OpenCV's 11-bit
 * fixed-point arithmetic; exact 2x shrinking takes cv::resize's INTER_AREA fast path) -- `get_patched_input`'s shrink of every
 * image patch (:361).  xtab / ytab: [ow] / [oh] entries of 4 int32 {tap index 0, tap index 1, coef 0, coef 1} built by the host
 * (may be NULL for the exact 2x shrink).
 * ssg_mask_vote_resized: ssg_mask_vote for map_size x map_size maps that `patch_merge` first quantises to uint8 and upsamples to
 * patch_size with cv2.resize (:150-152); the upsampled map is never materialised (tables over the patch_size outputs). */
int ssg_resize_u8_linear(const unsigned char* src, unsigned char* dst, int n, int h, int w, int c, int oh, int ow, const int* xtab,
                         const int* ytab, ssg_stream_t s);
int ssg_mask_vote_resized(const float* values, const int* windows, int patches, int classes, int map_size, int patch_size, int h, int w,
                          int apply_sigmoid, const int* xtab, const int* ytab, int* pos_votes, int* patch_count, ssg_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* SSUNET_B200_H */
