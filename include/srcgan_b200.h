/*
 * srcgan_b200 - C ABI of the B200-native (sm_100a) kernels behind the SRCGAN G+D training
 * hot path.  Plain pointers and sizes only: every buffer is DEVICE memory owned by the
 * caller (PyTorch on the Python side), every call is asynchronous on `stream`
 * (a cudaStream_t passed as void*), every call returns 0 on success or a non-zero
 * SRCGAN_E_* code, with a thread-local message available from srcgan_last_error().
 * The library never allocates user-visible memory and has no CPU fallback.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   srcgan_conv_fprop / _dgrad / _wgrad   nn.Conv2d forward / autograd backward used by
 *        ResidualDenseBlock_5 (src/model/model.py:193-211), RRDB (:214-233), RDDBNet /
 *        RDDBNetB trunk, upconv, HRconv, conv_first/last (:347-440), Decoder (:236-289) and
 *        NLayerDiscriminator's 4x4 convs (:612-635); torch.cat of the dense block
 *        (:207-210), the in-place LeakyReLU (:202,:410), the x0.2 residuals (:211,:233) and
 *        F.interpolate(nearest, x2) (:426-427) are folded into the operand addressing and
 *        the epilogue.
 *   srcgan_bn_*                           nn.BatchNorm2d train/eval forward + backward
 *        (model.py:240-260, :622, :630) fused with the following LeakyReLU.
 *   srcgan_upsample2x_adjoint             backward of F.interpolate(nearest, x2) (:426-427).
 *   srcgan_loss_fwd_bwd                   nn.L1Loss / nn.MSELoss forward+backward
 *        (src/losses.py:95-133, src/train.py:87 via GANLoss).
 *   srcgan_metrics_* / srcgan_ssim        src/metrics.py:10-144, src/losses.py:20-93,136-147.
 *   srcgan_rgb2lab / srcgan_lab2rgb       skimage.color calls at src/dataset.py:148-159,
 *        src/utils.py:22-26.
 */
#ifndef SRCGAN_B200_H
#define SRCGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRCGAN_OK 0
#define SRCGAN_E_INVALID 1      /* bad argument / unsupported shape (no fallback exists) */
#define SRCGAN_E_CUDA 2         /* CUDA runtime / driver error */
#define SRCGAN_E_WORKSPACE 3    /* workspace too small */

#define SRCGAN_DT_F32 0
#define SRCGAN_DT_BF16 1

/* packed-weight layouts produced by srcgan_pack_weights from fp32 OIHW */
#define SRCGAN_WL_RSCK 0        /* [kh][kw][cin][cout]  - fprop, SIMT engine             */
#define SRCGAN_WL_RSKC 1        /* [kh][kw][cout][cin]  - dgrad, SIMT engine             */
#define SRCGAN_WL_TC   2        /* tcgen05 engine, stride 1: [cout/BN][cin/64][kw][kh][BN][64] bf16, pre-swizzled */
#define SRCGAN_WL_TC_S2 3       /* tcgen05 engine, stride-2 fprop (taps grouped by row parity)                  */
#define SRCGAN_WL_TC_DGRAD_S2 4 /* tcgen05 engine, stride-2 dgrad (pad 1): four output-phase packs, transposed  */

#define SRCGAN_ENGINE_AUTO 0
#define SRCGAN_ENGINE_SIMT 1    /* fp32-accumulating FFMA implicit GEMM (fp32 parity mode) */
#define SRCGAN_ENGINE_TC   2    /* TMA + tcgen05/TMEM implicit GEMM (bf16)                 */

/*
 * One convolution over channel-sliced NHWC tensors.  A tensor is (ptr, ld): element
 * (n,y,x,c) lives at ptr[((n*H + y)*W + x)*ld + c], so a slice of a wider concat buffer
 * is expressed by offsetting ptr and keeping ld = the buffer's channel count.
 *
 * The struct always describes the FORWARD geometry:  X[n,h,w,cin] (*) W -> Y[n,ho,wo,cout],
 * ho = (h*(upsample?2:1) + 2*pad - kh)/stride + 1.
 *   fprop : reads x (=X) and wgt, writes y (=Y).
 *   dgrad : reads x (=dY, ho x wo x cout), writes y (=dX, h x w x cin); upsample must be 0.
 *           (SIMT engine: any stride, RSKC weights.  tcgen05 engine: stride 2 only, TC_DGRAD_S2 weights;
 *           a stride-1 dgrad on tcgen05 is an fprop over transposed+rotated weights.)
 *   wgrad : reads x (=X) and y (=dY); writes alpha * (the fp32 OIHW gradient) passed separately.
 * Epilogue applied to each written element v (fprop and dgrad):
 *   v = acc + bias[c];  if (act) v = v > 0 ? v : act_slope*v;
 *   v = alpha*v + beta1*r1 + beta2*r2;
 *   if (mask) v *= (mask > 0 ? 1 : mask_slope);
 */
typedef struct srcgan_conv_params {
  int32_t n, h, w;
  int32_t cin, cout;
  int32_t kh, kw, stride, pad;
  int32_t upsample;
  int32_t ho, wo;
  int32_t dtype;      /* SRCGAN_DT_* of x, y, wgt, r1, r2, mask */
  int32_t engine;     /* SRCGAN_ENGINE_* */
  const void* x;
  int32_t x_ld;
  const void* wgt;
  const float* bias;
  void* y;
  int32_t y_ld;
  int32_t act;
  float act_slope;
  float alpha;
  const void* r1;
  int32_t r1_ld;
  float beta1;
  const void* r2;
  int32_t r2_ld;
  float beta2;
  const void* mask;
  int32_t mask_ld;
  float mask_slope;
  /* Packed LeakyReLU masks (paired-sweep fprop kernel only; any other kernel returns SRCGAN_E_INVALID if one is set).
     Layout [n][ho][wo][cout/32] uint32, bit i of word j = channel 32j+i.
       signbits : OUT, bit = (value stored to y > 0) - what `mask` of the matching backward step would test
       maskbits : IN,  replaces `mask`: value *= bit ? 1 : mask_slope  (4 bytes instead of 64 per pixel and 32 channels) */
  void* signbits;
  const void* maskbits;
  /* "Tall image" batching of small maps (paired-sweep fprop kernel only): the caller stacks the images of a batch vertically
     in ONE image of n*(h+1)+1 rows with an all-zero separator row above, between and below them (the shared zero padding),
     so that 64x64 maps fill the 128-lane strips of the sweep kernel.  zero_row_period = h+1 makes the epilogue store zeros
     in the separator rows (row % period == 0) so that they stay zero from layer to layer.  0 = off. */
  int32_t zero_row_period;
  /* SRCGAN_CONV_FLAG_* bits.  REVERSE (paired-sweep fprop kernel; ignored elsewhere): process the work units from the last
     image to the first.  Alternating it between the consecutive layers of a dense block lets each layer start on the
     data the previous one touched last, i.e. on what the 126 MB L2 still holds.  Results do not depend on it beyond the
     fp32 summation order of a few columns (each setting is bit-reproducible). */
  int32_t flags;
  /* Planar concat buffers (paired-sweep fprop kernel only): x points at group 0 of a buffer laid out [group][n][h][w][64
     channels] (x_ld = 64); channel c of the slice lives in group c / 64, x_group_stride ELEMENTS further per group.  0 =
     ordinary interleaved NHWC.  A dense block's 192-channel concat buffer as three such groups turns the 64-byte slice writes
     at a 384-byte pitch (28 % slower on 64->32) into writes at a 128-byte pitch.  Outputs and side operands
     never span groups, so they stay (pointer, ld). */
  int64_t x_group_stride;
} srcgan_conv_params;
#define SRCGAN_CONV_FLAG_REVERSE 1

const char* srcgan_version(void);
const char* srcgan_last_error(void);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches) */
int64_t srcgan_launch_count(void);
/* name (a static string) of the kernel the calling thread's most recent call launched last, e.g.
   "conv3x3_sweep2_tc<32,2>" - bench.py keys its per-kernel timing on it */
const char* srcgan_last_kernel(void);

int srcgan_pack_weights(const float* w_oihw, int cout, int cin, int kh, int kw, int layout, int dtype,
                        void* out, void* stream);
size_t srcgan_packed_weight_bytes(int cout, int cin, int kh, int kw, int layout, int dtype);

/* Batched re-pack of tcgen05-layout (SRCGAN_WL_TC) weights: ONE launch for all packed tensors of a network.  A packed
   tensor is a list of source blocks concatenated along the GEMM-K channel axis; the table lives in DEVICE memory.
   srcgan_pack_slots() gives the slot table (tap offset kh*kw_size + kw per slot) and the N tile BN for a layer shape. */
typedef struct srcgan_pack_block {
  const float* src;       /* fp32 OIHW parameter                                                     */
  void* out;              /* base of the packed tensor (bf16); its K / N padding must already be zero */
  int64_t nstride, kstride;   /* element strides of the GEMM-N / GEMM-K channel index in src          */
  int64_t src_off;        /* element offset of this block's (n = 0, k = 0, tap 0)                     */
  int64_t elem0;          /* first global work index of this block (prefix sum of n_pad*slots*k_len)  */
  int32_t n_count;        /* GEMM-N channels that exist in src (rows n_count .. n_pad are written 0)  */
  int32_t k0, k_len;      /* the block fills packed K channels [k0, k0 + k_len)                       */
  int32_t bn, nchunks, total_slots;
  int32_t taps;           /* kh * kw                                                                  */
  int32_t flip;           /* 1: rotate the taps by 180 degrees (dgrad)                                */
  float scale;
  int32_t slot_off[16];
} srcgan_pack_block;
int srcgan_pack_slots(int cout, int kh, int kw, int layout, int32_t* slot_off16, int32_t* nslots, int32_t* bn);
int srcgan_pack_weights_batch(const srcgan_pack_block* blocks_dev, int nblocks, int64_t total_elems, void* stream);

int srcgan_conv_fprop(const srcgan_conv_params* p, void* stream);
int srcgan_conv_dgrad(const srcgan_conv_params* p, void* stream);
/* Two consecutive layers of a dense block in ONE launch (tcgen05 engine; replaces two srcgan_conv_fprop calls for
   x_k = lrelu(conv_k(cat(x, x1..x_(k-1)))), x_(k+1) = lrelu(conv_(k+1)(cat(x, x1..x_k))) of src/model/model.py:205-210, and for
   the same two pairs of the mirrored dense block of the backward pass).  pa reads the prefix x[:, :cin] (cin = 64 | 128) and
   writes the 32-channel slice right behind it; pb reads [prefix | that slice] (same x, x_ld; cin + 32) and writes another
   32-channel slice.  3x3 stride 1 pad 1, bf16, lane extent (the image width if >= 96, else the height) <= 256.  Epilogue fields
   honoured: bias, act / act_slope, signbits, maskbits / mask_slope.  The prefix is read from HBM once and x_k goes from layer A's
   epilogue to layer B's MMAs through shared memory.  Results equal the two separate calls up to fp32 summation order.
   srcgan_conv_fprop_pair_supported() -> 1 if the pair can run fused (else call srcgan_conv_fprop twice). */
int srcgan_conv_fprop_pair_supported(const srcgan_conv_params* pa, const srcgan_conv_params* pb);
int srcgan_conv_fprop_pair(const srcgan_conv_params* pa, const srcgan_conv_params* pb, void* stream);
size_t srcgan_conv_wgrad_workspace_bytes(const srcgan_conv_params* p);
/* dw: fp32 [cout][cin][kh][kw]; db: fp32 [cout] or NULL; accumulate != 0 adds into dw/db */
int srcgan_conv_wgrad(const srcgan_conv_params* p, float* dw, float* db, int accumulate,
                      void* workspace, size_t workspace_bytes, void* stream);

/* One wgrad launch, two destinations (tcgen05 engine, 3x3 stride 1 pad 1, cout 32 / 64): output channels [0, split_cout) of
   dY go to dw0 / db0, the rest to dw1 / db1.  dwX is an fp32 OIHW tensor with cin_ldX input channels; the launch writes its
   channels [ci0_X, ci0_X + p->cin).  Either destination (or bias) may be NULL.  Used for the paired dense-block weight
   gradients (conv_k and conv_(k+1) of src/model/model.py:205-210 share their input prefix). */
int srcgan_conv_wgrad_split(const srcgan_conv_params* p, float* dw0, int cin_ld0, int ci0_0, float* db0, float* dw1, int cin_ld1,
                            int ci0_1, float* db1, int split_cout, int accumulate, void* workspace, size_t workspace_bytes,
                            void* stream);

/* layout glue at the module boundary: NCHW fp32 <-> channel-sliced NHWC (f32 | bf16) */
int srcgan_nchw_to_nhwc(const float* src, int n, int c, int h, int w, void* dst, int dst_ld, int dtype, void* stream);
int srcgan_nhwc_to_nchw(const void* src, int src_ld, int dtype, float* dst, int n, int c, int h, int w, void* stream);

/* dst = a + b over channel-sliced NHWC rows (gradient merge at the 'fea + trunk' skip, model.py:421) */
/* dz = dy * (y > 0 ? 1 : slope): backward of an in-place (Leaky)ReLU whose output y was kept */
int srcgan_act_backward(const void* dy, int dy_ld, const void* y, int y_ld, void* dz, int dz_ld, int64_t npix, int c,
                        float slope, int dtype, void* stream);
int srcgan_add(const void* a, int a_ld, const void* b, int b_ld, void* dst, int dst_ld, int64_t npix, int c, int dtype,
               void* stream);
/* in place: y = leaky(y + bias[ch]) (bias may be NULL, has_act = 0 skips the activation) - nn.ConvTranspose2d's bias and the
   LeakyReLU / ReLU that follows it in the reference's generators (src/model/resdeconv.py, src/model/edsr.py) */
int srcgan_bias_act(void* y, int y_ld, int64_t npix, int c, const float* bias, int has_act, float slope, int dtype, void* stream);

/* out[ch] (+)= alpha * sum over the npix rows of x[row, ch]  (bias gradients of a whole dense block in one pass:
 * the gradient concat buffer [dOut | dZ4 | dZ3 | dZ2 | dZ1] holds the output gradients of all five convs) */
size_t srcgan_colsum_workspace_bytes(int64_t npix, int c);
int srcgan_colsum(const void* x, int x_ld, int dtype, int64_t npix, int c, float* out, float alpha, int accumulate,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Non-overlapping transposed convolution (ConvTranspose2d k2 s2, src/model/rddb.py:28-38,94-97) = 1x1 conv to
 * 4*c channels ordered (a,b,co) + this permutation: dst[n,2y+a,2x+b,co] = src[n,y,x,(a*2+b)*c+co].
 * space_to_depth is its adjoint with an optional LeakyReLU mask (mask indexed like dst). */
int srcgan_depth_to_space(const void* src, int src_ld, void* dst, int dst_ld, int n, int h, int w, int c, int dtype,
                          void* stream);
int srcgan_space_to_depth(const void* src, int src_ld, void* dst, int dst_ld, const void* mask, int mask_ld,
                          float mask_slope, int n, int h, int w, int c, int dtype, void* stream);

/* nn.PixelShuffle(r) (src/model/espcn.py:35,50) in NHWC: dst[n,h*r+a,w*r+b,c] = src[n,h,w,c*r*r+a*r+b];
 * (n,h,w) are the LOW-resolution dims, c the number of output channels; adjoint != 0 runs the inverse. */
int srcgan_pixel_shuffle(const void* src, int src_ld, void* dst, int dst_ld, int n, int h, int w, int c, int r,
                         int dtype, int adjoint, void* stream);

/* dst[n,2h,2w,c] = nearest x2 of src[n,h,w,c]  (F.interpolate(scale_factor=2, 'nearest'), model.py:426-427) */
int srcgan_upsample2x(const void* src, int src_ld, void* dst, int dst_ld, int n, int h, int w, int c, int dtype,
                      void* stream);
/* dst[n,h,w,c] = (sum of the 2x2 block of src[n,2h,2w,c]) * (mask>0 ? 1 : mask_slope); mask may be NULL */
int srcgan_upsample2x_adjoint(const void* src, int src_ld, void* dst, int dst_ld, const void* mask, int mask_ld,
                              float mask_slope, int n, int h, int w, int c, int dtype, void* stream);

/* BatchNorm2d over NHWC rows (npix = n*h*w).  Training: batch statistics -> save_mean/save_invstd
 * (fp32 [c]) and running-stat update (momentum, unbiased variance).  Eval: running stats.
 * y = lrelu(gamma*(x-mean)*invstd + beta, slope).  workspace: >= srcgan_bn_workspace_bytes. */
size_t srcgan_bn_workspace_bytes(int64_t npix, int c);
int srcgan_bn_forward(const void* x, int x_ld, void* y, int y_ld, int64_t npix, int c, int dtype,
                      const float* gamma, const float* beta, float* running_mean, float* running_var,
                      float* save_mean, float* save_invstd, int training, float momentum, float eps,
                      float slope, void* workspace, size_t workspace_bytes, void* stream);
/* dy_post: gradient w.r.t. y (after LeakyReLU); y: saved forward output (sign = LeakyReLU mask);
 * x: saved conv output.  Writes dx (may alias dy_post), dgamma/dbeta (+= if accumulate).
 * beta (optional): with it the bf16 kernels RECOMPUTE the LeakyReLU mask from x with the forward's own expression
 * (fmaf(x - mean, gamma * invstd, beta) > 0) and do not read y at all: 5 tensor sweeps instead of 7. */
int srcgan_bn_backward(const void* dy_post, int dy_ld, const void* y, int y_ld, const void* x, int x_ld,
                       void* dx, int dx_ld, int64_t npix, int c, int dtype, const float* gamma, const float* beta,
                       const float* save_mean, const float* save_invstd, float slope, int training,
                       float* dgamma, float* dbeta, int accumulate,
                       void* workspace, size_t workspace_bytes, void* stream);

/* GroupNorm(groups, c) over NHWC rows: statistics per (sample, group) over hw pixels x c/groups channels
 * (nn.GroupNorm(32, C), src/model/resdeconv.py:67,75 and src/model/edsr.py:45); save_mean / save_rstd: fp32 [n*groups].
 * forward fuses what follows the norm in the reference blocks: y = lrelu_slope(gn(x) + residual) (residual may be NULL,
 * act = 0 disables the activation; BasicBlock.forward resdeconv.py:80-98, ResnetBlock.forward edsr.py:48-54). */
size_t srcgan_gn_workspace_bytes(int n, int c);
int srcgan_gn_forward(const void* x, int x_ld, void* y, int y_ld, int n, int64_t hw, int c, int groups, int dtype,
                      const float* gamma, const float* beta, float* save_mean, float* save_rstd, float eps,
                      const void* residual, int res_ld, int act, float slope, void* workspace, size_t workspace_bytes,
                      void* stream);
int srcgan_gn_backward(const void* dy, int dy_ld, const void* x, int x_ld, void* dx, int dx_ld, int n, int64_t hw, int c,
                       int groups, int dtype, const float* gamma, const float* save_mean, const float* save_rstd,
                       float* dgamma, float* dbeta, int accumulate, void* workspace, size_t workspace_bytes,
                       void* stream);

/* kind 0 = L1 (mean |a-b|), 1 = MSE (mean (a-b)^2); b == NULL means the scalar `b_scalar`
 * (GANLoss's expanded label).  loss_out: one fp32 (overwritten).  grad_out (optional, fp32[n]):
 * d loss / d a.  Two-stage deterministic reduction; workspace >= srcgan_loss_workspace_bytes(n). */
size_t srcgan_loss_workspace_bytes(int64_t n);
int srcgan_loss_fwd_bwd(int kind, const float* a, const float* b, float b_scalar, int64_t n,
                        float* loss_out, float* grad_out, void* workspace, size_t workspace_bytes, void* stream);

/* metrics on NCHW fp32 (src/metrics.py): out[0]=sum (a-b)^2 (MSE*count); AE: out[b] = mean angular error (deg) */
int srcgan_metrics_sqerr(const float* a, const float* b, int64_t n, float* out_sum, void* workspace,
                         size_t workspace_bytes, void* stream);
int srcgan_metrics_ae(const float* pred, const float* truth, int n, int c, int h, int w, float* out_per_image,
                      void* stream);
/* SSIM, 11x11 sigma=1.5 'valid' window, data range L; out_per_image[b] = sum of ssim_map over (c,h',w') */
size_t srcgan_ssim_workspace_bytes(int n, int c, int h, int w);
int srcgan_ssim(const float* pred, const float* truth, int n, int c, int h, int w, float L, float* out_per_image,
                void* workspace, size_t workspace_bytes, void* stream);
int srcgan_minmax(const float* a, int64_t n, float* out_min_max, void* stream);

/* The eval sweep's four metrics (src/testCas.py:63-85: metrics.MSE, PSNR, AE, SSIM of src/metrics.py:10-144) in ONE kernel
   launch over pred / truth [n][c][h][w] fp32; SSIM's data range L is chosen on the device from pred's min / max exactly as
   metrics.py:102-111 does on the host.  out (8 + 5n floats): [0] MSE [1] PSNR [2] AE mean (degrees) [3] SSIM mean [4] L
   [5] min(pred) [6] max(pred) [7] 0, then n per-image SSIM means (batch-wide L), n per-image AE, n per-image MSE, n per-image
   SSIM with the image's own data range (what a one-tile-per-call loop like testCas.py:65-85 computes) and those n ranges.  The first 256 bytes of the workspace
   hold a ticket counter: pass a workspace that was zero-filled once (the kernel re-arms it). */
size_t srcgan_eval_metrics_workspace_bytes(int n, int c, int h, int w);
int srcgan_eval_metrics(const float* pred, const float* truth, int n, int c, int h, int w, float* out, void* workspace,
                        size_t workspace_bytes, void* stream);
/* d SSIM-sum / d pred (losses.SSIM / DSSIMLoss are differentiable, src/losses.py:40-93,170-180):
   dpred = *scale_dev * coef * d(sum of the SSIM map)/d pred; L_dev = device scalar holding the data range (out[4] above) */
size_t srcgan_ssim_backward_workspace_bytes(int n, int c, int h, int w);
int srcgan_ssim_backward(const float* pred, const float* truth, int n, int c, int h, int w, const float* L_dev,
                         const float* scale_dev, float coef, float* dpred, void* workspace, size_t workspace_bytes,
                         void* stream);

/* colour: NCHW fp32; rgb in [0,1]; lab normalised as dataset.py:154-157 (L/100,(a,b+128)/255) when normalised!=0 */
int srcgan_rgb2lab(const float* rgb, float* lab, int n, int h, int w, int normalised, void* stream);
int srcgan_lab2rgb(const float* lab, float* rgb, int n, int h, int w, int normalised, void* stream);
/* the dataset glue itself, exact (float64 arithmetic like scikit-image): uint8 [n][h][w][3] image -> normalised LAB
   [n][3][h][w] float32 tensor (Basic._arr2lab, src/dataset.py:148-159) and back with the uint8 truncation of
   Basic._lab2img / utils.tensor2img (src/dataset.py:94-104, src/utils.py:22-26).  Reproduces the reference's example
   PNG tiles bit for bit (tests/test_color.py). */
int srcgan_rgb2lab_u8(const uint8_t* rgb_nhwc, float* lab_nchw, int n, int h, int w, void* stream);
int srcgan_lab2rgb_u8(const float* lab_nchw, uint8_t* rgb_nhwc, int n, int h, int w, void* stream);

#ifdef __cplusplus
}
#endif
#endif
