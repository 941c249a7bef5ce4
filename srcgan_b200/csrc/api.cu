// extern "C" surface of libsrcgan_b200.so (see include/srcgan_b200.h).  No exceptions cross this
// boundary; errors are returned as codes with a thread-local message.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace srcgan {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static thread_local const char* g_last_kernel = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  g_last_kernel = what;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return SRCGAN_E_CUDA;
  }
  return SRCGAN_OK;
}

// engines (conv_simt.cu / conv_tc.cu)
int validate_conv(const srcgan_conv_params* p, bool need_w);
int conv_fprop_simt(const srcgan_conv_params* p, cudaStream_t st);
int conv_dgrad_simt(const srcgan_conv_params* p, cudaStream_t st);
size_t conv_wgrad_simt_workspace(const srcgan_conv_params* p);
int conv_wgrad_simt(const srcgan_conv_params* p, float* dw, float* db, int accumulate, void* ws, size_t ws_bytes,
                    cudaStream_t st);
int pack_weights_simt_host(const float* w, int cout, int cin, int kh, int kw, int layout, int dtype, void* out,
                           cudaStream_t st);
bool conv_tc_supported(const srcgan_conv_params* p);
int conv_fprop_tc(const srcgan_conv_params* p, cudaStream_t st);
bool conv_fprop_pair_supported(const srcgan_conv_params* pa, const srcgan_conv_params* pb);
int conv_fprop_pair_tc(const srcgan_conv_params* pa, const srcgan_conv_params* pb, cudaStream_t st);
int pack_weights_tc_host(const float* w, int cout, int cin, int kh, int kw, int layout, void* out, cudaStream_t st);
size_t packed_weight_bytes_tc(int cout, int cin, int kh, int kw, int layout);
int pack_slots_tc(int cout, int kh, int kw, int layout, int32_t* slot_off16, int32_t* nslots, int32_t* bn);
int pack_weights_batch(const srcgan_pack_block* blocks_dev, int nblocks, long long total, cudaStream_t st);
bool conv_dgrad_tc_supported(const srcgan_conv_params* p);
int conv_dgrad_tc(const srcgan_conv_params* p, cudaStream_t st);
bool conv_wgrad_tc_supported(const srcgan_conv_params* p);
size_t conv_wgrad_tc_workspace(const srcgan_conv_params* p);
int conv_wgrad_tc(const srcgan_conv_params* p, float* dw, float* db, int accumulate, void* ws, size_t ws_bytes,
                  cudaStream_t st);
int conv_wgrad_tc_split(const srcgan_conv_params* p, float* dw0, int ld0, int ci00, float* db0, float* dw1, int ld1, int ci01,
                        float* db1, int split, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);
// other kernels
int nchw_to_nhwc(const float* src, int n, int c, int h, int w, void* dst, int ld, int dtype, cudaStream_t st);
int nhwc_to_nchw(const void* src, int ld, int dtype, float* dst, int n, int c, int h, int w, cudaStream_t st);
int depth_to_space(const void* src, int src_ld, void* dst, int dst_ld, const void* mask, int mask_ld, float mslope,
                   int n, int h, int w, int c, int dtype, int adjoint, cudaStream_t st);
int pixel_shuffle(const void* src, int src_ld, void* dst, int dst_ld, int n, int h, int w, int c, int r, int dtype,
                  int adjoint, cudaStream_t st);
int upsample2x(const void* src, int src_ld, void* dst, int dst_ld, int n, int h, int w, int c, int dtype,
               cudaStream_t st);
int upsample2x_adjoint(const void* src, int src_ld, void* dst, int dst_ld, const void* mask, int mask_ld,
                       float mslope, int n, int h, int w, int c, int dtype, cudaStream_t st);
int bias_act(void* y, int ld, int64_t npix, int c, const float* bias, int has_act, float slope, int dtype, cudaStream_t st);
int add_slices(const void* a, int a_ld, const void* b, int b_ld, void* d, int d_ld, int64_t npix, int c, int dtype,
               cudaStream_t st);
int bias_grad_launch(const void* dy, int dy_ld, int dtype, long long M, int cout, float* db, int accumulate,
                     float alpha, void* ws, cudaStream_t st);
size_t bn_workspace_bytes(int64_t npix, int c);
int bn_forward(const void* x, int x_ld, void* y, int y_ld, int64_t npix, int c, int dtype, const float* gamma,
               const float* beta, float* rm, float* rv, float* save_mean, float* save_invstd, int training,
               float momentum, float eps, float slope, void* ws, size_t ws_bytes, cudaStream_t st);
int bn_backward(const void* dy, int dy_ld, const void* y, int y_ld, const void* x, int x_ld, void* dx, int dx_ld,
                int64_t npix, int c, int dtype, const float* gamma, const float* mean, const float* invstd,
                float slope, int training, float* dgamma, float* dbeta, int accumulate, void* ws, size_t ws_bytes,
                cudaStream_t st, const float* beta);
size_t gn_workspace_bytes(int n, int c);
int gn_forward(const void* x, int x_ld, void* y, int y_ld, int n, int64_t hw, int c, int groups, int dtype,
               const float* gamma, const float* beta, float* mean, float* rstd, float eps, const void* res, int res_ld,
               int act, float slope, void* ws, size_t ws_bytes, cudaStream_t st);
int act_backward(const void* dy, int dy_ld, const void* y, int y_ld, void* dz, int dz_ld, int64_t npix, int c,
                 float slope, int dtype, cudaStream_t st);
int gn_backward(const void* dy, int dy_ld, const void* x, int x_ld, void* dx, int dx_ld, int n, int64_t hw, int c,
                int groups, int dtype, const float* gamma, const float* mean, const float* rstd, float* dgamma,
                float* dbeta, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st);
size_t loss_workspace_bytes(int64_t n);
int loss_fwd_bwd(int kind, const float* a, const float* b, float b_scalar, int64_t n, float* loss_out,
                 float* grad_out, void* ws, size_t ws_bytes, int mean, cudaStream_t st);
size_t ssim_workspace_bytes(int n, int c, int h, int w);
int ssim(const float* p, const float* t, int n, int c, int h, int w, float L, float* out, void* ws, size_t ws_bytes,
         cudaStream_t st);
int metrics_ae(const float* p, const float* t, int n, int c, int h, int w, float* out, cudaStream_t st);
int minmax(const float* a, int64_t n, float* out, cudaStream_t st);
size_t eval_metrics_workspace_bytes(int n, int c, int h, int w);
int eval_metrics(const float* p, const float* t, int n, int c, int h, int w, float* out, void* ws, size_t ws_bytes,
                 cudaStream_t st);
size_t ssim_backward_workspace_bytes(int n, int c, int h, int w);
int ssim_backward(const float* p, const float* t, int n, int c, int h, int w, const float* L_dev, const float* scale_dev,
                  float coef, float* dp, void* ws, size_t ws_bytes, cudaStream_t st);
int rgb2lab(const float* rgb, float* lab, int n, int h, int w, int normalised, cudaStream_t st);
int lab2rgb(const float* lab, float* rgb, int n, int h, int w, int normalised, cudaStream_t st);
int rgb2lab_u8(const uint8_t* rgb, float* lab, int n, int h, int w, cudaStream_t st);
int lab2rgb_u8(const float* lab, uint8_t* rgb, int n, int h, int w, cudaStream_t st);

}  // namespace srcgan

using namespace srcgan;

extern "C" {

const char* srcgan_version(void) { return "srcgan_b200 0.1 (sm_100a)"; }
const char* srcgan_last_error(void) { return g_err; }
int64_t srcgan_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* srcgan_last_kernel(void) { return g_last_kernel; }

size_t srcgan_packed_weight_bytes(int cout, int cin, int kh, int kw, int layout, int dtype) {
  if (layout >= SRCGAN_WL_TC) return packed_weight_bytes_tc(cout, cin, kh, kw, layout);
  return (size_t)cout * cin * kh * kw * (dtype == SRCGAN_DT_F32 ? 4 : 2);
}

int srcgan_pack_weights(const float* w, int cout, int cin, int kh, int kw, int layout, int dtype, void* out,
                        void* stream) {
  SRCGAN_REQUIRE(w && out && cout > 0 && cin > 0 && kh > 0 && kw > 0, "pack_weights: bad arguments");
  if (layout >= SRCGAN_WL_TC) {
    SRCGAN_REQUIRE(dtype == SRCGAN_DT_BF16, "pack_weights: the tcgen05 layouts are bf16 only");
    return pack_weights_tc_host(w, cout, cin, kh, kw, layout, out, (cudaStream_t)stream);
  }
  SRCGAN_REQUIRE(layout == SRCGAN_WL_RSCK || layout == SRCGAN_WL_RSKC, "pack_weights: unknown layout %d", layout);
  return pack_weights_simt_host(w, cout, cin, kh, kw, layout, dtype, out, (cudaStream_t)stream);
}

int srcgan_pack_slots(int cout, int kh, int kw, int layout, int32_t* slot_off16, int32_t* nslots, int32_t* bn) {
  SRCGAN_REQUIRE(slot_off16 && nslots && bn && cout > 0 && kh > 0 && kw > 0, "pack_slots: bad arguments");
  return pack_slots_tc(cout, kh, kw, layout, slot_off16, nslots, bn);
}
int srcgan_pack_weights_batch(const srcgan_pack_block* blocks_dev, int nblocks, int64_t total_elems, void* stream) {
  return pack_weights_batch(blocks_dev, nblocks, (long long)total_elems, (cudaStream_t)stream);
}

int srcgan_conv_fprop(const srcgan_conv_params* p, void* stream) {
  int rc = validate_conv(p, true);
  if (rc) return rc;
  int engine = p->engine;
  if (engine == SRCGAN_ENGINE_AUTO) engine = conv_tc_supported(p) ? SRCGAN_ENGINE_TC : SRCGAN_ENGINE_SIMT;
  if (engine == SRCGAN_ENGINE_TC) {
    SRCGAN_REQUIRE(conv_tc_supported(p), "conv_fprop: shape not supported by the tcgen05 engine");
    return conv_fprop_tc(p, (cudaStream_t)stream);
  }
  if (p->signbits || p->maskbits || p->zero_row_period || p->x_group_stride) {
    set_error("conv_fprop: packed sign / mask bits, zero_row_period and planar inputs are only implemented by the paired-sweep "
              "tcgen05 kernel");
    return SRCGAN_E_INVALID;
  }
  return conv_fprop_simt(p, (cudaStream_t)stream);
}

int srcgan_conv_fprop_pair_supported(const srcgan_conv_params* pa, const srcgan_conv_params* pb) {
  if (!pa || !pb || validate_conv(pa, true) || validate_conv(pb, true)) return 0;
  return conv_fprop_pair_supported(pa, pb) ? 1 : 0;
}
int srcgan_conv_fprop_pair(const srcgan_conv_params* pa, const srcgan_conv_params* pb, void* stream) {
  SRCGAN_REQUIRE(pa && pb, "conv_fprop_pair: null parameters");
  int rc = validate_conv(pa, true);
  if (rc) return rc;
  rc = validate_conv(pb, true);
  if (rc) return rc;
  return conv_fprop_pair_tc(pa, pb, (cudaStream_t)stream);
}

int srcgan_conv_dgrad(const srcgan_conv_params* p, void* stream) {
  int rc = validate_conv(p, true);
  if (rc) return rc;
  if (p->signbits || p->maskbits || p->zero_row_period || p->x_group_stride) {
    set_error("conv_dgrad: packed sign / mask bits, zero_row_period and planar inputs are only implemented by the paired-sweep "
              "fprop kernel");
    return SRCGAN_E_INVALID;
  }
  if (p->engine == SRCGAN_ENGINE_TC) {
    SRCGAN_REQUIRE(conv_dgrad_tc_supported(p),
                   "conv_dgrad: tcgen05 engine handles stride-2 dgrad only (stride 1 = fprop over transposed weights)");
    return conv_dgrad_tc(p, (cudaStream_t)stream);
  }
  return conv_dgrad_simt(p, (cudaStream_t)stream);
}

size_t srcgan_conv_wgrad_workspace_bytes(const srcgan_conv_params* p) {
  if (!p) return 0;
  size_t a = conv_wgrad_simt_workspace(p);
  size_t b = conv_wgrad_tc_supported(p) ? conv_wgrad_tc_workspace(p) : 0;
  return a > b ? a : b;
}

int srcgan_conv_wgrad(const srcgan_conv_params* p, float* dw, float* db, int accumulate, void* workspace,
                      size_t workspace_bytes, void* stream) {
  int rc = validate_conv(p, false);
  if (rc) return rc;
  SRCGAN_REQUIRE(dw || db, "conv_wgrad: nothing to compute");
  SRCGAN_REQUIRE(p->x_group_stride == 0, "conv_wgrad: planar inputs are passed one 64-channel group at a time");
  int engine = p->engine;
  if (engine == SRCGAN_ENGINE_AUTO) engine = conv_wgrad_tc_supported(p) ? SRCGAN_ENGINE_TC : SRCGAN_ENGINE_SIMT;
  if (engine == SRCGAN_ENGINE_TC) {
    SRCGAN_REQUIRE(conv_wgrad_tc_supported(p), "conv_wgrad: shape not supported by the tcgen05 engine");
    return conv_wgrad_tc(p, dw, db, accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  return conv_wgrad_simt(p, dw, db, accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}

int srcgan_conv_wgrad_split(const srcgan_conv_params* p, float* dw0, int cin_ld0, int ci0_0, float* db0, float* dw1, int cin_ld1,
                            int ci0_1, float* db1, int split_cout, int accumulate, void* workspace, size_t workspace_bytes,
                            void* stream) {
  int rc = validate_conv(p, false);
  if (rc) return rc;
  SRCGAN_REQUIRE(p->engine == SRCGAN_ENGINE_TC && conv_wgrad_tc_supported(p),
                 "conv_wgrad_split: tcgen05 engine only (bf16, 3x3 stride 1)");
  SRCGAN_REQUIRE(p->x_group_stride == 0, "conv_wgrad_split: planar inputs are passed one 64-channel group at a time");
  return conv_wgrad_tc_split(p, dw0, cin_ld0, ci0_0, db0, dw1, cin_ld1, ci0_1, db1, split_cout, accumulate, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}

int srcgan_nchw_to_nhwc(const float* src, int n, int c, int h, int w, void* dst, int dst_ld, int dtype, void* stream) {
  SRCGAN_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && dst_ld >= c, "nchw_to_nhwc: bad arguments");
  return nchw_to_nhwc(src, n, c, h, w, dst, dst_ld, dtype, (cudaStream_t)stream);
}
int srcgan_nhwc_to_nchw(const void* src, int src_ld, int dtype, float* dst, int n, int c, int h, int w, void* stream) {
  SRCGAN_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && src_ld >= c, "nhwc_to_nchw: bad arguments");
  return nhwc_to_nchw(src, src_ld, dtype, dst, n, c, h, w, (cudaStream_t)stream);
}
int srcgan_upsample2x(const void* src, int src_ld, void* dst, int dst_ld, int n, int h, int w, int c, int dtype,
                      void* stream) {
  SRCGAN_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, "upsample2x: bad arguments");
  return upsample2x(src, src_ld, dst, dst_ld, n, h, w, c, dtype, (cudaStream_t)stream);
}
int srcgan_depth_to_space(const void* src, int src_ld, void* dst, int dst_ld, int n, int h, int w, int c, int dtype,
                          void* stream) {
  SRCGAN_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, "depth_to_space: bad arguments");
  return depth_to_space(src, src_ld, dst, dst_ld, nullptr, 0, 0.f, n, h, w, c, dtype, 0, (cudaStream_t)stream);
}
int srcgan_space_to_depth(const void* src, int src_ld, void* dst, int dst_ld, const void* mask, int mask_ld,
                          float mask_slope, int n, int h, int w, int c, int dtype, void* stream) {
  SRCGAN_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, "space_to_depth: bad arguments");
  return depth_to_space(src, src_ld, dst, dst_ld, mask, mask_ld, mask_slope, n, h, w, c, dtype, 1, (cudaStream_t)stream);
}
int srcgan_pixel_shuffle(const void* src, int src_ld, void* dst, int dst_ld, int n, int h, int w, int c, int r,
                         int dtype, int adjoint, void* stream) {
  SRCGAN_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && r > 0, "pixel_shuffle: bad arguments");
  return pixel_shuffle(src, src_ld, dst, dst_ld, n, h, w, c, r, dtype, adjoint, (cudaStream_t)stream);
}
int srcgan_upsample2x_adjoint(const void* src, int src_ld, void* dst, int dst_ld, const void* mask, int mask_ld,
                              float mask_slope, int n, int h, int w, int c, int dtype, void* stream) {
  SRCGAN_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, "upsample2x_adjoint: bad arguments");
  return upsample2x_adjoint(src, src_ld, dst, dst_ld, mask, mask_ld, mask_slope, n, h, w, c, dtype,
                            (cudaStream_t)stream);
}

int srcgan_add(const void* a, int a_ld, const void* b, int b_ld, void* dst, int dst_ld, int64_t npix, int c, int dtype,
               void* stream) {
  SRCGAN_REQUIRE(a && b && dst && npix > 0 && c > 0, "add: bad arguments");
  return add_slices(a, a_ld, b, b_ld, dst, dst_ld, npix, c, dtype, (cudaStream_t)stream);
}

int srcgan_bias_act(void* y, int y_ld, int64_t npix, int c, const float* bias, int has_act, float slope, int dtype, void* stream) {
  SRCGAN_REQUIRE(y && npix > 0 && c > 0 && (bias || has_act), "bias_act: bad arguments");
  return bias_act(y, y_ld, npix, c, bias, has_act, slope, dtype, (cudaStream_t)stream);
}

int srcgan_act_backward(const void* dy, int dy_ld, const void* y, int y_ld, void* dz, int dz_ld, int64_t npix, int c,
                        float slope, int dtype, void* stream) {
  SRCGAN_REQUIRE(dy && y && dz && npix > 0 && c > 0, "act_backward: bad arguments");
  return act_backward(dy, dy_ld, y, y_ld, dz, dz_ld, npix, c, slope, dtype, (cudaStream_t)stream);
}

size_t srcgan_colsum_workspace_bytes(int64_t, int c) { return (size_t)1024 * c * sizeof(float) + 256; }
int srcgan_colsum(const void* x, int x_ld, int dtype, int64_t npix, int c, float* out, float alpha, int accumulate,
                  void* workspace, size_t workspace_bytes, void* stream) {
  SRCGAN_REQUIRE(x && out && npix > 0 && c > 0, "colsum: bad arguments");
  SRCGAN_REQUIRE(workspace && workspace_bytes >= srcgan_colsum_workspace_bytes(npix, c), "colsum: workspace too small");
  return bias_grad_launch(x, x_ld, dtype, npix, c, out, accumulate, alpha, workspace, (cudaStream_t)stream);
}

size_t srcgan_bn_workspace_bytes(int64_t npix, int c) { return bn_workspace_bytes(npix, c); }
int srcgan_bn_forward(const void* x, int x_ld, void* y, int y_ld, int64_t npix, int c, int dtype, const float* gamma,
                      const float* beta, float* running_mean, float* running_var, float* save_mean,
                      float* save_invstd, int training, float momentum, float eps, float slope, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return bn_forward(x, x_ld, y, y_ld, npix, c, dtype, gamma, beta, running_mean, running_var, save_mean, save_invstd,
                    training, momentum, eps, slope, workspace, workspace_bytes, (cudaStream_t)stream);
}
int srcgan_bn_backward(const void* dy_post, int dy_ld, const void* y, int y_ld, const void* x, int x_ld, void* dx,
                       int dx_ld, int64_t npix, int c, int dtype, const float* gamma, const float* beta,
                       const float* save_mean, const float* save_invstd, float slope, int training, float* dgamma,
                       float* dbeta, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  return bn_backward(dy_post, dy_ld, y, y_ld, x, x_ld, dx, dx_ld, npix, c, dtype, gamma, save_mean, save_invstd,
                     slope, training, dgamma, dbeta, accumulate, workspace, workspace_bytes, (cudaStream_t)stream, beta);
}

size_t srcgan_gn_workspace_bytes(int n, int c) { return gn_workspace_bytes(n, c); }
int srcgan_gn_forward(const void* x, int x_ld, void* y, int y_ld, int n, int64_t hw, int c, int groups, int dtype,
                      const float* gamma, const float* beta, float* save_mean, float* save_rstd, float eps,
                      const void* residual, int res_ld, int act, float slope, void* workspace, size_t workspace_bytes,
                      void* stream) {
  return gn_forward(x, x_ld, y, y_ld, n, hw, c, groups, dtype, gamma, beta, save_mean, save_rstd, eps, residual, res_ld,
                    act, slope, workspace, workspace_bytes, (cudaStream_t)stream);
}
int srcgan_gn_backward(const void* dy, int dy_ld, const void* x, int x_ld, void* dx, int dx_ld, int n, int64_t hw, int c,
                       int groups, int dtype, const float* gamma, const float* save_mean, const float* save_rstd,
                       float* dgamma, float* dbeta, int accumulate, void* workspace, size_t workspace_bytes,
                       void* stream) {
  return gn_backward(dy, dy_ld, x, x_ld, dx, dx_ld, n, hw, c, groups, dtype, gamma, save_mean, save_rstd, dgamma, dbeta,
                     accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t srcgan_loss_workspace_bytes(int64_t n) { return loss_workspace_bytes(n); }
int srcgan_loss_fwd_bwd(int kind, const float* a, const float* b, float b_scalar, int64_t n, float* loss_out,
                        float* grad_out, void* workspace, size_t workspace_bytes, void* stream) {
  return loss_fwd_bwd(kind, a, b, b_scalar, n, loss_out, grad_out, workspace, workspace_bytes, 1,
                      (cudaStream_t)stream);
}
int srcgan_metrics_sqerr(const float* a, const float* b, int64_t n, float* out_sum, void* workspace,
                         size_t workspace_bytes, void* stream) {
  SRCGAN_REQUIRE(b != nullptr, "metrics_sqerr: null pointer");
  return loss_fwd_bwd(1, a, b, 0.f, n, out_sum, nullptr, workspace, workspace_bytes, 0, (cudaStream_t)stream);
}
int srcgan_metrics_ae(const float* pred, const float* truth, int n, int c, int h, int w, float* out_per_image,
                      void* stream) {
  return metrics_ae(pred, truth, n, c, h, w, out_per_image, (cudaStream_t)stream);
}
size_t srcgan_ssim_workspace_bytes(int n, int c, int h, int w) { return ssim_workspace_bytes(n, c, h, w); }
int srcgan_ssim(const float* pred, const float* truth, int n, int c, int h, int w, float L, float* out_per_image,
                void* workspace, size_t workspace_bytes, void* stream) {
  return ssim(pred, truth, n, c, h, w, L, out_per_image, workspace, workspace_bytes, (cudaStream_t)stream);
}
int srcgan_minmax(const float* a, int64_t n, float* out_min_max, void* stream) {
  return minmax(a, n, out_min_max, (cudaStream_t)stream);
}
size_t srcgan_eval_metrics_workspace_bytes(int n, int c, int h, int w) { return eval_metrics_workspace_bytes(n, c, h, w); }
int srcgan_eval_metrics(const float* pred, const float* truth, int n, int c, int h, int w, float* out, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return eval_metrics(pred, truth, n, c, h, w, out, workspace, workspace_bytes, (cudaStream_t)stream);
}
size_t srcgan_ssim_backward_workspace_bytes(int n, int c, int h, int w) { return ssim_backward_workspace_bytes(n, c, h, w); }
int srcgan_ssim_backward(const float* pred, const float* truth, int n, int c, int h, int w, const float* L_dev,
                         const float* scale_dev, float coef, float* dpred, void* workspace, size_t workspace_bytes,
                         void* stream) {
  return ssim_backward(pred, truth, n, c, h, w, L_dev, scale_dev, coef, dpred, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}
int srcgan_rgb2lab(const float* rgb, float* lab, int n, int h, int w, int normalised, void* stream) {
  return rgb2lab(rgb, lab, n, h, w, normalised, (cudaStream_t)stream);
}
int srcgan_lab2rgb(const float* lab, float* rgb, int n, int h, int w, int normalised, void* stream) {
  return lab2rgb(lab, rgb, n, h, w, normalised, (cudaStream_t)stream);
}
int srcgan_rgb2lab_u8(const uint8_t* rgb_nhwc, float* lab_nchw, int n, int h, int w, void* stream) {
  return rgb2lab_u8(rgb_nhwc, lab_nchw, n, h, w, (cudaStream_t)stream);
}
int srcgan_lab2rgb_u8(const float* lab_nchw, uint8_t* rgb_nhwc, int n, int h, int w, void* stream) {
  return lab2rgb_u8(lab_nchw, rgb_nhwc, n, h, w, (cudaStream_t)stream);
}

}  // extern "C"
