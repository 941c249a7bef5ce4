// GroupNorm over NHWC rows (per sample, per group of C/G channels), forward and backward.
// Reference: nn.GroupNorm(32, C) in src/model/resdeconv.py:67,75,120,156 and src/model/edsr.py:45
// (eps 1e-5, affine).  HBM-bound: statistics pass (read), apply pass (read + write).
#include "common.cuh"

namespace srcgan {

constexpr int kGnParts = 64;

size_t gn_workspace_bytes(int n, int c) { return ((size_t)2 * n * kGnParts * c + (size_t)2 * n * c) * sizeof(float) + 256; }

// per-(sample, channel) partial sums.  grid (ceil(c/32), parts, n), block (32, 8)
//   MODE 0: sum x, sum x^2          MODE 1: sum dy, sum dy * xhat
template <typename T, int MODE>
__global__ void gn_partial(const T* __restrict__ x, int x_ld, const T* __restrict__ dy, int dy_ld, int64_t hw, int c,
                           int cpg, const float* __restrict__ mean, const float* __restrict__ rstd, int groups,
                           float* __restrict__ p0, float* __restrict__ p1) {
  __shared__ float r0[8][33], r1[8][33];
  const int ch = blockIdx.x * 32 + threadIdx.x;
  const int n = blockIdx.z;
  float s = 0.f, q = 0.f;
  if (ch < c) {
    float mu = 0.f, rs = 1.f;
    if (MODE == 1) { mu = mean[n * groups + ch / cpg]; rs = rstd[n * groups + ch / cpg]; }
    for (int64_t m = (int64_t)blockIdx.y * 8 + threadIdx.y; m < hw; m += (int64_t)gridDim.y * 8) {
      const int64_t row = (int64_t)n * hw + m;
      const float v = to_f32(x[row * x_ld + ch]);
      if (MODE == 0) { s += v; q = fmaf(v, v, q); }
      else { const float d = to_f32(dy[row * dy_ld + ch]); s += d; q = fmaf(d, (v - mu) * rs, q); }
    }
  }
  r0[threadIdx.y][threadIdx.x] = s; r1[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += r0[k][threadIdx.x]; b += r1[k][threadIdx.x]; }
    const int64_t o = ((int64_t)n * gridDim.y + blockIdx.y) * c + ch;
    p0[o] = a; p1[o] = b;
  }
}

// one thread per (sample, group)
__global__ void gn_finalize(const float* __restrict__ p0, const float* __restrict__ p1, int parts, int n, int c, int cpg,
                            int64_t hw, float eps, float* __restrict__ mean, float* __restrict__ rstd) {
  const int groups = c / cpg;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * groups) return;
  const int s_ = i / groups, g = i - s_ * groups;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < parts; ++k)
    for (int j = 0; j < cpg; ++j) {
      const int64_t o = ((int64_t)s_ * parts + k) * c + g * cpg + j;
      s += (double)p0[o]; q += (double)p1[o];
    }
  const double m = (double)hw * cpg;
  const double mu = s / m;
  double var = q / m - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[i] = (float)mu;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

template <typename T>
__global__ void gn_apply(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld, int64_t hw, int64_t total, int c,
                         int cpg, const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ mean, const float* __restrict__ rstd, const T* __restrict__ res,
                         int res_ld, int act, float slope) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = (int)(i % c);
  const int64_t row = i / c;
  const int n = (int)(row / hw);
  const int sg = n * (c / cpg) + ch / cpg;
  float v = (to_f32(x[row * x_ld + ch]) - mean[sg]) * rstd[sg] * gamma[ch] + beta[ch];
  if (res) v += to_f32(res[row * res_ld + ch]);
  if (act && v < 0.f) v *= slope;
  y[row * y_ld + ch] = from_f32<T>(v);
}

// per (sample, channel) totals S1 = sum dy, S2 = sum dy*xhat -> sc[n][c][2]; dgamma/dbeta over samples
__global__ void gn_bwd_finalize(const float* __restrict__ p0, const float* __restrict__ p1, int parts, int n, int c,
                                float* __restrict__ sc, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                int accumulate) {
  int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double g = 0.0, b = 0.0;
  for (int s_ = 0; s_ < n; ++s_) {
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < parts; ++k) {
      const int64_t o = ((int64_t)s_ * parts + k) * c + ch;
      s1 += (double)p0[o]; s2 += (double)p1[o];
    }
    sc[((int64_t)s_ * c + ch) * 2] = (float)s1;
    sc[((int64_t)s_ * c + ch) * 2 + 1] = (float)s2;
    b += s1; g += s2;
  }
  if (dgamma) dgamma[ch] = accumulate ? dgamma[ch] + (float)g : (float)g;
  if (dbeta) dbeta[ch] = accumulate ? dbeta[ch] + (float)b : (float)b;
}

template <typename T>
__global__ void gn_bwd_apply(const T* __restrict__ dy, int dy_ld, const T* __restrict__ x, int x_ld, T* __restrict__ dx,
                             int dx_ld, int64_t hw, int64_t total, int c, int cpg, const float* __restrict__ gamma,
                             const float* __restrict__ mean, const float* __restrict__ rstd,
                             const float* __restrict__ sc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = (int)(i % c);
  const int64_t row = i / c;
  const int n = (int)(row / hw);
  const int g0 = (ch / cpg) * cpg;
  const int sg = n * (c / cpg) + ch / cpg;
  float A = 0.f, B = 0.f;
  for (int j = 0; j < cpg; ++j) {
    const float gm = gamma[g0 + j];
    A = fmaf(gm, sc[((int64_t)n * c + g0 + j) * 2], A);
    B = fmaf(gm, sc[((int64_t)n * c + g0 + j) * 2 + 1], B);
  }
  const float inv_m = 1.f / ((float)hw * cpg);
  const float rs = rstd[sg];
  const float xh = (to_f32(x[row * x_ld + ch]) - mean[sg]) * rs;
  const float v = rs * (to_f32(dy[row * dy_ld + ch]) * gamma[ch] - A * inv_m - xh * B * inv_m);
  dx[row * dx_ld + ch] = from_f32<T>(v);
}

template <typename T>
static int gn_forward_t(const T* x, int x_ld, T* y, int y_ld, int n, int64_t hw, int c, int groups, const float* gamma,
                        const float* beta, float* mean, float* rstd, float eps, const T* res, int res_ld, int act,
                        float slope, float* ws, cudaStream_t st) {
  const int cpg = c / groups;
  int parts = (int)((hw + 255) / 256);
  if (parts > kGnParts) parts = kGnParts;
  if (parts < 1) parts = 1;
  float* p0 = ws;
  float* p1 = ws + (size_t)n * kGnParts * c;
  dim3 grid(ceil_div(c, 32), parts, n), blk(32, 8);
  gn_partial<T, 0><<<grid, blk, 0, st>>>(x, x_ld, nullptr, 0, hw, c, cpg, nullptr, nullptr, groups, p0, p1);
  gn_finalize<<<ceil_div((int64_t)n * groups, 128), 128, 0, st>>>(p0, p1, parts, n, c, cpg, hw, eps, mean, rstd);
  const int64_t total = (int64_t)n * hw * c;
  gn_apply<T><<<ceil_div(total, 256), 256, 0, st>>>(x, x_ld, y, y_ld, hw, total, c, cpg, gamma, beta, mean, rstd, res,
                                                    res_ld, act, slope);
  count_launch(3);
  return check_launch("gn_forward");
}

int gn_forward(const void* x, int x_ld, void* y, int y_ld, int n, int64_t hw, int c, int groups, int dtype,
               const float* gamma, const float* beta, float* mean, float* rstd, float eps, const void* res, int res_ld,
               int act, float slope, void* ws, size_t ws_bytes, cudaStream_t st) {
  SRCGAN_REQUIRE(x && y && gamma && beta && mean && rstd && groups > 0 && c % groups == 0, "gn_forward: bad arguments");
  SRCGAN_REQUIRE(ws && ws_bytes >= gn_workspace_bytes(n, c), "gn_forward: workspace too small");
  if (dtype == SRCGAN_DT_F32)
    return gn_forward_t<float>((const float*)x, x_ld, (float*)y, y_ld, n, hw, c, groups, gamma, beta, mean, rstd, eps,
                               (const float*)res, res_ld, act, slope, (float*)ws, st);
  return gn_forward_t<__nv_bfloat16>((const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)y, y_ld, n, hw, c, groups, gamma,
                                     beta, mean, rstd, eps, (const __nv_bfloat16*)res, res_ld, act, slope, (float*)ws,
                                     st);
}

template <typename T>
static int gn_backward_t(const T* dy, int dy_ld, const T* x, int x_ld, T* dx, int dx_ld, int n, int64_t hw, int c,
                         int groups, const float* gamma, const float* mean, const float* rstd, float* dgamma,
                         float* dbeta, int accumulate, float* ws, cudaStream_t st) {
  const int cpg = c / groups;
  int parts = (int)((hw + 255) / 256);
  if (parts > kGnParts) parts = kGnParts;
  if (parts < 1) parts = 1;
  float* p0 = ws;
  float* p1 = ws + (size_t)n * kGnParts * c;
  float* sc = ws + (size_t)2 * n * kGnParts * c;
  dim3 grid(ceil_div(c, 32), parts, n), blk(32, 8);
  gn_partial<T, 1><<<grid, blk, 0, st>>>(x, x_ld, dy, dy_ld, hw, c, cpg, mean, rstd, groups, p0, p1);
  gn_bwd_finalize<<<ceil_div(c, 128), 128, 0, st>>>(p0, p1, parts, n, c, sc, dgamma, dbeta, accumulate);
  const int64_t total = (int64_t)n * hw * c;
  gn_bwd_apply<T><<<ceil_div(total, 256), 256, 0, st>>>(dy, dy_ld, x, x_ld, dx, dx_ld, hw, total, c, cpg, gamma, mean,
                                                        rstd, sc);
  count_launch(3);
  return check_launch("gn_backward");
}

int gn_backward(const void* dy, int dy_ld, const void* x, int x_ld, void* dx, int dx_ld, int n, int64_t hw, int c,
                int groups, int dtype, const float* gamma, const float* mean, const float* rstd, float* dgamma,
                float* dbeta, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  SRCGAN_REQUIRE(dy && x && dx && gamma && mean && rstd && groups > 0 && c % groups == 0, "gn_backward: bad arguments");
  SRCGAN_REQUIRE(ws && ws_bytes >= gn_workspace_bytes(n, c), "gn_backward: workspace too small");
  if (dtype == SRCGAN_DT_F32)
    return gn_backward_t<float>((const float*)dy, dy_ld, (const float*)x, x_ld, (float*)dx, dx_ld, n, hw, c, groups,
                                gamma, mean, rstd, dgamma, dbeta, accumulate, (float*)ws, st);
  return gn_backward_t<__nv_bfloat16>((const __nv_bfloat16*)dy, dy_ld, (const __nv_bfloat16*)x, x_ld,
                                      (__nv_bfloat16*)dx, dx_ld, n, hw, c, groups, gamma, mean, rstd, dgamma, dbeta,
                                      accumulate, (float*)ws, st);
}

}  // namespace srcgan
