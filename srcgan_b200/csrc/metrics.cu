// Evaluation metrics on NCHW fp32 (src/metrics.py:10-144; losses.SSIM src/losses.py:20-93):
//   SSIM : separable 11-tap gaussian (sigma 1.5, 'valid') producing the five moments in shared
//          memory in one pass, fused ssim-map evaluation and per-image deterministic reduction
//   AE   : mean angular error in degrees per image
//   minmax : the data-range heuristic inputs of SSIM (max>128 -> 255, min<-0.5 -> -1)
#include <math.h>

#include "common.cuh"

namespace srcgan {

constexpr int SS_T = 32;            // output tile
constexpr int SS_W = 11;            // window
constexpr int SS_IN = SS_T + SS_W - 1;

__constant__ float c_gauss[SS_W];

__device__ __forceinline__ float block_sum_any(float v, float* red) {
  v = warp_sum(v);
  int nw = (blockDim.x * blockDim.y + 31) >> 5;
  int t = threadIdx.y * blockDim.x + threadIdx.x;
  if ((t & 31) == 0) red[t >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (t == 0)
    for (int k = 0; k < nw; ++k) s += red[k];
  return s;
}

// grid (tiles_x, tiles_y, n*c), block (32, 8)
__global__ void __launch_bounds__(256)
ssim_tile(const float* __restrict__ p, const float* __restrict__ t, int h, int w, int ho, int wo, float C1,
          float C2, float* __restrict__ part) {
  __shared__ float sp[SS_IN][SS_IN + 1], st[SS_IN][SS_IN + 1];
  __shared__ float hb[5][SS_IN][SS_T + 1];
  __shared__ float red[8];
  const int64_t plane = blockIdx.z;
  const float* P = p + plane * h * w;
  const float* T = t + plane * h * w;
  const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int e = tid; e < SS_IN * SS_IN; e += 256) {
    int yy = e / SS_IN, xx = e - yy * SS_IN;
    int gy = y0 + yy, gx = x0 + xx;
    bool ok = gy < h && gx < w;
    sp[yy][xx] = ok ? __ldg(P + (int64_t)gy * w + gx) : 0.f;
    st[yy][xx] = ok ? __ldg(T + (int64_t)gy * w + gx) : 0.f;
  }
  __syncthreads();
  // horizontal blur of the five products
  for (int e = tid; e < SS_IN * SS_T; e += 256) {
    int yy = e / SS_T, xx = e - yy * SS_T;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int k = 0; k < SS_W; ++k) {
      float g = c_gauss[k], u = sp[yy][xx + k], v = st[yy][xx + k];
      a0 = fmaf(g, u, a0); a1 = fmaf(g, v, a1);
      a2 = fmaf(g, u * u, a2); a3 = fmaf(g, v * v, a3); a4 = fmaf(g, u * v, a4);
    }
    hb[0][yy][xx] = a0; hb[1][yy][xx] = a1; hb[2][yy][xx] = a2; hb[3][yy][xx] = a3; hb[4][yy][xx] = a4;
  }
  __syncthreads();
  float acc = 0.f;
  for (int yy = threadIdx.y; yy < SS_T; yy += 8) {
    int xx = threadIdx.x;
    if (y0 + yy < ho && x0 + xx < wo) {
      float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
      for (int k = 0; k < SS_W; ++k) {
        float g = c_gauss[k];
        m1 = fmaf(g, hb[0][yy + k][xx], m1); m2 = fmaf(g, hb[1][yy + k][xx], m2);
        e11 = fmaf(g, hb[2][yy + k][xx], e11); e22 = fmaf(g, hb[3][yy + k][xx], e22);
        e12 = fmaf(g, hb[4][yy + k][xx], e12);
      }
      float s11 = e11 - m1 * m1, s22 = e22 - m2 * m2, s12 = e12 - m1 * m2;
      float v1 = 2.f * s12 + C2, v2 = s11 + s22 + C2;
      acc += ((2.f * m1 * m2 + C1) * v1) / ((m1 * m1 + m2 * m2 + C1) * v2);
    }
  }
  float s = block_sum_any(acc, red);
  if (tid == 0) part[((int64_t)plane * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

// out[img] = sum of that image's tile partials (fixed order)
__global__ void ssim_final(const float* __restrict__ part, int per_image, float* __restrict__ out) {
  __shared__ double red[8];
  const float* p = part + (int64_t)blockIdx.x * per_image;
  double s = 0.0;
  for (int i = threadIdx.x; i < per_image; i += blockDim.x) s += (double)p[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    out[blockIdx.x] = (float)t;
  }
}

size_t ssim_workspace_bytes(int n, int c, int h, int w) {
  int ho = h - SS_W + 1, wo = w - SS_W + 1;
  if (ho < 1 || wo < 1) return 256;
  return (size_t)n * c * ceil_div(ho, SS_T) * ceil_div(wo, SS_T) * sizeof(float) + 256;
}

int ssim(const float* p, const float* t, int n, int c, int h, int w, float L, float* out, void* ws, size_t ws_bytes,
         cudaStream_t st) {
  SRCGAN_REQUIRE(p && t && out, "ssim: null pointer");
  SRCGAN_REQUIRE(h >= SS_W && w >= SS_W, "ssim: image smaller than the 11x11 window");
  SRCGAN_REQUIRE(ws && ws_bytes >= ssim_workspace_bytes(n, c, h, w), "ssim: workspace too small");
  static DeviceOnce init;     // __constant__ memory is per device
  int init_dev;
  if (init.needed(&init_dev)) {
    float g[SS_W];
    double s = 0.0;
    for (int i = 0; i < SS_W; ++i) { g[i] = (float)exp(-((i - SS_W / 2) * (i - SS_W / 2)) / (2.0 * 1.5 * 1.5)); s += g[i]; }
    // the reference normalises in fp32 (torch.Tensor / sum)
    float fs = 0.f;
    for (int i = 0; i < SS_W; ++i) fs += g[i];
    for (int i = 0; i < SS_W; ++i) g[i] = g[i] / fs;
    (void)s;
    SRCGAN_CUDA(cudaMemcpyToSymbolAsync(c_gauss, g, sizeof(g), 0, cudaMemcpyHostToDevice, st));
    init.mark(init_dev);
  }
  const int ho = h - SS_W + 1, wo = w - SS_W + 1;
  dim3 grid(ceil_div(wo, SS_T), ceil_div(ho, SS_T), n * c), blk(32, 8);
  const float C1 = (0.01f * L) * (0.01f * L), C2 = (0.03f * L) * (0.03f * L);
  float* part = reinterpret_cast<float*>(ws);
  ssim_tile<<<grid, blk, 0, st>>>(p, t, h, w, ho, wo, C1, C2, part);
  ssim_final<<<n, 256, 0, st>>>(part, c * grid.x * grid.y, out);
  count_launch(2);
  return check_launch("ssim");
}

// one block per image; deterministic
__global__ void __launch_bounds__(1024)
ae_kernel(const float* __restrict__ p, const float* __restrict__ t, int c, int64_t hw, float* __restrict__ out) {
  __shared__ float red[32];
  const float* P = p + (int64_t)blockIdx.x * c * hw;
  const float* T = t + (int64_t)blockIdx.x * c * hw;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < hw; i += blockDim.x) {
    float dot = 0.f, np = 0.f, nt = 0.f;
    for (int k = 0; k < c; ++k) {
      float a = __ldg(P + k * hw + i), b = __ldg(T + k * hw + i);
      dot = fmaf(a, b, dot); np = fmaf(a, a, np); nt = fmaf(b, b, nt);
    }
    acc += acosf(dot / (sqrtf(np) * sqrtf(nt) + 1e-6f)) * 57.29577951308232f;
  }
  float s = block_sum_any(acc, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s / (float)hw;
}

int metrics_ae(const float* p, const float* t, int n, int c, int h, int w, float* out, cudaStream_t st) {
  SRCGAN_REQUIRE(p && t && out && n > 0, "ae: null pointer");
  ae_kernel<<<n, 1024, 0, st>>>(p, t, c, (int64_t)h * w, out);
  count_launch();
  return check_launch("metrics_ae");
}

__global__ void minmax_init(float* out) { out[0] = INFINITY; out[1] = -INFINITY; }
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__global__ void minmax_kernel(const float* __restrict__ a, int64_t n, float* __restrict__ out) {
  float lo = INFINITY, hi = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = __ldg(a + i);
    lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomic_min_f(out, lo); atomic_max_f(out + 1, hi); }
}

int minmax(const float* a, int64_t n, float* out, cudaStream_t st) {
  SRCGAN_REQUIRE(a && out && n > 0, "minmax: null pointer / empty");
  minmax_init<<<1, 1, 0, st>>>(out);
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  minmax_kernel<<<blocks, 256, 0, st>>>(a, n, out);
  count_launch(2);
  return check_launch("minmax");
}

}  // namespace srcgan
