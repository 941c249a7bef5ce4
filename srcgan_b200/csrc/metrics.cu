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

static int upload_gauss(cudaStream_t st);

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
  { int rc_g = upload_gauss(st); if (rc_g) return rc_g; }
  const int ho = h - SS_W + 1, wo = w - SS_W + 1;
  dim3 grid(ceil_div(wo, SS_T), ceil_div(ho, SS_T), n * c), blk(32, 8);
  const float C1 = (0.01f * L) * (0.01f * L), C2 = (0.03f * L) * (0.03f * L);
  float* part = reinterpret_cast<float*>(ws);
  ssim_tile<<<grid, blk, 0, st>>>(p, t, h, w, ho, wo, C1, C2, part);
  ssim_final<<<n, 256, 0, st>>>(part, c * grid.x * grid.y, out);
  count_launch(2);
  return check_launch("ssim");
}

// ---------------------------------------------------------------------------------------------
// Fused evaluation sweep (src/testCas.py:63-85, src/metrics.py:10-144): MSE, PSNR, AE and SSIM of one prediction / truth
// pair in ONE pass over the two images and ONE kernel launch; the caller reads the scalars back with one D2H copy.
//   * SSIM's data range L depends on the global min / max of y_pred (metrics.py:102-111: max > 128 -> 255 else 1, min < -0.5
//     -> +1), which is only known after the pass; L can only be 1, 2, 255 or 256, so every block evaluates the SSIM map for
//     all four candidates (the 11x11 moments are shared) and the finalising block picks the one the reduced min / max select.
//   * a block owns one 32x32 tile of one image and loops over the channels: the squared error, the per-pixel dot products /
//     norms of the angular error and min / max are taken from the same shared-memory tiles the SSIM blur reads.  A pixel is
//     counted once, by the tile that owns it (the last tile row / column also owns the 10-pixel 'valid' margin).
//   * partials go to a workspace; the last block to finish (atomic ticket) reduces them in a fixed order -> deterministic.
// out: [0] MSE  [1] PSNR  [2] AE (mean over images)  [3] SSIM (mean)  [4] L  [5] min  [6] max  [7] unused
//      [8 .. 8+n)    per-image SSIM means  (size_average=False form)      [8+n .. 8+2n)  per-image AE (degrees)
//      [8+2n .. 8+3n) per-image MSE    [8+3n .. 8+4n) per-image SSIM with the image's OWN data range    [8+4n .. 8+5n) that range
// ---------------------------------------------------------------------------------------------
constexpr int EV_PART = 8;          // per block: ssim(L=1), ssim(2), ssim(255), ssim(256), sqerr, ae_sum, min, max

__device__ __forceinline__ float ssim_point(float m1, float m2, float e11, float e22, float e12, float L) {
  const float C1 = (0.01f * L) * (0.01f * L), C2 = (0.03f * L) * (0.03f * L);
  const float s11 = e11 - m1 * m1, s22 = e22 - m2 * m2, s12 = e12 - m1 * m2;
  const float v1 = 2.f * s12 + C2, v2 = s11 + s22 + C2;
  return ((2.f * m1 * m2 + C1) * v1) / ((m1 * m1 + m2 * m2 + C1) * v2);
}

// grid (tiles_x, tiles_y, n), block (32, 8)
__global__ void __launch_bounds__(256)
eval_metrics_k(const float* __restrict__ p, const float* __restrict__ t, int c, int h, int w, int ho, int wo,
               float* __restrict__ part, unsigned int* __restrict__ ticket, float* __restrict__ out, int n) {
  __shared__ float sp[SS_IN][SS_IN + 1], st[SS_IN][SS_IN + 1];
  __shared__ float hb[5][SS_IN][SS_T + 1];
  __shared__ float red[8][8];
  __shared__ bool is_last;
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  // ownership: pixels of this tile's 32x32 core; the last tile row / column extends to the image edge
  const int own_y1 = blockIdx.y == gridDim.y - 1 ? h : y0 + SS_T, own_x1 = blockIdx.x == gridDim.x - 1 ? w : x0 + SS_T;
  constexpr int PER = (SS_IN * SS_IN + 255) / 256;              // 7 region pixels per thread
  float dot[PER], np[PER], nt[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) dot[i] = np[i] = nt[i] = 0.f;
  float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f, sq = 0.f, lo = INFINITY, hi = -INFINITY;
  for (int ch = 0; ch < c; ++ch) {
    const float* P = p + ((int64_t)img * c + ch) * h * w;
    const float* T = t + ((int64_t)img * c + ch) * h * w;
    __syncthreads();                                             // previous channel's blur has finished with sp / st / hb
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int e = tid + i * 256;
      if (e < SS_IN * SS_IN) {
        const int yy = e / SS_IN, xx = e - yy * SS_IN;
        const int gy = y0 + yy, gx = x0 + xx;
        const bool ok = gy < h && gx < w;
        const float u = ok ? __ldg(P + (int64_t)gy * w + gx) : 0.f, v = ok ? __ldg(T + (int64_t)gy * w + gx) : 0.f;
        sp[yy][xx] = u;
        st[yy][xx] = v;
        if (ok && gy < own_y1 && gx < own_x1) {
          const float d = u - v;
          sq = fmaf(d, d, sq);
          dot[i] = fmaf(u, v, dot[i]); np[i] = fmaf(u, u, np[i]); nt[i] = fmaf(v, v, nt[i]);
          lo = fminf(lo, u); hi = fmaxf(hi, u);
        }
      }
    }
    __syncthreads();
    for (int e = tid; e < SS_IN * SS_T; e += 256) {
      const int yy = e / SS_T, xx = e - yy * SS_T;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
      for (int k = 0; k < SS_W; ++k) {
        const float g = c_gauss[k], u = sp[yy][xx + k], v = st[yy][xx + k];
        a0 = fmaf(g, u, a0); a1 = fmaf(g, v, a1);
        a2 = fmaf(g, u * u, a2); a3 = fmaf(g, v * v, a3); a4 = fmaf(g, u * v, a4);
      }
      hb[0][yy][xx] = a0; hb[1][yy][xx] = a1; hb[2][yy][xx] = a2; hb[3][yy][xx] = a3; hb[4][yy][xx] = a4;
    }
    __syncthreads();
    for (int yy = threadIdx.y; yy < SS_T; yy += 8) {
      const int xx = threadIdx.x;
      if (y0 + yy < ho && x0 + xx < wo) {
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < SS_W; ++k) {
          const float g = c_gauss[k];
          m1 = fmaf(g, hb[0][yy + k][xx], m1); m2 = fmaf(g, hb[1][yy + k][xx], m2);
          e11 = fmaf(g, hb[2][yy + k][xx], e11); e22 = fmaf(g, hb[3][yy + k][xx], e22);
          e12 = fmaf(g, hb[4][yy + k][xx], e12);
        }
        ss0 += ssim_point(m1, m2, e11, e22, e12, 1.f);
        ss1 += ssim_point(m1, m2, e11, e22, e12, 2.f);
        ss2 += ssim_point(m1, m2, e11, e22, e12, 255.f);
        ss3 += ssim_point(m1, m2, e11, e22, e12, 256.f);
      }
    }
  }
  float ae = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int e = tid + i * 256;
    if (e < SS_IN * SS_IN) {
      const int yy = e / SS_IN, xx = e - yy * SS_IN;
      const int gy = y0 + yy, gx = x0 + xx;
      if (gy < own_y1 && gx < own_x1 && gy < h && gx < w)
        ae += acosf(dot[i] / (sqrtf(np[i]) * sqrtf(nt[i]) + 1e-6f)) * 57.29577951308232f;
    }
  }
  // block reduction (fixed order): sums by warp shuffles + 8 warp partials; min / max likewise
  float v[6] = {ss0, ss1, ss2, ss3, sq, ae};
#pragma unroll
  for (int k = 0; k < 6; ++k) v[k] = warp_sum(v[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((tid & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) red[k][tid >> 5] = v[k];
    red[6][tid >> 5] = lo;
    red[7][tid >> 5] = hi;
  }
  __syncthreads();
  const int blk = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int nblk = gridDim.x * gridDim.y * gridDim.z;
  if (tid < EV_PART) {
    float s = red[tid][0];
    for (int k = 1; k < 8; ++k) s = tid == 6 ? fminf(s, red[6][k]) : (tid == 7 ? fmaxf(s, red[7][k]) : s + red[tid][k]);
    part[(int64_t)blk * EV_PART + tid] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(ticket, 1u) == (unsigned)nblk - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // ---- the last block: global min / max -> L, then the sums in block order (double accumulation, fixed order)
  const int per_img = gridDim.x * gridDim.y;
  __shared__ double dred[8];
  __shared__ float s_sel[2];
  float glo = INFINITY, ghi = -INFINITY;
  for (int b = tid; b < nblk; b += 256) {
    glo = fminf(glo, __ldcg(part + (int64_t)b * EV_PART + 6));
    ghi = fmaxf(ghi, __ldcg(part + (int64_t)b * EV_PART + 7));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    glo = fminf(glo, __shfl_xor_sync(0xffffffffu, glo, o));
    ghi = fmaxf(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
  }
  if ((tid & 31) == 0) { red[6][tid >> 5] = glo; red[7][tid >> 5] = ghi; }
  __syncthreads();
  if (tid == 0) {
    for (int k = 1; k < 8; ++k) { glo = fminf(glo, red[6][k]); ghi = fmaxf(ghi, red[7][k]); }
    s_sel[0] = glo; s_sel[1] = ghi;
  }
  __syncthreads();
  glo = s_sel[0]; ghi = s_sel[1];
  const float L = (ghi > 128.f ? 255.f : 1.f) - (glo < -0.5f ? -1.f : 0.f);
  const int cand = (ghi > 128.f ? 2 : 0) + (glo < -0.5f ? 1 : 0);
  auto block_total = [&](int first, int count, int field) -> double {    // deterministic: strided partials, fixed tree
    double s = 0.0;
    for (int b = tid; b < count; b += 256) s += (double)__ldcg(part + (int64_t)(first + b) * EV_PART + field);
    s = warp_sum(s);
    __syncthreads();
    if ((tid & 31) == 0) dred[tid >> 5] = s;
    __syncthreads();
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += dred[k];
    return tot;
  };
  const double count = (double)c * ho * wo, hw = (double)h * w;
  double ssim_all = 0.0, ae_all = 0.0;
  for (int i = 0; i < n; ++i) {
    const double s = block_total(i * per_img, per_img, cand), a = block_total(i * per_img, per_img, 5);
    const double sq_i = block_total(i * per_img, per_img, 4);
    // the image's own data range (what the reference's loop sees when it scores one tile per call, testCas.py:65-85)
    float ilo = INFINITY, ihi = -INFINITY;
    for (int b = tid; b < per_img; b += 256) {
      ilo = fminf(ilo, __ldcg(part + (int64_t)(i * per_img + b) * EV_PART + 6));
      ihi = fmaxf(ihi, __ldcg(part + (int64_t)(i * per_img + b) * EV_PART + 7));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ilo = fminf(ilo, __shfl_xor_sync(0xffffffffu, ilo, o));
      ihi = fmaxf(ihi, __shfl_xor_sync(0xffffffffu, ihi, o));
    }
    __syncthreads();
    if ((tid & 31) == 0) { red[6][tid >> 5] = ilo; red[7][tid >> 5] = ihi; }
    __syncthreads();
    for (int k = 0; k < 8; ++k) { ilo = fminf(ilo, red[6][k]); ihi = fmaxf(ihi, red[7][k]); }
    const int cand_i = (ihi > 128.f ? 2 : 0) + (ilo < -0.5f ? 1 : 0);
    const double s_own = cand_i == cand ? s : block_total(i * per_img, per_img, cand_i);
    if (tid == 0) {
      out[8 + i] = (float)(s / count);
      out[8 + n + i] = (float)(a / hw);
      out[8 + 2 * n + i] = (float)(sq_i / ((double)c * hw));
      out[8 + 3 * n + i] = (float)(s_own / count);
      out[8 + 4 * n + i] = (ihi > 128.f ? 255.f : 1.f) - (ilo < -0.5f ? -1.f : 0.f);
    }
    ssim_all += s;
    ae_all += a / hw;
  }
  const double sqsum = block_total(0, nblk, 4);
  if (tid == 0) {
    const float mse = (float)(sqsum / ((double)n * c * hw));
    out[0] = mse;
    out[1] = 10.f * log10f(1.f / mse);
    out[2] = (float)(ae_all / n);
    out[3] = (float)(ssim_all / (count * n));
    out[4] = L; out[5] = glo; out[6] = ghi; out[7] = 0.f;
    *ticket = 0u;                                                 // re-armed for the next call on this workspace
  }
}

size_t eval_metrics_workspace_bytes(int n, int c, int h, int w) {
  int ho = h - SS_W + 1, wo = w - SS_W + 1;
  if (ho < 1 || wo < 1) return 512;
  return 256 + (size_t)n * ceil_div(ho, SS_T) * ceil_div(wo, SS_T) * EV_PART * sizeof(float);
}

static int upload_gauss(cudaStream_t st) {
  static DeviceOnce init;     // __constant__ memory is per device
  int init_dev;
  if (init.needed(&init_dev)) {
    float g[SS_W];
    for (int i = 0; i < SS_W; ++i) g[i] = (float)exp(-((i - SS_W / 2) * (i - SS_W / 2)) / (2.0 * 1.5 * 1.5));
    float fs = 0.f;            // the reference normalises in fp32 (torch.Tensor / sum)
    for (int i = 0; i < SS_W; ++i) fs += g[i];
    for (int i = 0; i < SS_W; ++i) g[i] = g[i] / fs;
    SRCGAN_CUDA(cudaMemcpyToSymbolAsync(c_gauss, g, sizeof(g), 0, cudaMemcpyHostToDevice, st));
    init.mark(init_dev);
  }
  return SRCGAN_OK;
}

// ws: first 256 bytes = ticket counter (must be zero on first use: the caller passes a zero-initialised workspace)
int eval_metrics(const float* p, const float* t, int n, int c, int h, int w, float* out, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  SRCGAN_REQUIRE(p && t && out && n > 0 && c > 0, "eval_metrics: null pointer / empty batch");
  SRCGAN_REQUIRE(h >= SS_W && w >= SS_W, "eval_metrics: image smaller than the 11x11 SSIM window");
  SRCGAN_REQUIRE(ws && ws_bytes >= eval_metrics_workspace_bytes(n, c, h, w), "eval_metrics: workspace too small");
  int rc = upload_gauss(st);
  if (rc) return rc;
  const int ho = h - SS_W + 1, wo = w - SS_W + 1;
  dim3 grid(ceil_div(wo, SS_T), ceil_div(ho, SS_T), n), blk(32, 8);
  eval_metrics_k<<<grid, blk, 0, st>>>(p, t, c, h, w, ho, wo, reinterpret_cast<float*>((char*)ws + 256),
                                       reinterpret_cast<unsigned int*>(ws), out, n);
  count_launch();
  return check_launch("eval_metrics");
}

// ---------------------------------------------------------------------------------------------
// SSIM backward (losses.DSSIMLoss is differentiable through SSIM, src/losses.py:170-180, 40-93):
//   S = A1*A2 / (B1*B2),  A1 = 2 m1 m2 + C1, A2 = 2 s12 + C2, B1 = m1^2 + m2^2 + C1, B2 = s11 + s22 + C2,
//   m1 = G*p, e11 = G*(p^2), e12 = G*(p t)  ('valid' 11x11 gaussian), s11 = e11 - m1^2, s12 = e12 - m1 m2.
//   dS/dm1  = 2 m2 (A2 - A1)/(B1 B2) - 2 m1 A1 A2 (B2 - B1)/(B1 B2)^2,  dS/de11 = -A1 A2/(B1 B2^2),  dS/de12 = 2 A1/(B1 B2)
//   dp(j) = scale * sum_i g(i - j) [ dS/dm1(i) + 2 p(j) dS/de11(i) + t(j) dS/de12(i) ]        (adjoint of the valid blur)
// Kernel 1 writes the three derivative maps (ho x wo per plane), kernel 2 applies the adjoint blur.  `scale_dev` is a
// device scalar (the upstream gradient), `coef` a host constant (e.g. -0.5 / count for DSSIM with size_average).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ssim_bwd_maps(const float* __restrict__ p, const float* __restrict__ t, int h, int w, int ho, int wo,
              const float* __restrict__ Lp, float* __restrict__ maps) {
  __shared__ float sp[SS_IN][SS_IN + 1], st[SS_IN][SS_IN + 1];
  __shared__ float hb[5][SS_IN][SS_T + 1];
  const int64_t plane = blockIdx.z;
  const float* P = p + plane * h * w;
  const float* T = t + plane * h * w;
  const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const float L = __ldg(Lp);
  const float C1 = (0.01f * L) * (0.01f * L), C2 = (0.03f * L) * (0.03f * L);
  for (int e = tid; e < SS_IN * SS_IN; e += 256) {
    int yy = e / SS_IN, xx = e - yy * SS_IN;
    int gy = y0 + yy, gx = x0 + xx;
    bool ok = gy < h && gx < w;
    sp[yy][xx] = ok ? __ldg(P + (int64_t)gy * w + gx) : 0.f;
    st[yy][xx] = ok ? __ldg(T + (int64_t)gy * w + gx) : 0.f;
  }
  __syncthreads();
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
  const int64_t msz = (int64_t)ho * wo;
  float* M = maps + plane * 3 * msz;
  for (int yy = threadIdx.y; yy < SS_T; yy += 8) {
    const int xx = threadIdx.x;
    if (y0 + yy < ho && x0 + xx < wo) {
      float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
      for (int k = 0; k < SS_W; ++k) {
        float g = c_gauss[k];
        m1 = fmaf(g, hb[0][yy + k][xx], m1); m2 = fmaf(g, hb[1][yy + k][xx], m2);
        e11 = fmaf(g, hb[2][yy + k][xx], e11); e22 = fmaf(g, hb[3][yy + k][xx], e22);
        e12 = fmaf(g, hb[4][yy + k][xx], e12);
      }
      const float s11 = e11 - m1 * m1, s22 = e22 - m2 * m2, s12 = e12 - m1 * m2;
      const float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * s12 + C2, B1 = m1 * m1 + m2 * m2 + C1, B2 = s11 + s22 + C2;
      const float inv = 1.f / (B1 * B2);
      const int64_t o = (int64_t)(y0 + yy) * wo + x0 + xx;
      M[o] = 2.f * m2 * (A2 - A1) * inv - 2.f * m1 * A1 * A2 * (B2 - B1) * inv * inv;     // dS/dm1
      M[msz + o] = -A1 * A2 * inv / B2;                                                    // dS/de11
      M[2 * msz + o] = 2.f * A1 * inv;                                                     // dS/de12
    }
  }
}

// grid (ceil(w/32), ceil(h/32), n*c): dp over a 32x32 tile of INPUT pixels; reads the maps over [y-10, y] x [x-10, x]
__global__ void __launch_bounds__(256)
ssim_bwd_apply(const float* __restrict__ p, const float* __restrict__ t, int h, int w, int ho, int wo,
               const float* __restrict__ maps, const float* __restrict__ scale_dev, float coef, float* __restrict__ dp) {
  __shared__ float sm[3][SS_IN][SS_IN + 1];
  __shared__ float hv[3][SS_IN][SS_T + 1];
  const int64_t plane = blockIdx.z;
  const int64_t msz = (int64_t)ho * wo;
  const float* M = maps + plane * 3 * msz;
  const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int e = tid; e < SS_IN * SS_IN; e += 256) {
    const int yy = e / SS_IN, xx = e - yy * SS_IN;
    const int oy = y0 + yy - (SS_W - 1), ox = x0 + xx - (SS_W - 1);          // map coordinate
    const bool ok = oy >= 0 && ox >= 0 && oy < ho && ox < wo;
#pragma unroll
    for (int k = 0; k < 3; ++k) sm[k][yy][xx] = ok ? __ldg(M + k * msz + (int64_t)oy * wo + ox) : 0.f;
  }
  __syncthreads();
  // input pixel x receives map x - k with weight g[k]: region column (x - x0) + 10 - k
  for (int e = tid; e < SS_IN * SS_T; e += 256) {
    const int yy = e / SS_T, xx = e - yy * SS_T;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int k = 0; k < SS_W; ++k) {
      const float g = c_gauss[k];
      a0 = fmaf(g, sm[0][yy][xx + SS_W - 1 - k], a0);
      a1 = fmaf(g, sm[1][yy][xx + SS_W - 1 - k], a1);
      a2 = fmaf(g, sm[2][yy][xx + SS_W - 1 - k], a2);
    }
    hv[0][yy][xx] = a0; hv[1][yy][xx] = a1; hv[2][yy][xx] = a2;
  }
  __syncthreads();
  const float sc = __ldg(scale_dev) * coef;
  for (int yy = threadIdx.y; yy < SS_T; yy += 8) {
    const int xx = threadIdx.x;
    const int gy = y0 + yy, gx = x0 + xx;
    if (gy < h && gx < w) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int k = 0; k < SS_W; ++k) {
        const float g = c_gauss[k];
        a0 = fmaf(g, hv[0][yy + SS_W - 1 - k][xx], a0);
        a1 = fmaf(g, hv[1][yy + SS_W - 1 - k][xx], a1);
        a2 = fmaf(g, hv[2][yy + SS_W - 1 - k][xx], a2);
      }
      const int64_t o = plane * h * w + (int64_t)gy * w + gx;
      dp[o] = sc * (a0 + 2.f * __ldg(p + o) * a1 + __ldg(t + o) * a2);
    }
  }
}

size_t ssim_backward_workspace_bytes(int n, int c, int h, int w) {
  int ho = h - SS_W + 1, wo = w - SS_W + 1;
  if (ho < 1 || wo < 1) return 256;
  return (size_t)n * c * 3 * ho * wo * sizeof(float) + 256;
}

int ssim_backward(const float* p, const float* t, int n, int c, int h, int w, const float* L_dev, const float* scale_dev,
                  float coef, float* dp, void* ws, size_t ws_bytes, cudaStream_t st) {
  SRCGAN_REQUIRE(p && t && L_dev && scale_dev && dp, "ssim_backward: null pointer");
  SRCGAN_REQUIRE(h >= SS_W && w >= SS_W, "ssim_backward: image smaller than the 11x11 window");
  SRCGAN_REQUIRE(ws && ws_bytes >= ssim_backward_workspace_bytes(n, c, h, w), "ssim_backward: workspace too small");
  int rc = upload_gauss(st);
  if (rc) return rc;
  const int ho = h - SS_W + 1, wo = w - SS_W + 1;
  float* maps = reinterpret_cast<float*>(ws);
  dim3 blk(32, 8);
  ssim_bwd_maps<<<dim3(ceil_div(wo, SS_T), ceil_div(ho, SS_T), n * c), blk, 0, st>>>(p, t, h, w, ho, wo, L_dev, maps);
  ssim_bwd_apply<<<dim3(ceil_div(w, SS_T), ceil_div(h, SS_T), n * c), blk, 0, st>>>(p, t, h, w, ho, wo, maps, scale_dev, coef, dp);
  count_launch(2);
  return check_launch("ssim_backward");
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
