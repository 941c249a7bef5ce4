// BatchNorm2d (+ fused LeakyReLU) over channel-sliced NHWC rows, forward and backward.
// Reference: nn.BatchNorm2d(train) + nn.LeakyReLU in Decoder (src/model/model.py:240-289, slope 0.1)
// and NLayerDiscriminator (src/model/model.py:620-632, slope 0.2); eps 1e-5, momentum 0.1, running
// variance updated with the unbiased estimate.  HBM-bound: one read pass for the statistics, one
// read+write pass for the normalisation.
#include "common.cuh"

namespace srcgan {

constexpr int kBnMaxParts = 592;            // 4 blocks per SM for the HBM-bound statistics passes

static int bn_parts(int64_t npix) {
  int64_t p = (npix + 511) / 512;
  if (p < 1) p = 1;
  if (p > kBnMaxParts) p = kBnMaxParts;
  return (int)p;
}

size_t bn_workspace_bytes(int64_t npix, int c) {
  return ((size_t)2 * bn_parts(npix) * c + 2 * (size_t)c) * sizeof(float) + 256;
}

// block (32,8); grid (ceil(c/32), parts)
template <typename T>
__global__ void bn_stats_partial(const T* __restrict__ x, int ld, int64_t npix, int c, float* __restrict__ psum,
                                 float* __restrict__ psq) {
  __shared__ float r0[8][33], r1[8][33];
  int ch = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f, q = 0.f;
  if (ch < c)
    for (int64_t m = (int64_t)blockIdx.y * 8 + threadIdx.y; m < npix; m += (int64_t)gridDim.y * 8) {
      float v = to_f32(x[m * ld + ch]);
      s += v; q = fmaf(v, v, q);
    }
  r0[threadIdx.y][threadIdx.x] = s; r1[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += r0[k][threadIdx.x]; b += r1[k][threadIdx.x]; }
    psum[(int64_t)blockIdx.y * c + ch] = a;
    psq[(int64_t)blockIdx.y * c + ch] = b;
  }
}

// Fixed-order sum of `parts` partial rows for 32 channels by one block of 32 channels x 32 row lanes: lane rl adds rows
// rl, rl + 32, ... in double (the loads of a lane are independent and issued four at a time - the old 8-lane version was a
// serial chain of 74 dependent load+add steps and took 43 us, as long as the statistics pass itself at batch 16), then the
// 32 lane sums are added in lane order.  Result in s / q of the threads with rl == 0.
__device__ __forceinline__ void bn_reduce_parts(const float* __restrict__ a, const float* __restrict__ b, int parts, int c,
                                                int ch, int rl, int cl, double (&rs)[32][33], double (&rq)[32][33],
                                                double& s, double& q) {
  s = 0.0; q = 0.0;
  if (ch < c) {
    int k = rl;
    for (; k + 96 < parts; k += 128) {
      const float a0 = __ldg(a + (int64_t)k * c + ch), a1 = __ldg(a + (int64_t)(k + 32) * c + ch),
                  a2 = __ldg(a + (int64_t)(k + 64) * c + ch), a3 = __ldg(a + (int64_t)(k + 96) * c + ch);
      const float b0 = __ldg(b + (int64_t)k * c + ch), b1 = __ldg(b + (int64_t)(k + 32) * c + ch),
                  b2 = __ldg(b + (int64_t)(k + 64) * c + ch), b3 = __ldg(b + (int64_t)(k + 96) * c + ch);
      s += (double)a0; s += (double)a1; s += (double)a2; s += (double)a3;
      q += (double)b0; q += (double)b1; q += (double)b2; q += (double)b3;
    }
    for (; k < parts; k += 32) { s += (double)__ldg(a + (int64_t)k * c + ch); q += (double)__ldg(b + (int64_t)k * c + ch); }
  }
  rs[rl][cl] = s; rq[rl][cl] = q;
  __syncthreads();
  if (rl == 0) {
    s = 0.0; q = 0.0;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) { s += rs[r][cl]; q += rq[r][cl]; }
  }
}

__global__ void __launch_bounds__(1024)
bn_finalize(const float* __restrict__ psum, const float* __restrict__ psq, int parts, int c,
            int64_t npix, float eps, float momentum, float* __restrict__ running_mean,
            float* __restrict__ running_var, float* __restrict__ save_mean,
            float* __restrict__ save_invstd) {
  __shared__ double rs[32][33], rq[32][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cl;
  double s, q;
  bn_reduce_parts(psum, psq, parts, c, ch, rl, cl, rs, rq, s, q);
  if (rl != 0 || ch >= c) return;
  double mean = s / (double)npix;
  double var = q / (double)npix - mean * mean;
  if (var < 0.0) var = 0.0;
  save_mean[ch] = (float)mean;
  save_invstd[ch] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
  if (running_var) {
    double unb = npix > 1 ? var * (double)npix / (double)(npix - 1) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unb;
  }
}

__global__ void bn_eval_stats(const float* __restrict__ running_mean, const float* __restrict__ running_var, int c,
                              float eps, float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  save_mean[ch] = running_mean[ch];
  save_invstd[ch] = rsqrtf(running_var[ch] + eps);
}

template <typename T>
__global__ void bn_apply(const T* __restrict__ x, int x_ld, T* __restrict__ y, int y_ld, int64_t npix, int c,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ mean, const float* __restrict__ invstd, float slope) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * c) return;
  int ch = (int)(i % c);
  int64_t m = i / c;
  float sc = gamma[ch] * invstd[ch];
  float v = fmaf(to_f32(x[m * x_ld + ch]) - mean[ch], sc, beta[ch]);
  v = v > 0.f ? v : v * slope;
  y[m * y_ld + ch] = from_f32<T>(v);
}

// bf16 fast paths: 16-byte vectors, 8 channels per thread (c % 8 == 0, lds % 8 == 0, 16-byte aligned)
__device__ __forceinline__ void unpack8b(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ uint4 pack8b(const float (&f)[8]) {
  uint4 o;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return o;
}

// The apply passes are pure streaming (HBM-bound).  thread = (pixel lane, channel octet): the per-channel constants are
// loaded once into registers and the thread then walks pixels with four 16-byte loads in flight (the first version was
// one pixel-octet per thread and re-read 32 per-channel scalars for every 16 bytes of data: 2-5x off the roofline).
__global__ void __launch_bounds__(256)
bn_apply_vec(const __nv_bfloat16* __restrict__ x, int x_ld, __nv_bfloat16* __restrict__ y, int y_ld,
             int64_t npix, int c, const float* __restrict__ gamma, const float* __restrict__ beta,
             const float* __restrict__ mean, const float* __restrict__ invstd, float slope) {
  const int oct = c >> 3;
  const int rows_per_iter = 256 / oct;
  const int oc = threadIdx.x % oct, rl = threadIdx.x / oct;
  if (rl >= rows_per_iter) return;
  const int ch = oc * 8;
  float mu[8], gi[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { mu[k] = mean[ch + k]; gi[k] = gamma[ch + k] * invstd[ch + k]; be[k] = beta[ch + k]; }
  const int64_t step = (int64_t)gridDim.x * rows_per_iter;
  int64_t m = (int64_t)blockIdx.x * rows_per_iter + rl;
  for (; m + 3 * step < npix; m += 4 * step) {
    uint4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(x + (m + u * step) * x_ld + ch));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8b(q[u], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float v = fmaf(f[k] - mu[k], gi[k], be[k]);
        f[k] = v > 0.f ? v : v * slope;
      }
      *reinterpret_cast<uint4*>(y + (m + u * step) * y_ld + ch) = pack8b(f);
    }
  }
  for (; m < npix; m += step) {
    float f[8];
    unpack8b(__ldg(reinterpret_cast<const uint4*>(x + m * x_ld + ch)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = fmaf(f[k] - mu[k], gi[k], be[k]);
      f[k] = v > 0.f ? v : v * slope;
    }
    *reinterpret_cast<uint4*>(y + m * y_ld + ch) = pack8b(f);
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_vec(const __nv_bfloat16* dy /* may alias dx (in-place): plain loads, no __restrict__ */, int dy_ld,
                 const __nv_bfloat16* __restrict__ y,
                 int y_ld, const __nv_bfloat16* __restrict__ x, int x_ld, __nv_bfloat16* dx,
                 int dx_ld, int64_t npix, int c, const float* __restrict__ gamma,
                 const float* __restrict__ mean, const float* __restrict__ invstd,
                 const float* __restrict__ tot, float slope, int training, const float* __restrict__ beta) {
  // beta != nullptr: the LeakyReLU mask is RECOMPUTED from x with the forward's own expression (bn_apply_vec:
  // fmaf(x - mean, gamma * invstd, beta) > 0) instead of read from the saved activation y - 5 tensor sweeps instead of 7
  const bool recompute = beta != nullptr;
  const int oct = c >> 3;
  const int rows_per_iter = 256 / oct;
  const int oc = threadIdx.x % oct, rl = threadIdx.x / oct;
  if (rl >= rows_per_iter) return;
  const int ch = oc * 8;
  const float inv_n = 1.f / (float)npix;
  float mu[8], is[8], g[8], t0[8], t1[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    is[k] = invstd[ch + k]; g[k] = gamma[ch + k] * is[k]; mu[k] = mean[ch + k];
    t0[k] = training ? tot[ch + k] * inv_n : 0.f; t1[k] = training ? tot[c + ch + k] * inv_n : 0.f;
    be[k] = recompute ? beta[ch + k] : 0.f;
  }
  const int64_t step = (int64_t)gridDim.x * rows_per_iter;
  int64_t m = (int64_t)blockIdx.x * rows_per_iter + rl;
  for (; m < npix; m += 2 * step) {
    const bool two = m + step < npix;
    uint4 qd[2], qy[2], qx[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 0 || two) {
        const int64_t mm = m + u * step;
        qd[u] = *reinterpret_cast<const uint4*>(dy + mm * dy_ld + ch);
        if (!recompute) qy[u] = __ldg(reinterpret_cast<const uint4*>(y + mm * y_ld + ch));
        qx[u] = __ldg(reinterpret_cast<const uint4*>(x + mm * x_ld + ch));
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 0 || two) {
        float d[8], yy[8], xx[8];
        unpack8b(qd[u], d); unpack8b(qx[u], xx);
        if (!recompute) unpack8b(qy[u], yy);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float z = recompute ? fmaf(xx[k] - mu[k], g[k], be[k]) : yy[k];
          const float dz = z > 0.f ? d[k] : d[k] * slope;
          if (training) {
            const float xh = (xx[k] - mu[k]) * is[k];
            d[k] = g[k] * (dz - t0[k] - xh * t1[k]);
          } else {
            d[k] = g[k] * dz;
          }
        }
        *reinterpret_cast<uint4*>(dx + (m + u * step) * dx_ld + ch) = pack8b(d);
      }
    }
  }
}

// bf16, 16-byte loads: thread = (row lane, channel octet); two sums per channel, block tree-reduce
template <bool BWD>
__global__ void __launch_bounds__(256)
bn_partial_vec(const __nv_bfloat16* __restrict__ x, int x_ld, const __nv_bfloat16* __restrict__ dy, int dy_ld,
               const __nv_bfloat16* __restrict__ y, int y_ld, int64_t npix, int c, const float* __restrict__ mean,
               const float* __restrict__ invstd, float slope, float* __restrict__ p0, float* __restrict__ p1,
               const float* __restrict__ gamma, const float* __restrict__ beta) {
  __shared__ float r0[256][9], r1[256][9];
  const bool recompute = BWD && beta != nullptr;     // LeakyReLU mask from x (see bn_bwd_apply_vec), y is not read
  const int octets = c >> 3;
  const int rows_per_iter = 256 / octets;
  const int oc = threadIdx.x % octets, rl = threadIdx.x / octets;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (rl < rows_per_iter) {
    float mu[8], is[8], gi[8], be[8];
    if (BWD) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mu[i] = mean[oc * 8 + i]; is[i] = invstd[oc * 8 + i];
        gi[i] = recompute ? gamma[oc * 8 + i] * is[i] : 0.f; be[i] = recompute ? beta[oc * 8 + i] : 0.f;
      }
    }
    const int64_t step = (int64_t)gridDim.x * rows_per_iter;
    int64_t m = (int64_t)blockIdx.x * rows_per_iter + rl;
    if (!BWD) {                                   // four independent 16-byte loads in flight (HBM-bound pass)
      for (; m + 3 * step < npix; m += 4 * step) {
        uint4 qv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) qv[u] = __ldg(reinterpret_cast<const uint4*>(x + (m + u * step) * x_ld + oc * 8));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float xv[8];
          unpack8b(qv[u], xv);
#pragma unroll
          for (int i = 0; i < 8; ++i) { s[i] += xv[i]; q[i] = fmaf(xv[i], xv[i], q[i]); }
        }
      }
    }
    if (BWD) {                                    // two pixels (six 16-byte loads) in flight
      for (; m + step < npix; m += 2 * step) {
        uint4 qx[2], qd[2], qy[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          qx[u] = __ldg(reinterpret_cast<const uint4*>(x + (m + u * step) * x_ld + oc * 8));
          qd[u] = __ldg(reinterpret_cast<const uint4*>(dy + (m + u * step) * dy_ld + oc * 8));
          if (!recompute) qy[u] = __ldg(reinterpret_cast<const uint4*>(y + (m + u * step) * y_ld + oc * 8));
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float xv[8], dv[8], yv[8];
          unpack8b(qx[u], xv); unpack8b(qd[u], dv);
          if (!recompute) unpack8b(qy[u], yv);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float z = recompute ? fmaf(xv[i] - mu[i], gi[i], be[i]) : yv[i];
            const float dz = z > 0.f ? dv[i] : dv[i] * slope;
            s[i] += dz; q[i] = fmaf(dz, (xv[i] - mu[i]) * is[i], q[i]);
          }
        }
      }
    }
    for (; m < npix; m += step) {
      float xv[8];
      unpack8b(__ldg(reinterpret_cast<const uint4*>(x + m * x_ld + oc * 8)), xv);
      if (!BWD) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += xv[i]; q[i] = fmaf(xv[i], xv[i], q[i]); }
      } else {
        float dv[8], yv[8];
        unpack8b(__ldg(reinterpret_cast<const uint4*>(dy + m * dy_ld + oc * 8)), dv);
        if (!recompute) unpack8b(__ldg(reinterpret_cast<const uint4*>(y + m * y_ld + oc * 8)), yv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = recompute ? fmaf(xv[i] - mu[i], gi[i], be[i]) : yv[i];
          const float dz = z > 0.f ? dv[i] : dv[i] * slope;
          s[i] += dz; q[i] = fmaf(dz, (xv[i] - mu[i]) * is[i], q[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { r0[threadIdx.x][i] = s[i]; r1[threadIdx.x][i] = q[i]; }
  __syncthreads();
  if (threadIdx.x < c) {
    const int o = threadIdx.x >> 3, i = threadIdx.x & 7;
    float a = 0.f, bsum = 0.f;
    for (int r = 0; r < rows_per_iter; ++r) { a += r0[r * octets + o][i]; bsum += r1[r * octets + o][i]; }
    p0[(int64_t)blockIdx.x * c + threadIdx.x] = a;
    p1[(int64_t)blockIdx.x * c + threadIdx.x] = bsum;
  }
}

// partial sums of dz and dz*xhat, dz = dy_post * lrelu'(y)
template <typename T>
__global__ void bn_bwd_partial(const T* __restrict__ dy, int dy_ld, const T* __restrict__ y, int y_ld,
                               const T* __restrict__ x, int x_ld, int64_t npix, int c,
                               const float* __restrict__ mean, const float* __restrict__ invstd, float slope,
                               float* __restrict__ p0, float* __restrict__ p1) {
  __shared__ float r0[8][33], r1[8][33];
  int ch = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f, q = 0.f;
  if (ch < c) {
    float mu = mean[ch], is = invstd[ch];
    for (int64_t m = (int64_t)blockIdx.y * 8 + threadIdx.y; m < npix; m += (int64_t)gridDim.y * 8) {
      float dz = to_f32(dy[m * dy_ld + ch]);
      if (!(to_f32(y[m * y_ld + ch]) > 0.f)) dz *= slope;
      float xh = (to_f32(x[m * x_ld + ch]) - mu) * is;
      s += dz; q = fmaf(dz, xh, q);
    }
  }
  r0[threadIdx.y][threadIdx.x] = s; r1[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += r0[k][threadIdx.x]; b += r1[k][threadIdx.x]; }
    p0[(int64_t)blockIdx.y * c + ch] = a;
    p1[(int64_t)blockIdx.y * c + ch] = b;
  }
}

__global__ void __launch_bounds__(1024)
bn_bwd_finalize(const float* __restrict__ p0, const float* __restrict__ p1, int parts, int c,
                float* __restrict__ tot, float* __restrict__ dgamma, float* __restrict__ dbeta,
                int accumulate) {
  __shared__ double rs[32][33], rq[32][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cl;
  double s, q;
  bn_reduce_parts(p0, p1, parts, c, ch, rl, cl, rs, rq, s, q);
  if (rl != 0 || ch >= c) return;
  tot[ch] = (float)s; tot[c + ch] = (float)q;
  if (dbeta) dbeta[ch] = accumulate ? dbeta[ch] + (float)s : (float)s;
  if (dgamma) dgamma[ch] = accumulate ? dgamma[ch] + (float)q : (float)q;
}

template <typename T>
__global__ void bn_bwd_apply(const T* dy /* may alias dx */, int dy_ld, const T* __restrict__ y, int y_ld,
                             const T* __restrict__ x, int x_ld, T* dx, int dx_ld, int64_t npix, int c,
                             const float* __restrict__ gamma, const float* __restrict__ mean,
                             const float* __restrict__ invstd, const float* __restrict__ tot, float slope,
                             int training) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * c) return;
  int ch = (int)(i % c);
  int64_t m = i / c;
  float dz = to_f32(dy[m * dy_ld + ch]);
  if (!(to_f32(y[m * y_ld + ch]) > 0.f)) dz *= slope;
  float is = invstd[ch];
  float g = gamma[ch] * is;
  float v;
  if (training) {
    float inv_n = 1.f / (float)npix;
    float xh = (to_f32(x[m * x_ld + ch]) - mean[ch]) * is;
    v = g * (dz - tot[ch] * inv_n - xh * tot[c + ch] * inv_n);
  } else {
    v = g * dz;
  }
  dx[m * dx_ld + ch] = from_f32<T>(v);
}

// grid of the streaming apply kernels: every block covers 256 / (c/8) pixels per iteration, four iterations per trip
static unsigned bn_stream_blocks(int64_t npix, int c) {
  const int rows_per_iter = 256 / (c >> 3);
  int64_t nb = (npix + (int64_t)rows_per_iter * 4 - 1) / ((int64_t)rows_per_iter * 4);
  if (nb > 16 * kNumSMs) nb = 16 * kNumSMs;
  return (unsigned)(nb < 1 ? 1 : nb);
}

template <typename T>
static int bn_forward_t(const T* x, int x_ld, T* y, int y_ld, int64_t npix, int c, const float* gamma,
                        const float* beta, float* rm, float* rv, float* save_mean, float* save_invstd, int training,
                        float momentum, float eps, float slope, float* ws, cudaStream_t st) {
  if (training) {
    int parts = bn_parts(npix);
    float* psum = ws;
    float* psq = ws + (size_t)parts * c;
    dim3 grid(ceil_div(c, 32), parts), blk(32, 8);
    if (sizeof(T) == 2 && c % 8 == 0 && c <= 256 && x_ld % 8 == 0 && ((uintptr_t)x) % 16 == 0)
      bn_partial_vec<false><<<parts, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), x_ld, nullptr, 0, nullptr, 0,
                                                   npix, c, nullptr, nullptr, 0.f, psum, psq, nullptr, nullptr);
    else
      bn_stats_partial<T><<<grid, blk, 0, st>>>(x, x_ld, npix, c, psum, psq);
    bn_finalize<<<ceil_div(c, 32), 1024, 0, st>>>(psum, psq, parts, c, npix, eps, momentum, rm, rv, save_mean,
                                                  save_invstd);
    count_launch(2);
  } else {
    bn_eval_stats<<<ceil_div(c, 128), 128, 0, st>>>(rm, rv, c, eps, save_mean, save_invstd);
    count_launch();
  }
  if (sizeof(T) == 2 && c % 8 == 0 && c <= 2048 && x_ld % 8 == 0 && y_ld % 8 == 0 && ((uintptr_t)x) % 16 == 0 &&
      ((uintptr_t)y) % 16 == 0)
    bn_apply_vec<<<bn_stream_blocks(npix, c), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), x_ld, reinterpret_cast<__nv_bfloat16*>(y), y_ld, npix, c, gamma,
        beta, save_mean, save_invstd, slope);
  else
    bn_apply<T><<<ceil_div(npix * c, 256), 256, 0, st>>>(x, x_ld, y, y_ld, npix, c, gamma, beta, save_mean,
                                                         save_invstd, slope);
  count_launch();
  return check_launch("bn_forward");
}

int bn_forward(const void* x, int x_ld, void* y, int y_ld, int64_t npix, int c, int dtype, const float* gamma,
               const float* beta, float* rm, float* rv, float* save_mean, float* save_invstd, int training,
               float momentum, float eps, float slope, void* ws, size_t ws_bytes, cudaStream_t st) {
  SRCGAN_REQUIRE(x && y && gamma && beta && save_mean && save_invstd, "bn_forward: null pointer");
  SRCGAN_REQUIRE(training || (rm && rv), "bn_forward: eval mode needs running stats");
  SRCGAN_REQUIRE(ws && ws_bytes >= bn_workspace_bytes(npix, c), "bn_forward: workspace too small");
  if (dtype == SRCGAN_DT_F32)
    return bn_forward_t<float>((const float*)x, x_ld, (float*)y, y_ld, npix, c, gamma, beta, rm, rv, save_mean,
                               save_invstd, training, momentum, eps, slope, (float*)ws, st);
  return bn_forward_t<__nv_bfloat16>((const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)y, y_ld, npix, c, gamma, beta, rm,
                                     rv, save_mean, save_invstd, training, momentum, eps, slope, (float*)ws, st);
}

template <typename T>
static int bn_backward_t(const T* dy, int dy_ld, const T* y, int y_ld, const T* x, int x_ld, T* dx, int dx_ld,
                         int64_t npix, int c, const float* gamma, const float* mean, const float* invstd, float slope,
                         int training, float* dgamma, float* dbeta, int accumulate, float* ws, cudaStream_t st,
                         const float* beta) {
  int parts = bn_parts(npix);
  float* p0 = ws;
  float* p1 = ws + (size_t)parts * c;
  float* tot = ws + (size_t)2 * parts * c;
  dim3 grid(ceil_div(c, 32), parts), blk(32, 8);
  if (sizeof(T) == 2 && c % 8 == 0 && c <= 256 && x_ld % 8 == 0 && dy_ld % 8 == 0 && y_ld % 8 == 0 &&
      ((uintptr_t)x) % 16 == 0 && ((uintptr_t)dy) % 16 == 0 && ((uintptr_t)y) % 16 == 0)
    bn_partial_vec<true><<<parts, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), x_ld,
                                                reinterpret_cast<const __nv_bfloat16*>(dy), dy_ld,
                                                reinterpret_cast<const __nv_bfloat16*>(y), y_ld, npix, c, mean, invstd,
                                                slope, p0, p1, gamma, beta);
  else
    bn_bwd_partial<T><<<grid, blk, 0, st>>>(dy, dy_ld, y, y_ld, x, x_ld, npix, c, mean, invstd, slope, p0, p1);
  bn_bwd_finalize<<<ceil_div(c, 32), 1024, 0, st>>>(p0, p1, parts, c, tot, dgamma, dbeta, accumulate);
  if (sizeof(T) == 2 && c % 8 == 0 && c <= 2048 && dy_ld % 8 == 0 && y_ld % 8 == 0 && x_ld % 8 == 0 && dx_ld % 8 == 0 &&
      ((uintptr_t)dy) % 16 == 0 && ((uintptr_t)y) % 16 == 0 && ((uintptr_t)x) % 16 == 0 && ((uintptr_t)dx) % 16 == 0)
    bn_bwd_apply_vec<<<bn_stream_blocks(npix, c), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(dy), dy_ld, reinterpret_cast<const __nv_bfloat16*>(y), y_ld,
        reinterpret_cast<const __nv_bfloat16*>(x), x_ld, reinterpret_cast<__nv_bfloat16*>(dx), dx_ld, npix, c, gamma,
        mean, invstd, tot, slope, training, beta);
  else
    bn_bwd_apply<T><<<ceil_div(npix * c, 256), 256, 0, st>>>(dy, dy_ld, y, y_ld, x, x_ld, dx, dx_ld, npix, c, gamma,
                                                             mean, invstd, tot, slope, training);
  count_launch(3);
  return check_launch("bn_backward");
}

int bn_backward(const void* dy, int dy_ld, const void* y, int y_ld, const void* x, int x_ld, void* dx, int dx_ld,
                int64_t npix, int c, int dtype, const float* gamma, const float* mean, const float* invstd,
                float slope, int training, float* dgamma, float* dbeta, int accumulate, void* ws, size_t ws_bytes,
                cudaStream_t st, const float* beta) {
  SRCGAN_REQUIRE(dy && y && x && dx && gamma && mean && invstd, "bn_backward: null pointer");
  SRCGAN_REQUIRE(ws && ws_bytes >= bn_workspace_bytes(npix, c), "bn_backward: workspace too small");
  if (dtype == SRCGAN_DT_F32)
    return bn_backward_t<float>((const float*)dy, dy_ld, (const float*)y, y_ld, (const float*)x, x_ld, (float*)dx,
                                dx_ld, npix, c, gamma, mean, invstd, slope, training, dgamma, dbeta, accumulate,
                                (float*)ws, st, nullptr);
  return bn_backward_t<__nv_bfloat16>((const __nv_bfloat16*)dy, dy_ld, (const __nv_bfloat16*)y, y_ld,
                                      (const __nv_bfloat16*)x, x_ld, (__nv_bfloat16*)dx, dx_ld, npix, c, gamma, mean,
                                      invstd, slope, training, dgamma, dbeta, accumulate, (float*)ws, st, beta);
}

}  // namespace srcgan
