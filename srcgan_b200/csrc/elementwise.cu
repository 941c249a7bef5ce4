// Bandwidth-bound glue kernels: layout conversion at the module boundary (NCHW fp32 <-> NHWC),
// adjoint of nearest x2 upsampling (backward of F.interpolate at src/model/model.py:426-427).
#include "common.cuh"

namespace srcgan {

// NCHW fp32 -> NHWC T.  One thread per (pixel, channel) with channel fastest in the write; the
// read is strided by H*W per channel which is fine for the tiny C (1..3) this is used for, and a
// 32x32 smem transpose is used for wide C.
template <typename T>
__global__ void nchw_to_nhwc_small(const float* __restrict__ src, int c, int64_t hw, int64_t npix,
                                   T* __restrict__ dst, int ld) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  int64_t n = p / hw, q = p - n * hw;
  const float* s = src + n * c * hw + q;
  T* d = dst + p * ld;
  for (int k = 0; k < c; ++k) d[k] = from_f32<T>(__ldg(s + (int64_t)k * hw));
}

template <typename T>
__global__ void nhwc_to_nchw_small(const T* __restrict__ src, int ld, int c, int64_t hw, int64_t npix,
                                   float* __restrict__ dst) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  int64_t n = p / hw, q = p - n * hw;
  const T* s = src + p * ld;
  float* d = dst + n * c * hw + q;
  for (int k = 0; k < c; ++k) d[(int64_t)k * hw] = to_f32(s[k]);
}

// wide-C variants: tile of 32 pixels x 32 channels through shared memory, both sides coalesced
template <typename T>
__global__ void nchw_to_nhwc_tiled(const float* __restrict__ src, int c, int64_t hw, T* __restrict__ dst, int ld) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t q0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int ch = c0 + j;
    int64_t q = q0 + threadIdx.x;
    tile[j][threadIdx.x] = (ch < c && q < hw) ? __ldg(src + (n * c + ch) * hw + q) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t q = q0 + j;
    int ch = c0 + threadIdx.x;
    if (q < hw && ch < c) dst[(n * hw + q) * ld + ch] = from_f32<T>(tile[threadIdx.x][j]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_tiled(const T* __restrict__ src, int ld, int c, int64_t hw, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t q0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t q = q0 + j;
    int ch = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (q < hw && ch < c) ? to_f32(src[(n * hw + q) * ld + ch]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int ch = c0 + j;
    int64_t q = q0 + threadIdx.x;
    if (ch < c && q < hw) dst[(n * c + ch) * hw + q] = tile[threadIdx.x][j];
  }
}

template <typename T>
__global__ void upsample2x_adjoint_k(const T* __restrict__ src, int src_ld, T* __restrict__ dst, int dst_ld,
                                     const T* __restrict__ mask, int mask_ld, float mslope, int n, int h, int w,
                                     int c) {
  int64_t total = (int64_t)n * h * w * c;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ch = (int)(i % c);
  int64_t p = i / c;
  int x = (int)(p % w);
  int64_t t = p / w;
  int y = (int)(t % h);
  int64_t b = t / h;
  const int W2 = 2 * w;
  int64_t s00 = ((b * 2 * h + 2 * y) * W2 + 2 * x) * src_ld + ch;
  float v = to_f32(src[s00]) + to_f32(src[s00 + src_ld]) + to_f32(src[s00 + (int64_t)W2 * src_ld]) +
            to_f32(src[s00 + (int64_t)(W2 + 1) * src_ld]);
  if (mask) v *= (to_f32(mask[p * mask_ld + ch]) > 0.f ? 1.f : mslope);
  dst[p * dst_ld + ch] = from_f32<T>(v);
}

// bf16, 16-byte vectors (8 channels per thread): the four source pixels are summed in fp32 in the scalar kernel's order
__global__ void __launch_bounds__(256)
upsample2x_adjoint_vec_bf16(const uint4* __restrict__ src, int src_ld16, uint4* __restrict__ dst, int dst_ld16,
                            const uint4* __restrict__ mask, int mask_ld16, float mslope, int n, int h, int w, int c16) {
  const int64_t total = (int64_t)n * h * w * c16;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = (int)(i % c16);
  const int64_t p = i / c16;
  const int x = (int)(p % w);
  const int64_t t = p / w;
  const int y = (int)(t % h);
  const int64_t b = t / h;
  const int W2 = 2 * w;
  const int64_t s00 = ((b * 2 * h + 2 * y) * W2 + 2 * x) * src_ld16 + v;
  const uint4 q0 = __ldg(src + s00), q1 = __ldg(src + s00 + src_ld16), q2 = __ldg(src + s00 + (int64_t)W2 * src_ld16),
              q3 = __ldg(src + s00 + (int64_t)(W2 + 1) * src_ld16);
  uint4 qm = make_uint4(0, 0, 0, 0);
  if (mask) qm = __ldg(mask + p * mask_ld16 + v);
  const __nv_bfloat162 *a0 = reinterpret_cast<const __nv_bfloat162*>(&q0), *a1 = reinterpret_cast<const __nv_bfloat162*>(&q1),
                       *a2 = reinterpret_cast<const __nv_bfloat162*>(&q2), *a3 = reinterpret_cast<const __nv_bfloat162*>(&q3),
                       *am = reinterpret_cast<const __nv_bfloat162*>(&qm);
  uint4 o;
  __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float lo = __low2float(a0[k]) + __low2float(a1[k]) + __low2float(a2[k]) + __low2float(a3[k]);
    float hi = __high2float(a0[k]) + __high2float(a1[k]) + __high2float(a2[k]) + __high2float(a3[k]);
    if (mask) {
      lo *= __low2float(am[k]) > 0.f ? 1.f : mslope;
      hi *= __high2float(am[k]) > 0.f ? 1.f : mslope;
    }
    oh[k] = __floats2bfloat162_rn(lo, hi);
  }
  dst[p * dst_ld16 + v] = o;
}

// nearest x2 upsampling, 16-byte vectors: dst[n,2y+a,2x+b,:] = src[n,y,x,:]
__global__ void upsample2x_vec(const uint4* __restrict__ src, int src_ld16, uint4* __restrict__ dst, int dst_ld16, int n,
                               int h, int w, int c16) {
  int64_t total = (int64_t)n * 2 * h * 2 * w * c16;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int v = (int)(i % c16);
  int64_t p = i / c16;
  int X = (int)(p % (2 * w));
  int64_t t = p / (2 * w);
  int Y = (int)(t % (2 * h));
  int64_t b = t / (2 * h);
  dst[p * dst_ld16 + v] = __ldg(src + ((b * h + (Y >> 1)) * w + (X >> 1)) * src_ld16 + v);
}

int upsample2x(const void* src, int src_ld, void* dst, int dst_ld, int n, int h, int w, int c, int dtype,
               cudaStream_t st) {
  const int es = dtype == SRCGAN_DT_F32 ? 4 : 2;
  const int per16 = 16 / es;
  SRCGAN_REQUIRE(c % per16 == 0 && src_ld % per16 == 0 && dst_ld % per16 == 0 && ((uintptr_t)src % 16 == 0) &&
                     ((uintptr_t)dst % 16 == 0),
                 "upsample2x: channels / strides must be multiples of 16 bytes");
  int64_t total = (int64_t)n * 4 * h * w * (c / per16);
  upsample2x_vec<<<ceil_div(total, 256), 256, 0, st>>>((const uint4*)src, src_ld / per16, (uint4*)dst, dst_ld / per16, n,
                                                       h, w, c / per16);
  count_launch();
  return check_launch("upsample2x");
}

// depth-to-space for the k2 s2 transposed convolution (rddb.py:94-97) computed as a 1x1 conv to 4*c channels:
//   forward : dst[n,2y+a,2x+b,co] = src[n,y,x,(a*2+b)*c+co]
//   adjoint : dst[n,y,x,(a*2+b)*c+co] = src[n,2y+a,2x+b,co] * (mask[n,y,x,(a*2+b)*c+co] > 0 ? 1 : slope)
template <typename T, bool ADJ>
__global__ void d2s_k(const T* __restrict__ src, int src_ld, T* __restrict__ dst, int dst_ld, const T* __restrict__ mask,
                      int mask_ld, float mslope, int n, int h, int w, int c) {
  int64_t total = (int64_t)n * h * w * 4 * c;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int wc = (int)(i % (4 * c));
  int64_t p = i / (4 * c);
  int x = (int)(p % w);
  int64_t t = p / w;
  int y = (int)(t % h);
  int64_t b = t / h;
  int ab = wc / c, co = wc - ab * c;
  int64_t big = ((b * 2 * h + 2 * y + (ab >> 1)) * (2 * w) + 2 * x + (ab & 1));
  if (!ADJ) {
    dst[big * dst_ld + co] = src[p * src_ld + wc];
  } else {
    float v = to_f32(src[big * src_ld + co]);
    if (mask) v *= (to_f32(mask[p * mask_ld + wc]) > 0.f ? 1.f : mslope);
    dst[p * dst_ld + wc] = from_f32<T>(v);
  }
}

// nn.PixelShuffle(r) in NHWC (src/model/espcn.py:35,50): dst[n,h*r+a,w*r+b,c] = src[n,h,w,c*r*r+a*r+b]; ADJ = inverse
template <typename T, bool ADJ>
__global__ void pixel_shuffle_k(const T* __restrict__ src, int src_ld, T* __restrict__ dst, int dst_ld, int n, int h,
                                int w, int c, int r) {
  const int rr = r * r;
  int64_t total = (int64_t)n * h * w * c * rr;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int wc = (int)(i % (c * rr));
  int64_t p = i / (c * rr);
  int x = (int)(p % w);
  int64_t t = p / w;
  int y = (int)(t % h);
  int64_t b = t / h;
  int ch = wc / rr, ab = wc - ch * rr, a_ = ab / r, b_ = ab - a_ * r;
  int64_t big = ((b * h * r + (int64_t)y * r + a_) * ((int64_t)w * r) + (int64_t)x * r + b_);
  if (!ADJ) dst[big * dst_ld + ch] = src[p * src_ld + wc];
  else dst[p * dst_ld + wc] = src[big * src_ld + ch];
}

int pixel_shuffle(const void* src, int src_ld, void* dst, int dst_ld, int n, int h, int w, int c, int r, int dtype,
                  int adjoint, cudaStream_t st) {
  int64_t total = (int64_t)n * h * w * c * r * r;
  unsigned grid = (unsigned)ceil_div(total, 256);
  if (dtype == SRCGAN_DT_F32) {
    if (adjoint) pixel_shuffle_k<float, true><<<grid, 256, 0, st>>>((const float*)src, src_ld, (float*)dst, dst_ld, n, h, w, c, r);
    else pixel_shuffle_k<float, false><<<grid, 256, 0, st>>>((const float*)src, src_ld, (float*)dst, dst_ld, n, h, w, c, r);
  } else {
    if (adjoint) pixel_shuffle_k<__nv_bfloat16, true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, src_ld, (__nv_bfloat16*)dst, dst_ld, n, h, w, c, r);
    else pixel_shuffle_k<__nv_bfloat16, false><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, src_ld, (__nv_bfloat16*)dst, dst_ld, n, h, w, c, r);
  }
  count_launch();
  return check_launch("pixel_shuffle");
}

int depth_to_space(const void* src, int src_ld, void* dst, int dst_ld, const void* mask, int mask_ld, float mslope,
                   int n, int h, int w, int c, int dtype, int adjoint, cudaStream_t st) {
  int64_t total = (int64_t)n * h * w * 4 * c;
  unsigned grid = (unsigned)ceil_div(total, 256);
  if (dtype == SRCGAN_DT_F32) {
    if (adjoint) d2s_k<float, true><<<grid, 256, 0, st>>>((const float*)src, src_ld, (float*)dst, dst_ld, (const float*)mask, mask_ld, mslope, n, h, w, c);
    else d2s_k<float, false><<<grid, 256, 0, st>>>((const float*)src, src_ld, (float*)dst, dst_ld, nullptr, 0, 0.f, n, h, w, c);
  } else {
    if (adjoint) d2s_k<__nv_bfloat16, true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, src_ld, (__nv_bfloat16*)dst, dst_ld, (const __nv_bfloat16*)mask, mask_ld, mslope, n, h, w, c);
    else d2s_k<__nv_bfloat16, false><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, src_ld, (__nv_bfloat16*)dst, dst_ld, nullptr, 0, 0.f, n, h, w, c);
  }
  count_launch();
  return check_launch("depth_to_space");
}

template <typename T>
__global__ void add_k(const T* a, int a_ld, const T* b, int b_ld, T* d /* may alias a or b */, int d_ld,
                      int64_t npix, int c) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * c) return;
  int ch = (int)(i % c);
  int64_t m = i / c;
  d[m * d_ld + ch] = from_f32<T>(to_f32(a[m * a_ld + ch]) + to_f32(b[m * b_ld + ch]));
}

__global__ void add_vec_bf16(const __nv_bfloat16* a, int a_ld, const __nv_bfloat16* b, int b_ld,
                             __nv_bfloat16* d /* may alias a or b */, int d_ld, int64_t npix, int oct) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * oct) return;
  const int ch = (int)(i % oct) * 8;
  const int64_t m = i / oct;
  uint4 qa = *reinterpret_cast<const uint4*>(a + m * a_ld + ch), qb = *reinterpret_cast<const uint4*>(b + m * b_ld + ch);
  const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&qa);
  const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&qb);
  uint4 o;
  __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    ho[k] = __floats2bfloat162_rn(__low2float(ha[k]) + __low2float(hb[k]), __high2float(ha[k]) + __high2float(hb[k]));
  *reinterpret_cast<uint4*>(d + m * d_ld + ch) = o;
}

int add_slices(const void* a, int a_ld, const void* b, int b_ld, void* d, int d_ld, int64_t npix, int c, int dtype,
               cudaStream_t st) {
  if (dtype == SRCGAN_DT_BF16 && c % 8 == 0 && a_ld % 8 == 0 && b_ld % 8 == 0 && d_ld % 8 == 0 &&
      ((uintptr_t)a) % 16 == 0 && ((uintptr_t)b) % 16 == 0 && ((uintptr_t)d) % 16 == 0) {
    add_vec_bf16<<<ceil_div(npix * (c / 8), 256), 256, 0, st>>>((const __nv_bfloat16*)a, a_ld, (const __nv_bfloat16*)b, b_ld,
                                                                (__nv_bfloat16*)d, d_ld, npix, c / 8);
    count_launch();
    return check_launch("add");
  }
  if (dtype == SRCGAN_DT_F32)
    add_k<float><<<ceil_div(npix * c, 256), 256, 0, st>>>((const float*)a, a_ld, (const float*)b, b_ld, (float*)d, d_ld,
                                                          npix, c);
  else
    add_k<__nv_bfloat16><<<ceil_div(npix * c, 256), 256, 0, st>>>((const __nv_bfloat16*)a, a_ld,
                                                                  (const __nv_bfloat16*)b, b_ld, (__nv_bfloat16*)d,
                                                                  d_ld, npix, c);
  count_launch();
  return check_launch("add");
}

// y[px][ch] = leaky(y[px][ch] + bias[ch]) in place: bias / activation after a gather-form transposed convolution
// (functional._ConvTranspose2d), instead of materialising a broadcast bias tensor for the add kernel
template <typename T>
__global__ void bias_act_k(T* y, int ld, int64_t npix, int c, const float* __restrict__ bias, int has_act, float slope) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * c) return;
  const int ch = (int)(i % c);
  const int64_t m = i / c;
  float v = to_f32(y[m * ld + ch]);
  if (bias) v += __ldg(bias + ch);
  if (has_act) v = v > 0.f ? v : v * slope;
  y[m * ld + ch] = from_f32<T>(v);
}

int bias_act(void* y, int ld, int64_t npix, int c, const float* bias, int has_act, float slope, int dtype, cudaStream_t st) {
  if (dtype == SRCGAN_DT_F32) bias_act_k<float><<<ceil_div(npix * c, 256), 256, 0, st>>>((float*)y, ld, npix, c, bias, has_act, slope);
  else bias_act_k<__nv_bfloat16><<<ceil_div(npix * c, 256), 256, 0, st>>>((__nv_bfloat16*)y, ld, npix, c, bias, has_act, slope);
  count_launch();
  return check_launch("bias_act");
}

template <typename T>
__global__ void act_bwd_k(const T* dy, int dy_ld, const T* y, int y_ld, T* d /* may alias dy or y */, int d_ld,
                          int64_t npix, int c, float slope) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * c) return;
  int ch = (int)(i % c);
  int64_t m = i / c;
  const float g = to_f32(dy[m * dy_ld + ch]);
  d[m * d_ld + ch] = from_f32<T>(to_f32(y[m * y_ld + ch]) > 0.f ? g : g * slope);
}

__global__ void act_bwd_vec_bf16(const __nv_bfloat16* dy, int dy_ld, const __nv_bfloat16* y,
                                 int y_ld, __nv_bfloat16* d /* may alias dy or y */, int d_ld, int64_t npix, int oct, float slope) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * oct) return;
  const int ch = (int)(i % oct) * 8;
  const int64_t m = i / oct;
  uint4 qa = *reinterpret_cast<const uint4*>(dy + m * dy_ld + ch), qb = *reinterpret_cast<const uint4*>(y + m * y_ld + ch);
  const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&qa);
  const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&qb);
  uint4 o;
  __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float g0 = __low2float(ha[k]), g1 = __high2float(ha[k]);
    ho[k] = __floats2bfloat162_rn(__low2float(hb[k]) > 0.f ? g0 : g0 * slope, __high2float(hb[k]) > 0.f ? g1 : g1 * slope);
  }
  *reinterpret_cast<uint4*>(d + m * d_ld + ch) = o;
}

int act_backward(const void* dy, int dy_ld, const void* y, int y_ld, void* d, int d_ld, int64_t npix, int c, float slope,
                 int dtype, cudaStream_t st) {
  if (dtype == SRCGAN_DT_BF16 && c % 8 == 0 && dy_ld % 8 == 0 && y_ld % 8 == 0 && d_ld % 8 == 0 &&
      ((uintptr_t)dy) % 16 == 0 && ((uintptr_t)y) % 16 == 0 && ((uintptr_t)d) % 16 == 0) {
    act_bwd_vec_bf16<<<ceil_div(npix * (c / 8), 256), 256, 0, st>>>((const __nv_bfloat16*)dy, dy_ld, (const __nv_bfloat16*)y,
                                                                    y_ld, (__nv_bfloat16*)d, d_ld, npix, c / 8, slope);
  } else if (dtype == SRCGAN_DT_F32) {
    act_bwd_k<float><<<ceil_div(npix * c, 256), 256, 0, st>>>((const float*)dy, dy_ld, (const float*)y, y_ld, (float*)d,
                                                              d_ld, npix, c, slope);
  } else {
    act_bwd_k<__nv_bfloat16><<<ceil_div(npix * c, 256), 256, 0, st>>>((const __nv_bfloat16*)dy, dy_ld,
                                                                      (const __nv_bfloat16*)y, y_ld, (__nv_bfloat16*)d,
                                                                      d_ld, npix, c, slope);
  }
  count_launch();
  return check_launch("act_backward");
}

template <typename T>
static int nchw_to_nhwc_t(const float* src, int n, int c, int h, int w, T* dst, int ld, cudaStream_t st) {
  int64_t hw = (int64_t)h * w, npix = hw * n;
  if (c <= 8) {
    nchw_to_nhwc_small<T><<<ceil_div(npix, 256), 256, 0, st>>>(src, c, hw, npix, dst, ld);
  } else {
    dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), n), blk(32, 8);
    nchw_to_nhwc_tiled<T><<<grid, blk, 0, st>>>(src, c, hw, dst, ld);
  }
  count_launch();
  return check_launch("nchw_to_nhwc");
}
template <typename T>
static int nhwc_to_nchw_t(const T* src, int ld, float* dst, int n, int c, int h, int w, cudaStream_t st) {
  int64_t hw = (int64_t)h * w, npix = hw * n;
  if (c <= 8) {
    nhwc_to_nchw_small<T><<<ceil_div(npix, 256), 256, 0, st>>>(src, ld, c, hw, npix, dst);
  } else {
    dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), n), blk(32, 8);
    nhwc_to_nchw_tiled<T><<<grid, blk, 0, st>>>(src, ld, c, hw, dst);
  }
  count_launch();
  return check_launch("nhwc_to_nchw");
}

int nchw_to_nhwc(const float* src, int n, int c, int h, int w, void* dst, int ld, int dtype, cudaStream_t st) {
  if (dtype == SRCGAN_DT_F32) return nchw_to_nhwc_t<float>(src, n, c, h, w, (float*)dst, ld, st);
  return nchw_to_nhwc_t<__nv_bfloat16>(src, n, c, h, w, (__nv_bfloat16*)dst, ld, st);
}
int nhwc_to_nchw(const void* src, int ld, int dtype, float* dst, int n, int c, int h, int w, cudaStream_t st) {
  if (dtype == SRCGAN_DT_F32) return nhwc_to_nchw_t<float>((const float*)src, ld, dst, n, c, h, w, st);
  return nhwc_to_nchw_t<__nv_bfloat16>((const __nv_bfloat16*)src, ld, dst, n, c, h, w, st);
}

int upsample2x_adjoint(const void* src, int src_ld, void* dst, int dst_ld, const void* mask, int mask_ld,
                       float mslope, int n, int h, int w, int c, int dtype, cudaStream_t st) {
  int64_t total = (int64_t)n * h * w * c;
  auto al16 = [](const void* ptr, int ld) { return ptr == nullptr || (((uintptr_t)ptr) % 16 == 0 && ld % 8 == 0); };
  if (dtype == SRCGAN_DT_BF16 && c % 8 == 0 && al16(src, src_ld) && al16(dst, dst_ld) && al16(mask, mask_ld)) {
    upsample2x_adjoint_vec_bf16<<<ceil_div(total / 8, 256), 256, 0, st>>>((const uint4*)src, src_ld / 8, (uint4*)dst, dst_ld / 8,
                                                                          (const uint4*)mask, mask_ld / 8, mslope, n, h, w, c / 8);
    count_launch();
    return check_launch("upsample2x_adjoint");
  }
  if (dtype == SRCGAN_DT_F32)
    upsample2x_adjoint_k<float><<<ceil_div(total, 256), 256, 0, st>>>((const float*)src, src_ld, (float*)dst, dst_ld,
                                                                     (const float*)mask, mask_ld, mslope, n, h, w, c);
  else
    upsample2x_adjoint_k<__nv_bfloat16><<<ceil_div(total, 256), 256, 0, st>>>(
        (const __nv_bfloat16*)src, src_ld, (__nv_bfloat16*)dst, dst_ld, (const __nv_bfloat16*)mask, mask_ld, mslope,
        n, h, w, c);
  count_launch();
  return check_launch("upsample2x_adjoint");
}

}  // namespace srcgan
