// SIMT (FFMA, fp32-accumulate) implicit-GEMM convolution engine: fprop, dgrad (gather form of the
// transposed convolution, any stride) and split-K wgrad over channel-sliced NHWC tensors.
// This is the fp32 parity engine (exact fp32 math) and the engine for the layers tensor cores
// cannot help (3-channel image convs, 1-channel patch logits, strided layers).
//
// Reference semantics: nn.Conv2d fwd/bwd as used at src/model/model.py:193-211 (dense block),
// :236-289 (Decoder), :396-440 (RDDBNetB), :612-635 (NLayerDiscriminator).
#include "common.cuh"

namespace srcgan {

struct ConvDev {
  int n, h, w, cin, cout, kh, kw, stride, pad, up, ho, wo;
  const void* x; int x_ld;
  const void* wgt;
  const float* bias;
  void* y; int y_ld;
  int act; float act_slope, alpha;
  const void* r1; int r1_ld; float beta1;
  const void* r2; int r2_ld; float beta2;
  const void* mask; int mask_ld; float mask_slope;
  int vec_a;   // K-dim channel loads may use 4-wide vectors
  int vec_b;   // weight rows may use 4-wide vectors
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}

constexpr int BK = 16;

// One CTA computes a BM (pixels) x BN (channels) tile of the output; K = taps * kch.
//   DGRAD == false : output pixel = (n,oy,ox) of Y, K-channels = cin, N-channels = cout
//   DGRAD == true  : output pixel = (n,iy,ix) of dX, K-channels = cout, N-channels = cin
template <typename T, int BM, int BN, int TM, int TN, bool DGRAD>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_igemm_simt(const ConvDev a) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int rowN[BM], rowY[BM], rowX[BM];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int OH = DGRAD ? a.h : a.ho, OW = DGRAD ? a.w : a.wo;
  const int64_t M = (int64_t)a.n * OH * OW;
  const int kch = DGRAD ? a.cout : a.cin;
  const int nch = DGRAD ? a.cin : a.cout;
  const int K = a.kh * a.kw * kch;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const T* __restrict__ X = reinterpret_cast<const T*>(a.x);
  const T* __restrict__ Wt = reinterpret_cast<const T*>(a.wgt);

  for (int r = tid; r < BM; r += NT) {
    int64_t m = m0 + r;
    if (m < M) {
      int ox = (int)(m % OW);
      int64_t t = m / OW;
      rowX[r] = ox; rowY[r] = (int)(t % OH); rowN[r] = (int)(t / OH);
    } else {
      rowN[r] = -1; rowY[r] = 0; rowX[r] = 0;
    }
  }
  __syncthreads();

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int SH = DGRAD ? a.ho : a.h, SW = DGRAD ? a.wo : a.w;   // stored dims of the source tensor
  const int HV = a.up ? 2 * a.h : a.h, WV = a.up ? 2 * a.w : a.w;

  // returns element offset of source pixel for (row r, tap), or -1
  auto src_pixel = [&](int r, int tap) -> int64_t {
    int n = rowN[r];
    if (n < 0) return -1;
    int fr = tap / a.kw, fs = tap - fr * a.kw;
    int sy, sx;
    if (!DGRAD) {
      int iy = rowY[r] * a.stride - a.pad + fr, ix = rowX[r] * a.stride - a.pad + fs;
      if (iy < 0 || iy >= HV || ix < 0 || ix >= WV) return -1;
      sy = a.up ? (iy >> 1) : iy; sx = a.up ? (ix >> 1) : ix;
    } else {
      int ty_ = rowY[r] + a.pad - fr, tx_ = rowX[r] + a.pad - fs;
      if (ty_ < 0 || tx_ < 0) return -1;
      if (a.stride > 1) {
        if ((ty_ % a.stride) | (tx_ % a.stride)) return -1;
        ty_ /= a.stride; tx_ /= a.stride;
      }
      if (ty_ >= SH || tx_ >= SW) return -1;
      sy = ty_; sx = tx_;
    }
    return (((int64_t)n * SH + sy) * SW + sx) * a.x_ld;
  };

  for (int kc = 0; kc < K; kc += BK) {
    // ---- A tile: BM x BK, 4 consecutive k per item
    for (int e = tid; e < BM * (BK / 4); e += NT) {
      int r = e / (BK / 4), kq = (e % (BK / 4)) * 4;
      int k = kc + kq;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.vec_a) {
        if (k < K) {
          int tap = k / kch, c = k - tap * kch;
          int64_t off = src_pixel(r, tap);
          if (off >= 0) load4<T>(X + off + c, v);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int kk = k + q;
          if (kk < K) {
            int tap = kk / kch, c = kk - tap * kch;
            int64_t off = src_pixel(r, tap);
            if (off >= 0) v[q] = to_f32(X[off + c]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) As[kq + q][r] = v[q];
    }
    // ---- B tile: BK x BN from wgt[k][nch]
    for (int e = tid; e < BK * (BN / 4); e += NT) {
      int kr = e / (BN / 4), nq = (e % (BN / 4)) * 4;
      int k = kc + kr, nn = n0 + nq;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < K) {
        if (a.vec_b && nn + 3 < nch) {
          load4<T>(Wt + (int64_t)k * nch + nn, v);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (nn + q < nch) v[q] = to_f32(Wt[(int64_t)k * nch + nn + q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) Bs[kr][nq + q] = v[q];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  T* __restrict__ Y = reinterpret_cast<T*>(a.y);
  const T* R1 = reinterpret_cast<const T*>(a.r1);
  const T* R2 = reinterpret_cast<const T*>(a.r2);
  const T* MK = reinterpret_cast<const T*>(a.mask);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int c = n0 + tx * TN + j;
      if (c >= nch) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[c];
      if (a.act) v = v > 0.f ? v : v * a.act_slope;
      v *= a.alpha;
      if (R1) v = fmaf(a.beta1, to_f32(R1[m * a.r1_ld + c]), v);
      if (R2) v = fmaf(a.beta2, to_f32(R2[m * a.r2_ld + c]), v);
      if (MK) v *= (to_f32(MK[m * a.mask_ld + c]) > 0.f ? 1.f : a.mask_slope);
      Y[m * a.y_ld + c] = from_f32<T>(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: part[split][tap][ci][co] = sum over the split's pixels of X[src(m,tap)][ci] * dY[m][co]
// ---------------------------------------------------------------------------------------------
template <typename T, int BI, int BO, int TI, int TO>
__global__ void __launch_bounds__((BI / TI) * (BO / TO))
conv_wgrad_simt(const ConvDev a, float* __restrict__ part, int pix_per_split) {
  constexpr int NT = (BI / TI) * (BO / TO);
  __shared__ float Xs[BK][BI + 4];
  __shared__ float Gs[BK][BO + 4];
  const int tid = threadIdx.x;
  const int to_ = tid % (BO / TO), ti = tid / (BO / TO);
  const int n_it = (a.cin + BI - 1) / BI, n_ot = (a.cout + BO - 1) / BO;
  int b = blockIdx.x;
  const int ot = b % n_ot; b /= n_ot;
  const int it = b % n_it; b /= n_it;
  const int tap = b;
  const int fr = tap / a.kw, fs = tap - fr * a.kw;
  const int ci0 = it * BI, co0 = ot * BO;
  const int64_t M = (int64_t)a.n * a.ho * a.wo;
  const int64_t mbeg = (int64_t)blockIdx.y * pix_per_split;
  const int64_t mend = (mbeg + pix_per_split < M) ? mbeg + pix_per_split : M;
  const T* __restrict__ X = reinterpret_cast<const T*>(a.x);
  const T* __restrict__ G = reinterpret_cast<const T*>(a.y);
  const int HV = a.up ? 2 * a.h : a.h, WV = a.up ? 2 * a.w : a.w;

  float acc[TI][TO];
#pragma unroll
  for (int i = 0; i < TI; ++i)
#pragma unroll
    for (int j = 0; j < TO; ++j) acc[i][j] = 0.f;

  for (int64_t mc = mbeg; mc < mend; mc += BK) {
    for (int e = tid; e < BK * BI; e += NT) {
      int kr = e / BI, c = e - kr * BI;
      int64_t m = mc + kr;
      float v = 0.f;
      if (m < mend && ci0 + c < a.cin) {
        int ox = (int)(m % a.wo);
        int64_t t = m / a.wo;
        int oy = (int)(t % a.ho), n = (int)(t / a.ho);
        int iy = oy * a.stride - a.pad + fr, ix = ox * a.stride - a.pad + fs;
        if (iy >= 0 && iy < HV && ix >= 0 && ix < WV) {
          int sy = a.up ? (iy >> 1) : iy, sx = a.up ? (ix >> 1) : ix;
          v = to_f32(X[(((int64_t)n * a.h + sy) * a.w + sx) * a.x_ld + ci0 + c]);
        }
      }
      Xs[kr][c] = v;
    }
    for (int e = tid; e < BK * BO; e += NT) {
      int kr = e / BO, c = e - kr * BO;
      int64_t m = mc + kr;
      float v = 0.f;
      if (m < mend && co0 + c < a.cout) v = to_f32(G[m * a.y_ld + co0 + c]);
      Gs[kr][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float xv[TI], gv[TO];
#pragma unroll
      for (int i = 0; i < TI; ++i) xv[i] = Xs[kk][ti * TI + i];
#pragma unroll
      for (int j = 0; j < TO; ++j) gv[j] = Gs[kk][to_ * TO + j];
#pragma unroll
      for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TO; ++j) acc[i][j] = fmaf(xv[i], gv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = part + ((int64_t)blockIdx.y * a.kh * a.kw + tap) * a.cin * a.cout;
#pragma unroll
  for (int i = 0; i < TI; ++i) {
    int ci = ci0 + ti * TI + i;
    if (ci >= a.cin) continue;
#pragma unroll
    for (int j = 0; j < TO; ++j) {
      int co = co0 + to_ * TO + j;
      if (co < a.cout) out[(int64_t)ci * a.cout + co] = acc[i][j];
    }
  }
}

// dw[co][ci][tap] (+)= sum_s part[s][tap][ci][co]    (deterministic, fixed order)
__global__ void wgrad_reduce(const float* __restrict__ part, int splits, int taps, int cin, int cout,
                             float* __restrict__ dw, int accumulate, float alpha) {
  int64_t total = (int64_t)taps * cin * cout;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int co = (int)(i % cout);
  int64_t t = i / cout;
  int ci = (int)(t % cin);
  int tap = (int)(t / cin);
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(int64_t)k * total + i];
  s *= alpha;
  int64_t o = ((int64_t)co * cin + ci) * taps + tap;
  dw[o] = accumulate ? dw[o] + s : s;
}

// db[c] (+)= sum_m G[m][c] : stage 1 partial column sums, stage 2 fixed-order reduce
template <typename T>
__global__ void colsum_partial(const T* __restrict__ g, int ld, int64_t M, int c, float* __restrict__ part) {
  // block (32, 8): thread.x -> channel, thread.y -> row lane
  __shared__ float red[8][33];
  int ch = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (ch < c)
    for (int64_t m = (int64_t)blockIdx.y * 8 + threadIdx.y; m < M; m += (int64_t)gridDim.y * 8)
      s += to_f32(g[m * ld + ch]);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    part[(int64_t)blockIdx.y * c + ch] = t;
  }
}
__global__ void colsum_final(const float* __restrict__ part, int nparts, int c, float* __restrict__ out,
                             int accumulate, float alpha) {
  int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double s = 0.0;
  for (int k = 0; k < nparts; ++k) s += (double)part[(int64_t)k * c + ch];
  s *= (double)alpha;
  out[ch] = accumulate ? out[ch] + (float)s : (float)s;
}

// OIHW fp32 -> packed [tap][a][b] with (a,b) = (cin,cout) for RSCK or (cout,cin) for RSKC
template <typename T>
__global__ void pack_weights_simt(const float* __restrict__ w, int cout, int cin, int taps, int layout,
                                  T* __restrict__ out) {
  int64_t total = (int64_t)taps * cin * cout;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ci, co, tap;
  if (layout == SRCGAN_WL_RSCK) {
    co = (int)(i % cout); int64_t t = i / cout; ci = (int)(t % cin); tap = (int)(t / cin);
  } else {
    ci = (int)(i % cin); int64_t t = i / cin; co = (int)(t % cout); tap = (int)(t / cout);
  }
  out[i] = from_f32<T>(w[((int64_t)co * cin + ci) * taps + tap]);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static ConvDev to_dev(const srcgan_conv_params* p, bool dgrad) {
  ConvDev a;
  a.n = p->n; a.h = p->h; a.w = p->w; a.cin = p->cin; a.cout = p->cout; a.kh = p->kh; a.kw = p->kw;
  a.stride = p->stride; a.pad = p->pad; a.up = p->upsample; a.ho = p->ho; a.wo = p->wo;
  a.x = p->x; a.x_ld = p->x_ld; a.wgt = p->wgt; a.bias = p->bias; a.y = p->y; a.y_ld = p->y_ld;
  a.act = p->act; a.act_slope = p->act_slope; a.alpha = p->alpha;
  a.r1 = p->r1; a.r1_ld = p->r1_ld; a.beta1 = p->beta1;
  a.r2 = p->r2; a.r2_ld = p->r2_ld; a.beta2 = p->beta2;
  a.mask = p->mask; a.mask_ld = p->mask_ld; a.mask_slope = p->mask_slope;
  const int es = p->dtype == SRCGAN_DT_F32 ? 4 : 2;
  const int kch = dgrad ? p->cout : p->cin, nch = dgrad ? p->cin : p->cout;
  a.vec_a = (kch % 4 == 0) && (p->x_ld % 4 == 0) && (((uintptr_t)p->x) % (4 * es) == 0);
  a.vec_b = (nch % 4 == 0) && (((uintptr_t)p->wgt) % (4 * es) == 0);
  return a;
}

int validate_conv(const srcgan_conv_params* p, bool need_w) {
  SRCGAN_REQUIRE(p != nullptr, "conv: null params");
  SRCGAN_REQUIRE(p->n > 0 && p->h > 0 && p->w > 0 && p->cin > 0 && p->cout > 0, "conv: non-positive dims");
  SRCGAN_REQUIRE(p->kh > 0 && p->kw > 0 && p->stride > 0 && p->pad >= 0, "conv: bad filter geometry");
  SRCGAN_REQUIRE(p->dtype == SRCGAN_DT_F32 || p->dtype == SRCGAN_DT_BF16, "conv: bad dtype %d", p->dtype);
  const int hv = p->upsample ? 2 * p->h : p->h, wv = p->upsample ? 2 * p->w : p->w;
  SRCGAN_REQUIRE((hv + 2 * p->pad - p->kh) / p->stride + 1 == p->ho && (wv + 2 * p->pad - p->kw) / p->stride + 1 == p->wo,
                 "conv: ho/wo (%d,%d) inconsistent with geometry", p->ho, p->wo);
  SRCGAN_REQUIRE(p->x && p->y, "conv: null tensor");
  SRCGAN_REQUIRE(!need_w || p->wgt, "conv: null weights");
  return SRCGAN_OK;
}

template <typename T, bool DGRAD>
static int launch_igemm(const srcgan_conv_params* p, cudaStream_t st) {
  ConvDev a = to_dev(p, DGRAD);
  const int64_t M = (int64_t)p->n * (DGRAD ? p->h * p->w : p->ho * p->wo);
  const int nch = DGRAD ? p->cin : p->cout;
  if (nch <= 4) {
    dim3 grid(ceil_div(M, 256), 1);
    conv_igemm_simt<T, 256, 4, 1, 4, DGRAD><<<grid, 256, 0, st>>>(a);
  } else if (nch <= 32) {
    dim3 grid(ceil_div(M, 128), 1);
    conv_igemm_simt<T, 128, 32, 4, 4, DGRAD><<<grid, 256, 0, st>>>(a);
  } else {
    dim3 grid(ceil_div(M, 128), ceil_div(nch, 64));
    conv_igemm_simt<T, 128, 64, 8, 4, DGRAD><<<grid, 256, 0, st>>>(a);
  }
  count_launch();
  return check_launch(DGRAD ? "conv_dgrad_simt" : "conv_fprop_simt");
}

int conv_fprop_simt(const srcgan_conv_params* p, cudaStream_t st) {
  if (p->dtype == SRCGAN_DT_F32) return launch_igemm<float, false>(p, st);
  return launch_igemm<__nv_bfloat16, false>(p, st);
}
int conv_dgrad_simt(const srcgan_conv_params* p, cudaStream_t st) {
  SRCGAN_REQUIRE(!p->upsample, "conv_dgrad: fused upsample is not supported (use srcgan_upsample2x_adjoint)");
  if (p->dtype == SRCGAN_DT_F32) return launch_igemm<float, true>(p, st);
  return launch_igemm<__nv_bfloat16, true>(p, st);
}

static void wgrad_plan(const srcgan_conv_params* p, int& bi, int& bo, int& splits, int& pix_per_split) {
  bi = p->cin <= 4 ? 4 : 64;
  bo = p->cout <= 4 ? 4 : 64;
  const int tiles = p->kh * p->kw * ceil_div(p->cin, bi) * ceil_div(p->cout, bo);
  const int64_t M = (int64_t)p->n * p->ho * p->wo;
  int want = ceil_div(kNumSMs * 8, tiles);
  int64_t maxs = (M + 255) / 256;
  splits = (int)(want < 1 ? 1 : (want > maxs ? maxs : want));
  pix_per_split = (int)((M + splits - 1) / splits);
  pix_per_split = (pix_per_split + BK - 1) / BK * BK;
  splits = (int)((M + pix_per_split - 1) / pix_per_split);
}

size_t conv_wgrad_simt_workspace(const srcgan_conv_params* p) {
  int bi, bo, splits, pps;
  wgrad_plan(p, bi, bo, splits, pps);
  size_t wbytes = (size_t)splits * p->kh * p->kw * p->cin * p->cout * sizeof(float);
  size_t bbytes = (size_t)1024 * p->cout * sizeof(float);
  return wbytes + bbytes + 256;
}

template <typename T>
static int launch_wgrad(const srcgan_conv_params* p, float* dw, float* db, int accumulate, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  ConvDev a = to_dev(p, false);
  int bi, bo, splits, pps;
  wgrad_plan(p, bi, bo, splits, pps);
  const int taps = p->kh * p->kw;
  const size_t wbytes = (size_t)splits * taps * p->cin * p->cout * sizeof(float);
  SRCGAN_REQUIRE(ws && ws_bytes >= conv_wgrad_simt_workspace(p), "conv_wgrad: workspace too small");
  float* part = reinterpret_cast<float*>(ws);
  const int64_t M = (int64_t)p->n * p->ho * p->wo;
  if (dw) {
    dim3 grid(taps * ceil_div(p->cin, bi) * ceil_div(p->cout, bo), splits);
    if (bi == 64 && bo == 64) conv_wgrad_simt<T, 64, 64, 4, 4><<<grid, 256, 0, st>>>(a, part, pps);
    else if (bi == 64 && bo == 4) conv_wgrad_simt<T, 64, 4, 1, 4><<<grid, 64, 0, st>>>(a, part, pps);
    else if (bi == 4 && bo == 64) conv_wgrad_simt<T, 4, 64, 4, 1><<<grid, 64, 0, st>>>(a, part, pps);
    else conv_wgrad_simt<T, 4, 4, 1, 1><<<grid, 16, 0, st>>>(a, part, pps);
    count_launch();
    int64_t total = (int64_t)taps * p->cin * p->cout;
    wgrad_reduce<<<ceil_div(total, 256), 256, 0, st>>>(part, splits, taps, p->cin, p->cout, dw, accumulate,
                                                           p->alpha);
    count_launch();
  }
  if (db) {
    float* bpart = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ((wbytes + 255) / 256) * 256);
    int ny = (int)((M + 2047) / 2048);
    if (ny > 1024) ny = 1024;
    if (ny < 1) ny = 1;
    dim3 grid(ceil_div(p->cout, 32), ny), blk(32, 8);
    colsum_partial<T><<<grid, blk, 0, st>>>(reinterpret_cast<const T*>(p->y), p->y_ld, M, p->cout, bpart);
    colsum_final<<<ceil_div(p->cout, 128), 128, 0, st>>>(bpart, ny, p->cout, db, accumulate, p->alpha);
    count_launch(2);
  }
  return check_launch("conv_wgrad_simt");
}

int conv_wgrad_simt(const srcgan_conv_params* p, float* dw, float* db, int accumulate, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
  if (p->dtype == SRCGAN_DT_F32) return launch_wgrad<float>(p, dw, db, accumulate, ws, ws_bytes, st);
  return launch_wgrad<__nv_bfloat16>(p, dw, db, accumulate, ws, ws_bytes, st);
}

int wgrad_reduce_launch(const float* part, int splits, int taps, int cin, int cout, float* dw, int accumulate,
                        float alpha, cudaStream_t st) {
  int64_t total = (int64_t)taps * cin * cout;
  wgrad_reduce<<<ceil_div(total, 256), 256, 0, st>>>(part, splits, taps, cin, cout, dw, accumulate, alpha);
  count_launch();
  return check_launch("wgrad_reduce");
}

int bias_grad_launch(const void* dy, int dy_ld, int dtype, long long M, int cout, float* db, int accumulate,
                     float alpha, void* ws, cudaStream_t st) {
  float* bpart = reinterpret_cast<float*>(ws);
  int ny = (int)((M + 2047) / 2048);
  if (ny > 1024) ny = 1024;
  if (ny < 1) ny = 1;
  dim3 grid(ceil_div(cout, 32), ny), blk(32, 8);
  if (dtype == SRCGAN_DT_F32)
    colsum_partial<float><<<grid, blk, 0, st>>>(reinterpret_cast<const float*>(dy), dy_ld, M, cout, bpart);
  else
    colsum_partial<__nv_bfloat16><<<grid, blk, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy), dy_ld, M, cout,
                                                        bpart);
  colsum_final<<<ceil_div(cout, 128), 128, 0, st>>>(bpart, ny, cout, db, accumulate, alpha);
  count_launch(2);
  return check_launch("bias_grad");
}

int pack_weights_simt_host(const float* w, int cout, int cin, int kh, int kw, int layout, int dtype, void* out,
                           cudaStream_t st) {
  int64_t total = (int64_t)kh * kw * cin * cout;
  if (dtype == SRCGAN_DT_F32)
    pack_weights_simt<float><<<ceil_div(total, 256), 256, 0, st>>>(w, cout, cin, kh * kw, layout, (float*)out);
  else
    pack_weights_simt<__nv_bfloat16><<<ceil_div(total, 256), 256, 0, st>>>(w, cout, cin, kh * kw, layout,
                                                                            (__nv_bfloat16*)out);
  count_launch();
  return check_launch("pack_weights");
}

}  // namespace srcgan
