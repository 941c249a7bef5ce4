// SIMT (FFMA, fp32-accumulate) implicit-GEMM convolution engine: fprop, dgrad (gather form of the
// transposed convolution, any stride) and split-K wgrad over channel-sliced NHWC tensors.
// This is the fp32 parity engine (exact fp32 math) and the engine for the layers tensor cores
// cannot help (3-channel image convs, 1-channel patch logits, strided layers).
//
// Reference semantics: nn.Conv2d fwd/bwd as used at src/model/model.py:193-211 (dense block),
// :236-289 (Decoder), :396-440 (RDDBNetB), :612-635 (NLayerDiscriminator).
#include <type_traits>

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

  // Two-level accumulation: each BK = 16 chunk of K is summed in fp32 (FFMA), the chunk sums are added to a double master.
  // A single fp32 accumulator over K = 9 x 192 terms carried 4x the rounding error of the reference's blocked CPU kernels;
  // at full size that flipped 4x as many LeakyReLU gates and doubled the gradient noise against an fp64 evaluation
  // (profiles/r2_fullsize_parity_fp32.json).  This engine is the parity path: accuracy first.
  double acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;

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
    float blk[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) blk[i][j] = 0.f;
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
        for (int j = 0; j < TN; ++j) blk[i][j] = fmaf(av[i], bv[j], blk[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] += (double)blk[i][j];
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
      float v = (float)acc[i][j];
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

  double acc[TI][TO];        // double master over the pixel chunks (see conv_igemm_simt)
#pragma unroll
  for (int i = 0; i < TI; ++i)
#pragma unroll
    for (int j = 0; j < TO; ++j) acc[i][j] = 0.0;

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
    float blk[TI][TO];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
      for (int j = 0; j < TO; ++j) blk[i][j] = 0.f;
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
        for (int j = 0; j < TO; ++j) blk[i][j] = fmaf(xv[i], gv[j], blk[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
      for (int j = 0; j < TO; ++j) acc[i][j] += (double)blk[i][j];
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
      if (co < a.cout) out[(int64_t)ci * a.cout + co] = (float)acc[i][j];
    }
  }
}

// dw[co][ci][tap] (+)= alpha * sum_s part[s][tap][ci][co]   - the split-K reduction behind every wgrad kernel.
// Deterministic (fixed partition, fixed order) and built for latency: a block is 32 outputs x 8 split lanes; lane y adds
// splits y, y+8, ... with four independent accumulators (4 loads in flight per thread, ~1 150 blocks for a 64x64x9 gradient),
// the eight lane sums are added in lane order.  The one-thread-per-output loop it replaces was a chain of up to 148
// dependent load+add steps: 41-46 us per launch, 22 ms of a 299 ms step (profiles/r2_profile_step.txt).
// DBL: double accumulation for the fp32 parity engine.  Two destinations (see srcgan_conv_wgrad_split): output channels
// [0, split) go to d0, the rest to d1; each is an OIHW tensor with `ld` input channels, filled from channel ci0.
template <bool DBL>
__global__ void __launch_bounds__(256)
wgrad_reduce_k(const float* __restrict__ part, int splits, int taps, int cin, int cout, int split, float* __restrict__ d0,
               int ld0, int ci00, float* __restrict__ d1, int ld1, int ci01, int accumulate, float alpha,
               const float* __restrict__ dbpart, float* __restrict__ db0, float* __restrict__ db1, int nred, int ilp8) {
  using Acc = typename std::conditional<DBL, double, float>::type;
  __shared__ Acc red[8][33];
  pdl_launch_dependents();
  pdl_wait();                                                          // the partials are the wgrad kernel's output
  if ((int)blockIdx.x >= nred) {
    // blocks behind the weight-gradient ones: the bias gradients of the same launch, db[c] (+)= alpha * sum_s dbpart[s][c]
    // (what colsum_final does - same partition, same order, double accumulation - without its own launch)
    __shared__ double redd[8][33];
    const int ch = ((int)blockIdx.x - nred) * 32 + (int)threadIdx.x, rl = threadIdx.y;
    double s = 0.0;
    if (ch < cout)
      for (int k = rl; k < splits; k += 8) s += (double)dbpart[(int64_t)k * cout + ch];
    redd[rl][threadIdx.x] = s;
    __syncthreads();
    if (rl == 0 && ch < cout) {
      float* out = ch < split ? db0 : db1;
      if (out != nullptr) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += redd[r][threadIdx.x];
        t *= (double)alpha;
        const int c = ch < split ? ch : ch - split;
        out[c] = accumulate ? out[c] + (float)t : (float)t;
      }
    }
    return;
  }
  const int64_t total = (int64_t)taps * cin * cout;
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int y = threadIdx.y;
  Acc a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  if (i < total) {
    const float* p = part + i;
    int k = y;
    // eight loads in flight per thread while there are that many splits left (148 splits: 3 dependent rounds instead of 5 - the
    // kernel is a chain of L2 latencies, three waves of blocks deep), then four, then one
    for (; ilp8 && k + 56 < splits; k += 64) {
      const float v0 = __ldg(p + (int64_t)k * total), v1 = __ldg(p + (int64_t)(k + 8) * total),
                  v2 = __ldg(p + (int64_t)(k + 16) * total), v3 = __ldg(p + (int64_t)(k + 24) * total),
                  v4 = __ldg(p + (int64_t)(k + 32) * total), v5 = __ldg(p + (int64_t)(k + 40) * total),
                  v6 = __ldg(p + (int64_t)(k + 48) * total), v7 = __ldg(p + (int64_t)(k + 56) * total);
      a0 += (Acc)v0; a1 += (Acc)v1; a2 += (Acc)v2; a3 += (Acc)v3;
      a0 += (Acc)v4; a1 += (Acc)v5; a2 += (Acc)v6; a3 += (Acc)v7;
    }
    for (; k + 24 < splits; k += 32) {
      const float v0 = __ldg(p + (int64_t)k * total), v1 = __ldg(p + (int64_t)(k + 8) * total),
                  v2 = __ldg(p + (int64_t)(k + 16) * total), v3 = __ldg(p + (int64_t)(k + 24) * total);
      a0 += (Acc)v0; a1 += (Acc)v1; a2 += (Acc)v2; a3 += (Acc)v3;
    }
    for (; k < splits; k += 8) a0 += (Acc)__ldg(p + (int64_t)k * total);
  }
  red[y][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (y != 0 || i >= total) return;
  Acc t = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
  const float sv = (float)t * alpha;
  const int co = (int)(i % cout);
  const int64_t q = i / cout;
  const int ci = (int)(q % cin);
  const int tap = (int)(q / cin);
  float* dst = co < split ? d0 : d1;
  if (dst == nullptr) return;
  const int c = co < split ? co : co - split, ld = co < split ? ld0 : ld1, c0 = co < split ? ci00 : ci01;
  const int64_t o = ((int64_t)c * ld + c0 + ci) * taps + tap;
  dst[o] = accumulate ? dst[o] + sv : sv;
}

template <bool DBL>
static void wgrad_reduce_go(const float* part, int splits, int taps, int cin, int cout, int split, float* d0, int ld0, int ci00,
                            float* d1, int ld1, int ci01, int accumulate, float alpha, cudaStream_t st,
                            const float* dbpart = nullptr, float* db0 = nullptr, float* db1 = nullptr) {
  const int64_t total = (int64_t)taps * cin * cout;
  const int nred = ceil_div(total, 32);
  const int ndb = (dbpart && (db0 || db1)) ? ceil_div(cout, 32) : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nred + ndb)); cfg.blockDim = dim3(32, 8); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  cfg.attrs = at; cfg.numAttrs = pdl_attr(&at[0]);
  const int ilp8 = getenv("SRCGAN_B200_REDUCE_ILP4") ? 0 : 1;          // A/B switch: four loads in flight per thread as before
  (void)cudaLaunchKernelEx(&cfg, wgrad_reduce_k<DBL>, part, splits, taps, cin, cout, split, d0, ld0, ci00, d1, ld1, ci01, accumulate,
                           alpha, dbpart, db0, db1, nred, ilp8);
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
// bf16, 16-byte loads: thread = (row lane, channel octet); block-level tree reduce over the row lanes
__global__ void __launch_bounds__(256)
colsum_partial_vec(const __nv_bfloat16* __restrict__ g, int ld, int64_t M, int c, float* __restrict__ part) {
  __shared__ float red[256][9];
  const int octets = c >> 3;
  const int rows_per_iter = 256 / octets;
  const int oc = threadIdx.x % octets, rl = threadIdx.x / octets;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rl < rows_per_iter) {
    // four independent 16-byte loads in flight per thread (the pass is HBM-bound: one load per iteration left it at ~1 TB/s)
    const int64_t step = (int64_t)gridDim.x * rows_per_iter;
    int64_t m = (int64_t)blockIdx.x * rows_per_iter + rl;
    for (; m + 3 * step < M; m += 4 * step) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(g + (m + u * step) * ld + oc * 8));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q[u]);
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[2 * i] += __low2float(h[i]); acc[2 * i + 1] += __high2float(h[i]); }
      }
    }
    for (; m < M; m += step) {
      uint4 q = __ldg(reinterpret_cast<const uint4*>(g + m * ld + oc * 8));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[2 * i] += __low2float(h[i]); acc[2 * i + 1] += __high2float(h[i]); }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = acc[i];
  __syncthreads();
  // thread t < c sums channel t over the row lanes
  if (threadIdx.x < c) {
    const int o = threadIdx.x >> 3, i = threadIdx.x & 7;
    float s = 0.f;
    for (int r = 0; r < rows_per_iter; ++r) s += red[r * octets + o][i];
    part[(int64_t)blockIdx.x * c + threadIdx.x] = s;
  }
}

// block = 32 channels x 8 row lanes; each lane sums every 8th partial row, the lanes are then added in a fixed order
// (deterministic; a single thread per channel walking up to 1024 rows took 50 us per launch)
__global__ void __launch_bounds__(256)
colsum_final(const float* __restrict__ part, int nparts, int c, float* __restrict__ out, int accumulate, float alpha,
             int ld) {
  __shared__ double red[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cl;
  double s = 0.0;
  if (ch < c)
    for (int k = rl; k < nparts; k += 8) s += (double)part[(int64_t)k * ld + ch];
  red[rl][cl] = s;
  __syncthreads();
  if (rl == 0 && ch < c) {
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][cl];
    t *= (double)alpha;
    out[ch] = accumulate ? out[ch] + (float)t : (float)t;
  }
}

// ---------------------------------------------------------------------------------------------
// thin fprop / dgrad (bandwidth-bound): one side of the GEMM has <= 4 channels.
//   thin_out_conv : <= 4 output channels (64->3 image conv, 256->1 patch logits, dgrad of 3->64):
//                   one thread per output pixel, 4-wide vector loads of the wide source, weights
//                   [K][4] in shared memory read as warp-uniform float4 broadcasts
//   thin_in_conv  : <= 4 GEMM-K channels (3->64 image conv, dgrad of 64->3 / 256->1): one thread per
//                   (output pixel, 8 output channels), weights [K][N] in shared memory, 16/32-byte stores
// Pixel mapping (fprop vs gather-form dgrad, stride, upsample) is the same as in conv_igemm_simt.
// ---------------------------------------------------------------------------------------------
template <bool DGRAD>
__device__ __forceinline__ int64_t thin_src_pixel(const ConvDev& a, int n, int oy, int ox, int fr, int fs, int SH,
                                                  int SW, int HV, int WV) {
  int sy, sx;
  if (!DGRAD) {
    int iy = oy * a.stride - a.pad + fr, ix = ox * a.stride - a.pad + fs;
    if (iy < 0 || iy >= HV || ix < 0 || ix >= WV) return -1;
    sy = a.up ? (iy >> 1) : iy; sx = a.up ? (ix >> 1) : ix;
  } else {
    int ty_ = oy + a.pad - fr, tx_ = ox + a.pad - fs;
    if (ty_ < 0 || tx_ < 0) return -1;
    if (a.stride > 1) {
      if ((ty_ % a.stride) | (tx_ % a.stride)) return -1;
      ty_ /= a.stride; tx_ /= a.stride;
    }
    if (ty_ >= SH || tx_ >= SW) return -1;
    sy = ty_; sx = tx_;
  }
  return (((int64_t)n * SH + sy) * SW + sx) * a.x_ld;
}

template <typename T, bool DGRAD>
__global__ void __launch_bounds__(256)
thin_out_conv(const ConvDev a) {
  extern __shared__ float4 wsm[];                 // [taps*kch] float4 (N padded to 4)
  const int OH = DGRAD ? a.h : a.ho, OW = DGRAD ? a.w : a.wo;
  const int64_t M = (int64_t)a.n * OH * OW;
  const int kch = DGRAD ? a.cout : a.cin, nch = DGRAD ? a.cin : a.cout;
  const int taps = a.kh * a.kw, K = taps * kch;
  const T* __restrict__ Wt = reinterpret_cast<const T*>(a.wgt);
  for (int k = threadIdx.x; k < K; k += 256) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = to_f32(Wt[(int64_t)k * nch]);
    if (nch > 1) v.y = to_f32(Wt[(int64_t)k * nch + 1]);
    if (nch > 2) v.z = to_f32(Wt[(int64_t)k * nch + 2]);
    if (nch > 3) v.w = to_f32(Wt[(int64_t)k * nch + 3]);
    wsm[k] = v;
  }
  __syncthreads();
  const int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (m >= M) return;
  const int ox = (int)(m % OW);
  const int64_t t = m / OW;
  const int oy = (int)(t % OH), n = (int)(t / OH);
  const int SH = DGRAD ? a.ho : a.h, SW = DGRAD ? a.wo : a.w;
  const int HV = a.up ? 2 * a.h : a.h, WV = a.up ? 2 * a.w : a.w;
  const T* __restrict__ X = reinterpret_cast<const T*>(a.x);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int fr = 0; fr < a.kh; ++fr)
    for (int fs = 0; fs < a.kw; ++fs) {
      const int64_t off = thin_src_pixel<DGRAD>(a, n, oy, ox, fr, fs, SH, SW, HV, WV);
      if (off < 0) continue;
      const float4* wrow = wsm + (fr * a.kw + fs) * kch;
      const T* src = X + off;
#pragma unroll 4
      for (int c = 0; c < kch; c += 4) {
        float v[4];
        load4<T>(src + c, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = wrow[c + q];
          acc[0] = fmaf(v[q], w4.x, acc[0]); acc[1] = fmaf(v[q], w4.y, acc[1]);
          acc[2] = fmaf(v[q], w4.z, acc[2]); acc[3] = fmaf(v[q], w4.w, acc[3]);
        }
      }
    }
  T* __restrict__ Y = reinterpret_cast<T*>(a.y);
  const T* R1 = reinterpret_cast<const T*>(a.r1);
  const T* R2 = reinterpret_cast<const T*>(a.r2);
  const T* MK = reinterpret_cast<const T*>(a.mask);
  for (int c = 0; c < nch; ++c) {
    float v = acc[c];
    if (a.bias) v += a.bias[c];
    if (a.act) v = v > 0.f ? v : v * a.act_slope;
    v *= a.alpha;
    if (R1) v = fmaf(a.beta1, to_f32(R1[m * a.r1_ld + c]), v);
    if (R2) v = fmaf(a.beta2, to_f32(R2[m * a.r2_ld + c]), v);
    if (MK) v *= (to_f32(MK[m * a.mask_ld + c]) > 0.f ? 1.f : a.mask_slope);
    Y[m * a.y_ld + c] = from_f32<T>(v);
  }
}

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 o;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}

template <typename T, bool DGRAD>
__global__ void __launch_bounds__(256)
thin_in_conv(const ConvDev a) {
  extern __shared__ float wsf[];                  // [taps*kch][nch] fp32
  const int OH = DGRAD ? a.h : a.ho, OW = DGRAD ? a.w : a.wo;
  const int64_t M = (int64_t)a.n * OH * OW;
  const int kch = DGRAD ? a.cout : a.cin, nch = DGRAD ? a.cin : a.cout;
  const int taps = a.kh * a.kw, K = taps * kch;
  const T* __restrict__ Wt = reinterpret_cast<const T*>(a.wgt);
  for (int i = threadIdx.x; i < K * nch; i += 256) wsf[i] = to_f32(Wt[i]);
  __syncthreads();
  const int octs = nch >> 3;
  const int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (gid >= M * octs) return;
  const int c0 = (int)(gid % octs) * 8;
  const int64_t m = gid / octs;
  const int ox = (int)(m % OW);
  const int64_t t = m / OW;
  const int oy = (int)(t % OH), n = (int)(t / OH);
  const int SH = DGRAD ? a.ho : a.h, SW = DGRAD ? a.wo : a.w;
  const int HV = a.up ? 2 * a.h : a.h, WV = a.up ? 2 * a.w : a.w;
  const T* __restrict__ X = reinterpret_cast<const T*>(a.x);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int fr = 0; fr < a.kh; ++fr)
    for (int fs = 0; fs < a.kw; ++fs) {
      const int64_t off = thin_src_pixel<DGRAD>(a, n, oy, ox, fr, fs, SH, SW, HV, WV);
      if (off < 0) continue;
      for (int k = 0; k < kch; ++k) {
        const float xv = to_f32(X[off + k]);
        const float* wr = wsf + ((fr * a.kw + fs) * kch + k) * nch + c0;
        const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + 4);
        acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]); acc[2] = fmaf(xv, w0.z, acc[2]);
        acc[3] = fmaf(xv, w0.w, acc[3]); acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
        acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
      }
    }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v = acc[i];
    if (a.bias) v += a.bias[c0 + i];
    if (a.act) v = v > 0.f ? v : v * a.act_slope;
    acc[i] = v * a.alpha;
  }
  if (a.r1) {
    float r[8];
    load8<T>(reinterpret_cast<const T*>(a.r1) + m * a.r1_ld + c0, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(a.beta1, r[i], acc[i]);
  }
  if (a.r2) {
    float r[8];
    load8<T>(reinterpret_cast<const T*>(a.r2) + m * a.r2_ld + c0, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(a.beta2, r[i], acc[i]);
  }
  if (a.mask) {
    float r[8];
    load8<T>(reinterpret_cast<const T*>(a.mask) + m * a.mask_ld + c0, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= (r[i] > 0.f ? 1.f : a.mask_slope);
  }
  store8<T>(reinterpret_cast<T*>(a.y) + m * a.y_ld + c0, acc);
}

// ---------------------------------------------------------------------------------------------
// thin_in_tiled : <= 4 GEMM-K channels and a multiple of 64 output channels (first discriminator layer 3 -> 64 4x4 stride 2,
// ResDeconv's 7x7 stride-2 stem, gradient of the 256 -> 1 patch-logit layer), FFMA-bound instead of load-bound.
// thin_in_conv above issues one scalar global load and two shared-memory loads per 8 FMAs and recomputes the source
// pixel per tap (7x7 stem at batch 64: 2.3 ms for 9.9 GFMA).  Here a block owns a tile of 8 x 32 output pixels x 64 channels:
//   * the haloed input patch of the tile sits in shared memory as fp32 planes [k channel][row][col] (zero outside the image),
//     the layer's weights as [tap * kch][64] fp32, loaded once per block (persistent over tiles);
//   * a warp = 32 columns x one 16-channel group x 4 rows: per (tap, k channel) a thread reads 4 patch values (consecutive
//     lanes -> consecutive or every-second bank) and 4 x LDS.128 of weights (warp-uniform broadcast) for 64 FMAs.
// Accumulation order per output = (filter row, filter column, k channel), exactly thin_in_conv's; taps outside the image
// contribute fma(0, w, acc) = acc.  DGRAD (gather form) is supported for stride 1: tap (fr, fs) reads the patch mirrored.
// ---------------------------------------------------------------------------------------------
constexpr int TI_TH = 8, TI_TW = 32;

template <typename T, bool DGRAD>
__global__ void __launch_bounds__(256, 2)
thin_in_tiled(const ConvDev a, int tiles_x, int tiles_y, long long num_tiles, int ph, int pw) {
  extern __shared__ float ti_smem[];
  const int kch = DGRAD ? a.cout : a.cin, nch = DGRAD ? a.cin : a.cout;
  const int OH = DGRAD ? a.h : a.ho, OW = DGRAD ? a.w : a.wo;          // output of this launch
  const int SH = DGRAD ? a.ho : a.h, SW = DGRAD ? a.wo : a.w;          // its source
  const int taps = a.kh * a.kw, K = taps * kch;
  const int s = DGRAD ? 1 : a.stride;
  float* wsm = ti_smem;                                                // [K][64]
  float* xs = ti_smem + K * 64;                                        // [kch][ph][pw]
  const int n0 = blockIdx.y * 64;
  const T* __restrict__ Wt = reinterpret_cast<const T*>(a.wgt);
  for (int i = threadIdx.x; i < K * 64; i += 256) wsm[i] = to_f32(Wt[(long long)(i >> 6) * nch + n0 + (i & 63)]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = (warp & 3) * 16, r0 = (warp >> 2) * 4;
  const T* __restrict__ X = reinterpret_cast<const T*>(a.x);
  // patch origin in source coordinates for output tile origin (oy0, ox0): fprop oy0*s - pad; dgrad oy0 + pad - (kh - 1)
  const int org_y = DGRAD ? a.pad - (a.kh - 1) : -a.pad, org_x = DGRAD ? a.pad - (a.kw - 1) : -a.pad;
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    long long r = tile;
    const int bx = (int)(r % tiles_x); r /= tiles_x;
    const int by = (int)(r % tiles_y);
    const int img = (int)(r / tiles_y);
    const int oy0 = by * TI_TH, ox0 = bx * TI_TW;
    const int sy0 = oy0 * s + org_y, sx0 = ox0 * s + org_x;
    __syncthreads();                                                   // previous tile's readers are done (and wsm is written)
    for (int e = threadIdx.x; e < ph * pw; e += 256) {
      const int ry = e / pw, rx = e - ry * pw;
      const int gy = sy0 + ry, gx = sx0 + rx;
      const bool in = gy >= 0 && gy < SH && gx >= 0 && gx < SW;
      const T* src = X + (((long long)img * SH + (in ? gy : 0)) * SW + (in ? gx : 0)) * a.x_ld;
      for (int k = 0; k < kch; ++k) xs[(k * ph + ry) * pw + rx] = in ? to_f32(src[k]) : 0.f;
    }
    __syncthreads();
    float acc[4][16];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[p][c] = 0.f;
    for (int fr = 0; fr < a.kh; ++fr) {
      const int pr = DGRAD ? a.kh - 1 - fr : fr;
      for (int fs = 0; fs < a.kw; ++fs) {
        const int pc = (DGRAD ? a.kw - 1 - fs : fs) + lane * s;
        const float* wrow = wsm + ((fr * a.kw + fs) * kch) * 64 + cg;
        for (int k = 0; k < kch; ++k) {
          const float* xp = xs + (k * ph + r0 * s + pr) * pw + pc;
          float xv[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) xv[p] = xp[p * s * pw];
          const float4 w0 = *reinterpret_cast<const float4*>(wrow + k * 64);
          const float4 w1 = *reinterpret_cast<const float4*>(wrow + k * 64 + 4);
          const float4 w2 = *reinterpret_cast<const float4*>(wrow + k * 64 + 8);
          const float4 w3 = *reinterpret_cast<const float4*>(wrow + k * 64 + 12);
          const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[p][c] = fmaf(xv[p], wv[c], acc[p][c]);
        }
      }
    }
    const int ox = ox0 + lane;
    if (ox < OW) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int oy = oy0 + r0 + p;
        if (oy >= OH) continue;
        const long long m = ((long long)img * OH + oy) * OW + ox;
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          const int c0 = n0 + cg + h8 * 8;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float t = acc[p][h8 * 8 + i];
            if (a.bias) t += a.bias[c0 + i];
            if (a.act) t = t > 0.f ? t : t * a.act_slope;
            v[i] = t * a.alpha;
          }
          if (a.r1) {
            float rr[8];
            load8<T>(reinterpret_cast<const T*>(a.r1) + m * a.r1_ld + c0, rr);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(a.beta1, rr[i], v[i]);
          }
          if (a.r2) {
            float rr[8];
            load8<T>(reinterpret_cast<const T*>(a.r2) + m * a.r2_ld + c0, rr);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(a.beta2, rr[i], v[i]);
          }
          if (a.mask) {
            float rr[8];
            load8<T>(reinterpret_cast<const T*>(a.mask) + m * a.mask_ld + c0, rr);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= (rr[i] > 0.f ? 1.f : a.mask_slope);
          }
          store8<T>(reinterpret_cast<T*>(a.y) + m * a.y_ld + c0, v);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// thin wgrad: one side of the convolution has <= 4 channels (image convs 3->64 / 64->3, patch logits
// 256->1).  Bandwidth-bound: every element of the wide tensor is read exactly once, by the thread that
// owns its channel; the thin tensor's haloed tile sits in shared memory as float4 per pixel and is
// read as warp-uniform broadcasts.  dW[tap][c][k] = sum_q wide[q][c] * thin[ts*q + sign*(tap-pad)][k]
//   case A (thin = dY, wide = X, stride 1):  sign = -1, ts = 1
//   case B (thin = X, wide = dY, stride s):  sign = +1, ts = s
// ---------------------------------------------------------------------------------------------
struct ThinArgs {
  const void* wide; int wide_ld; int cw;          // wide tensor (n, hw, ww, cw)
  const void* thin; int thin_ld; int kt;          // thin tensor (n, ht, wt, kt<=4)
  int n, hw, ww, ht, wt;
  int ts, sign, pad;
  int tiles_x, tiles_y;
  long long num_tiles;
};
constexpr int THIN_TH = 8, THIN_TW = 32;

// FRN = filter rows handled by one launch, starting at fr_begin: a 7x7 filter (ResDeconv's stem, src/model/resdeconv.py) would
// need 49 x 4 accumulators per thread, so it runs as seven launches of one filter row each (the wide tensor is re-read seven
// times: 0.9 GB at batch 64, against 9 ms per launch of the generic FFMA wgrad it replaces).
template <typename T, int KH, int KW, int FRN>
__global__ void __launch_bounds__(256)
thin_wgrad_kernel(const ThinArgs a, float* __restrict__ part, int fr_begin) {
  constexpr int TAPS = KH * KW, NACC = FRN * KW;
  constexpr int TTH = 2 * (THIN_TH - 1) + KH, TTW = 2 * (THIN_TW - 1) + KW;   // sized for ts = 2
  __shared__ float4 st[TTH][TTW];
  const int t = threadIdx.x, c = t & 63, g = t >> 6;
  const int c0 = blockIdx.y * 64;
  const T* __restrict__ W = reinterpret_cast<const T*>(a.wide);
  const T* __restrict__ Th = reinterpret_cast<const T*>(a.thin);
  const int tth = a.ts * (THIN_TH - 1) + KH, ttw = a.ts * (THIN_TW - 1) + KW;
  const int omin_y = a.sign > 0 ? -a.pad : -(KH - 1 - a.pad);
  const int omin_x = a.sign > 0 ? -a.pad : -(KW - 1 - a.pad);
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

  for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    long long r = tile;
    const int bx = (int)(r % a.tiles_x); r /= a.tiles_x;
    const int by = (int)(r % a.tiles_y);
    const int img = (int)(r / a.tiles_y);
    const int y0 = by * THIN_TH, x0 = bx * THIN_TW;
    const int gy0 = a.ts * y0 + omin_y, gx0 = a.ts * x0 + omin_x;
    __syncthreads();
    for (int e = t; e < tth * ttw; e += 256) {
      int ry = e / ttw, rx = e - ry * ttw;
      int gy = gy0 + ry, gx = gx0 + rx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < a.ht && gx >= 0 && gx < a.wt) {
        const T* src = Th + (((long long)img * a.ht + gy) * a.wt + gx) * a.thin_ld;
        v.x = to_f32(src[0]);
        if (a.kt > 1) v.y = to_f32(src[1]);
        if (a.kt > 2) v.z = to_f32(src[2]);
        if (a.kt > 3) v.w = to_f32(src[3]);
      }
      st[ry][rx] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int rr = 0; rr < 2; ++rr) {
      const int ly = 2 * g + rr, y = y0 + ly;
      if (y >= a.hw) continue;
      const T* wrow = W + (((long long)img * a.hw + y) * a.ww) * a.wide_ld + c0 + c;
#pragma unroll 4
      for (int lx = 0; lx < THIN_TW; ++lx) {
        const int x = x0 + lx;
        const float wv = x < a.ww ? to_f32(wrow[(long long)x * a.wide_ld]) : 0.f;
#pragma unroll
        for (int f = 0; f < FRN; ++f) {
          const int fr = fr_begin + f;
          const int ry = a.ts * ly + (a.sign > 0 ? fr : KH - 1 - fr);
#pragma unroll
          for (int fs = 0; fs < KW; ++fs) {
            const int rx = a.ts * lx + (a.sign > 0 ? fs : KW - 1 - fs);
            const float4 tv = st[ry][rx];
            float* ac = acc[f * KW + fs];
            ac[0] = fmaf(wv, tv.x, ac[0]); ac[1] = fmaf(wv, tv.y, ac[1]);
            ac[2] = fmaf(wv, tv.z, ac[2]); ac[3] = fmaf(wv, tv.w, ac[3]);
          }
        }
      }
    }
  }
  // reduce the four row-pair subgroups through shared memory (reusing the thin tile), then one partial per block:
  // part[block][tap][cw][4]
  __syncthreads();
  float4* red = &st[0][0];                       // TTH*TTW float4 >= 3*64 entries
  float4* out = reinterpret_cast<float4*>(part) + ((long long)blockIdx.x * TAPS + fr_begin * KW) * a.cw + c0 + c;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    if (g > 0) red[(g - 1) * 64 + c] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    if (g == 0) {
      float4 s = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float4 v = red[k * 64 + c];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      out[(long long)i * a.cw] = s;
    }
    __syncthreads();
  }
}

// 7x7 filters (ResDeconv's stem, case B) in ONE launch: the filter row is a thread dimension - a block is 64 wide channels x 7
// filter rows x 2 groups of four tile rows, a thread keeps the 7 x 4 accumulators of its filter row.  Same tile and partial
// layout as seven thin_wgrad_kernel<T, 7, 7, 1> launches (3.7 ms at batch 64 for 13 GFMA: each re-read the wide tensor with
// 16 warps per SM; SRCGAN_B200_THIN_WGRAD7_ROWS=1 brings them back); a thread sums four tile rows instead of two, so the fp32
// results differ in the last bits.
template <typename T>
__global__ void __launch_bounds__(896)
thin_wgrad7_kernel(const ThinArgs a, float* __restrict__ part) {
  constexpr int K7 = 7, TAPS = 49;
  constexpr int TTH = 2 * (THIN_TH - 1) + K7, TTW = 2 * (THIN_TW - 1) + K7;   // sized for ts = 2
  __shared__ float4 st[TTH][TTW];
  const int t = threadIdx.x, c = t & 63, fr = (t >> 6) % K7, g = t / (64 * K7);
  const int c0 = blockIdx.y * 64;
  const T* __restrict__ W = reinterpret_cast<const T*>(a.wide);
  const T* __restrict__ Th = reinterpret_cast<const T*>(a.thin);
  const int tth = a.ts * (THIN_TH - 1) + K7, ttw = a.ts * (THIN_TW - 1) + K7;
  const int omin_y = -a.pad, omin_x = -a.pad;                           // case B: sign = +1
  float acc[K7][4];
#pragma unroll
  for (int i = 0; i < K7; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

  for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    long long r = tile;
    const int bx = (int)(r % a.tiles_x); r /= a.tiles_x;
    const int by = (int)(r % a.tiles_y);
    const int img = (int)(r / a.tiles_y);
    const int y0 = by * THIN_TH, x0 = bx * THIN_TW;
    const int gy0 = a.ts * y0 + omin_y, gx0 = a.ts * x0 + omin_x;
    __syncthreads();
    for (int e = t; e < tth * ttw; e += 896) {
      int ry = e / ttw, rx = e - ry * ttw;
      int gy = gy0 + ry, gx = gx0 + rx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < a.ht && gx >= 0 && gx < a.wt) {
        const T* src = Th + (((long long)img * a.ht + gy) * a.wt + gx) * a.thin_ld;
        v.x = to_f32(src[0]);
        if (a.kt > 1) v.y = to_f32(src[1]);
        if (a.kt > 2) v.z = to_f32(src[2]);
        if (a.kt > 3) v.w = to_f32(src[3]);
      }
      st[ry][rx] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int rr = 0; rr < 4; ++rr) {
      const int ly = 4 * g + rr, y = y0 + ly;
      if (y >= a.hw) continue;
      const T* wrow = W + (((long long)img * a.hw + y) * a.ww) * a.wide_ld + c0 + c;
      const float4* trow = &st[a.ts * ly + fr][0];
#pragma unroll 4
      for (int lx = 0; lx < THIN_TW; ++lx) {
        const int x = x0 + lx;
        const float wv = x < a.ww ? to_f32(wrow[(long long)x * a.wide_ld]) : 0.f;
#pragma unroll
        for (int fs = 0; fs < K7; ++fs) {
          const float4 tv = trow[a.ts * lx + fs];
          acc[fs][0] = fmaf(wv, tv.x, acc[fs][0]); acc[fs][1] = fmaf(wv, tv.y, acc[fs][1]);
          acc[fs][2] = fmaf(wv, tv.z, acc[fs][2]); acc[fs][3] = fmaf(wv, tv.w, acc[fs][3]);
        }
      }
    }
  }
  // fold the two row groups through shared memory (reusing the thin tile), one partial per block: part[block][tap][cw][4]
  __syncthreads();
  float4* red = &st[0][0];                       // TTH*TTW float4 >= 7*64 entries
  float4* out = reinterpret_cast<float4*>(part) + ((long long)blockIdx.x * TAPS + fr * K7) * a.cw + c0 + c;
#pragma unroll
  for (int i = 0; i < K7; ++i) {
    if (g > 0) red[fr * 64 + c] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    if (g == 0) {
      const float4 v = red[fr * 64 + c];
      out[(long long)i * a.cw] = make_float4(acc[i][0] + v.x, acc[i][1] + v.y, acc[i][2] + v.z, acc[i][3] + v.w);
    }
    __syncthreads();
  }
}

// dw (OIHW) (+)= alpha * sum_parts part[p][tap][c][k];  thin_is_out: k indexes cout (case A) else cin (case B)
__global__ void thin_wgrad_reduce(const float* __restrict__ part, int nparts, int taps, int cw, int kt, int thin_is_out,
                                  float* __restrict__ dw, int accumulate, float alpha) {
  long long total = (long long)taps * cw * kt;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int k = (int)(i % kt);
  long long r = i / kt;
  int c = (int)(r % cw);
  int tap = (int)(r / cw);
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(((long long)p * taps + tap) * cw + c) * 4 + k];
  s *= alpha;
  long long o = thin_is_out ? ((long long)k * cw + c) * taps + tap : ((long long)c * kt + k) * taps + tap;
  dw[o] = accumulate ? dw[o] + s : s;
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
  const int kch = DGRAD ? p->cout : p->cin;
  const int es = (int)sizeof(T);
  auto al = [&](const void* ptr, int ld) { return ptr == nullptr || (((uintptr_t)ptr) % (8 * es) == 0 && ld % 8 == 0); };
  if (nch <= 4 && kch % 4 == 0 && a.vec_a && (size_t)p->kh * p->kw * kch * 16 <= 160 * 1024) {
    const size_t smem = (size_t)p->kh * p->kw * kch * sizeof(float4);
    static DeviceOnce attr_done;
    int attr_done_dev;
    if (attr_done.needed(&attr_done_dev)) {
      cudaFuncSetAttribute(thin_out_conv<T, DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      attr_done.mark(attr_done_dev);
    }
    thin_out_conv<T, DGRAD><<<ceil_div(M, 256), 256, smem, st>>>(a);
  } else if (kch <= 4 && nch % 64 == 0 && !p->upsample && (!DGRAD || p->stride == 1) && al(p->y, p->y_ld) &&
             al(p->r1, p->r1_ld) && al(p->r2, p->r2_ld) && al(p->mask, p->mask_ld) && p->kh * p->kw * kch <= 256 &&
             !getenv("SRCGAN_B200_NO_THIN_TILED")) {
    const int s = DGRAD ? 1 : p->stride;
    const int ph = (TI_TH - 1) * s + p->kh, pw = ((TI_TW - 1) * s + p->kw) | 1;     // odd row pitch: rows land on different banks
    const size_t smem = ((size_t)p->kh * p->kw * kch * 64 + (size_t)kch * ph * pw) * sizeof(float);
    static DeviceOnce attr_done3;
    int attr_done3_dev;
    if (attr_done3.needed(&attr_done3_dev)) {
      cudaFuncSetAttribute(thin_in_tiled<T, DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      attr_done3.mark(attr_done3_dev);
    }
    SRCGAN_REQUIRE(smem <= 100 * 1024, "thin_in_tiled: %zu bytes of shared memory", smem);
    const int OH = DGRAD ? p->h : p->ho, OW = DGRAD ? p->w : p->wo;
    const int tx = (OW + TI_TW - 1) / TI_TW, ty = (OH + TI_TH - 1) / TI_TH;
    const long long tiles = (long long)tx * ty * p->n;
    const int ngroups = nch / 64;
    long long gx = (2LL * kNumSMs + ngroups - 1) / ngroups;
    if (gx > tiles) gx = tiles;
    thin_in_tiled<T, DGRAD><<<dim3((unsigned)gx, (unsigned)ngroups), 256, smem, st>>>(a, tx, ty, tiles, ph, pw);
  } else if (kch <= 4 && nch % 8 == 0 && al(p->y, p->y_ld) && al(p->r1, p->r1_ld) && al(p->r2, p->r2_ld) &&
             al(p->mask, p->mask_ld) && (size_t)p->kh * p->kw * kch * nch * 4 <= 160 * 1024) {
    const size_t smem = (size_t)p->kh * p->kw * kch * nch * sizeof(float);
    static DeviceOnce attr_done2;
    int attr_done2_dev;
    if (attr_done2.needed(&attr_done2_dev)) {
      cudaFuncSetAttribute(thin_in_conv<T, DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      attr_done2.mark(attr_done2_dev);
    }
    thin_in_conv<T, DGRAD><<<ceil_div(M * (nch / 8), 256), 256, smem, st>>>(a);
  } else if (nch <= 4) {
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

int bias_grad_launch(const void* dy, int dy_ld, int dtype, long long M, int cout, float* db, int accumulate,
                     float alpha, void* ws, cudaStream_t st);

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

static bool thin_wgrad_ok(const srcgan_conv_params* p) {
  if (p->upsample || p->kh != p->kw || (p->kh != 3 && p->kh != 4 && p->kh != 7)) return false;
  if (p->kh == 7) return p->cin <= 4 && p->cout % 64 == 0 && (p->stride == 1 || p->stride == 2);   // case B only
  if (p->cout <= 4 && p->cin % 64 == 0 && p->stride == 1) return true;                       // case A
  if (p->cin <= 4 && p->cout % 64 == 0 && (p->stride == 1 || p->stride == 2)) return true;   // case B
  return false;
}
static int thin_grid(const srcgan_conv_params* p, long long& tiles, int& tx, int& ty) {
  const bool case_a = p->cout <= 4;
  const int hw = case_a ? p->h : p->ho, ww = case_a ? p->w : p->wo;
  tx = (ww + THIN_TW - 1) / THIN_TW; ty = (hw + THIN_TH - 1) / THIN_TH;
  tiles = (long long)tx * ty * p->n;
  return (int)(tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs);
}

size_t conv_wgrad_simt_workspace(const srcgan_conv_params* p) {
  size_t bbytes = (size_t)1024 * p->cout * sizeof(float);
  if (thin_wgrad_ok(p)) {
    long long tiles; int tx, ty;
    int grid = thin_grid(p, tiles, tx, ty);
    const int cw = p->cout <= 4 ? p->cin : p->cout;
    size_t wbytes = (size_t)grid * 4 * p->kh * p->kw * cw * 4 * sizeof(float);
    return ((wbytes + 255) / 256) * 256 + bbytes + 256;
  }
  int bi, bo, splits, pps;
  wgrad_plan(p, bi, bo, splits, pps);
  size_t wbytes = (size_t)splits * p->kh * p->kw * p->cin * p->cout * sizeof(float);
  return ((wbytes + 255) / 256) * 256 + bbytes + 256;
}

template <typename T>
static int launch_thin_wgrad(const srcgan_conv_params* p, float* dw, int accumulate, void* ws, cudaStream_t st) {
  const bool case_a = p->cout <= 4;
  ThinArgs a;
  a.n = p->n;
  if (case_a) {
    a.wide = p->x; a.wide_ld = p->x_ld; a.cw = p->cin; a.hw = p->h; a.ww = p->w;
    a.thin = p->y; a.thin_ld = p->y_ld; a.kt = p->cout; a.ht = p->ho; a.wt = p->wo;
    a.ts = 1; a.sign = -1;
  } else {
    a.wide = p->y; a.wide_ld = p->y_ld; a.cw = p->cout; a.hw = p->ho; a.ww = p->wo;
    a.thin = p->x; a.thin_ld = p->x_ld; a.kt = p->cin; a.ht = p->h; a.wt = p->w;
    a.ts = p->stride; a.sign = 1;
  }
  a.pad = p->pad;
  int grid = thin_grid(p, a.num_tiles, a.tiles_x, a.tiles_y);
  float* part = reinterpret_cast<float*>(ws);
  dim3 g(grid, a.cw / 64);
  int nk = 1;
  if (p->kh == 3) thin_wgrad_kernel<T, 3, 3, 3><<<g, 256, 0, st>>>(a, part, 0);
  else if (p->kh == 4) thin_wgrad_kernel<T, 4, 4, 4><<<g, 256, 0, st>>>(a, part, 0);
  else if (!getenv("SRCGAN_B200_THIN_WGRAD7_ROWS")) {
    thin_wgrad7_kernel<T><<<g, 896, 0, st>>>(a, part);
  } else {
    nk = 7;
    for (int fr = 0; fr < 7; ++fr) thin_wgrad_kernel<T, 7, 7, 1><<<g, 256, 0, st>>>(a, part, fr);
  }
  count_launch(nk - 1);
  const int taps = p->kh * p->kw;
  long long total = (long long)taps * a.cw * a.kt;
  thin_wgrad_reduce<<<ceil_div(total, 128), 128, 0, st>>>(part, grid, taps, a.cw, a.kt, case_a ? 1 : 0, dw,
                                                          accumulate, p->alpha);
  count_launch(2);
  return check_launch("thin_wgrad");
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
  if (dw && thin_wgrad_ok(p)) {
    int rc = launch_thin_wgrad<T>(p, dw, accumulate, ws, st);
    if (rc) return rc;
    long long tiles; int tx_, ty_;
    int tg = thin_grid(p, tiles, tx_, ty_);
    const int cw = p->cout <= 4 ? p->cin : p->cout;
    size_t tb = (size_t)tg * 4 * taps * cw * 4 * sizeof(float);
    if (db) {
      int rc2 = bias_grad_launch(p->y, p->y_ld, p->dtype, M, p->cout, db, accumulate, p->alpha,
                                 reinterpret_cast<char*>(ws) + ((tb + 255) / 256) * 256, st);
      if (rc2) return rc2;
    }
    return SRCGAN_OK;
  }
  if (dw) {
    dim3 grid(taps * ceil_div(p->cin, bi) * ceil_div(p->cout, bo), splits);
    if (bi == 64 && bo == 64) conv_wgrad_simt<T, 64, 64, 4, 4><<<grid, 256, 0, st>>>(a, part, pps);
    else if (bi == 64 && bo == 4) conv_wgrad_simt<T, 64, 4, 1, 4><<<grid, 64, 0, st>>>(a, part, pps);
    else if (bi == 4 && bo == 64) conv_wgrad_simt<T, 4, 64, 4, 1><<<grid, 64, 0, st>>>(a, part, pps);
    else conv_wgrad_simt<T, 4, 4, 1, 1><<<grid, 16, 0, st>>>(a, part, pps);
    count_launch();
    int64_t total = (int64_t)taps * p->cin * p->cout;
    wgrad_reduce_go<true>(part, splits, taps, p->cin, p->cout, p->cout, dw, p->cin, 0, nullptr, 0, 0, accumulate, p->alpha, st);
    count_launch();
  }
  if (db) {
    int rc = bias_grad_launch(p->y, p->y_ld, p->dtype, M, p->cout, db, accumulate, p->alpha,
                              reinterpret_cast<char*>(ws) + ((wbytes + 255) / 256) * 256, st);
    if (rc) return rc;
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
  wgrad_reduce_go<false>(part, splits, taps, cin, cout, cout, dw, cin, 0, nullptr, 0, 0, accumulate, alpha, st);   // tcgen05 partials: float
  count_launch();
  return check_launch("wgrad_reduce");
}

// fixed-order sum of `nparts` partial rows [nparts][c] -> out[c] (* alpha, optionally accumulated)
int colsum_final_launch(const float* part, int nparts, int c, float* out, int accumulate, float alpha, cudaStream_t st) {
  colsum_final<<<ceil_div(c, 32), 256, 0, st>>>(part, nparts, c, out, accumulate, alpha, c);
  count_launch();
  return check_launch("colsum_final");
}

// channels [0, c) of partial rows that are `ld` floats apart
int colsum_final_ld_launch(const float* part, int nparts, int c, int ld, float* out, int accumulate, float alpha, cudaStream_t st) {
  colsum_final<<<ceil_div(c, 32), 256, 0, st>>>(part, nparts, c, out, accumulate, alpha, ld);
  count_launch();
  return check_launch("colsum_final");
}

int wgrad_reduce_split_launch(const float* part, int splits, int taps, int cin, int cout, int split, float* d0, int ld0,
                              int ci00, float* d1, int ld1, int ci01, int accumulate, float alpha, cudaStream_t st) {
  wgrad_reduce_go<false>(part, splits, taps, cin, cout, split, d0, ld0, ci00, d1, ld1, ci01, accumulate, alpha, st);
  count_launch();
  return check_launch("wgrad_reduce");
}

// the same launch also finishes the bias gradients the kw-stacked wgrad kernel summed per split: dbpart [splits][cout] ->
// db0 (channels [0, split)) / db1 (the rest); either may be null
int wgrad_reduce_db_launch(const float* part, int splits, int taps, int cin, int cout, int split, float* d0, int ld0, int ci00,
                           float* d1, int ld1, int ci01, int accumulate, float alpha, const float* dbpart, float* db0, float* db1,
                           cudaStream_t st) {
  wgrad_reduce_go<false>(part, splits, taps, cin, cout, split, d0, ld0, ci00, d1, ld1, ci01, accumulate, alpha, st, dbpart, db0,
                         db1);
  count_launch();
  return check_launch("wgrad_reduce");
}

int bias_grad_launch(const void* dy, int dy_ld, int dtype, long long M, int cout, float* db, int accumulate,
                     float alpha, void* ws, cudaStream_t st) {
  float* bpart = reinterpret_cast<float*>(ws);
  if (dtype == SRCGAN_DT_BF16 && cout % 8 == 0 && cout <= 256 && dy_ld % 8 == 0 &&
      ((uintptr_t)dy) % 16 == 0) {
    const int rows_per_iter = 256 / (cout / 8);
    long long nb = (M + (long long)rows_per_iter * 8 - 1) / ((long long)rows_per_iter * 8);
    if (nb > 1024) nb = 1024;                 // the workspace holds 1024 partial rows (conv_wgrad_*_workspace)
    if (nb < 1) nb = 1;
    colsum_partial_vec<<<(unsigned)nb, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy), dy_ld, M, cout, bpart);
    colsum_final<<<ceil_div(cout, 32), 256, 0, st>>>(bpart, (int)nb, cout, db, accumulate, alpha, cout);
    count_launch(2);
    return check_launch("bias_grad");
  }
  int ny = (int)((M + 2047) / 2048);
  if (ny > 1024) ny = 1024;
  if (ny < 1) ny = 1;
  dim3 grid(ceil_div(cout, 32), ny), blk(32, 8);
  if (dtype == SRCGAN_DT_F32)
    colsum_partial<float><<<grid, blk, 0, st>>>(reinterpret_cast<const float*>(dy), dy_ld, M, cout, bpart);
  else
    colsum_partial<__nv_bfloat16><<<grid, blk, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy), dy_ld, M, cout,
                                                        bpart);
  colsum_final<<<ceil_div(cout, 32), 256, 0, st>>>(bpart, ny, cout, db, accumulate, alpha, cout);
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
