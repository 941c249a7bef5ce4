// Fused loss forward+backward reductions (HBM-bound): one pass over (a, b) produces the scalar
// loss AND d loss / d a.  Replaces nn.L1Loss / nn.MSELoss (size_average=True) of
// src/losses.py:95-133 and the lsgan MSE-vs-expanded-label of GANLoss (src/train.py:112-128).
// Deterministic two-stage reduction (fixed block partial order), 128-bit loads, warp shuffles.
#include "common.cuh"

namespace srcgan {

constexpr int kLossBlocks = kNumSMs * 4;
constexpr int kLossThreads = 256;

size_t loss_workspace_bytes(int64_t) { return (size_t)kLossBlocks * sizeof(float) + 256; }

__device__ __forceinline__ float block_sum_256(float v) {
  __shared__ float red[8];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 8) t = red[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid on thread 0
}

template <int KIND>
__device__ __forceinline__ void loss_elem(float a, float b, float inv_n, float& acc, float& g) {
  float d = a - b;
  if (KIND == 0) {
    acc += fabsf(d);
    g = d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f);
  } else {
    acc = fmaf(d, d, acc);
    g = 2.f * d * inv_n;
  }
}

template <int KIND>
__global__ void __launch_bounds__(kLossThreads)
loss_partial(const float* __restrict__ a, const float* __restrict__ b, float b_scalar, int64_t n,
             float* __restrict__ grad, float* __restrict__ part, int vec) {
  const float inv_n = 1.f / (float)n;
  float acc = 0.f;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nthreads) {
      float4 va = __ldg(reinterpret_cast<const float4*>(a) + i);
      float4 vb = b ? __ldg(reinterpret_cast<const float4*>(b) + i) : make_float4(b_scalar, b_scalar, b_scalar, b_scalar);
      float4 g;
      loss_elem<KIND>(va.x, vb.x, inv_n, acc, g.x);
      loss_elem<KIND>(va.y, vb.y, inv_n, acc, g.y);
      loss_elem<KIND>(va.z, vb.z, inv_n, acc, g.z);
      loss_elem<KIND>(va.w, vb.w, inv_n, acc, g.w);
      if (grad) reinterpret_cast<float4*>(grad)[i] = g;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthreads) {
      float g;
      loss_elem<KIND>(a[i], b ? b[i] : b_scalar, inv_n, acc, g);
      if (grad) grad[i] = g;
    }
  } else {
    for (int64_t i = tid; i < n; i += nthreads) {
      float g;
      loss_elem<KIND>(a[i], b ? b[i] : b_scalar, inv_n, acc, g);
      if (grad) grad[i] = g;
    }
  }
  float t = block_sum_256(acc);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

__global__ void loss_final(const float* __restrict__ part, int nparts, double scale, float* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += (double)part[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
    out[0] = (float)(t * scale);
  }
}

int loss_fwd_bwd(int kind, const float* a, const float* b, float b_scalar, int64_t n, float* loss_out,
                 float* grad_out, void* ws, size_t ws_bytes, int mean, cudaStream_t st) {
  SRCGAN_REQUIRE(kind == 0 || kind == 1, "loss: kind must be 0 (L1) or 1 (MSE)");
  SRCGAN_REQUIRE(a && loss_out && n > 0, "loss: null pointer / empty input");
  SRCGAN_REQUIRE(ws && ws_bytes >= loss_workspace_bytes(n), "loss: workspace too small");
  float* part = reinterpret_cast<float*>(ws);
  int blocks = (int)((n + kLossThreads * 4 - 1) / (kLossThreads * 4));
  if (blocks > kLossBlocks) blocks = kLossBlocks;
  if (blocks < 1) blocks = 1;
  int vec = ((uintptr_t)a % 16 == 0) && (!b || (uintptr_t)b % 16 == 0) && (!grad_out || (uintptr_t)grad_out % 16 == 0);
  if (kind == 0)
    loss_partial<0><<<blocks, kLossThreads, 0, st>>>(a, b, b_scalar, n, grad_out, part, vec);
  else
    loss_partial<1><<<blocks, kLossThreads, 0, st>>>(a, b, b_scalar, n, grad_out, part, vec);
  loss_final<<<1, 256, 0, st>>>(part, blocks, mean ? 1.0 / (double)n : 1.0, loss_out);
  count_launch(2);
  return check_launch("loss_fwd_bwd");
}

}  // namespace srcgan
