// Hardware experiment for the 2-CTA column-sweep kernel:
//  (1) semantics of tcgen05.mma.cta_group::2 (M = 256: A rows split across the CTA pair, B rows = N/2 per CTA, D = 128 x N per CTA),
//  (2) whether the D column address wraps modulo 512 when base + N runs past the end of TMEM (ring without split pieces),
//  (3) clocks per MMA for the sweep kernel's operand pattern (A 128 x 16, B N x 16) in 1-CTA and 2-CTA mode,
//      optionally with a background of extra shared-memory traffic (st.shared by idle warps).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_pair exp_pair.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t en) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(en) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(en) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// mode 0: one MMA at D column dcol, dump all 512 columns of lanes 0 and 127.  mode 1: rate loop.
template <int CG, int N>
__global__ void __launch_bounds__(192, 1) pair_kernel(float* out, long long* clocks, int mode, int dcol, int rounds, int noise) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 4 A slabs of 130 rows x 128 B (17408 B stride)
  uint8_t* sb = smem + 4 * 17408;     // 4 B tiles (one per k-step) of up to 192 rows x 128 B
  uint8_t* sn = sb + 4 * 24576;      // noise target (16 KB)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cta_rank() : 0;
  constexpr int NB = N / CG;          // B rows held by this CTA
  // A: every element = rank + 1.  B tile t: row j holds (rank * NB + j + 1) / 64
  for (int i = threadIdx.x; i < 4 * 17408 / 2; i += blockDim.x) ((__nv_bfloat16*)sa)[i] = __float2bfloat16((float)(rank + 1));
  for (int i = threadIdx.x; i < 4 * 24576 / 2; i += blockDim.x) {
    const int row = (i % (24576 / 2)) / 64;
    ((__nv_bfloat16*)sb)[i] = __float2bfloat16(row < NB ? (float)(rank * NB + row + 1) / 64.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp < 4) {                      // zero all of TMEM
    const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 512; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tq + c), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
  if (warp == 4 && rank == 0) {
    const long long t0 = clock64();
    const int reps = mode == 0 ? 1 : rounds;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (mode == 0 && (kh | ks)) continue;
          const uint64_t ad = desc(smem_u32(sa) + (r & 3) * 17408 + kh * 128 + ks * 32, 1024);
          const uint64_t bd = desc(smem_u32(sb) + ks * 24576, 1024);
          const uint32_t d = tmem + (mode == 0 ? dcol : ((r * 32) & 255));
          if (elect_one()) mma<CG>(d, ad, bd, idesc, 1u);
        }
    }
    if (elect_one()) commit<CG>(&bar);
    __syncwarp();
    while (!mbar_try(&bar, 0)) {}
    const long long t1 = clock64();
    if (lane == 0) clocks[blockIdx.x] = t1 - t0;
  } else if (warp == 5 && mode == 1 && noise > 0) {
    // background shared-memory traffic: `noise` bytes per 64 clocks, roughly (st.shared.v4 by 32 lanes = 512 B)
    const uint32_t base = smem_u32(sn) + lane * 16;
    while (!mbar_try(&bar, 0)) {
      for (int i = 0; i < noise; i += 512)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + (i & 0x3e00)), "r"(0u) : "memory");
      __nanosleep(0);
    }
  }
  if (warp < 4) {
    while (!mbar_try(&bar, 0)) {}
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mode == 0 && blockIdx.x < 2) {
      for (int c = 0; c < 512; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[((size_t)blockIdx.x * 128 + warp * 32 + lane) * 512 + c + i] = __uint_as_float(v[i]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int CG, int N>
void launch(float* out, long long* clk, int mode, int dcol, int rounds, int noise) {
  const int smem = 4 * 17408 + 4 * 24576 + 16384 + 2048;
  cudaFuncSetAttribute(pair_kernel<CG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, pair_kernel<CG, N>, out, clk, mode, dcol, rounds, noise);
  if (e != cudaSuccess) printf("launch failed: %s\n", cudaGetErrorString(e));
}

template <int CG, int N>
void semantics(int dcol) {
  float* out; long long* clk;
  cudaMalloc(&out, 2 * 128 * 512 * sizeof(float)); cudaMalloc(&clk, 148 * sizeof(long long));
  cudaMemset(out, 0, 2 * 128 * 512 * sizeof(float));
  launch<CG, N>(out, clk, 0, dcol, 1, 0);
  cudaError_t err = cudaDeviceSynchronize();
  std::vector<float> h(2 * 128 * 512);
  cudaMemcpy(h.data(), out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
  printf("semantics CG=%d N=%d dcol=%d: %s\n", CG, N, dcol, cudaGetErrorString(err));
  for (int cta = 0; cta < CG; ++cta)
    for (int ln : {0, 127}) {
      printf("  cta %d lane %3d nonzero runs:", cta, ln);
      const float* row = &h[((size_t)cta * 128 + ln) * 512];
      int c = 0;
      while (c < 512) {
        if (row[c] == 0.f) { ++c; continue; }
        int e = c;
        while (e < 512 && row[e] != 0.f) ++e;
        printf(" [%d..%d] first %.3f last %.3f;", c, e - 1, row[c], row[e - 1]);
        c = e;
      }
      printf("\n");
    }
  // expected D[m][n] = 16 * (rank + 1) * (n + 1) / 64
  int bad = 0;
  for (int cta = 0; cta < CG; ++cta)
    for (int ln = 0; ln < 128; ++ln)
      for (int n = 0; n < N; ++n) {
        const float want = 16.f * (cta + 1) * (n + 1) / 64.f;
        const float got = h[((size_t)cta * 128 + ln) * 512 + ((dcol + n) & 511)];
        if (got != want) ++bad;
      }
  printf("  mismatches vs 16*(rank+1)*(n+1)/64 at column (dcol+n) mod 512: %d\n", bad);
  cudaFree(out); cudaFree(clk);
}

template <int CG, int N>
void rate(int noise) {
  float* out; long long* clk;
  cudaMalloc(&out, 2 * 128 * 512 * sizeof(float)); cudaMalloc(&clk, 148 * sizeof(long long));
  const int rounds = 2000;
  launch<CG, N>(out, clk, 1, 0, 10, noise);
  launch<CG, N>(out, clk, 1, 0, rounds, noise);
  cudaError_t err = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  const double c = (double)h[0] / (rounds * 12.0);
  const double flop = 2.0 * 128 * CG * N * 16;     // per instruction (pair-wide for CG=2)
  printf("rate CG=%d N=%3d noise %5d B/iter: %s  clk/MMA %.1f  FLOP/clk/SM %.0f (%.0f%% of 8192)\n", CG, N, noise, cudaGetErrorString(err), c,
         flop / c / CG, 100 * flop / c / CG / 8192);
  cudaFree(out); cudaFree(clk);
}

int main() {
  semantics<1, 96>(0);
  semantics<2, 96>(0);
  semantics<2, 192>(64);
  for (int noise : {0, 2048}) {
    rate<1, 32>(noise); rate<1, 96>(noise); rate<1, 192>(noise);
    rate<2, 32>(noise); rate<2, 96>(noise); rate<2, 192>(noise);
  }
  return 0;
}
