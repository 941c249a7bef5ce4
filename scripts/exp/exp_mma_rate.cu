// Hardware experiment: issue rate of tcgen05.mma (SS mode, kind::f16, M=128, K=16) as a function of N, operands
// resident in shared memory (contents irrelevant).  Two accumulators are interleaved like the conv kernels do.
// Prints clocks per MMA and the implied fraction of the dense bf16 peak (8192 FLOP/clk/SM nominal).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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

template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* clocks, int rounds, int a_stride, int a_sbo) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 4 A slabs of 16 KB (128 rows x 128 B)
  uint8_t* sb = smem + 98304;         // B: N rows x 128 B (up to 32 KB) x 2
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (warp == 1) {
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
      for (int t = 0; t < 3; ++t) {           // "taps": different A slab start
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int acc = 0; acc < NACC; ++acc) {
            const uint64_t ad = desc(smem_u32(sa) + acc * 24576 + t * a_stride + ks * 32, a_sbo);
            const uint64_t bd = desc(smem_u32(sb) + t * 0 + ks * 32, 1024);
            const uint32_t d = tmem + acc * (NACC > 2 ? 128 : 256);
            const uint32_t en = (r | t | ks) != 0;
            if (elect_one())
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                           ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(en) : "memory");
          }
        }
      }
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    while (!mbar_try(&bar, 0)) {}
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) clocks[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int NACC>
void run(int a_stride, int a_sbo = 1024) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int rounds = 2000;
  const int smem = 98304 + 65536 + 2048;
  cudaFuncSetAttribute(rate_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<N, NACC><<<148, 128, smem>>>(d, 10, a_stride, a_sbo);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<N, NACC><<<148, 128, smem>>>(d, rounds, a_stride, a_sbo);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double mm = (double)rounds * 12 * NACC;
  double clk = (double)h[0] / mm;
  double flop = 2.0 * 128 * N * 16;
  printf("SBO=%d NACC=%d N=%3d a_stride=%5d  %s  clk/MMA %.1f  FLOP/clk/SM %.0f (%.0f%% of 8192)  smemB/clk %.1f  chip %.0f TFLOP/s  (%.3f ms)\n", a_sbo, NACC, N, a_stride,
         cudaGetErrorString(err), clk, flop / clk, 100 * flop / clk / 8192, (4096 + N * 32) / clk, 148 * mm * flop / (ms * 1e-3) / 1e12, ms);
  cudaFree(d);
}

int main() {
  for (int sbo : {1024, 1280, 2048, 1152}) {
    run<32, 2>(128, sbo); run<64, 2>(128, sbo); run<32, 2>(1280, sbo); run<96, 2>(128, sbo);
  }
  return 0;
}
