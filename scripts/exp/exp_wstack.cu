// Hardware experiment for an N-stacked wgrad: can the MN-major B operand of tcgen05.mma be an OVERLAPPING view
//   B[k][(j, co)] = dY[pixel k + j][co],  j = 0..2
// of a [pixels][BN channels] slab, expressed as N atoms LBO = one pixel apart (LBO < atom size)?  Tested for
// BN = 32 (64-byte pixels, SWIZZLE_64B) and BN = 64 (128-byte pixels, SWIZZLE_128B); A = X^T, MN-major SWIZZLE_128B,
// M = 128 = two 64-channel atoms.  The slabs are written with the address-based swizzle TMA would apply.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o exp_wstack exp_wstack.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok;
}
// layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}

constexpr int NPIX = 40;       // pixels in the slabs (K up to 32 + shifts)
constexpr int KSTEPS = 2;      // K = 32 pixels

template <int BN>
__global__ void __launch_bounds__(128, 1) wstack_kernel(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* out, int kbase) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // two atoms [NPIX][64 ch] (128 B rows), 8 KB apart
  uint8_t* sb = smem + 16384;               // [NPIX][BN ch]
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  // x: [NPIX][128 ch], dy: [NPIX][BN]
  for (int i = threadIdx.x; i < NPIX * 128; i += blockDim.x) {
    const int p = i / 128, c = i % 128, atom = c / 64, cc = c % 64;
    uint32_t off = (uint32_t)(atom * 8192 + p * 128 + cc * 2);
    const uint32_t a = smem_u32(sa) + off;
    const uint32_t sw = a ^ (((a >> 7) & 7) << 4);                       // SWIZZLE_128B on the absolute address
    *(__nv_bfloat16*)(sa + (sw - smem_u32(sa))) = x[i];
  }
  for (int i = threadIdx.x; i < NPIX * BN; i += blockDim.x) {
    const int p = i / BN, c = i % BN;
    const uint32_t a = smem_u32(sb) + (uint32_t)(p * BN * 2 + c * 2);
    const uint32_t sw = BN == 64 ? (a ^ (((a >> 7) & 7) << 4)) : (a ^ (((a >> 7) & 3) << 4));   // SW128 / SW64
    *(__nv_bfloat16*)(sb + (sw - smem_u32(sb))) = dy[i];
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  constexpr int N = 3 * BN;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const int k0 = kbase + ks * 16;       // first X pixel of this k-step
      const uint64_t ad = desc(smem_u32(sa) + k0 * 128, 8192, 1024, 2);                 // LBO = atom stride (M), SBO = 8 K-rows
      const uint64_t bd = desc(smem_u32(sb) + k0 * BN * 2, BN * 2, 8 * BN * 2, BN == 64 ? 2 : 4);   // LBO = one pixel
      const uint32_t en = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(en) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  while (!mbar_try(&bar, 0)) {}
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(size_t)threadIdx.x * N + c + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// issue rate of the same instruction shape (MN-major A and B, N = 3*BN): rounds x 24 MMAs over 8 slab rows
template <int BN>
__global__ void __launch_bounds__(128, 1) wrate_kernel(long long* clocks, int rounds) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // two atoms of 10 rows x 16 px x 128 B, 20480 B apart
  uint8_t* sb = smem + 40960;               // [8 rows][18 px][BN ch]
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
  constexpr int N = 3 * BN;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < rounds; ++it)
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint64_t ad = desc(smem_u32(sa) + (r + g) * 2048, 20480, 1024, 2);
          const uint64_t bd = desc(smem_u32(sb) + r * 18 * BN * 2, BN * 2, 8 * BN * 2, BN == 64 ? 2 : 4);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + g * N), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    while (!mbar_try(&bar, 0)) {}
    clocks[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int BN>
void rate() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(wrate_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  wrate_kernel<BN><<<148, 128, 65536>>>(d, 10);
  wrate_kernel<BN><<<148, 128, 65536>>>(d, 1000);
  cudaError_t err = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double c = (double)h[0] / (1000.0 * 24);
  printf("rate MN-major A,B  N=%d: %s  %.1f clk/MMA  (%.0f%% of 8192 FLOP/clk)\n", 3 * BN, cudaGetErrorString(err), c, 100.0 * 2 * 128 * 3 * BN * 16 / c / 8192);
  cudaFree(d);
}

template <int BN>
void run(int kbase) {
  constexpr int N = 3 * BN;
  std::vector<__nv_bfloat16> hx(NPIX * 128), hdy(NPIX * BN);
  std::vector<float> fx(NPIX * 128), fdy(NPIX * BN);
  srand(7);
  for (size_t i = 0; i < hx.size(); ++i) { fx[i] = (float)(rand() % 17 - 8) / 8.f; hx[i] = __float2bfloat16(fx[i]); }
  for (size_t i = 0; i < hdy.size(); ++i) { fdy[i] = (float)(rand() % 13 - 6) / 4.f; hdy[i] = __float2bfloat16(fdy[i]); }
  __nv_bfloat16 *dx, *ddy; float* dout;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&ddy, hdy.size() * 2); cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(ddy, hdy.data(), hdy.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(wstack_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  wstack_kernel<BN><<<1, 128, 32768>>>(dx, ddy, dout, kbase);
  cudaError_t err = cudaDeviceSynchronize();
  std::vector<float> ho(128 * N);
  cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; double maxerr = 0;
  for (int ci = 0; ci < 128; ++ci)
    for (int j = 0; j < 3; ++j)
      for (int co = 0; co < BN; ++co) {
        double want = 0;
        for (int k = 0; k < 16 * KSTEPS; ++k) want += (double)fx[(kbase + k) * 128 + ci] * fdy[(kbase + k + j) * BN + co];
        const double got = ho[(size_t)ci * N + j * BN + co];
        if (fabs(got - want) > 1e-3) ++bad;
        if (fabs(got - want) > maxerr) maxerr = fabs(got - want);
      }
  printf("BN=%d kbase=%d: %s  mismatches %d of %d  max |err| %.4f\n", BN, kbase, cudaGetErrorString(err), bad, 128 * N, maxerr);
  cudaFree(dx); cudaFree(ddy); cudaFree(dout);
}

int main() {
  for (int kb : {0, 1, 3, 5}) { run<32>(kb); run<64>(kb); }
  rate<32>(); rate<64>();
  return 0;
}
