// Hardware experiment: can a K-major SWIZZLE_128B UMMA A-descriptor start at an address that is only
// 128-byte (not 1024-byte) aligned, i.e. can the kw taps of a 3x3 conv read ONE haloed slab shifted by
// kw pixels?  Variants: slab row pitch (box width) 16 px (group bases stay congruent mod 1024) or 10 px,
// descriptor base_offset 0 or (start>>7)&7.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  long long t0 = clock64();
  while (!mbar_try(b, ph)) if (clock64() - t0 > 2000000000LL) __trap();
}

__global__ void __launch_bounds__(128, 1)
halo_kernel(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16* wswz, float* out, int boxw, int kh, int kw,
            int bo, int slab_bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sb = smem + 40960;
  uint64_t* bar = (uint64_t*)(sb + 8192);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar[0], slab_bytes + 8192);
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(&tmap), "r"(smem_u32(&bar[0])), "r"(0), "r"(-1), "r"(-1), "r"(0) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sb)), "l"(wswz), "r"(8192), "r"(smem_u32(&bar[0])) : "memory");
  }
  mbar_wait(&bar[0], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t sa = smem_u32(smem) + kw * 128 + kh * boxw * 128 + ks * 32;
      uint64_t ad = (uint64_t)((sa & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((boxw * 128) >> 4) << 32) |
                    ((uint64_t)1 << 46) | ((uint64_t)(bo & 7) << 49) | ((uint64_t)2 << 61);
      uint32_t sbb = smem_u32(sb) + ks * 32;
      uint64_t bd = (uint64_t)((sbb & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                    ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
  }
  mbar_wait(&bar[1], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cb = 0; cb < 64; cb += 32) {
    uint32_t v[32];
    uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + cb;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + cb + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int exp_halo(const void* x, int H, int W, const void* wswz, float* out, int boxw, int kh, int kw, int bo) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) return -1;
  EncodeTiledFn encode = (EncodeTiledFn)fp;
  CUtensorMap tm;
  cuuint64_t gdim[4] = {64, (cuuint64_t)W, (cuuint64_t)H, 1};
  cuuint64_t gstr[3] = {128, (cuuint64_t)128 * W, (cuuint64_t)128 * W * H};
  cuuint32_t box[4] = {64, (cuuint32_t)boxw, 18, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)x, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return -2;
  int smem = 40960 + 8192 + 64 + 1024;
  cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  halo_kernel<<<1, 128, smem>>>(tm, (const __nv_bfloat16*)wswz, out, boxw, kh, kw, bo, 18 * boxw * 128);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("exp_halo: %s\n", cudaGetErrorString(e)); return -3; }
  return 0;
}
