// Hardware experiment: per-SM throughput of TMA tensor loads / stores for the box shapes of the sweep kernels
// (rows of 128 B / 64 B out of a 192-channel NHWC bf16 tensor), with nothing else running in the CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o exp_tma exp_tma.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) { while (!mbar_try(b, ph)) {} }
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

constexpr int NSLOT = 8;
constexpr int SLOT = 17408;

// mode bit 0: loads, bit 1: stores.  Each CTA sweeps `cols` columns of its own 128-row strip: per column one load box per chunk
// (nchunks) and one store box.  lane-dim coordinate = strip * 128, sweep coordinate = column.
__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap my,
                                                     int mode, int nchunks, int cols, int strips_per_img, int load_bytes, long long* clocks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[NSLOT], empty[NSLOT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int img = blockIdx.x / strips_per_img, y0 = (blockIdx.x % strips_per_img) * 128;
  const long long t0 = clock64();
  if (warp == 0 && (mode & 1)) {
    int s = 0; uint32_t ph = 0;
    for (int c = 0; c < cols; ++c)
      for (int k = 0; k < nchunks; ++k) {
        mbar_wait(&empty[s], ph ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&full[s], load_bytes);
          tma_load_4d(&mx, &full[s], smem + s * SLOT, k * 64, c, y0 - 1, img);
        }
        __syncwarp();
        if (++s == NSLOT) { s = 0; ph ^= 1; }
      }
  } else if (warp == 1 && (mode & 1)) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < cols * nchunks; ++i) {
      mbar_wait(&full[s], ph);
      if (lane == 0) mbar_arrive(&empty[s]);
      __syncwarp();
      if (++s == NSLOT) { s = 0; ph ^= 1; }
    }
  } else if (warp == 2 && (mode & 2)) {
    uint8_t* st = smem + NSLOT * SLOT;
    for (int c = 0; c < cols; ++c) {
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        tma_store_4d(&my, st + (c & 1) * 16384, 0, c, y0, img);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return (EncodeTiledFn)p;
}

// layout 0: column sweep (lane dim = image rows), 1: row sweep (lane dim = pixels of a row), 2: row sweep with the 130 extent in box dim 1
static void make(CUtensorMap* tm, void* ptr, int c, int w, int h, int n, int ld, int lanes, int layout, int boxc, CUtensorMapSwizzle sw) {
  cuuint64_t gdim[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * w, (cuuint64_t)ld * 2 * w * h};
  cuuint32_t box[4] = {(cuuint32_t)boxc, 1, (cuuint32_t)lanes, 1};
  if (layout == 1) { gdim[1] = h; gdim[2] = w; gstr[0] = (cuuint64_t)ld * 2 * w; gstr[1] = (cuuint64_t)ld * 2; }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
}

int main() {
  const int n = 37, h = 256, w = 256;            // 74 strips x 2 = 148 CTAs, one strip each
  for (int ld : {192, 64}) {
    __nv_bfloat16 *x, *y;
    cudaMalloc(&x, (size_t)n * h * w * ld * 2); cudaMalloc(&y, (size_t)n * h * w * ld * 2);
    cudaMemset(x, 0, (size_t)n * h * w * ld * 2);
    long long* clk; cudaMalloc(&clk, 148 * 8);
    const int smem = NSLOT * SLOT + 32768 + 2048;
    cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int layout : {0, 1})
      for (int boxc : {64, 32})                   // load 64-channel rows (128 B); store 64- or 32-channel rows
        for (int mode : {1, 2, 3}) {
          if (boxc == 32 && mode == 1) continue;
          CUtensorMap mx, my;
          make(&mx, x, ld < 64 ? ld : 64, w, h, n, ld, 130, layout, 64, CU_TENSOR_MAP_SWIZZLE_128B);
          make(&my, y, boxc, w, h, n, ld, 128, layout, boxc, boxc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
          const int cols = 256, nchunks = 1;
          tma_kernel<<<148, 128, smem>>>(mx, my, mode, nchunks, 32, 2, 130 * 128, clk);
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          tma_kernel<<<148, 128, smem>>>(mx, my, mode, nchunks, cols, 2, 130 * 128, clk);
          cudaEventRecord(e1);
          cudaError_t err = cudaDeviceSynchronize();
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          long long hc[148]; cudaMemcpy(hc, clk, sizeof(hc), cudaMemcpyDeviceToHost);
          const double lrows = (mode & 1) ? 130.0 * cols * nchunks : 0, srows = (mode & 2) ? 128.0 * cols : 0;
          const double bytes = 148.0 * (lrows * 128 + srows * boxc * 2);
          printf("ld=%3d layout=%d store_ch=%d mode=%d: %s  %.3f ms  %.2f TB/s  %.1f clk/col  %.2f clk/row\n", ld, layout, boxc, mode,
                 cudaGetErrorString(err), ms, bytes / ms / 1e9, (double)hc[0] / cols, (double)hc[0] / (lrows + srows));
        }
    cudaFree(x); cudaFree(y); cudaFree(clk);
  }
  return 0;
}
