// Shared helpers for the srcgan_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/srcgan_b200.h"

namespace srcgan {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

#define SRCGAN_REQUIRE(cond, ...)              \
  do {                                         \
    if (!(cond)) {                             \
      srcgan::set_error(__VA_ARGS__);          \
      return SRCGAN_E_INVALID;                 \
    }                                          \
  } while (0)

#define SRCGAN_CUDA(call)                                                                 \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      srcgan::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                        \
      return SRCGAN_E_CUDA;                                                               \
    }                                                                                     \
  } while (0)

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kNumSMs = 148;  // B200

// Guard for per-DEVICE CUDA state (cudaFuncSetAttribute, __constant__ uploads): one bit per device ordinal, so a process
// that touches a second GPU initialises it there too.  A race between the main and the autograd thread only repeats an
// idempotent call.
struct DeviceOnce {
  unsigned long long done = 0ull;
  // -> true if the calling thread's current device still needs the initialisation; call mark() after it succeeded
  bool needed(int* dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev;
    if (dev < 0 || dev >= 64) return true;
    return !((__atomic_load_n(&done, __ATOMIC_ACQUIRE) >> dev) & 1ull);
  }
  void mark(int dev) {
    if (dev >= 0 && dev < 64) __atomic_fetch_or(&done, 1ull << dev, __ATOMIC_RELEASE);
  }
};

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch (sm_90+).  A kernel launched with pdl_attr() may start while its predecessor in the stream is
// still draining: its CTAs are scheduled as soon as every CTA of the predecessor has executed pdl_launch_dependents() (or
// exited) and an SM has room, run their prologue (barrier init, TMEM allocation, weight / bias loads - data no triggering
// kernel writes) and block in pdl_wait() until the predecessor grid has completed and its writes are visible.  EVERY global
// access to data another kernel of the step produces or still reads - activations in, outputs, workspaces - must come after
// pdl_wait().  Kernels that write parameters or packed weights never call pdl_launch_dependents(), so nothing can start (and
// pre-load weights) before they are done.  Both instructions are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// fills one launch attribute; returns the number of attributes written (0 when SRCGAN_B200_NO_PDL is set)
static inline int pdl_attr(cudaLaunchAttribute* at) {
  if (getenv("SRCGAN_B200_NO_PDL")) return 0;
  at->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at->val.programmaticStreamSerializationAllowed = 1;
  return 1;
}

}  // namespace srcgan
