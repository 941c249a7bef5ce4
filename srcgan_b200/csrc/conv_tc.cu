// tcgen05 / TMEM / TMA implicit-GEMM convolution engine for sm_100a (bf16 in, fp32 accumulate).
//
//   Y[n,y,x,co] = epilogue( sum_{kh,kw,ci} X[n, s*y+kh-pad, s*x+kw-pad, ci] * W[co,ci,kh,kw] )   s = 1, 2
//
// GEMM view: M = output pixels (one CTA tile = 16 rows x 8 columns = 128 pixels = UMMA M),
//            N = output channels (BN = 32/64/128 per tile), K = taps x input channels.
// Pipeline (warp-specialised, persistent CTAs, one per SM):
//   warp 0  TMA producer : per (64-channel chunk, kw) ONE 4-D tensor-map box load of the haloed
//                          activation slab  [16+KH-1 rows][8 px][64 ch]  (SWIZZLE_128B, OOB zero fill
//                          = the convolution padding), plus one bulk copy of the KH pre-swizzled
//                          weight tiles [BN][64].  The KH taps of a column reuse the SAME slab:
//                          a row shift is +8 px = +1024 B = one swizzle atom, so the UMMA descriptor
//                          start address just moves by kh*1024 (3 L2->SMEM loads per chunk, not 9).
//   warp 1  MMA issuer   : tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 per instruction,
//                          accumulators in TMEM (double buffered: epilogue of tile i overlaps the
//                          MMAs of tile i+1), tcgen05.commit releases SMEM stages / publishes TMEM.
//   warps 2-5 epilogue   : tcgen05.ld 32x32b -> bias, LeakyReLU, alpha, two residuals, LeakyReLU-mask
//                          (see include/srcgan_b200.h) -> bf16 -> 16-byte stores into the channel
//                          slice of the NHWC (concat) buffer.
// Reference semantics replaced: nn.Conv2d 3x3 of the dense blocks / HRconv / trunk / Decoder
// (src/model/model.py:193-211, :236-289, :396-440) and their dgrad (as an fprop over transposed weights).
#include <cuda.h>
#include <stdio.h>

#include <mutex>
#include <stdlib.h>

#include "common.cuh"

namespace srcgan {
namespace tc {

constexpr int TILE_H = 16, TILE_W = 8, TILE_M = TILE_H * TILE_W;
constexpr int KCH = 64;                 // channels per K chunk = one 128-byte swizzle row
constexpr int NUM_THREADS = 192;        // warp0 TMA, warp1 MMA, warps 2..5 epilogue
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int SMEM_AUX = 2048;          // barriers + tmem ptr + bias

// A "load" is one TMA box of the (possibly stride-sampled) input: [16+MAXT-1 rows][8 px][64 ch].
// The taps that differ only by a row shift share it (tap j reads the slab shifted by j rows = j*1024 B).
//   stride 1, KxK : K loads (one per kw), K taps each (kh)
//   stride 2, KxK : the box samples every 2nd pixel (tensor-map elementStrides = 2); taps whose kh have
//                   the same parity share a load: 3x3 -> 6 loads (2+1 taps), 4x4 -> 8 loads (2 taps)
//   stride-2 dgrad: four output phases, each a stride-1 conv over dY with 1x1 / 1x2 / 2x1 / 2x2 taps
constexpr int MAX_LOADS = 8, MAX_SLOTS = 16;
struct Plan {
  int nloads, total_slots, in_scale;     // slab origin = in_scale * (x0, y0) + (dx, dy)
  int dx[MAX_LOADS], dy[MAX_LOADS], ntaps[MAX_LOADS], slot0[MAX_LOADS];
};

struct TcArgs {
  int n, cin, cout;
  int gh, gw;                           // extent of the tile grid (output positions of this launch)
  int oh, ow, os, oa, ob;               // output tensor dims; position (y,x) is written at (y*os+oa, x*os+ob)
  Plan plan;
  int nchunks, n_blocks;                // K chunks of 64 channels; N blocks of BN channels
  int tiles_x, tiles_y;
  long long num_tiles;
  const __nv_bfloat16* wgt;             // packed [n_block][chunk][slot][BN][64] (pre-swizzled); 3x3 s1 cout<=64: slot = kh*3+(2-kw)
  const float* bias;
  __nv_bfloat16* y; int y_ld;
  int act; float act_slope, alpha;
  const __nv_bfloat16* r1; int r1_ld; float beta1;
  const __nv_bfloat16* r2; int r2_ld; float beta2;
  const __nv_bfloat16* mask; int mask_ld; float mask_slope;
  int dbg;                              // experiment switches (SRCGAN_B200_DBG), 0 in production
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {       // non-blocking peek
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must fault (trap), never hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(const void* src, uint64_t* bar, void* dst, uint32_t bytes) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset = 1024 B between 8-row groups
  d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4)                            // D format  F32
         | (1u << 7)                          // A format  BF16
         | (1u << 10)                         // B format  BF16
         | ((uint32_t)(N >> 3) << 17)         // N
         | ((uint32_t)(M >> 4) << 24);        // M         (A, B K-major: bits 15,16 = 0)
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Descriptor words kept separately: every descriptor of a stage differs from the stage's base only in
// the low word (start address >> 4), so an MMA costs one uniform add per operand at issue.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes = 16) {
  return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) {
  return (sbo_bytes >> 4) | (1u << 14) | (2u << 29);      // SBO, descriptor version 1, SWIZZLE_128B
}
__device__ __forceinline__ void umma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// One lane of a fully converged warp.  The producer / MMA warps keep their control flow warp-uniform and
// predicate only the issuing instructions with this: addresses, descriptors and loop counters then live in
// uniform registers (no per-MMA R2UR round trips; a lone divergent lane cost ~150 cycles per UTCHMMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}

// Shared-memory / TMEM budget of one CTA.  A CTA works on a SUPER-TILE of MT = 256/BN adjacent pixel
// tiles that share every weight stage (weights are the larger half of the L2->SMEM traffic at BN <= 64,
// and L2->SMEM bandwidth, not the tensor pipe, bounds this kernel): the weight ring is filled once per
// (chunk, load) and reused by MT activation slabs from a separate, deeper ring.  TMEM holds two sets of
// MT accumulators (2 x 256 columns): the epilogue drains one set while the MMAs fill the other.
template <int BN, int MAXT>
struct Cfg {
  static constexpr int MT = 256 / BN;
  static constexpr int A_BYTES = (TILE_H + MAXT - 1) * TILE_W * 128;
  static constexpr int B_TAP_BYTES = BN * 128;
  static constexpr int B_BYTES = MAXT * B_TAP_BYTES;
  static constexpr int NB = (3 * B_BYTES <= 80 * 1024) ? 3 : 2;
  static constexpr int NA_RAW = (SMEM_BUDGET - SMEM_AUX - 1024 - NB * B_BYTES) / A_BYTES;
  static constexpr int NA = NA_RAW > 10 ? 10 : NA_RAW;
  static constexpr int SMEM_BYTES = NA * A_BYTES + NB * B_BYTES + SMEM_AUX + 1024;
  static constexpr int TMEM_COLS = 512;
  static_assert(A_BYTES % 1024 == 0 && B_TAP_BYTES % 1024 == 0, "swizzle atoms must stay 1024-B aligned");
  static_assert(NA >= 3, "not enough shared memory for the activation ring");
};

// two epilogue groups of four warps (warps 2..5, 6..9) share every accumulator: group g takes the 32-column chunks with
// (chunk index % 2) == g.  One group needed ~1 200 clocks per (tile, chunk) = ~10 000 per super-tile of BN = 128, more than the
// 9 216 clocks of MMAs of a 3x3 128 -> 128 layer: the low-K layers of the Decoder / discriminators were bound by it.
constexpr int IG_GROUPS = 2;
constexpr int IG_THREADS = 64 + 128 * IG_GROUPS;
template <int BN, int MAXT>
__global__ void __launch_bounds__(IG_THREADS, 1)
conv_igemm_tc(const __grid_constant__ CUtensorMap tmap_x, const TcArgs a) {
  using C = Cfg<BN, MAXT>;
  constexpr int MT = C::MT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::NA * C::A_BYTES;
  uint8_t* aux = smem_b + C::NB * C::B_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);              // [NA]
  uint64_t* a_empty = a_full + C::NA;                                // [NA]
  uint64_t* b_full = a_empty + C::NA;                                // [NB]
  uint64_t* b_empty = b_full + C::NB;                                // [NB]
  uint64_t* tfull_bar = b_empty + C::NB;                             // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>(aux + 512);                // [<=256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C::NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128 * IG_GROUPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += IG_THREADS) sbias[i] = (a.bias && i < a.cout) ? a.bias[i] : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // super-tile t -> (n block, x group of MT tiles, tile row, image); identical decode in all roles
  const int groups_x = (a.tiles_x + MT - 1) / MT;
#define SRCGAN_DECODE_SUPERTILE(t)                                     \
  const int nb = (int)((t) % a.n_blocks);                              \
  long long r_ = (t) / a.n_blocks;                                     \
  const int bxg = (int)(r_ % groups_x); r_ /= groups_x;                \
  const int by = (int)(r_ % a.tiles_y);                                \
  const int img = (int)(r_ / a.tiles_y);                               \
  const int mcount = (a.tiles_x - bxg * MT) < MT ? (a.tiles_x - bxg * MT) : MT;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform)
    {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
        SRCGAN_DECODE_SUPERTILE(t)
        const int y0 = by * TILE_H;
        for (int c = 0; c < a.nchunks; ++c) {
          for (int l = 0; l < a.plan.nloads; ++l) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            const uint32_t bbytes = (uint32_t)(a.plan.ntaps[l] * C::B_TAP_BYTES);
            const __nv_bfloat16* wsrc =
                a.wgt + (((size_t)nb * a.nchunks + c) * a.plan.total_slots + a.plan.slot0[l]) * (size_t)(BN * KCH);
            if (elect_one()) {
              mbar_expect_tx(&b_full[bs], bbytes);
              bulk_load(wsrc, &b_full[bs], smem_b + bs * C::B_BYTES, bbytes);
            }
            __syncwarp();
            if (++bs == C::NB) { bs = 0; bph ^= 1; }
            for (int m = 0; m < mcount; ++m) {
              const int x0 = (bxg * MT + m) * TILE_W;
              mbar_wait(&a_empty[as], aph ^ 1);
              if (elect_one()) {
                mbar_expect_tx(&a_full[as], C::A_BYTES);
                tma_load_4d(&tmap_x, &a_full[as], smem_a + as * C::A_BYTES, c * KCH,
                            a.plan.in_scale * x0 + a.plan.dx[l], a.plan.in_scale * y0 + a.plan.dy[l], img);
              }
              __syncwarp();
              if (++as == C::NA) { as = 0; aph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform)
    {
      constexpr uint32_t idesc = umma_idesc(TILE_M, BN);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int set = 0;
      uint32_t set_phase = 0;
      for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
        SRCGAN_DECODE_SUPERTILE(t)
        (void)nb; (void)by; (void)img;
        mbar_wait(&tempty_bar[set], set_phase ^ 1);
        tc_fence_after();
        uint32_t started = 0;                       // 0 while the first (chunk, load) initialises the accumulators
        for (int c = 0; c < a.nchunks; ++c) {
          const int rem = a.cin - c * KCH;
          const int ksteps = ((rem >= KCH ? KCH : rem) + 15) >> 4;   // OOB channels are TMA zero-filled
          for (int l = 0; l < a.plan.nloads; ++l) {
            mbar_wait(&b_full[bs], bph);
            const uint32_t sb = smem_u32(smem_b + bs * C::B_BYTES);
            const int nt = a.plan.ntaps[l];
            for (int m = 0; m < mcount; ++m) {
              mbar_wait(&a_full[as], aph);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem_a + as * C::A_BYTES);
              const uint32_t tmem_d = tmem_base + (uint32_t)((set * MT + m) * BN);
              if (elect_one()) {
                const uint32_t a_lo = desc_lo(sa), b_lo = desc_lo(sb);
                constexpr uint32_t d_hi = desc_hi(1024);
                uint32_t accumulate = started;
                for (int j = 0; j < nt; ++j) {
                  if (ksteps == 4) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                      umma_bf16_w(tmem_d, a_lo + (uint32_t)(j * (TILE_W * 8) + ks * 2), d_hi,
                                  b_lo + (uint32_t)(j * (C::B_TAP_BYTES >> 4) + ks * 2), d_hi, idesc, accumulate);
                      accumulate = 1;
                    }
                  } else {
                    for (int ks = 0; ks < ksteps; ++ks) {
                      umma_bf16_w(tmem_d, a_lo + (uint32_t)(j * (TILE_W * 8) + ks * 2), d_hi,
                                  b_lo + (uint32_t)(j * (C::B_TAP_BYTES >> 4) + ks * 2), d_hi, idesc, accumulate);
                      accumulate = 1;
                    }
                  }
                }
                umma_commit(&a_empty[as]);         // frees the activation slab when these MMAs retire
              }
              __syncwarp();
              if (++as == C::NA) { as = 0; aph ^= 1; }
            }
            if (elect_one()) umma_commit(&b_empty[bs]);   // frees the weight stage
            __syncwarp();
            if (++bs == C::NB) { bs = 0; bph ^= 1; }
            started = 1;
          }
        }
        if (elect_one()) umma_commit(&tfull_bar[set]);    // all MT accumulators complete -> epilogue
        __syncwarp();
        if (++set == 2) { set = 0; set_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5 and 6..9)
    const int q = warp & 3;                        // TMEM lane quadrant this warp may access
    const int eg = (warp - 2) >> 2;                // epilogue group: its share of the 32-column chunks
    const int row = q * 32 + lane;                 // GEMM row = tile pixel
    const int ty = row >> 3, tx = row & 7;
    int set = 0;
    uint32_t set_phase = 0;
    for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      SRCGAN_DECODE_SUPERTILE(t)
      const int y = by * TILE_H + ty;
      mbar_wait(&tfull_bar[set], set_phase);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < mcount; ++m) {
        const int x = (bxg * MT + m) * TILE_W + tx;
        const bool valid = (y < a.gh) && (x < a.gw);
        const long long pix = ((long long)img * a.oh + (y * a.os + a.oa)) * a.ow + (x * a.os + a.ob);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((set * MT + m) * BN);
#pragma unroll 1
        for (int cb = eg * 32; cb < BN; cb += 32 * IG_GROUPS) {
          uint32_t v[32];
          tmem_ld32(taddr + cb, v);
          if (valid) {
            const int c0 = nb * BN + cb;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float t0 = __uint_as_float(v[g * 8 + i]) + sbias[c0 + g * 8 + i];
                if (a.act) t0 = t0 > 0.f ? t0 : t0 * a.act_slope;
                f[i] = t0 * a.alpha;
              }
              const int cc = c0 + g * 8;
              if (a.r1) {
                float rr[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(a.r1 + pix * a.r1_ld + cc)), rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaf(a.beta1, rr[i], f[i]);
              }
              if (a.r2) {
                float rr[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(a.r2 + pix * a.r2_ld + cc)), rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaf(a.beta2, rr[i], f[i]);
              }
              if (a.mask) {
                float mm[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(a.mask + pix * a.mask_ld + cc)), mm);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] *= (mm[i] > 0.f ? 1.f : a.mask_slope);
              }
              uint4 o;
              __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              *reinterpret_cast<uint4*>(a.y + pix * a.y_ld + cc) = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[set]);
      if (++set == 2) { set = 0; set_phase ^= 1; }
    }
  }
#undef SRCGAN_DECODE_SUPERTILE

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// 3x3 stride-1 fprop (and dgrad-as-fprop) for BN = 32 / 64 - the dense-block, HRconv, trunk and Decoder
// shapes that carry >90 % of the step's FLOPs.  Measured on B200 (scripts/exp/exp_halo.cu): the UMMA
// SWIZZLE_128B pattern is a pure function of the shared-memory ADDRESS, so a K-major A descriptor may
// start at any 128-byte row, not only at a 1024-byte atom.  One haloed slab [18 rows][10 px][64 ch]
// (23 KB, one TMA box) therefore feeds all nine taps of a tile: tap (kh,kw) = start + (kh*10+kw)*128 B,
// SBO = 1280 B.  That is 2.4x less L2->SMEM traffic than three kw-shifted slabs - and L2->SMEM feed is
// what bounds these kernels.  The chunk's nine weight tiles form ONE stage shared by the 256/BN pixel
// tiles of a super-tile.
// ---------------------------------------------------------------------------------------------
constexpr int HALO_W = TILE_W + 2, HALO_H = TILE_H + 2;
template <int BN>
struct HCfg {
  static constexpr int MT = 256 / BN > 8 ? 8 : 256 / BN;
  static constexpr int A_BOX_BYTES = HALO_H * HALO_W * 128;                  // 23040
  static constexpr int A_STRIDE = (A_BOX_BYTES + 1023) / 1024 * 1024;        // 23552
  static constexpr int B_TAP_BYTES = BN * 128;
  static constexpr int B_BYTES = 9 * B_TAP_BYTES;
  static constexpr int NB = 2;
  static constexpr int NA_RAW = (SMEM_BUDGET - SMEM_AUX - 1024 - NB * B_BYTES) / A_STRIDE;
  static constexpr int NA = NA_RAW > 8 ? 8 : NA_RAW;
  static constexpr int SMEM_BYTES = NA * A_STRIDE + NB * B_BYTES + SMEM_AUX + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int ILV = NA >= 4 ? 2 : 1;      // pixel tiles whose MMA chains are interleaved
  static_assert(NA >= 3, "not enough shared memory for the activation ring");
};

__device__ __forceinline__ uint64_t umma_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_halo_tc(const __grid_constant__ CUtensorMap tmap_x, const TcArgs a) {
  using C = HCfg<BN>;
  constexpr int MT = C::MT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::NA * C::A_STRIDE;
  uint8_t* aux = smem_b + C::NB * C::B_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* a_empty = a_full + C::NA;
  uint64_t* b_full = a_empty + C::NA;
  uint64_t* b_empty = b_full + C::NB;
  uint64_t* tfull_bar = b_empty + C::NB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>(aux + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C::NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += NUM_THREADS) sbias[i] = (a.bias && i < a.cout) ? a.bias[i] : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int groups_x = (a.tiles_x + MT - 1) / MT;
  const int pad = -a.plan.dy[0];
#define SRCGAN_DECODE_SUPERTILE(t)                                     \
  const int nb = (int)((t) % a.n_blocks);                              \
  long long r_ = (t) / a.n_blocks;                                     \
  const int bxg = (int)(r_ % groups_x); r_ /= groups_x;                \
  const int by = (int)(r_ % a.tiles_y);                                \
  const int img = (int)(r_ / a.tiles_y);                               \
  const int mcount = (a.tiles_x - bxg * MT) < MT ? (a.tiles_x - bxg * MT) : MT;

  if (warp == 0) {
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      SRCGAN_DECODE_SUPERTILE(t)
      const int y0 = by * TILE_H;
      for (int c = 0; c < a.nchunks; ++c) {
        mbar_wait(&b_empty[bs], bph ^ 1);
        const __nv_bfloat16* wsrc = a.wgt + (((size_t)nb * a.nchunks + c) * 9) * (size_t)(BN * KCH);
        if (elect_one()) {
          mbar_expect_tx(&b_full[bs], C::B_BYTES);
          bulk_load(wsrc, &b_full[bs], smem_b + bs * C::B_BYTES, C::B_BYTES);
        }
        __syncwarp();
        if (++bs == C::NB) { bs = 0; bph ^= 1; }
        for (int m = 0; m < mcount; ++m) {
          const int x0 = (bxg * MT + m) * TILE_W;
          mbar_wait(&a_empty[as], aph ^ 1);
          if (elect_one()) {
            if (a.dbg & 4) {
              mbar_arrive(&a_full[as]);
            } else {
              mbar_expect_tx(&a_full[as], C::A_BOX_BYTES);
              tma_load_4d(&tmap_x, &a_full[as], smem_a + as * C::A_STRIDE, c * KCH, x0 - pad, y0 - pad, img);
            }
          }
          __syncwarp();
          if (++as == C::NA) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc(TILE_M, BN);
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    int set = 0;
    uint32_t set_phase = 0;
    for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      SRCGAN_DECODE_SUPERTILE(t)
      (void)nb; (void)by; (void)img;
      mbar_wait(&tempty_bar[set], set_phase ^ 1);
      tc_fence_after();
      for (int c = 0; c < a.nchunks; ++c) {
        const int rem = a.cin - c * KCH;
        const int ksteps = ((rem >= KCH ? KCH : rem) + 15) >> 4;   // OOB channels are TMA zero-filled
        mbar_wait(&b_full[bs], bph);
        const uint32_t sb = smem_u32(smem_b + bs * C::B_BYTES);
        // Consecutive MMAs into the SAME accumulator form a dependent chain that the tensor pipe does not
        // overlap (short N=32/64 MMAs then pay the full issue-to-retire latency each).  Two pixel tiles are
        // therefore processed together, their MMAs interleaved: two independent accumulation chains.
        for (int m = 0; m < mcount; m += C::ILV) {
          const int cnt = (mcount - m) < C::ILV ? (mcount - m) : C::ILV;
          const int as0 = as;
          mbar_wait(&a_full[as], aph);
          if (++as == C::NA) { as = 0; aph ^= 1; }
          const int as1 = as;
          if (cnt == 2) {
            mbar_wait(&a_full[as], aph);
            if (++as == C::NA) { as = 0; aph ^= 1; }
          }
          tc_fence_after();
          const uint32_t sa0 = smem_u32(smem_a + as0 * C::A_STRIDE), sa1 = smem_u32(smem_a + as1 * C::A_STRIDE);
          const uint32_t tmem_d0 = tmem_base + (uint32_t)((set * MT + m) * BN), tmem_d1 = tmem_d0 + BN;
          if (elect_one()) {
            const uint32_t a_lo0 = desc_lo(sa0), a_lo1 = desc_lo(sa1), b_lo = desc_lo(sb);
            constexpr uint32_t a_hi = desc_hi(HALO_W * 128), b_hi = desc_hi(1024);
            uint32_t accumulate = c > 0 ? 1u : 0u;
            if (a.dbg & 2) {
            } else if (ksteps == 4 && cnt == 2) {
#pragma unroll
              for (int fw = 0; fw < 3; ++fw)
#pragma unroll
                for (int fh = 0; fh < 3; ++fh)
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ao = (uint32_t)((fh * HALO_W + fw) * 8 + ks * 2);
                    const uint32_t bo = (uint32_t)(((fh * 3 + 2 - fw) * C::B_TAP_BYTES >> 4) + ks * 2);
                    umma_bf16_w(tmem_d0, a_lo0 + ao, a_hi, b_lo + bo, b_hi, idesc, accumulate);
                    umma_bf16_w(tmem_d1, a_lo1 + ao, a_hi, b_lo + bo, b_hi, idesc, accumulate);
                    accumulate = 1;
                  }
            } else {
#pragma unroll 1
              for (int tap = 0; tap < 9; ++tap) {
                const int fh = tap / 3, fw = 2 - (tap - fh * 3);     // weight slot = kh*3 + (2 - kw)
                for (int ks = 0; ks < ksteps; ++ks) {
                  const uint32_t ao = (uint32_t)((fh * HALO_W + fw) * 8 + ks * 2);
                  const uint32_t bo = (uint32_t)((tap * C::B_TAP_BYTES >> 4) + ks * 2);
                  umma_bf16_w(tmem_d0, a_lo0 + ao, a_hi, b_lo + bo, b_hi, idesc, accumulate);
                  if (cnt == 2) umma_bf16_w(tmem_d1, a_lo1 + ao, a_hi, b_lo + bo, b_hi, idesc, accumulate);
                  accumulate = 1;
                }
              }
            }
            umma_commit(&a_empty[as0]);
            if (cnt == 2) umma_commit(&a_empty[as1]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&b_empty[bs]);
        __syncwarp();
        if (++bs == C::NB) { bs = 0; bph ^= 1; }
      }
      if (elect_one()) umma_commit(&tfull_bar[set]);
      __syncwarp();
      if (++set == 2) { set = 0; set_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ty = row >> 3, tx = row & 7;
    int set = 0;
    uint32_t set_phase = 0;
    for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      SRCGAN_DECODE_SUPERTILE(t)
      const int y = by * TILE_H + ty;
      mbar_wait(&tfull_bar[set], set_phase);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < mcount; ++m) {
        const int x = (bxg * MT + m) * TILE_W + tx;
        const bool valid = (y < a.gh) && (x < a.gw) && !(a.dbg & 1);
        const long long pix = ((long long)img * a.oh + y) * a.ow + x;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((set * MT + m) * BN);
        if (BN == 16) {
          // thin outputs (cout <= 16, e.g. the 3-channel image): scalar bf16 stores, bias / activation / alpha only
          uint32_t v[32];
          tmem_ld32(taddr, v);                     // 16 valid columns (+16 of the neighbouring accumulator, ignored)
          if (valid) {
            for (int c = 0; c < a.cout; ++c) {
              float t0 = __uint_as_float(v[c]) + sbias[c];
              if (a.act) t0 = t0 > 0.f ? t0 : t0 * a.act_slope;
              a.y[pix * a.y_ld + c] = __float2bfloat16_rn(t0 * a.alpha);
            }
          }
          continue;
        }
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + cb, v);
          if (valid) {
            const int c0 = nb * BN + cb;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float t0 = __uint_as_float(v[g * 8 + i]) + sbias[c0 + g * 8 + i];
                if (a.act) t0 = t0 > 0.f ? t0 : t0 * a.act_slope;
                f[i] = t0 * a.alpha;
              }
              const int cc = c0 + g * 8;
              if (a.r1) {
                float rr[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(a.r1 + pix * a.r1_ld + cc)), rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaf(a.beta1, rr[i], f[i]);
              }
              if (a.r2) {
                float rr[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(a.r2 + pix * a.r2_ld + cc)), rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaf(a.beta2, rr[i], f[i]);
              }
              if (a.mask) {
                float mm[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(a.mask + pix * a.mask_ld + cc)), mm);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] *= (mm[i] > 0.f ? 1.f : a.mask_slope);
              }
              uint4 o;
              __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              *reinterpret_cast<uint4*>(a.y + pix * a.y_ld + cc) = o;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[set]);
      if (++set == 2) { set = 0; set_phase ^= 1; }
    }
  }
#undef SRCGAN_DECODE_SUPERTILE
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// v4: kw-stacked 3x3 stride-1 pad-1 convolution, weights resident in shared memory.
//
// tcgen05.mma in SS mode re-reads its A tile (128 x 16 bf16 = 4 KB) from shared memory for every
// instruction and shared memory delivers 128 B/clk/SM, so an MMA takes max(N/2, 32 + N/4) clocks
// (measured, scripts/exp/exp_mma_rate.cu): N = 32 -> 40 clk (40 % of the tensor peak), N = 64 -> 48 (67 %),
// N = 96 -> 56 (86 %), N >= 128 -> full rate.  The dense-block layers have cout = 32 / 64, so instead of nine
// taps x N = cout this kernel stacks the three kw taps into the N dimension:
//     Z[p, (kw, co)] = sum_{kh, ci} X[p + (kh-1) rows, ci] * W[kh, kw, ci, co]        (N = 3*cout, 3x fewer MMAs)
//     Y[p, co]       = Z[p - 1, (0, co)] + Z[p, (1, co)] + Z[p + 1, (2, co)]          (epilogue, warp shuffles)
// A pixel tile is 4 rows x 32 columns (M = 128, one TMEM lane per pixel, warp q of the epilogue = tile row q,
// lane = column): the column shifts stay inside a warp.  Lanes 0 and 31 are halo columns - a tile produces
// 30 output columns - and R vertically adjacent tiles share one haloed slab [4R+2 rows][32 px][64 ch].
// The kh shift is a descriptor start-address offset of one slab row (4096 B).
// All nine weight tiles of every K chunk ([chunk][kh][3*BN rows][64] pre-swizzled) are loaded ONCE per CTA.
// Epilogue: TMEM -> registers -> shuffles -> bias / LeakyReLU / residuals / mask -> bf16 -> swizzled staging
// tile in shared memory -> ONE TMA tensor store per tile (clipped at the image edge by the tensor map).
// ---------------------------------------------------------------------------------------------
constexpr int KW_TW = 32, KW_TH = 4, KW_VX = 30;         // tile width / height (M = 128), valid output columns
struct KwArgs {
  int n, cin, cout, h, w;
  int nchunks, na;                      // K chunks of 64 channels; activation ring depth (host-sized)
  int tiles_x, tiles_y;
  long long num_tiles;
  const __nv_bfloat16* wgt;
  const float* bias;
  int act; float act_slope, alpha;
  const __nv_bfloat16* r1; int r1_ld; float beta1;
  const __nv_bfloat16* r2; int r2_ld; float beta2;
  const __nv_bfloat16* mask; int mask_ld; float mask_slope;
  int dbg;
};

template <int BN, int R>
struct KwCfg {
  static constexpr int N = 3 * BN;
  static constexpr int SLAB_ROWS = KW_TH * R + 2;
  static constexpr int SLAB_BYTES = SLAB_ROWS * KW_TW * 128;          // 24576 (R=1) / 40960 (R=2)
  static constexpr int W_KH_BYTES = N * 128;                          // one (chunk, kh) weight tile
  static constexpr int W_CHUNK_BYTES = 3 * W_KH_BYTES;
  static constexpr int STAGE_BYTES = ((KW_TH * R * KW_VX * BN * 2) + 1023) / 1024 * 1024;
  static constexpr int NSTAGE = 2;                                    // one staging tile per epilogue group
  static constexpr int TMEM_SET = R * N;                              // columns per accumulator set
  static_assert(2 * TMEM_SET <= 512, "two accumulator sets must fit in TMEM");
  static size_t smem_bytes(int nchunks, int na) {
    return (size_t)nchunks * W_CHUNK_BYTES + (size_t)na * SLAB_BYTES + NSTAGE * STAGE_BYTES + SMEM_AUX + 1024;
  }
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

constexpr int KW_THREADS = 320;          // warp0 TMA, warp1 MMA, warps 2..5 / 6..9 the two epilogue groups
template <int BN, int R>
__global__ void __launch_bounds__(KW_THREADS, 1)
conv3x3_kws_tc(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y, const KwArgs a) {
  using C = KwCfg<BN, R>;
  constexpr int MAX_NA = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + (size_t)a.nchunks * C::W_CHUNK_BYTES;
  uint8_t* smem_o = smem_a + (size_t)a.na * C::SLAB_BYTES;
  uint8_t* aux = smem_o + C::NSTAGE * C::STAGE_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* a_empty = a_full + MAX_NA;
  uint64_t* w_full = a_empty + MAX_NA;
  uint64_t* tfull_bar = w_full + 1;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>(aux + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.na; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < BN; i += KW_THREADS) sbias[i] = (a.bias && i < a.cout) ? a.bias[i] : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

#define SRCGAN_DECODE_KW_TILE(t)                                       \
  const uint32_t t32_ = (uint32_t)(t);                                 \
  const uint32_t r_ = t32_ / (uint32_t)a.tiles_x;                      \
  const int bx = (int)(t32_ - r_ * (uint32_t)a.tiles_x);               \
  const int img = (int)(r_ / (uint32_t)a.tiles_y);                     \
  const int by = (int)(r_ - (uint32_t)img * (uint32_t)a.tiles_y);      \
  const int x0 = bx * KW_VX, y0 = by * (KW_TH * R);

  if (warp == 0) {
    // ---- producer: resident weights once, then one haloed slab per (tile, K chunk)
    if (elect_one()) {
      mbar_expect_tx(w_full, (uint32_t)(a.nchunks * C::W_CHUNK_BYTES));
      for (int i = 0; i < a.nchunks * 3; ++i)
        bulk_load(a.wgt + (size_t)i * (C::W_KH_BYTES / 2), w_full, smem_w + (size_t)i * C::W_KH_BYTES, C::W_KH_BYTES);
    }
    __syncwarp();
    int as = 0;
    uint32_t aph = 0;
    for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      SRCGAN_DECODE_KW_TILE(t)
      for (int c = 0; c < a.nchunks; ++c) {
        mbar_wait(&a_empty[as], aph ^ 1);
        if (elect_one()) {
          if (a.dbg & 4) {
            mbar_arrive(&a_full[as]);
          } else {
            mbar_expect_tx(&a_full[as], C::SLAB_BYTES);
            tma_load_4d(&tmap_x, &a_full[as], smem_a + (size_t)as * C::SLAB_BYTES, c * KCH, x0 - 1, y0 - 1, img);
          }
        }
        __syncwarp();
        if (++as == a.na) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: per slab R x 3 (kh) x ksteps instructions of N = 3*BN
    constexpr uint32_t idesc = umma_idesc(TILE_M, C::N);
    constexpr uint32_t hi = desc_hi(1024);
    int as = 0;
    uint32_t aph = 0;
    int set = 0;
    uint32_t set_phase = 0;
    mbar_wait(w_full, 0);
    for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      mbar_wait(&tempty_bar[set], set_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(set * C::TMEM_SET);
      for (int c = 0; c < a.nchunks; ++c) {
        const int rem = a.cin - c * KCH;
        const int ksteps = ((rem >= KCH ? KCH : rem) + 15) >> 4;       // channels past cin are TMA zero-filled
        mbar_wait(&a_full[as], aph);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(smem_u32(smem_a + (size_t)as * C::SLAB_BYTES));
        const uint32_t b_lo = desc_lo(smem_u32(smem_w + (size_t)c * C::W_CHUNK_BYTES));
        if (elect_one()) {
          if (!(a.dbg & 2)) {
            if (ksteps == 4) {
#pragma unroll
              for (int s = 0; s < R; ++s)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_bf16_w(tmem_d + s * C::N, a_lo + (uint32_t)((s * KW_TH + kh) * (KW_TW * 8) + ks * 2), hi,
                                b_lo + (uint32_t)(kh * (C::W_KH_BYTES >> 4) + ks * 2), hi, idesc,
                                (c | kh | ks) != 0 ? 1u : 0u);
            } else {
#pragma unroll 1
              for (int s = 0; s < R; ++s)
#pragma unroll 1
                for (int kh = 0; kh < 3; ++kh)
                  for (int ks = 0; ks < ksteps; ++ks)
                    umma_bf16_w(tmem_d + s * C::N, a_lo + (uint32_t)((s * KW_TH + kh) * (KW_TW * 8) + ks * 2), hi,
                                b_lo + (uint32_t)(kh * (C::W_KH_BYTES >> 4) + ks * 2), hi, idesc,
                                (c | kh | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[as]);
          if (c == a.nchunks - 1) umma_commit(&tfull_bar[set]);
        }
        __syncwarp();
        if (++as == a.na) { as = 0; aph ^= 1; }
      }
      if (++set == 2) { set = 0; set_phase ^= 1; }
    }
  } else {
    // ---- epilogue: two groups of four warps take alternate tiles (group g drains accumulator set g), so a
    // tile's epilogue may last two tile periods.  Warp q = tile row q, lane = slab column; output column
    // x0 + lane - 1 for lanes 1..30.
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;
    const bool issuer = threadIdx.x == 64 + g * 128;
    uint8_t* so = smem_o + g * C::STAGE_BYTES;
    const uint32_t so_addr = smem_u32(so), sbias_addr = smem_u32(sbias);
    const uint32_t bar_id = 1 + g;
    long long it = 0;
    long long tp[6] = {0, 0, 0, 0, 0, 0}, tc0 = 0;
    const bool prof = (a.dbg & 32) != 0;
#define SRCGAN_TICK(k) if (prof) { const long long n_ = clock64(); tp[k] += n_ - tc0; tc0 = n_; }
    for (long long t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++it) {
      if ((int)(it & 1) != g) continue;
      if (prof) tc0 = clock64();
      const uint32_t set_phase = (uint32_t)((it >> 1) & 1);
      SRCGAN_DECODE_KW_TILE(t)
      // this group's previous TMA store must have finished reading the staging tile
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      SRCGAN_TICK(0)
      mbar_wait(&tfull_bar[g], set_phase);
      tc_fence_after();
      SRCGAN_TICK(1)
      const int x = x0 + lane - 1;
#pragma unroll 1
      for (int s = 0; s < R; ++s) {
        const int y = y0 + s * KW_TH + q;
        const bool valid = lane >= 1 && lane <= KW_VX && x < a.w && y < a.h && !(a.dbg & 1);
        const long long pix = ((long long)img * a.h + y) * a.w + x;
        const int sidx = (s * KW_TH + q) * KW_VX + (lane - 1);           // pixel slot in the staging tile
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * C::TMEM_SET + s * C::N);
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 32) {
          uint32_t z0[32], z1[32], z2[32];
          tmem_ld32_nowait(taddr + 2 * BN + cb, z0);        // accumulator columns are [kw2 | kw1 | kw0]
          tmem_ld32_nowait(taddr + BN + cb, z1);
          tmem_ld32_nowait(taddr + cb, z2);
          tmem_ld_wait();
          SRCGAN_TICK(2)
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = lds_f4(sbias_addr + (uint32_t)(cb + i) * 4);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float l = __shfl_up_sync(0xffffffffu, __uint_as_float(z0[i + k]), 1);
              const float r = __shfl_down_sync(0xffffffffu, __uint_as_float(z2[i + k]), 1);
              f[i + k] = (__uint_as_float(z1[i + k]) + bb[k]) + (l + r);
            }
          }
          if (a.act) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = f[i] > 0.f ? f[i] : f[i] * a.act_slope;
          }
          if (a.alpha != 1.f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] *= a.alpha;
          }
          SRCGAN_TICK(3)
          if (valid) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              const int cc = cb + gq * 8;
              if (a.r1) {
                float rr[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(a.r1 + pix * a.r1_ld + cc)), rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[gq * 8 + i] = fmaf(a.beta1, rr[i], f[gq * 8 + i]);
              }
              if (a.r2) {
                float rr[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(a.r2 + pix * a.r2_ld + cc)), rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[gq * 8 + i] = fmaf(a.beta2, rr[i], f[gq * 8 + i]);
              }
              if (a.mask) {
                float mm[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(a.mask + pix * a.mask_ld + cc)), mm);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[gq * 8 + i] *= (mm[i] > 0.f ? 1.f : a.mask_slope);
              }
              uint4 o;
              __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(f[gq * 8 + 2 * i], f[gq * 8 + 2 * i + 1]);
              // staging tile = the TMA box image [rows][30 px][BN ch] with the tensor map's swizzle
              // (BN = 64: SWIZZLE_128B, 16-byte chunk j of pixel slot i at j ^ (i & 7); BN = 32: SWIZZLE_64B, j ^ ((i >> 1) & 3))
              const int j = cc >> 3;
              const int js = BN == 64 ? (j ^ (sidx & 7)) : (j ^ ((sidx >> 1) & 3));
              sts_u4(so_addr + (uint32_t)(sidx * (BN * 2) + js * 16), o);
            }
          }
        }
      }
      SRCGAN_TICK(4)
      tc_fence_before();
      mbar_arrive(&tempty_bar[g]);                                       // accumulators drained
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // staging writes -> visible to the TMA engine
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (issuer && !(a.dbg & 1)) {
        tma_store_4d(&tmap_y, so, 0, x0, y0, img);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      SRCGAN_TICK(5)
    }
    if (prof && blockIdx.x == 0 && lane == 0)
      printf("kws epi warp %d: tiles %lld  bar %lld  tfull %lld  tmem_ld %lld  math %lld  stage %lld  tail %lld (clk)\n", warp,
             (it + 1 - g) / 2, tp[0], tp[1], tp[2], tp[3], tp[4], tp[5]);
#undef SRCGAN_TICK
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
#undef SRCGAN_DECODE_KW_TILE
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


constexpr int SW_ROWS = 128, SW_SLAB_ROWS = SW_ROWS + 2;               // a strip: 128 lanes + one halo row either side
constexpr int SW_SLAB_BYTES = SW_SLAB_ROWS * 128;                      // 16640
constexpr int SW_SLAB_STRIDE = (SW_SLAB_BYTES + 1023) / 1024 * 1024;   // 17408
constexpr int SW_MAX_NA = 10;

__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// v5: paired sweep.  Same N-stacking as v4 (three taps of one direction stacked into N = 3*BN, A read once per
// (tap of the other direction, k-step)), but the stacked taps' shift is applied to the ACCUMULATOR address, not the data:
//   * a work unit is a strip of 128 lanes (image rows, or the pixels of an image row: see `tr`) swept over a segment
//     of the other image dimension ("columns" below); TMEM is a ring of 512/BN accumulator blocks, one per output column;
//   * the MMAs of INPUT column c (A = its [130 lanes][64 ch] slab, lane tap = +128 B start offset) accumulate into the
//     three consecutive blocks [Y_{c-1} | Y_c | Y_{c+1}] with one N = 3*BN instruction;
//   * blocks are zeroed by the epilogue when it drains them, so every MMA accumulates (no first-touch flag);
//   * no shuffles, BN TMEM columns read per pixel, halo = 130/128 lanes x (seg+2)/seg columns.
// SS-mode MMAs re-read A (4 KB) and B (N x 32 B) for every instruction; at 128 B/clk/SM an N = 96 MMA is SMEM-bound at
// 56 clk against 48 clk of math (scripts/exp/exp_mma_rate.cu).  Two design points follow:
//   * cta_group::2 - a CTA pair sweeps two 128-row strips in lock step (M = 256).  Each CTA feeds its own A slab
//     but only HALF of the stacked weight rows (N/2 x 32 B per MMA; measured 49 clk at N = 96, 96 clk at N = 192,
//     scripts/exp/exp_pair.cu), so the weights of 192->64 (221 KB stacked) fit as 2 x 110 KB and every layer of a
//     dense block runs on this kernel.  CG = 1 instantiations exist for A/B tests.
//   * a ring without split instructions - TMEM faults when base + N crosses column 512 (exp_pair.cu), and a
//     2-CTA instruction cannot be cut at the ring wrap (each CTA would need a different half of B).  The ring is
//     therefore run as RUN = NBLK - 2 input columns per lap: input column k of a lap accumulates into blocks
//     [k, k+1, k+2] (one N = 3*BN instruction, always), the lap restarts at block 0.  The two output columns that
//     straddle a lap boundary own two partial blocks each (block RUN of lap r + block 0 of lap r+1; block RUN+1 of
//     lap r + block 1 of lap r+1); the epilogue adds them.  Segment ends are handled the same way: every input
//     column (including the two halo columns of a segment, zero-filled by TMA outside the image) issues the full
//     instruction, and the two blocks per segment boundary that collect only halo contributions are drained as
//     dummies.  The MMA warp's loop is one wait + 3*ksteps instructions + commits per column, no index arithmetic.
// Roles: warp 0 TMA producer (both CTAs; the peer's loads signal the leader's mbarrier), warp 1 MMA issuer (leader
// CTA only; commits are multicast to both CTAs), then SW2_GROUPS epilogue groups of four warps taking output columns in turn.
// ---------------------------------------------------------------------------------------------
struct Sw2Args {
  int n, cin, cout, h, w;
  int nchunks, na;
  int wseg, segs_x, strips_y, strips;   // strips = n * strips_y, flattened over images
  int num_units;                        // ceil(strips / CG) * segs_x
  int pfd;                              // L2 prefetch distance of the producer, in columns (0 = off)
  unsigned int* sched;                  // {next unit, finished clusters} of the dynamic unit queue, or nullptr (static round-robin)
  int tr;                               // 0: lanes = image rows, sweep over x;  1: lanes = pixels of a row, sweep over y
  long long lane_stride, sweep_stride, img_stride;   // pixel index = img * img_stride + lane * lane_stride + sweep * sweep_stride
  const __nv_bfloat16* wgt;
  const float* bias;
  __nv_bfloat16* y; int y_ld;           // output channel slice (pixel pitch y_ld elements)
  int act; float act_slope, alpha;
  const __nv_bfloat16* r1; int r1_ld; float beta1;
  const __nv_bfloat16* r2; int r2_ld; float beta2;
  const __nv_bfloat16* mask; int mask_ld; float mask_slope;
  uint32_t* signbits;                   // OUT [pixel][BN/32]: bit = stored value > 0 (the mask of the matching backward step)
  const uint32_t* maskbits;             // IN  [pixel][BN/32]: replaces `mask` (4 bytes instead of 64 per pixel and 32 channels)
  int zero_period;                      // > 0: image rows with row % zero_period == 0 are stored as zeros ("tall image" separators)
  int planar;                           // 1: x is a planar concat buffer [group][n][h][w][64]: tmap_x is 5-D, K chunk k = group k
  int rev;                              // 1: work units in descending order (last image first): the layer then starts on the data the
                                        //    previous layer of a dense block touched last, which is what still sits in the 126 MB L2
  int per, total;                       // per > 0: "range mode" - the total = (strip groups) x w input columns of the launch are dealt to
                                        //    the clusters as equal contiguous ranges of `per` columns, cut into units at group boundaries
                                        //    (64 images x 256 columns on 74 CTA pairs: 222 columns each instead of 64 units of 256)
  int dbg;
};

template <int BN, int CG>
struct Sw2Cfg {
  static constexpr int NBLK = 512 / BN;                               // accumulator blocks in TMEM
  static constexpr int RUN = NBLK - 2;                                // input columns per lap
  static constexpr int N = 3 * BN;
  static constexpr int W_KH_FULL = N * 128;                           // one (chunk, kh) weight tile [kw2 | kw1 | kw0] x BN rows
  static constexpr int W_KH_BYTES = W_KH_FULL / CG;                   // the rows this CTA feeds
  static constexpr int W_CHUNK_BYTES = 3 * W_KH_BYTES;
  static size_t smem_bytes(int nchunks, int na) {
    return (size_t)nchunks * W_CHUNK_BYTES + (size_t)na * SW_SLAB_STRIDE + SMEM_AUX + 1024;
  }
};

constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's even (leader) CTA

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {       // arrive on the leader CTA's copy of `bar`
  // default (.release.cta) semantics: the TMEM hand-over is ordered by tcgen05.wait::st + tcgen05.fence, and a
  // .release.cluster arrive costs a MEMBAR.ALL.GPU + ERRBAR per call (37 % of the kernel's stall samples in ncu)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// 5-D forms for planar concat buffers ([group][n][h][w][64 ch]: coordinate 4 = the 64-channel group = the K chunk)
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_w2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void stg_u8(void* p, const uint4& lo, const uint4& hi) {      // one 32-byte sector per lane
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
               "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {           // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

// Dynamic unit queue of a CTA pair (opt-in, SRCGAN_B200_SWEEP_DYNAMIC=1).  The pairs of one launch take 127-216 us for equal
// shares of the work (the HBM-bound layers see very different TMA latencies per SM), so units can be claimed from a global
// counter instead of dealt round-robin.
// The leader's producer warp claims one unit ahead and publishes it to both CTAs' `unit_list` (one single-use mbarrier per
// list slot); every role of both CTAs walks the list.  A pair stops claiming after SW_MAXU - 1 units (host guarantees that
// ncl * (SW_MAXU - 1) >= num_units, else the launch is static).
constexpr int SW_MAXU = 32;
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // observes a peer CTA's release
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void publish_unit(int* list, uint64_t* bars, int i, int u, bool pair) {
  list[i] = u;
  mbar_arrive(&bars[i]);
  if (pair) {
    uint32_t rl, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rl) : "r"(smem_u32(&list[i])), "r"(1u));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(&bars[i])), "r"(1u));
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(rl), "r"(u) : "memory");
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rb) : "memory");
  }
}

// warp 0 producer, warp 1 MMA issuer, then SW2_GROUPS epilogue groups of four warps.  A group has SW2_GROUPS column periods
// to drain one column (one warp needs ~2 000 clocks per [32 lanes][32 channels] block: the thin layers are bound by this
// epilogue, not by their MMAs).  With the residual / bf16-mask operands compiled in (SIDE = true, 165 registers) a third
// group caps the kernel at 128 registers and spills (64 -> 64: 0.30 -> 0.48 ms), so that variant runs two groups.
// SIDE = false compiles those operands out (122 registers): three groups without spills, used for 64 output channels
// (0.308 -> 0.258 ms); for 32 output channels a third group loses (0.236 -> 0.257 ms).
constexpr int sw2_threads(int groups) { return 64 + 128 * groups; }
template <int BN, int CG, int SW2_GROUPS = 2, bool SIDE = true>
__global__ void __launch_bounds__(sw2_threads(SW2_GROUPS), 1)
conv3x3_sweep2_tc(const __grid_constant__ CUtensorMap tmap_x, const Sw2Args a) {
  using C = Sw2Cfg<BN, CG>;
  constexpr int SW2_THREADS = sw2_threads(SW2_GROUPS);
  constexpr int NBLK = C::NBLK, RUN = C::RUN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + (size_t)a.nchunks * C::W_CHUNK_BYTES;
  uint8_t* aux = smem_a + (size_t)a.na * SW_SLAB_STRIDE;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);                 // [SW_MAX_NA]  (leader's copy is the live one)
  uint64_t* a_empty = a_full + SW_MAX_NA;                              // [SW_MAX_NA]
  uint64_t* y_full = a_empty + SW_MAX_NA;                              // [NBLK <= 16]
  uint64_t* y_empty = y_full + 16;                                     // [NBLK <= 16] (leader's copy is the live one)
  uint64_t* w_full = y_empty + 16;
  uint64_t* w_pair = w_full + 1;                                       // leader: the peer's weights have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_pair + 1);
  float* sbias = reinterpret_cast<float*>(aux + 768);
  uint64_t* u_full = reinterpret_cast<uint64_t*>(aux + 1024);          // [SW_MAXU] single-use: unit_list[i] is valid
  int* unit_list = reinterpret_cast<int*>(aux + 1024 + SW_MAXU * 8);   // [SW_MAXU]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int cid = (int)(blockIdx.x / CG), ncl = (int)(gridDim.x / CG);
  pdl_launch_dependents();                                             // the next kernel's prologue may overlap this kernel's tail
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.na; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NBLK; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 4 * CG); }
    mbar_init(w_full, 1);
    mbar_init(w_pair, 1);
    for (int s = 0; s < SW_MAXU; ++s) mbar_init(&u_full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < BN; i += SW2_THREADS) sbias[i] = (a.bias && i < a.cout) ? a.bias[i] : 0.f;
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 2 && warp < 6) {                                         // every accumulator block starts at zero
    const uint32_t tq = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int cc = 0; cc < 512; cc += 32) tmem_st32_zero(tq + cc);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                                     // barriers initialised and TMEM zeroed in both CTAs
  tc_fence_after();

  // i-th unit of this CTA pair (-1: none left)
  auto get_unit = [&](int i) -> int {
    if (a.per > 0) {                                                   // range mode: the i-th strip group this cluster's range touches
      const int g0 = cid * a.per, g1 = (g0 + a.per) < a.total ? (g0 + a.per) : a.total;
      return (g0 < g1 && (g0 / a.w + i) * a.w < g1) ? i : -1;
    }
    if (a.sched == nullptr) { const int u = cid + i * ncl; return u < a.num_units ? (a.rev ? a.num_units - 1 - u : u) : -1; }
    if (CG == 2 && rank == 1) mbar_wait_cluster(&u_full[i], 0); else mbar_wait(&u_full[i], 0);
    return reinterpret_cast<volatile int*>(unit_list)[i];
  };
  auto claim_unit = [&](int i) {                                       // leader's producer warp, one lane
    int u = -1;
    if (i < SW_MAXU - 1) {
      u = (int)atomicAdd(a.sched, 1u);
      if (u >= a.num_units) u = -1;
      else if (a.rev) u = a.num_units - 1 - u;
    }
    publish_unit(unit_list, u_full, i, u, CG == 2);
  };

  // unit u -> column segment [x_start, x_end) of strip (u / segs_x) * CG + rank; an absent strip (odd total) sweeps
  // rows below the image: its loads are zero-filled, its stores suppressed
#define SRCGAN_DECODE_SW2_RANGE(u)                                                 \
  uint32_t sg_;                                                                    \
  int x_start, x_end;                                                              \
  if (a.per > 0) {                                                                 \
    const int g0_ = cid * a.per, g1_ = (g0_ + a.per) < a.total ? (g0_ + a.per) : a.total; \
    sg_ = (uint32_t)(g0_ / a.w + (u));                                             \
    const int base_ = (int)sg_ * a.w;                                              \
    x_start = (u) == 0 ? g0_ - base_ : 0;                                          \
    x_end = (g1_ - base_) < a.w ? (g1_ - base_) : a.w;                             \
  } else {                                                                         \
    sg_ = (uint32_t)(u) / (uint32_t)a.segs_x;                                      \
    const int seg_ = (int)((uint32_t)(u) - sg_ * (uint32_t)a.segs_x);              \
    x_start = seg_ * a.wseg;                                                       \
    x_end = (x_start + a.wseg) < a.w ? (x_start + a.wseg) : a.w;                   \
  }
#define SRCGAN_DECODE_SW2_UNIT(u)                                                  \
  SRCGAN_DECODE_SW2_RANGE(u)                                                       \
  const int sidx = (int)sg_ * CG + (int)rank;                                      \
  const bool strip_ok = sidx < a.strips;                                           \
  const int img = strip_ok ? sidx / a.strips_y : 0;                                \
  const int y0 = strip_ok ? (sidx - img * a.strips_y) * SW_ROWS : a.h + 1;

  if (warp == 0) {
    // ---- producer: this CTA's weight rows once, then one column slab per (input column, K chunk)
    if (elect_one()) {
      // tile (chunk, lane tap t) = the three sweep taps stacked [s = 2 | 1 | 0] x BN rows; packed slot of tap (kh, kw) = kh*3 + (2 - kw).
      // A CTA of a pair takes its half of the stacked rows (three half-slots).
      mbar_expect_tx(w_full, (uint32_t)(a.nchunks * C::W_CHUNK_BYTES));
      constexpr int PIECE = BN * 128 / CG;                              // bytes per copy: a slot (CG = 1) or half a slot
      for (int i = 0; i < a.nchunks; ++i)
        for (int t = 0; t < 3; ++t)
          for (int h3 = 0; h3 < 3; ++h3) {
            const int hs = CG == 1 ? h3 : (int)rank * 3 + h3;          // piece index in the stacked tile
            const int si = CG == 1 ? hs : hs >> 1, part = CG == 1 ? 0 : hs & 1;
            const int sw = 2 - si;                                     // sweep-direction tap
            const int kh = a.tr ? sw : t, kw = a.tr ? t : sw;
            const int slot = kh * 3 + (2 - kw);
            bulk_load(a.wgt + ((size_t)(i * 9 + slot) * (BN * 128) + (size_t)part * PIECE) / 2, w_full,
                      smem_w + (size_t)(i * 3 + t) * C::W_KH_BYTES + (size_t)h3 * PIECE, PIECE);
          }
    }
    __syncwarp();
    if (CG == 2 && rank == 1) {
      mbar_wait(w_full, 0);
      if (elect_one()) mbar_arrive_leader(w_pair);
      __syncwarp();
    }
    int as = 0;
    uint32_t aph = 0;
    pdl_wait();                                                        // the activations are the previous kernels' outputs
    const bool fetcher = a.sched != nullptr && rank == 0;
    if (fetcher && elect_one()) claim_unit(0);
    __syncwarp();
    for (int ui = 0;; ++ui) {
      const int u = get_unit(ui);
      if (u < 0) break;
      if (fetcher && elect_one()) claim_unit(ui + 1);                  // one unit ahead: the claim's round trip is off the path
      __syncwarp();
      SRCGAN_DECODE_SW2_UNIT(u)
      (void)strip_ok;
      const int pfd = a.pfd;                                           // L2 prefetch distance in columns
      if (pfd > 0 && elect_one())
        for (int c = x_start - 1; c < x_start - 1 + pfd && c <= x_end; ++c)
          for (int k = 0; k < a.nchunks; ++k) tma_prefetch_4d(&tmap_x, k * KCH, c, y0 - 1, img);
      __syncwarp();
      for (int c = x_start - 1; c <= x_end; ++c)
        for (int k = 0; k < a.nchunks; ++k) {
          mbar_wait(&a_empty[as], aph ^ 1);
          if (elect_one()) {
            if (pfd > 0 && c + pfd <= x_end) tma_prefetch_4d(&tmap_x, k * KCH, c + pfd, y0 - 1, img);
            uint8_t* dst = smem_a + (size_t)as * SW_SLAB_STRIDE;
            if (a.dbg & 4) {
              if (rank == 0) mbar_arrive(&a_full[as]);
            } else if (CG == 1) {
              mbar_expect_tx(&a_full[as], SW_SLAB_BYTES);
              if (a.planar) tma_load_5d(&tmap_x, &a_full[as], dst, 0, c, y0 - 1, img, k);
              else tma_load_4d(&tmap_x, &a_full[as], dst, k * KCH, c, y0 - 1, img);
            } else {
              if (rank == 0) mbar_expect_tx(&a_full[as], 2 * SW_SLAB_BYTES);
              if (a.planar) tma_load_5d_pair(&tmap_x, &a_full[as], dst, 0, c, y0 - 1, img, k);
              else tma_load_4d_pair(&tmap_x, &a_full[as], dst, k * KCH, c, y0 - 1, img);
            }
          }
          __syncwarp();
          if (++as == a.na) { as = 0; aph ^= 1; }
        }
    }
    if (fetcher && elect_one()) {                                      // the last pair to finish re-arms the queue
      __threadfence();
      if (atomicAdd(a.sched + 1, 1u) == (unsigned)ncl - 1) {
        a.sched[0] = 0;
        a.sched[1] = 0;
        __threadfence();
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA of a pair)
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc(TILE_M * CG, C::N);
      constexpr uint32_t hi = desc_hi(1024);
      int as = 0;
      uint32_t aph = 0;
      int k = 0;                                                       // position in the lap
      uint32_t lap_par = 0;
      bool ye_ready = false, af_ready = false;
      mbar_wait(w_full, 0);
      if (CG == 2) mbar_wait(w_pair, 0);
      const uint32_t w_lo = desc_lo(smem_u32(smem_w));
      long long tp[4] = {0, 0, 0, 0}, tc0 = 0, ncol = 0;
      const bool prof = (a.dbg & 32) != 0;
      const long long t_begin = clock64();
      unsigned long long g_begin;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_begin));
#define SRCGAN_TICK(i) if (prof) { const long long n_ = clock64(); tp[i] += n_ - tc0; tc0 = n_; }
      for (int ui = 0;; ++ui) {
        const int u = get_unit(ui);
        if (u < 0) break;
        SRCGAN_DECODE_SW2_RANGE(u)
        (void)sg_;
        for (int c = x_start - 1; c <= x_end; ++c) {
          if (prof) { tc0 = clock64(); ++ncol; }
          // blocks touched for the first time in this lap must have been drained (and zeroed) by the epilogues
          // (the states of this column's barriers were peeked while the previous column's MMAs were being issued)
          if (k == 0) { mbar_wait(&y_empty[0], lap_par ^ 1); mbar_wait(&y_empty[1], lap_par ^ 1); }
          if (!ye_ready) mbar_wait(&y_empty[k + 2], lap_par ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(k * BN);
          SRCGAN_TICK(0)
          for (int kc = 0; kc < a.nchunks; ++kc) {
            const int rem = a.cin - kc * KCH;
            const int ksteps = ((rem >= KCH ? KCH : rem) + 15) >> 4;     // channels past cin are TMA zero-filled
            if (!af_ready) mbar_wait(&a_full[as], aph);
            tc_fence_after();
            SRCGAN_TICK(1)
            const uint32_t a_lo = desc_lo(smem_u32(smem_a + (size_t)as * SW_SLAB_STRIDE));
            const uint32_t b_lo = w_lo + (uint32_t)(kc * (C::W_CHUNK_BYTES >> 4));
            if (elect_one()) {
              if (!(a.dbg & 2)) {
                if (ksteps == 4) {
#pragma unroll
                  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                      if (CG == 1) umma_bf16_w(d, a_lo + (uint32_t)(kh * 8 + ks * 2), hi,
                                               b_lo + (uint32_t)(kh * (C::W_KH_BYTES >> 4) + ks * 2), hi, idesc, 1u);
                      else umma_bf16_w2(d, a_lo + (uint32_t)(kh * 8 + ks * 2), hi,
                                        b_lo + (uint32_t)(kh * (C::W_KH_BYTES >> 4) + ks * 2), hi, idesc, 1u);
                    }
                } else {
#pragma unroll 1
                  for (int kh = 0; kh < 3; ++kh)
                    for (int ks = 0; ks < ksteps; ++ks) {
                      if (CG == 1) umma_bf16_w(d, a_lo + (uint32_t)(kh * 8 + ks * 2), hi,
                                               b_lo + (uint32_t)(kh * (C::W_KH_BYTES >> 4) + ks * 2), hi, idesc, 1u);
                      else umma_bf16_w2(d, a_lo + (uint32_t)(kh * 8 + ks * 2), hi,
                                        b_lo + (uint32_t)(kh * (C::W_KH_BYTES >> 4) + ks * 2), hi, idesc, 1u);
                    }
                }
              }
              if (kc == a.nchunks - 1) {
                // block k has received its last contribution; the lap's last column also completes its two tail blocks.
                // (The epilogue group that waits for this commit releases the column's slabs.)
                if (CG == 1) umma_commit(&y_full[k]); else umma_commit_pair(&y_full[k]);
                if (k == RUN - 1) {
                  if (CG == 1) { umma_commit(&y_full[RUN]); umma_commit(&y_full[RUN + 1]); }
                  else { umma_commit_pair(&y_full[RUN]); umma_commit_pair(&y_full[RUN + 1]); }
                }
              }
            }
            __syncwarp();
            if (++as == a.na) { as = 0; aph ^= 1; }
            // peek the next slab's barrier (and, after the last chunk, the next column's block) while the tensor pipe works
            af_ready = mbar_test_wait(&a_full[as], aph);
            if (kc == a.nchunks - 1) {
              const int kn = k + 1 == RUN ? 0 : k + 1;
              ye_ready = mbar_test_wait(&y_empty[kn + 2], (k + 1 == RUN ? lap_par ^ 1 : lap_par) ^ 1);
            }
            SRCGAN_TICK(2)
          }
          if (++k == RUN) { k = 0; lap_par ^= 1; }
        }
      }
      if (prof && (blockIdx.x % 16 == 0 || blockIdx.x == gridDim.x - 2) && lane == 0) {
        unsigned long long g_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
        printf("sweep2 mma (cta %d of %d, units %d, wseg %d): cols %lld  y_empty %lld  a_full %lld  issue+commit %lld (clk/col)  loop total %lld clk = %llu ns, began at %llu us\n",
               (int)blockIdx.x, (int)gridDim.x, a.num_units, a.wseg, ncol, tp[0] / ncol, tp[1] / ncol, tp[2] / ncol, clock64() - t_begin, g_end - g_begin,
               (g_begin / 1000ull) % 1000000ull);
      }
#undef SRCGAN_TICK
    }
  } else {
    // ---- epilogue: TMEM lane = image row y0 + q*32 + lane; group g takes the outputs with running index % SW2_GROUPS == g;
    // the four warps of a group never synchronise with each other (registers -> global memory, no staging).
    // Running output index o = running input index of the column it is centred on; lap r = o / RUN, m = o % RUN:
    // main block m + 1 of lap r, plus block 0 of lap r + 1 when m == RUN - 1, plus block RUN + 1 of lap r - 1 when m == 0.
    // Every output's wait set contains the commit that follows input column o + 1, so the group's first warp also
    // hands that column's slabs back to the producer (the MMA warp issues one commit per column, not two).
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;
    const uint32_t sbias_addr = smem_u32(sbias);
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool wide_st = (a.y_ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 31) == 0);
    pdl_wait();                                                        // side operands in, outputs the previous kernels may still read
    auto arrive_empty = [&](int blk) {
      if (CG == 1) mbar_arrive(&y_empty[blk]); else mbar_arrive_leader(&y_empty[blk]);
    };
    // slots of input column t: (t * nchunks + kc) % na
    auto release_slabs = [&](uint32_t t) {
      if (q == 0 && lane == 0) {
        uint32_t s = (t * (uint32_t)a.nchunks) % (uint32_t)a.na;
        for (int kc = 0; kc < a.nchunks; ++kc) {
          mbar_arrive(&a_empty[s]);
          if (++s == (uint32_t)a.na) s = 0;
        }
      }
    };
    if (g == SW2_GROUPS - 1) {                                         // "output -1": block 0 of lap 0 collects only a halo tap
      mbar_wait(&y_full[0], 0);
      tc_fence_after();
      release_slabs(0);
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) tmem_st32_zero(lane_base + cb);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_empty(0);
    }
    long long tp[4] = {0, 0, 0, 0}, tc0 = 0, ncol = 0;
    const bool prof = (a.dbg & 32) != 0;
#define SRCGAN_TICK(i) if (prof) { const long long n_ = clock64(); tp[i] += n_ - tc0; tc0 = n_; }
    uint32_t t0 = 0;                                                   // running index of the unit's first input column
    uint32_t r = 0;                                                    // lap and position of this group's next output
    int m = g;
    for (int ui = 0;; ++ui) {
      const int u = get_unit(ui);
      if (u < 0) break;
      SRCGAN_DECODE_SW2_UNIT(u)
      const bool last_unit = get_unit(ui + 1) < 0;
      const int n_in = x_end - x_start + 2;
      const int y = y0 + row;
      const bool row_ok = strip_ok && y < a.h && !(a.dbg & 1);
      for (int j = (g + SW2_GROUPS - (int)(t0 % (uint32_t)SW2_GROUPS)) % SW2_GROUPS; j < n_in; j += SW2_GROUPS) {
        if (last_unit && j == n_in - 1) break;                         // centred on the stream's last input: never completes
        const uint32_t o = t0 + (uint32_t)j;
        const int pb = m + 1;
        const uint32_t ppar = r & 1u;
        const int sb = m == RUN - 1 ? 0 : ((m == 0 && o > 0) ? RUN + 1 : -1);
        const uint32_t spar = ppar ^ 1u;                               // lap r + 1 or r - 1
        const bool real = j >= 1 && j <= n_in - 2;
        const int x = x_start + j - 1;
        // separator row of a tall image (the zero padding shared by two stacked images): stored as zeros
        const bool zrow = a.zero_period > 0 && ((a.tr ? x : y) % a.zero_period) == 0;
        if (prof) { tc0 = clock64(); ++ncol; }
        // residual / mask operands of this pixel do not depend on the accumulator: request them before waiting for it
        // (issued behind the wait they cost one DRAM round trip per column and group - the dgrad launches all have a mask)
        const long long pix = (long long)img * a.img_stride + (long long)y * a.lane_stride + (long long)x * a.sweep_stride;
        uint4 pr1[SIDE ? 4 : 1], pr2[SIDE ? 4 : 1], pmk[SIDE ? 4 : 1];
        uint32_t pbits = 0;
        auto fetch_side = [&](int cb) {
          if (a.maskbits) pbits = __ldg(a.maskbits + pix * (BN / 32) + (cb >> 5));
          if constexpr (SIDE) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              if (a.r1) pr1[gq] = __ldg(reinterpret_cast<const uint4*>(a.r1 + pix * a.r1_ld + cb + gq * 8));
              if (a.r2) pr2[gq] = __ldg(reinterpret_cast<const uint4*>(a.r2 + pix * a.r2_ld + cb + gq * 8));
              if (a.mask) pmk[gq] = __ldg(reinterpret_cast<const uint4*>(a.mask + pix * a.mask_ld + cb + gq * 8));
            }
          }
        };
        if (real && row_ok) fetch_side(0);
        mbar_wait(&y_full[pb], ppar);
        if (sb >= 0) mbar_wait(&y_full[sb], spar);
        tc_fence_after();
        SRCGAN_TICK(0)
        release_slabs(o + 1);
        const uint32_t taddr = lane_base + (uint32_t)(pb * BN);
        const uint32_t saddr = lane_base + (uint32_t)((sb >= 0 ? sb : 0) * BN);
        if (!real) {
#pragma unroll 1
          for (int cb = 0; cb < BN; cb += 32) {
            tmem_st32_zero(taddr + cb);
            if (sb >= 0) tmem_st32_zero(saddr + cb);
          }
        } else {
#pragma unroll 1
          for (int cb = 0; cb < BN; cb += 32) {
            uint32_t z[32];
            float f[32];

            if (sb >= 0) {
              uint32_t z2[32];
              tmem_ld32_nowait(taddr + cb, z);
              tmem_ld32_nowait(saddr + cb, z2);
              tmem_ld_wait();
              tmem_st32_zero(taddr + cb);
              tmem_st32_zero(saddr + cb);
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(z[i]) + __uint_as_float(z2[i]);
            } else {
              tmem_ld32(taddr + cb, z);
              tmem_st32_zero(taddr + cb);                              // the block is reused one lap later
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(z[i]);
            }
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = lds_f4(sbias_addr + (uint32_t)(cb + i) * 4);
              f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
            }
            if (a.act) {                                               // LeakyReLU, 0 <= slope <= 1: max(f, f * slope)
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], f[i] * a.act_slope);
            }
            if (a.alpha != 1.f) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] *= a.alpha;
            }
            if (row_ok && a.maskbits) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] *= ((pbits >> i) & 1u) ? 1.f : a.mask_slope;
            }
            if (row_ok && a.signbits) {
              uint32_t sb32 = 0;
#pragma unroll
              for (int i = 0; i < 32; ++i) sb32 |= (f[i] > 0.f ? 1u : 0u) << i;
              a.signbits[pix * (BN / 32) + (cb >> 5)] = zrow ? 0u : sb32;
            }
            if (row_ok) {
              uint4 ov[4];
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                if constexpr (SIDE) {
                  if (a.r1) {
                    float rr[8];
                    unpack8(pr1[gq], rr);
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[gq * 8 + i] = fmaf(a.beta1, rr[i], f[gq * 8 + i]);
                  }
                  if (a.r2) {
                    float rr[8];
                    unpack8(pr2[gq], rr);
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[gq * 8 + i] = fmaf(a.beta2, rr[i], f[gq * 8 + i]);
                  }
                  if (a.mask) {
                    float mm[8];
                    unpack8(pmk[gq], mm);
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[gq * 8 + i] *= (mm[i] > 0.f ? 1.f : a.mask_slope);
                  }
                }
                __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ov[gq]);
#pragma unroll
                for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(f[gq * 8 + 2 * i], f[gq * 8 + 2 * i + 1]);
                if (zrow) ov[gq] = make_uint4(0u, 0u, 0u, 0u);
              }
              __nv_bfloat16* dst = a.y + pix * a.y_ld + cb;            // this pixel's 32 channels: 64 contiguous bytes
              if (wide_st) {
                stg_u8(dst, ov[0], ov[1]);
                stg_u8(dst + 16, ov[2], ov[3]);
              } else {
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) *reinterpret_cast<uint4*>(dst + gq * 8) = ov[gq];
              }
              if (cb + 32 < BN) fetch_side(cb + 32);                   // in flight during the next chunk's TMEM load
            }
          }
        }
        SRCGAN_TICK(1)
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                               // drained and zeroed (one arrival per warp)
          arrive_empty(pb);
          if (sb >= 0) arrive_empty(sb);
        }
        SRCGAN_TICK(2)
        m += SW2_GROUPS;
        if (m >= RUN) { m -= RUN; ++r; }
      }
      t0 += (uint32_t)n_in;
    }
    if (prof && blockIdx.x == 0 && lane == 0 && ncol)
      printf("sweep2 epi warp %d: cols %lld  y_full %lld  ld+math+store %lld  st_wait+arrive %lld (clk/col)\n", warp, ncol, tp[0] / ncol,
             tp[1] / ncol, tp[2] / ncol);
#undef SRCGAN_TICK
  }
#undef SRCGAN_DECODE_SW2_UNIT
#undef SRCGAN_DECODE_SW2_RANGE
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                                     // the peer may still signal this CTA's barriers
  if (warp == 1) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

#include "conv_pair.cuh"

// fp32 OIHW -> bf16 [n_block][chunk][slot][BN][64]: each [BN][64] tile is stored as its SWIZZLE_128B
// shared-memory image (16-byte chunk j of row r lives at chunk j ^ (r & 7)), zero padded in K.
// Rows are the GEMM-N channels, columns the GEMM-K channels: (co, ci) for fprop, (ci, co) for dgrad.
struct PackArgs {
  int n_total, k_total, bn, nchunks, total_slots;
  long long nstride, kstride;            // element strides of the N / K channel index in the OIHW tensor
  int slot_off[MAX_SLOTS];               // kh*KW + kw of each slot
};
__global__ void pack_weights_tc(const float* __restrict__ w, const PackArgs a, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)((a.n_total + a.bn - 1) / a.bn) * a.nchunks * a.total_slots * a.bn * KCH;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = (int)(i % KCH);
  long long t = i / KCH;
  const int r = (int)(t % a.bn); t /= a.bn;
  const int slot = (int)(t % a.total_slots); t /= a.total_slots;
  const int c = (int)(t % a.nchunks);
  const int nb = (int)(t / a.nchunks);
  const int nch = nb * a.bn + r, kch = c * KCH + j;
  float v = 0.f;
  if (kch < a.k_total && nch < a.n_total) v = w[nch * a.nstride + kch * a.kstride + a.slot_off[slot]];
  const long long tile = i - (long long)r * KCH - j;   // start of this [BN][64] tile
  const int chunk16 = (j >> 3) ^ (r & 7);
  out[tile + (long long)r * KCH + chunk16 * 8 + (j & 7)] = __float2bfloat16_rn(v);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// NHWC bf16 tensor (c channels at ptr, pixel pitch ld) -> tensor map whose box is the sampled slab
// [rows][8 px][64 ch]; sample = 1 (dense) or 2 (every other pixel, for stride-2 convolutions).
static int make_tmap(CUtensorMap* tm, const void* ptr, int c, int w, int h, int n, int ld, int rows, int sample,
                     const char* what, int box_w = TILE_W, bool transposed = false, bool promote256 = true) {
  EncodeTiledFn encode = get_encode_fn();
  SRCGAN_REQUIRE(encode != nullptr, "%s: cuTensorMapEncodeTiled is not available from the driver", what);
  cuuint64_t gdim[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * w, (cuuint64_t)ld * 2 * w * h};
  if (transposed) {                       // coordinate 1 walks image rows, coordinate 2 walks pixels of a row
    gdim[1] = (cuuint64_t)h; gdim[2] = (cuuint64_t)w;
    gstr[0] = (cuuint64_t)ld * 2 * w; gstr[1] = (cuuint64_t)ld * 2;
  }
  cuuint32_t box[4] = {(cuuint32_t)KCH, (cuuint32_t)(box_w * sample), (cuuint32_t)(rows * sample), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)sample, (cuuint32_t)sample, 1};
  CUresult cr = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       promote256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)cr);
    return SRCGAN_E_CUDA;
  }
  return SRCGAN_OK;
}

// Planar concat buffer [groups][n][h][w][64 ch] (group g at ptr + g * group_stride elements, pixel pitch ld = 64) -> 5-D tensor
// map whose box is one sweep slab [rows lanes][64 ch] of one group; coordinate order (ch, lane-or-sweep, sweep-or-lane, image,
// group) like make_tmap, `transposed` swaps the two image coordinates.  `c` = channels of the slice: the last group may be
// half used - its unused channels are loaded but never multiplied (k-steps stop at cin).
static int make_tmap_planar(CUtensorMap* tm, const void* ptr, int c, int w, int h, int n, int ld, long long group_stride, int rows,
                            const char* what, bool transposed) {
  EncodeTiledFn encode = get_encode_fn();
  SRCGAN_REQUIRE(encode != nullptr, "%s: cuTensorMapEncodeTiled is not available from the driver", what);
  SRCGAN_REQUIRE(ld == KCH && group_stride >= (long long)n * h * w * ld && group_stride % 8 == 0,
                 "%s: a planar buffer has 64-channel groups (ld 64) at a 16-byte aligned stride", what);
  const cuuint64_t groups = (cuuint64_t)((c + KCH - 1) / KCH);
  cuuint64_t gdim[5] = {(cuuint64_t)KCH, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n, groups};
  cuuint64_t gstr[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * w, (cuuint64_t)ld * 2 * w * h, (cuuint64_t)group_stride * 2};
  if (transposed) {
    gdim[1] = (cuuint64_t)h; gdim[2] = (cuuint64_t)w;
    gstr[0] = (cuuint64_t)ld * 2 * w; gstr[1] = (cuuint64_t)ld * 2;
  }
  cuuint32_t box[5] = {(cuuint32_t)KCH, 1, (cuuint32_t)rows, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult cr = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled (5-D) failed with CUresult %d", what, (int)cr);
    return SRCGAN_E_CUDA;
  }
  return SRCGAN_OK;
}

static int bn_for(int cout) { return cout >= 128 ? 128 : (cout <= 16 ? 16 : cout); }   // thin outputs pad to N=16
static long long supertiles(int tiles_x, int bn) {
  int mt = 256 / bn;
  if (mt > 8) mt = 8;
  return (tiles_x + mt - 1) / mt;
}

struct HostPlan {
  Plan plan;
  int maxt;
  int slot_kh[MAX_SLOTS], slot_kw[MAX_SLOTS];
};

// forward convolution KxK, stride s (1|2), padding p
static HostPlan fprop_plan(int K, int s, int p) {
  HostPlan hp = {};
  Plan& pl = hp.plan;
  pl.in_scale = s;
  int slot = 0;
  if (s == 1) {
    hp.maxt = K < 2 ? 2 : K;             // 1x1: the two-tap instantiations, one tap used (plan.ntaps) - one slab row in vain
    for (int kw = 0; kw < K; ++kw) {
      pl.dx[pl.nloads] = kw - p; pl.dy[pl.nloads] = -p; pl.ntaps[pl.nloads] = K; pl.slot0[pl.nloads] = slot;
      for (int kh = 0; kh < K; ++kh) { hp.slot_kh[slot] = kh; hp.slot_kw[slot] = kw; ++slot; }
      ++pl.nloads;
    }
  } else {
    hp.maxt = 2;
    for (int kw = 0; kw < K; ++kw)
      for (int first = 0; first < 2 && first < K; ++first) {       // kh = first, first+2  (same row parity)
        int l = pl.nloads++;
        pl.dx[l] = kw - p; pl.dy[l] = first - p; pl.slot0[l] = slot; pl.ntaps[l] = 0;
        for (int kh = first; kh < K; kh += 2) { hp.slot_kh[slot] = kh; hp.slot_kw[slot] = kw; ++slot; ++pl.ntaps[l]; }
      }
  }
  pl.total_slots = slot;
  return hp;
}

// output phase (a,b) of the stride-2 dgrad of a KxK pad-p convolution: a stride-1 conv over dY
static HostPlan dgrad2_phase_plan(int K, int p, int a, int b) {
  HostPlan hp = {};
  Plan& pl = hp.plan;
  pl.in_scale = 1;
  hp.maxt = 2;
  int khs[4], nkh = 0, kws[4], nkw = 0;
  for (int kh = K - 1; kh >= 0; --kh)              // descending kh = ascending row offset (a+p-kh)/2
    if (((a + p - kh) & 1) == 0) khs[nkh++] = kh;
  for (int kw = K - 1; kw >= 0; --kw)
    if (((b + p - kw) & 1) == 0) kws[nkw++] = kw;
  int slot = 0;
  for (int i = 0; i < nkw; ++i) {
    int l = pl.nloads++;
    pl.dx[l] = (b + p - kws[i]) / 2;               // exact: numerator is even (may be negative)
    pl.dy[l] = nkh ? (a + p - khs[0]) / 2 : 0;
    pl.ntaps[l] = nkh; pl.slot0[l] = slot;
    for (int j = 0; j < nkh; ++j) { hp.slot_kh[slot] = khs[j]; hp.slot_kw[slot] = kws[i]; ++slot; }
  }
  pl.total_slots = slot;
  return hp;
}

static size_t packed_elems(int n_total, int k_total, int slots) {
  { const int bn = bn_for(n_total); n_total = (n_total + bn - 1) / bn * bn; }
  return (size_t)n_total * ((k_total + KCH - 1) / KCH) * KCH * slots;
}

static int pack_launch(const float* w, int n_total, int k_total, long long nstride, long long kstride, int KW,
                       const HostPlan& hp, __nv_bfloat16* out, cudaStream_t st) {
  PackArgs a;
  a.n_total = n_total; a.k_total = k_total; a.bn = bn_for(n_total); a.nchunks = (k_total + KCH - 1) / KCH;
  a.total_slots = hp.plan.total_slots; a.nstride = nstride; a.kstride = kstride;
  for (int s = 0; s < MAX_SLOTS; ++s) a.slot_off[s] = s < a.total_slots ? hp.slot_kh[s] * KW + hp.slot_kw[s] : 0;
  const long long total = (long long)packed_elems(n_total, k_total, a.total_slots);
  if (total == 0) return SRCGAN_OK;
  pack_weights_tc<<<ceil_div(total, 256), 256, 0, st>>>(w, a, out);
  count_launch();
  return check_launch("pack_weights_tc");
}

template <int BN, int MAXT>
static int launch(const CUtensorMap& tmap, const TcArgs& a, cudaStream_t st) {
  using C = Cfg<BN, MAXT>;
  static DeviceOnce attr_set;
  int attr_set_dev;
  if (attr_set.needed(&attr_set_dev)) {
    SRCGAN_CUDA(cudaFuncSetAttribute(conv_igemm_tc<BN, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
    attr_set.mark(attr_set_dev);
  }
  long long grid = a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs;
  conv_igemm_tc<BN, MAXT><<<(unsigned)grid, IG_THREADS, C::SMEM_BYTES, st>>>(tmap, a);
  count_launch();
  return check_launch(BN == 128 ? "conv_igemm_tc<128>" : (BN == 64 ? "conv_igemm_tc<64>" : "conv_igemm_tc<32>"));
}

template <int BN>
static int launch_halo(const CUtensorMap& tmap, const TcArgs& a, cudaStream_t st) {
  using C = HCfg<BN>;
  static DeviceOnce attr_set;
  int attr_set_dev;
  if (attr_set.needed(&attr_set_dev)) {
    SRCGAN_CUDA(cudaFuncSetAttribute(conv3x3_halo_tc<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set.mark(attr_set_dev);
  }
  long long grid = a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs;
  conv3x3_halo_tc<BN><<<(unsigned)grid, NUM_THREADS, C::SMEM_BYTES, st>>>(tmap, a);
  count_launch();
  return check_launch(BN == 64 ? "conv3x3_halo_tc<64>" : (BN == 32 ? "conv3x3_halo_tc<32>" : "conv3x3_halo_tc<16>"));
}


// NHWC bf16 output slice (c channels at ptr, pixel pitch ld) -> tensor map whose box is one kw-stacked tile's
// output [rows][30 px][c ch]; the shared-memory image is swizzled (128B for 64 channels, 64B for 32).
static int make_tmap_out(CUtensorMap* tm, const void* ptr, int c, int w, int h, int n, int ld, int rows, const char* what,
                         int box_w = KW_VX, bool transposed = false) {
  EncodeTiledFn encode = get_encode_fn();
  SRCGAN_REQUIRE(encode != nullptr, "%s: cuTensorMapEncodeTiled is not available from the driver", what);
  cuuint64_t gdim[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * w, (cuuint64_t)ld * 2 * w * h};
  if (transposed) {
    gdim[1] = (cuuint64_t)h; gdim[2] = (cuuint64_t)w;
    gstr[0] = (cuuint64_t)ld * 2 * w; gstr[1] = (cuuint64_t)ld * 2;
  }
  cuuint32_t box[4] = {(cuuint32_t)c, (cuuint32_t)box_w, (cuuint32_t)rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled(out) failed with CUresult %d", what, (int)cr);
    return SRCGAN_E_CUDA;
  }
  return SRCGAN_OK;
}

// ring depth that fits next to the resident weights; 0 = does not fit
template <int BN, int R>
static int kws_ring_depth(int nchunks) {
  using C = KwCfg<BN, R>;
  const long long fixed = (long long)nchunks * C::W_CHUNK_BYTES + C::NSTAGE * C::STAGE_BYTES + SMEM_AUX + 1024;
  long long na = (SMEM_BUDGET - fixed) / C::SLAB_BYTES;
  return na > 8 ? 8 : (int)(na < 0 ? 0 : na);
}

template <int BN, int R>
static int launch_kws(const CUtensorMap& tx, const CUtensorMap& ty, KwArgs& a, cudaStream_t st) {
  using C = KwCfg<BN, R>;
  a.na = kws_ring_depth<BN, R>(a.nchunks);
  a.tiles_x = (a.w + KW_VX - 1) / KW_VX;
  a.tiles_y = (a.h + KW_TH * R - 1) / (KW_TH * R);
  a.num_tiles = (long long)a.tiles_x * a.tiles_y * a.n;
  static DeviceOnce attr_set;
  int attr_set_dev;
  if (attr_set.needed(&attr_set_dev)) {
    SRCGAN_CUDA(cudaFuncSetAttribute(conv3x3_kws_tc<BN, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
    attr_set.mark(attr_set_dev);
  }
  long long grid = a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs;
  conv3x3_kws_tc<BN, R><<<(unsigned)grid, KW_THREADS, C::smem_bytes(a.nchunks, a.na), st>>>(tx, ty, a);
  count_launch();
  return check_launch(BN == 64 ? "conv3x3_kws_tc<64>" : "conv3x3_kws_tc<32>");
}


template <int BN, int CG>
static int sweep2_ring_depth(int nchunks) {
  using C = Sw2Cfg<BN, CG>;
  const long long fixed = (long long)nchunks * C::W_CHUNK_BYTES + SMEM_AUX + 1024;
  long long na = (SMEM_BUDGET - fixed) / SW_SLAB_STRIDE;
  return na > SW_MAX_NA ? SW_MAX_NA : (int)(na < 0 ? 0 : na);
}

// co-resident CTA pairs (clusters of CG) for the kernel's shared-memory footprint; queried once per instantiation
template <int BN, int CG, int G, bool SIDE>
static int sweep2_max_clusters(size_t smem) {
  static int cached = 0;
  if (cached) return cached;
  int ncl = kNumSMs / CG;
  if (CG > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)kNumSMs); cfg.blockDim = dim3(sw2_threads(G)); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int q = 0;
    const cudaError_t qe = cudaOccupancyMaxActiveClusters(&q, conv3x3_sweep2_tc<BN, CG, G, SIDE>, &cfg);
    if (getenv("SRCGAN_B200_DBG")) fprintf(stderr, "sweep2<%d,%d>: cudaOccupancyMaxActiveClusters -> %s, %d clusters\n", BN, CG, cudaGetErrorString(qe), q);
    if (qe == cudaSuccess && q > 0 && q < ncl) ncl = q;
    (void)cudaGetLastError();
  }
  cached = ncl;
  return ncl;
}

// {next, done} counter pairs of the dynamic unit queue: a small per-device buffer owned by the library (allocated on first
// use, zero-initialised, re-armed by each launch's last CTA pair); launches take the slots round-robin
static unsigned int* sweep2_sched_slot() {
  constexpr int SLOTS = 64, MAXDEV = 16;
  static std::mutex mu;
  static unsigned int* base[MAXDEV] = {};
  static unsigned next[MAXDEV] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!base[dev]) {
    unsigned int* p = nullptr;
    if (cudaMalloc(&p, SLOTS * 2 * sizeof(unsigned int)) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    if (cudaMemset(p, 0, SLOTS * 2 * sizeof(unsigned int)) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    base[dev] = p;
  }
  return base[dev] + 2 * (next[dev]++ % SLOTS);
}

template <int BN, int CG, int G, bool SIDE>
static int launch_sweep2_g(const CUtensorMap& tx, Sw2Args& a, cudaStream_t st) {
  using C = Sw2Cfg<BN, CG>;
  a.na = sweep2_ring_depth<BN, CG>(a.nchunks);
  a.strips_y = (a.h + SW_ROWS - 1) / SW_ROWS;
  a.strips = a.n * a.strips_y;
  static DeviceOnce attr_set;
  int attr_set_dev;
  if (attr_set.needed(&attr_set_dev)) {
    SRCGAN_CUDA((cudaFuncSetAttribute(conv3x3_sweep2_tc<BN, CG, G, SIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET)));
    attr_set.mark(attr_set_dev);
  }
  const size_t smem = C::smem_bytes(a.nchunks, a.na);
  const int ncl_max = sweep2_max_clusters<BN, CG, G, SIDE>(SMEM_BUDGET);
  // segment width: fewest waves x (segment + 2 halo columns) over the co-resident clusters
  const long long groups = (a.strips + CG - 1) / CG;
  long long best = -1;
  for (int ws = 8; ws <= a.w; ++ws) {
    if (ws != a.w && ws % 2) continue;
    const long long segs = (a.w + ws - 1) / ws;
    const long long units = segs * groups;
    const long long waves = (units + ncl_max - 1) / ncl_max;
    const long long cost = waves * (ws + 2);
    if (best < 0 || cost < best) { best = cost; a.wseg = ws; a.segs_x = (int)segs; a.num_units = (int)units; }
  }
  int ncl = a.num_units < ncl_max ? a.num_units : ncl_max;
  // range mode: equal contiguous column ranges instead of whole units when that shortens the longest cluster's sweep
  // (every unit a range is cut into sweeps 2 halo columns)
  a.per = 0;
  a.total = (int)(groups * a.w);
  if (!getenv("SRCGAN_B200_NO_SWEEP_RANGES") && !getenv("SRCGAN_B200_SWEEP_DYNAMIC") && !a.rev) {
    int per = (a.total + ncl_max - 1) / ncl_max;
    if (per < 8) per = 8;
    const long long range_cost = per + 2 * ((per + a.w - 1) / a.w + 1);
    if (range_cost < best) { a.per = per; ncl = (a.total + per - 1) / per; }
  }
  // The dynamic unit queue is opt-in: which output columns straddle a lap of the accumulator ring (two partial blocks added
  // in the epilogue instead of one accumulation chain) depends on a pair's position in its unit sequence, so claiming units
  // at run time changes the fp32 summation order of a few columns from run to run (last-bit differences after the bf16
  // rounding).  Round-robin units keep the kernel bit-reproducible; the queue is worth +15 % on 64->32, +1 % on a dense block.
  a.sched = nullptr;
  if (a.per == 0 && a.num_units > ncl && (long long)ncl * (SW_MAXU - 1) >= a.num_units && getenv("SRCGAN_B200_SWEEP_DYNAMIC"))
    a.sched = sweep2_sched_slot();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ncl * CG)); cfg.blockDim = dim3(sw2_threads(G)); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1 + pdl_attr(&at[1]);
  SRCGAN_CUDA((cudaLaunchKernelEx(&cfg, conv3x3_sweep2_tc<BN, CG, G, SIDE>, tx, a)));
  count_launch();
  return check_launch(BN == 64 ? (CG == 2 ? "conv3x3_sweep2_tc<64,2>" : "conv3x3_sweep2_tc<64,1>")
                               : (CG == 2 ? "conv3x3_sweep2_tc<32,2>" : "conv3x3_sweep2_tc<32,1>"));
}
// epilogue groups (SRCGAN_B200_SWEEP_GROUPS=2 turns the third group off for A/B tests)
template <int BN, int CG>
static int launch_sweep2(const CUtensorMap& tx, Sw2Args& a, cudaStream_t st) {
  if constexpr (CG == 2 && BN == 64) {
    // 64 -> 64 layers (HRconv, upconv) are bound by the epilogue (1 152 clocks of MMAs per column against ~2 600 per group and
    // column): launches without residual / bf16-mask operands run the lean epilogue (122 registers) with THREE groups - 0.308 ->
    // 0.258 ms at 64 x 256^2.  With the side operands three groups spill (0.30 -> 0.48 ms), and the 32-channel layers lose with
    // three groups (0.236 -> 0.257 ms), so both stay at two.
    const char* e = getenv("SRCGAN_B200_SWEEP_GROUPS");
    const int groups = e ? atoi(e) : 3;
    if (!a.r1 && !a.r2 && !a.mask && groups == 3) return launch_sweep2_g<BN, CG, 3, false>(tx, a, st);
  }
  return launch_sweep2_g<BN, CG, 2, true>(tx, a, st);
}

// fused pair of dense-block layers (conv_pair.cuh): depth of the P slab ring that fits beside both layers' weights + the XK ring
static int swf_ring_depth(int nchunks) {
  const long long fixed = (long long)(2 * nchunks + 1) * SWF_W_CHUNK_BYTES + SWF_XK_BYTES + SMEM_AUX + 1024;
  long long na = (SMEM_BUDGET - fixed) / SW_SLAB_STRIDE;
  return na > SW_MAX_NA ? SW_MAX_NA : (int)(na < 0 ? 0 : na);
}
static int swf_max_clusters() {
  static int cached = 0;
  if (cached) return cached;
  int ncl = kNumSMs / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)kNumSMs); cfg.blockDim = dim3(SWF_THREADS); cfg.dynamicSmemBytes = SMEM_BUDGET;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int q = 0;
  const cudaError_t qe = cudaOccupancyMaxActiveClusters(&q, conv3x3_pair_sweep_tc, &cfg);
  if (getenv("SRCGAN_B200_DBG")) fprintf(stderr, "pair_sweep: cudaOccupancyMaxActiveClusters -> %s, %d clusters\n", cudaGetErrorString(qe), q);
  if (qe == cudaSuccess && q > 0 && q < ncl) ncl = q;
  (void)cudaGetLastError();
  cached = ncl;
  return ncl;
}
static int launch_pair_sweep(const CUtensorMap& tx, SwfArgs& a, cudaStream_t st) {
  a.na = swf_ring_depth(a.nchunks);
  static DeviceOnce attr_set;
  int attr_set_dev;
  if (attr_set.needed(&attr_set_dev)) {
    SRCGAN_CUDA(cudaFuncSetAttribute(conv3x3_pair_sweep_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
    attr_set.mark(attr_set_dev);
  }
  const int ncl_max = swf_max_clusters();
  a.total = a.n * a.w;
  a.per = (a.total + ncl_max - 1) / ncl_max;
  if (a.per < 8) a.per = 8;                                            // a unit sweeps 4 halo columns: keep ranges worth it
  const int ncl = (a.total + a.per - 1) / a.per;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ncl * 2)); cfg.blockDim = dim3(SWF_THREADS); cfg.dynamicSmemBytes = swf_smem_bytes(a.nchunks, a.na);
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1 + pdl_attr(&at[1]);
  SRCGAN_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_pair_sweep_tc, tx, a));
  count_launch();
  return check_launch("conv3x3_pair_sweep_tc");
}

static int dispatch(int bn, int maxt, const CUtensorMap& tmap, const TcArgs& a, cudaStream_t st) {
#define SRCGAN_TC_CASE(B, T) if (bn == B && maxt == T) return launch<B, T>(tmap, a, st)
  SRCGAN_TC_CASE(32, 2); SRCGAN_TC_CASE(64, 2); SRCGAN_TC_CASE(128, 2);
  SRCGAN_TC_CASE(32, 3); SRCGAN_TC_CASE(64, 3); SRCGAN_TC_CASE(128, 3);
  SRCGAN_TC_CASE(32, 4); SRCGAN_TC_CASE(64, 4); SRCGAN_TC_CASE(128, 4);
#undef SRCGAN_TC_CASE
  set_error("conv tc: no kernel for BN=%d MAXT=%d", bn, maxt);
  return SRCGAN_E_INVALID;
}

static void fill_epilogue(TcArgs& a, const srcgan_conv_params* p) {
  a.wgt = (const __nv_bfloat16*)p->wgt; a.bias = p->bias;
  a.y = (__nv_bfloat16*)p->y; a.y_ld = p->y_ld;
  a.act = p->act; a.act_slope = p->act_slope; a.alpha = p->alpha;
  a.r1 = (const __nv_bfloat16*)p->r1; a.r1_ld = p->r1_ld; a.beta1 = p->beta1;
  a.r2 = (const __nv_bfloat16*)p->r2; a.r2_ld = p->r2_ld; a.beta2 = p->beta2;
  a.mask = (const __nv_bfloat16*)p->mask; a.mask_ld = p->mask_ld; a.mask_slope = p->mask_slope;
}

}  // namespace tc

static bool tc_common_ok(const srcgan_conv_params* p) {
  if (p->dtype != SRCGAN_DT_BF16 || p->upsample) return false;
  if (p->kh != p->kw || (p->kh != 1 && p->kh != 3 && p->kh != 4)) return false;
  if (p->kh == 1 && p->pad != 0) return false;
  if (p->stride != 1 && p->stride != 2) return false;
  if (p->x_ld % 8 || p->y_ld % 8) return false;
  if (((uintptr_t)p->x) % 16 || ((uintptr_t)p->y) % 16) return false;
  if (p->r1 && (p->r1_ld % 8 || ((uintptr_t)p->r1) % 16)) return false;
  if (p->r2 && (p->r2_ld % 8 || ((uintptr_t)p->r2) % 16)) return false;
  if (p->mask && (p->mask_ld % 8 || ((uintptr_t)p->mask) % 16)) return false;
  return true;
}
static bool tc_nch_ok(int c) { return c == 32 || c == 64 || c == 128 || c == 256; }

bool conv_tc_supported(const srcgan_conv_params* p) {
  if (!tc_common_ok(p)) return false;
  // thin outputs (<= 16 channels, e.g. the 64->3 image conv): 3x3 stride-1 halo kernel only, plain epilogue
  if (p->cout <= 16) return p->kh == 3 && p->stride == 1 && !p->r1 && !p->r2 && !p->mask;
  return tc_nch_ok(p->cout);      // any cin: channels beyond cin are zero-filled by TMA (K padded to 16)
}
// stride-2 dgrad on the tensor-core engine (stride-1 dgrad is an fprop over transposed weights)
bool conv_dgrad_tc_supported(const srcgan_conv_params* p) {
  return tc_common_ok(p) && p->stride == 2 && p->cout >= 64 && p->cout % 16 == 0 && tc_nch_ok(p->cin);
}

size_t packed_weight_bytes_tc(int cout, int cin, int kh, int kw, int layout) {
  if (layout == SRCGAN_WL_TC) return tc::packed_elems(cout, cin, kh * kw) * sizeof(__nv_bfloat16);
  if (layout == SRCGAN_WL_TC_S2) return tc::packed_elems(cout, cin, kh * kw) * sizeof(__nv_bfloat16);
  return tc::packed_elems(cin, cout, kh * kw) * sizeof(__nv_bfloat16);     // SRCGAN_WL_TC_DGRAD_S2: 4 phases
}

// slot table + N tile of a SRCGAN_WL_TC pack, exactly what pack_weights_tc_host() below uses (for the batched packer)
int pack_slots_tc(int cout, int kh, int kw, int layout, int32_t* slot_off16, int32_t* nslots, int32_t* bn) {
  SRCGAN_REQUIRE(layout == SRCGAN_WL_TC && kh == kw, "pack_slots: only the stride-1 tcgen05 layout is batched");
  SRCGAN_REQUIRE(tc_nch_ok(cout) || cout <= 16, "pack_slots: cout %d unsupported", cout);
  tc::HostPlan hp = tc::fprop_plan(kh, 1, 0);
  if (kh == 3 && cout <= 64)
    for (int t = 0; t < 9; ++t) { hp.slot_kh[t] = t / 3; hp.slot_kw[t] = 2 - t % 3; }
  SRCGAN_REQUIRE(hp.plan.total_slots <= 16, "pack_slots: too many taps");
  for (int s2 = 0; s2 < 16; ++s2) slot_off16[s2] = s2 < hp.plan.total_slots ? hp.slot_kh[s2] * kw + hp.slot_kw[s2] : 0;
  *nslots = hp.plan.total_slots;
  *bn = tc::bn_for(cout);
  return SRCGAN_OK;
}

int pack_weights_tc_host(const float* w, int cout, int cin, int kh, int kw, int layout, void* out, cudaStream_t st) {
  SRCGAN_REQUIRE(kh == kw, "pack_weights(tc): square filters only");
  const long long taps = (long long)kh * kw;
  if (layout == SRCGAN_WL_TC || layout == SRCGAN_WL_TC_S2) {
    SRCGAN_REQUIRE(tc_nch_ok(cout) || (cout <= 16 && layout == SRCGAN_WL_TC), "pack_weights(tc): cout %d unsupported",
                   cout);
    tc::HostPlan hp = tc::fprop_plan(kh, layout == SRCGAN_WL_TC ? 1 : 2, 0);   // slot order does not depend on pad
    if (layout == SRCGAN_WL_TC && kh == 3 && cout <= 64)       // halo / stacked kernels: slot = kh*3 + (2 - kw)
      for (int t = 0; t < 9; ++t) { hp.slot_kh[t] = t / 3; hp.slot_kw[t] = 2 - t % 3; }
    return tc::pack_launch(w, cout, cin, (long long)cin * taps, taps, kw, hp, (__nv_bfloat16*)out, st);
  }
  SRCGAN_REQUIRE(layout == SRCGAN_WL_TC_DGRAD_S2, "pack_weights(tc): unknown layout %d", layout);
  SRCGAN_REQUIRE(tc_nch_ok(cin), "pack_weights(tc dgrad): cin %d unsupported", cin);
  __nv_bfloat16* o = (__nv_bfloat16*)out;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      // phase tap sets depend on pad only through parity; the packer is told the pad via kh's parity class:
      // caller convention: pad = 1 (the only padding the stride-2 layers of the reference use)
      tc::HostPlan hp = tc::dgrad2_phase_plan(kh, 1, a, b);
      int rc = tc::pack_launch(w, cin, cout, taps, (long long)cin * taps, kw, hp, o, st);
      if (rc) return rc;
      o += tc::packed_elems(cin, cout, hp.plan.total_slots);
    }
  return SRCGAN_OK;
}

// kw-stacked kernel: 3x3 stride 1 pad 1, cout 32 / 64, all weights resident in shared memory
static int kws_variant(const srcgan_conv_params* p) {
  if (p->kh != 3 || p->stride != 1 || p->pad != 1 || (p->cout != 32 && p->cout != 64)) return 0;
  if (p->y_ld % 8 || ((uintptr_t)p->y) % 16 || getenv("SRCGAN_B200_NO_KWS")) return 0;
  const int nchunks = (p->cin + tc::KCH - 1) / tc::KCH;
  if (p->cout == 64) return tc::kws_ring_depth<64, 1>(nchunks) >= 3 ? 641 : 0;
  if (tc::kws_ring_depth<32, 2>(nchunks) >= 3 && p->h > 4) return 322;
  return tc::kws_ring_depth<32, 1>(nchunks) >= 3 ? 321 : 0;
}

static int conv_fprop_kws(const srcgan_conv_params* p, int variant, cudaStream_t st) {
  const int rows = variant == 322 ? 2 : 1;
  CUtensorMap tx, ty;
  int rc = tc::make_tmap(&tx, p->x, p->cin, p->w, p->h, p->n, p->x_ld, tc::KW_TH * rows + 2, 1, "conv_fprop_tc(kws x)",
                         tc::KW_TW);
  if (rc) return rc;
  rc = tc::make_tmap_out(&ty, p->y, p->cout, p->wo, p->ho, p->n, p->y_ld, tc::KW_TH * rows, "conv_fprop_tc(kws y)");
  if (rc) return rc;
  tc::KwArgs a;
  a.n = p->n; a.cin = p->cin; a.cout = p->cout; a.h = p->ho; a.w = p->wo;
  a.nchunks = (p->cin + tc::KCH - 1) / tc::KCH;
  a.wgt = (const __nv_bfloat16*)p->wgt; a.bias = p->bias;
  a.act = p->act; a.act_slope = p->act_slope; a.alpha = p->alpha;
  a.r1 = (const __nv_bfloat16*)p->r1; a.r1_ld = p->r1_ld; a.beta1 = p->beta1;
  a.r2 = (const __nv_bfloat16*)p->r2; a.r2_ld = p->r2_ld; a.beta2 = p->beta2;
  a.mask = (const __nv_bfloat16*)p->mask; a.mask_ld = p->mask_ld; a.mask_slope = p->mask_slope;
  { const char* d = getenv("SRCGAN_B200_DBG"); a.dbg = d ? atoi(d) : 0; }
  if (variant == 641) return tc::launch_kws<64, 1>(tx, ty, a, st);
  return variant == 322 ? tc::launch_kws<32, 2>(tx, ty, a, st) : tc::launch_kws<32, 1>(tx, ty, a, st);
}

// paired column-sweep kernel (cta_group::2): same shapes as the sweep; the halved weight rows make 192 -> 64 resident
static int sweep2_cg(const srcgan_conv_params* p) {
  if (p->kh != 3 || p->stride != 1 || p->pad != 1 || (p->cout != 32 && p->cout != 64)) return 0;
  if (p->y_ld % 8 || ((uintptr_t)p->y) % 16 || p->h < 96 || getenv("SRCGAN_B200_NO_SWEEP2")) return 0;
  const int nchunks = (p->cin + tc::KCH - 1) / tc::KCH;
  const char* e = getenv("SRCGAN_B200_SWEEP2_CG");
  const int want = e ? atoi(e) : 2;
  // a column's slabs are handed back when its last chunk has been multiplied: the ring holds at least one column + 2 slabs
  if (want == 2) {
    if ((p->cout == 64 ? tc::sweep2_ring_depth<64, 2>(nchunks) : tc::sweep2_ring_depth<32, 2>(nchunks)) >= nchunks + 2) return 2;
    return 0;
  }
  return (p->cout == 64 ? tc::sweep2_ring_depth<64, 1>(nchunks) : tc::sweep2_ring_depth<32, 1>(nchunks)) >= nchunks + 2 ? 1 : 0;
}

static int conv_fprop_sweep2(const srcgan_conv_params* p, int cg, cudaStream_t st) {
  CUtensorMap tx;
  // orientation: lanes along the image rows' pixels (sweep over y) when the image is wide enough - a slab is then 130
  // neighbouring pixels and a warp's stores land on neighbouring pixels - else lanes along a column (sweep over x)
  bool tr = p->wo >= 96;
  { const char* e = getenv("SRCGAN_B200_SWEEP_TR"); if (e) tr = (atoi(e) != 0 && p->wo >= 96) || p->ho < 96; }
  int rc = p->x_group_stride
               ? tc::make_tmap_planar(&tx, p->x, p->cin, p->w, p->h, p->n, p->x_ld, p->x_group_stride, tc::SW_SLAB_ROWS,
                                      "conv_fprop_tc(sweep2 planar x)", tr)
               : tc::make_tmap(&tx, p->x, p->cin, p->w, p->h, p->n, p->x_ld, tc::SW_SLAB_ROWS, 1, "conv_fprop_tc(sweep2 x)", 1, tr,
                               /*promote256=*/p->cin % 128 == 0 && p->x_ld == p->cin);
  if (rc) return rc;
  tc::Sw2Args a;
  a.planar = p->x_group_stride ? 1 : 0;
  a.n = p->n; a.cin = p->cin; a.cout = p->cout;
  a.h = tr ? p->wo : p->ho; a.w = tr ? p->ho : p->wo;              // a.h = lane extent, a.w = sweep extent
  a.tr = tr ? 1 : 0;
  a.img_stride = (long long)p->ho * p->wo;
  a.lane_stride = tr ? 1 : p->wo; a.sweep_stride = tr ? p->wo : 1;
  a.y = (__nv_bfloat16*)p->y; a.y_ld = p->y_ld;
  a.nchunks = (p->cin + tc::KCH - 1) / tc::KCH;
  a.wgt = (const __nv_bfloat16*)p->wgt; a.bias = p->bias;
  a.act = p->act; a.act_slope = p->act_slope; a.alpha = p->alpha;
  a.r1 = (const __nv_bfloat16*)p->r1; a.r1_ld = p->r1_ld; a.beta1 = p->beta1;
  a.r2 = (const __nv_bfloat16*)p->r2; a.r2_ld = p->r2_ld; a.beta2 = p->beta2;
  a.mask = (const __nv_bfloat16*)p->mask; a.mask_ld = p->mask_ld; a.mask_slope = p->mask_slope;
  SRCGAN_REQUIRE(!((p->signbits || p->maskbits) && (p->r1 || p->r2 || p->mask)),
                 "conv_fprop_tc: packed sign / mask bits cannot be combined with residual or bf16 mask operands");
  a.signbits = (uint32_t*)p->signbits; a.maskbits = (const uint32_t*)p->maskbits;
  a.zero_period = p->zero_row_period;
  a.rev = (p->flags & SRCGAN_CONV_FLAG_REVERSE) ? 1 : 0;
  { const char* d = getenv("SRCGAN_B200_DBG"); a.dbg = d ? atoi(d) : 0; }
  { const char* d = getenv("SRCGAN_B200_PFD"); a.pfd = d ? atoi(d) : 0; }
  if (cg == 2) return p->cout == 64 ? tc::launch_sweep2<64, 2>(tx, a, st) : tc::launch_sweep2<32, 2>(tx, a, st);
  return p->cout == 64 ? tc::launch_sweep2<64, 1>(tx, a, st) : tc::launch_sweep2<32, 1>(tx, a, st);
}

// Two consecutive dense-block layers in one launch (conv_pair.cuh).  pa: P = x[:, :cin] -> the 32-channel slice right behind it;
// pb: [P | that slice] -> another 32-channel slice.  Epilogues: bias, LeakyReLU, packed masks in / out.
static const char* pair_why_not(const srcgan_conv_params* pa, const srcgan_conv_params* pb) {
  const srcgan_conv_params* ps[2] = {pa, pb};
  for (int i = 0; i < 2; ++i) {
    const srcgan_conv_params* p = ps[i];
    if (!tc_common_ok(p)) return "not a bf16 tcgen05 layer";
    if (p->kh != 3 || p->stride != 1 || p->pad != 1 || p->cout != 32) return "layers must be 3x3 stride 1 pad 1 with 32 output channels";
    if (p->r1 || p->r2 || p->mask) return "residual / bf16 mask operands are not supported (packed maskbits are)";
    if (p->alpha != 1.f) return "alpha must be 1";
    if (p->zero_row_period || p->x_group_stride) return "tall-image / planar buffers are not supported";
    if (p->y_ld % 8 || ((uintptr_t)p->y) % 16) return "output slice alignment";
  }
  if (pa->n != pb->n || pa->h != pb->h || pa->w != pb->w || pa->ho != pb->ho || pa->wo != pb->wo) return "geometry differs";
  if (pa->cin != 64 && pa->cin != 128) return "the shared prefix must have 64 or 128 channels";
  if (pb->cin != pa->cin + 32 || pb->x != pa->x || pb->x_ld != pa->x_ld) return "layer B must read [prefix | layer A's output]";
  if ((const char*)pa->y != (const char*)pa->x + (size_t)pa->cin * 2 || pa->y_ld != pa->x_ld)
    return "layer A must write the 32-channel slice right behind the prefix";
  const bool tr = pa->wo >= 96;
  const int lanes = tr ? pa->wo : pa->ho;
  if (lanes > 2 * tc::SW_ROWS) return "lane extent above 256";
  if (lanes < 96) return "map too small";
  if (tc::swf_ring_depth(pa->cin / tc::KCH) < pa->cin / tc::KCH + 1) return "shared memory";
  return nullptr;
}
bool conv_fprop_pair_supported(const srcgan_conv_params* pa, const srcgan_conv_params* pb) { return pair_why_not(pa, pb) == nullptr; }

int conv_fprop_pair_tc(const srcgan_conv_params* pa, const srcgan_conv_params* pb, cudaStream_t st) {
  const char* why = pair_why_not(pa, pb);
  SRCGAN_REQUIRE(why == nullptr, "conv_fprop_pair: %s", why ? why : "");
  const bool tr = pa->wo >= 96;
  CUtensorMap tx;
  int rc = tc::make_tmap(&tx, pa->x, pa->cin, pa->w, pa->h, pa->n, pa->x_ld, tc::SW_SLAB_ROWS, 1, "conv_fprop_pair(x)", 1, tr,
                         /*promote256=*/pa->cin % 128 == 0 && pa->x_ld == pa->cin);
  if (rc) return rc;
  tc::SwfArgs a;
  a.n = pa->n; a.cin = pa->cin;
  a.h = tr ? pa->wo : pa->ho; a.w = tr ? pa->ho : pa->wo;
  a.tr = tr ? 1 : 0;
  a.img_stride = (long long)pa->ho * pa->wo;
  a.lane_stride = tr ? 1 : pa->wo; a.sweep_stride = tr ? pa->wo : 1;
  a.nchunks = pa->cin / tc::KCH;
  a.wgt_a = (const __nv_bfloat16*)pa->wgt; a.wgt_b = (const __nv_bfloat16*)pb->wgt;
  const srcgan_conv_params* ps[2] = {pa, pb};
  for (int i = 0; i < 2; ++i) {
    tc::SwfLayer& L = a.L[i];
    L.bias = ps[i]->bias;
    L.y = (__nv_bfloat16*)ps[i]->y; L.y_ld = ps[i]->y_ld;
    L.act = ps[i]->act; L.act_slope = ps[i]->act_slope;
    L.signbits = (uint32_t*)ps[i]->signbits; L.maskbits = (const uint32_t*)ps[i]->maskbits;
    L.mask_slope = ps[i]->mask_slope;
  }
  { const char* d = getenv("SRCGAN_B200_DBG"); a.dbg = d ? atoi(d) : 0; }
  { const char* d = getenv("SRCGAN_B200_PAIR_PFD"); a.pfd = d ? atoi(d) : 0; }
  { const char* d = getenv("SRCGAN_B200_PAIR_ISSUERS"); a.issuers = (d && atoi(d) == 1) ? 1 : 2; }
  return tc::launch_pair_sweep(tx, a, st);
}

int conv_fprop_tc(const srcgan_conv_params* p, cudaStream_t st) {
  if (const int cg = sweep2_cg(p)) return conv_fprop_sweep2(p, cg, st);
  SRCGAN_REQUIRE(!p->signbits && !p->maskbits,
                 "conv_fprop_tc: packed sign / mask bits need the paired-sweep kernel (3x3 s1 p1, cout 32/64, map >= 96)");
  SRCGAN_REQUIRE(p->zero_row_period == 0,
                 "conv_fprop_tc: zero_row_period (tall-image separators) needs the paired-sweep kernel (3x3 s1 p1, cout 32/64)");
  SRCGAN_REQUIRE(p->x_group_stride == 0,
                 "conv_fprop_tc: a planar input (x_group_stride) needs the paired-sweep kernel (3x3 s1 p1, cout 32/64, map >= 96)");
  if (const int v = kws_variant(p)) return conv_fprop_kws(p, v, st);
  tc::HostPlan hp = tc::fprop_plan(p->kh, p->stride, p->pad);
  const bool halo = p->kh == 3 && p->stride == 1 && (p->cout <= 16 || p->cout == 32 || p->cout == 64);
  CUtensorMap tmap;
  int rc = halo ? tc::make_tmap(&tmap, p->x, p->cin, p->w, p->h, p->n, p->x_ld, tc::HALO_H, 1, "conv_fprop_tc",
                                tc::HALO_W)
                : tc::make_tmap(&tmap, p->x, p->cin, p->w, p->h, p->n, p->x_ld, tc::TILE_H + hp.maxt - 1, p->stride,
                                "conv_fprop_tc");
  if (rc) return rc;
  tc::TcArgs a;
  a.n = p->n; a.cin = p->cin; a.cout = p->cout;
  a.gh = p->ho; a.gw = p->wo; a.oh = p->ho; a.ow = p->wo; a.os = 1; a.oa = 0; a.ob = 0;
  a.plan = hp.plan;
  a.nchunks = (p->cin + tc::KCH - 1) / tc::KCH;
  const int bn = tc::bn_for(p->cout);
  a.n_blocks = (p->cout + bn - 1) / bn;
  a.tiles_x = (a.gw + tc::TILE_W - 1) / tc::TILE_W;
  a.tiles_y = (a.gh + tc::TILE_H - 1) / tc::TILE_H;
  a.num_tiles = tc::supertiles(a.tiles_x, bn) * a.tiles_y * p->n * a.n_blocks;
  tc::fill_epilogue(a, p);
  { const char* d = getenv("SRCGAN_B200_DBG"); a.dbg = d ? atoi(d) : 0; }
  if (halo) {
    if (bn == 16) return tc::launch_halo<16>(tmap, a, st);
    return bn == 32 ? tc::launch_halo<32>(tmap, a, st) : tc::launch_halo<64>(tmap, a, st);
  }
  return tc::dispatch(bn, hp.maxt, tmap, a, st);
}

// stride-2 dgrad: dX (h x w x cin) from dY (ho x wo x cout); x = dY, y = dX, wgt = SRCGAN_WL_TC_DGRAD_S2 pack
int conv_dgrad_tc(const srcgan_conv_params* p, cudaStream_t st) {
  SRCGAN_REQUIRE(p->pad == 1, "conv_dgrad_tc: stride-2 dgrad is built for pad 1 (got %d)", p->pad);
  CUtensorMap tmap;
  int rc = tc::make_tmap(&tmap, p->x, p->cout, p->wo, p->ho, p->n, p->x_ld, tc::TILE_H + 1, 1, "conv_dgrad_tc");
  if (rc) return rc;
  const int bn = tc::bn_for(p->cin);
  const __nv_bfloat16* w = (const __nv_bfloat16*)p->wgt;
  for (int pa = 0; pa < 2; ++pa)
    for (int pb = 0; pb < 2; ++pb) {
      tc::HostPlan hp = tc::dgrad2_phase_plan(p->kh, p->pad, pa, pb);
      tc::TcArgs a;
      a.dbg = 0;
      a.n = p->n; a.cin = p->cout; a.cout = p->cin;          // GEMM K channels = dY channels, N = dX channels
      a.gh = (p->h - pa + 1) / 2; a.gw = (p->w - pb + 1) / 2;
      a.oh = p->h; a.ow = p->w; a.os = 2; a.oa = pa; a.ob = pb;
      a.plan = hp.plan;
      a.nchunks = (p->cout + tc::KCH - 1) / tc::KCH;
      a.n_blocks = p->cin / bn;
      a.tiles_x = (a.gw + tc::TILE_W - 1) / tc::TILE_W;
      a.tiles_y = (a.gh + tc::TILE_H - 1) / tc::TILE_H;
      a.num_tiles = tc::supertiles(a.tiles_x, bn) * a.tiles_y * p->n * a.n_blocks;
      tc::fill_epilogue(a, p);
      a.wgt = w;
      a.bias = nullptr;
      if (a.num_tiles > 0 && hp.plan.total_slots > 0) {
        rc = tc::dispatch(bn, 2, tmap, a, st);
        if (rc) return rc;
      }
      w += tc::packed_elems(p->cin, p->cout, hp.plan.total_slots);
    }
  return SRCGAN_OK;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 wgrad:  dW[co][ci][kh][kw] = sum_{n,y,x} X[n,y+kh-pad,x+kw-pad,ci] * dY[n,y,x,co]
//
// GEMM view per tap: D[ci][co] += X_tap^T[ci][pixels] * dY[pixels][co] :  M = 128 input channels (two
// 64-channel swizzle atoms), N = BN output channels, K = pixels.  Both operands are "MN-major" for the
// tensor core: NHWC keeps the channel (M resp. N) contiguous and the pixel (K) strided, which is
// exactly the SWIZZLE_128B slab a TMA box load produces - no transposes anywhere.
// One CTA owns (kw, 128-channel ci block, co block, pixel split): it streams its pixel tiles
// (16x8 pixels = 8 MMA K-steps) through a TMA ring, accumulates the KH taps of its column into KH
// TMEM accumulators (a tap's row shift is again +1024 B on the descriptor), and finally writes fp32
// partials [split][tap][ci][co]; a fixed-order reduce kernel sums the splits into the OIHW gradient.
// ---------------------------------------------------------------------------------------------
namespace tcw {
using namespace tc;

struct WgArgs {
  int n, ho, wo, cin, cout;
  Plan plan;                 // the forward convolution's load plan (X slabs; stride-sampled for s=2)
  int tapidx[MAX_SLOTS];     // slot -> kh*KW + kw
  int ktaps;                 // KH*KW
  int cblocks, nblocks, splits;
  int tiles_x, tiles_y;
  long long num_tiles, tiles_per_split;
  float* part;        // [splits][KH*KW][cin][cout]
};

// MN-major SWIZZLE_128B descriptor: 64-element (128 B) rows along M/N, 8 K-rows per 1024-B atom,
// next 8 K-rows at SBO, next 64 M/N elements at LBO.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_mn(int M, int N) {
  return umma_idesc(M, N) | (1u << 15) | (1u << 16);      // A and B MN-major
}

template <int BN, int KH>      // KH = max taps per load (MAXT)
struct WCfg {
  static constexpr int A_SLAB = (TILE_H + KH - 1) * TILE_W * 128;   // one 64-channel chunk of X (haloed rows)
  static constexpr int A_BYTES = 2 * A_SLAB;
  static constexpr int G_SLAB = TILE_M * 128;                        // 128 pixels x 64 channels of dY
  static constexpr int G_BYTES = (BN / 64) * G_SLAB;
  static constexpr int STAGE_BYTES = A_BYTES + G_BYTES;
  static constexpr int STAGES_RAW = (SMEM_BUDGET - SMEM_AUX - 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + SMEM_AUX + 1024;
  static constexpr int TMEM_COLS = KH * BN <= 256 ? 256 : 512;
  static_assert(KH * BN <= 512, "accumulators exceed TMEM");
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
};

template <int BN, int KH>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_g,
                     const WgArgs a) {
  using C = WCfg<BN, KH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* aux = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* done_bar = empty_bar + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // CTA -> (load, ci block, co block, split)
  int b = blockIdx.x;
  const int split = b % a.splits; b /= a.splits;
  const int nb = b % a.nblocks; b /= a.nblocks;
  const int cb = b % a.cblocks; b /= a.cblocks;
  const int ld = b;
  const int nt = a.plan.ntaps[ld];
  const int ox = a.plan.dx[ld], oy = a.plan.dy[ld], sc = a.plan.in_scale;
  const long long t_beg = (long long)split * a.tiles_per_split;
  long long t_end = t_beg + a.tiles_per_split;
  if (t_end > a.num_tiles) t_end = a.num_tiles;

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = t_beg; t < t_end; ++t) {
        long long r = t;
        const int bx = (int)(r % a.tiles_x); r /= a.tiles_x;
        const int by = (int)(r % a.tiles_y);
        const int img = (int)(r / a.tiles_y);
        const int x0 = bx * TILE_W, y0 = by * TILE_H;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * C::STAGE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          tma_load_4d(&tmap_x, &full_bar[stage], sa, cb * 128, sc * x0 + ox, sc * y0 + oy, img);
          tma_load_4d(&tmap_x, &full_bar[stage], sa + C::A_SLAB, cb * 128 + 64, sc * x0 + ox, sc * y0 + oy, img);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(&tmap_g, &full_bar[stage], sa + C::A_BYTES + j * C::G_SLAB, nb * BN + j * 64, x0, y0, img);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = umma_idesc_mn(128, BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accumulate = 0;
      for (long long t = t_beg; t < t_end; ++t) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
        const uint32_t sg = sa + C::A_BYTES;
        if (elect_one()) {
          const uint32_t a_lo = desc_lo(sa, C::A_SLAB), g_lo = desc_lo(sg, C::G_SLAB);
          constexpr uint32_t d_hi = desc_hi(1024);
          for (int kh = 0; kh < nt; ++kh) {
#pragma unroll
            for (int ks = 0; ks < TILE_M / 16; ++ks) {
              umma_bf16_w(tmem_base + (uint32_t)(kh * BN), a_lo + (uint32_t)(kh * (TILE_W * 8) + ks * 128), d_hi,
                          g_lo + (uint32_t)(ks * 128), d_hi, idesc, accumulate | (uint32_t)(ks > 0));
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        accumulate = 1;
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(done_bar);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ci = cb * 128 + row;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const bool has_work = t_end > t_beg;
#pragma unroll 1
    for (int kh = 0; kh < nt; ++kh) {
      const int tap = a.tapidx[a.plan.slot0[ld] + kh];
      float* dst = a.part + (((long long)split * a.ktaps + tap) * a.cin + ci) * a.cout;
#pragma unroll 1
      for (int cb32 = 0; cb32 < BN; cb32 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kh * BN + cb32), v);
        const int co0 = nb * BN + cb32;
        if (ci < a.cin && co0 < a.cout) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 o;
            o.x = has_work ? __uint_as_float(v[4 * g + 0]) : 0.f;
            o.y = has_work ? __uint_as_float(v[4 * g + 1]) : 0.f;
            o.z = has_work ? __uint_as_float(v[4 * g + 2]) : 0.f;
            o.w = has_work ? __uint_as_float(v[4 * g + 3]) : 0.f;
            if (co0 + 4 * g < a.cout) *reinterpret_cast<float4*>(dst + co0 + 4 * g) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

static void plan(const srcgan_conv_params* p, int& bn, int& cblocks, int& nblocks, int& splits, long long& tiles,
                 long long& tps) {
  bn = p->cout > 64 ? 128 : 64;
  cblocks = (p->cin + 127) / 128;
  nblocks = (p->cout + bn - 1) / bn;
  const int tx = (p->wo + TILE_W - 1) / TILE_W, ty = (p->ho + TILE_H - 1) / TILE_H;
  tiles = (long long)tx * ty * p->n;
  const int groups = tc::fprop_plan(p->kh, p->stride, p->pad).plan.nloads * cblocks * nblocks;
  long long s = (2 * kNumSMs + groups - 1) / groups;     // ~2 CTAs per SM worth of work, 1 resident
  if (s > tiles) s = tiles;
  if (s < 1) s = 1;
  tps = (tiles + s - 1) / s;
  splits = (int)((tiles + tps - 1) / tps);
}

template <int BN, int KH>
static int launch(const CUtensorMap& tx, const CUtensorMap& tg, const WgArgs& a, cudaStream_t st) {
  using C = WCfg<BN, KH>;
  static DeviceOnce attr_set;
  int attr_set_dev;
  if (attr_set.needed(&attr_set_dev)) {
    SRCGAN_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel<BN, KH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
    attr_set.mark(attr_set_dev);
  }
  const unsigned grid = (unsigned)(a.plan.nloads * a.cblocks * a.nblocks * a.splits);
  conv_wgrad_tc_kernel<BN, KH><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(tx, tg, a);
  count_launch();
  return check_launch(BN == 128 ? "conv_wgrad_tc<128>" : "conv_wgrad_tc<64>");
}
}  // namespace tcw

// ---------------------------------------------------------------------------------------------
// 3x3 stride-1 wgrad over haloed slabs.  As in the fprop kernel, ONE haloed X slab [18][10][64] per
// 64-channel chunk serves all nine taps (MN-major A descriptor: start = slab + (kh*10+kw)*128 B,
// 8-pixel K groups SBO = 1280 B apart), so a CTA owns every tap its TMEM can hold:
//   cout = 32      : N = 32 (first half of the 64-wide dY rows), 9 accumulators x 32 columns
//   cout % 64 == 0 : N = 64, taps split into two groups (5 + 4 accumulators)
//   cin <= 64      : "tap pairing" - the M = 128 rows of one MMA are two taps x 64 channels (the second
//                    swizzle atom is simply another tap of the same slab: LBO = tap offset difference),
//                    so no half-empty M blocks: 5 accumulators, one group
// CTA = (tap group, 128-channel ci block, 64-channel co block, pixel split); fp32 partials
// [split][tap][ci][co] are summed in fixed order by wgrad_reduce.
// ---------------------------------------------------------------------------------------------
namespace tcw3 {
using namespace tc;

struct Wg3Args {
  int n, ho, wo, cin, cout, pad;
  int cblocks, nblocks, splits, ngroups, slots, slots_per_group, pair;
  int tiles_x, tiles_y;
  long long num_tiles, tiles_per_split;
  float* part;
};

template <int BN>
struct W3Cfg {
  static constexpr int X_BOX = HALO_H * HALO_W * 128;                 // 23040
  static constexpr int X_SLAB = (X_BOX + 1023) / 1024 * 1024;         // 23552
  static constexpr int G_BYTES = TILE_M * 128;                         // 128 px x 64 ch
  static constexpr int STAGE_BYTES = 2 * X_SLAB + G_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET - SMEM_AUX - 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + SMEM_AUX + 1024;
  static constexpr int MAX_SLOTS = 512 / BN;
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_wgrad_halo_tc(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_g,
                      const Wg3Args a) {
  using C = W3Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* aux = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* done_bar = empty_bar + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int b = blockIdx.x;
  const int split = b % a.splits; b /= a.splits;
  const int nb = b % a.nblocks; b /= a.nblocks;
  const int cb = b % a.cblocks; b /= a.cblocks;
  const int grp = b;
  const int slot_beg = grp * a.slots_per_group;
  const int slot_end = (slot_beg + a.slots_per_group) < a.slots ? (slot_beg + a.slots_per_group) : a.slots;
  const long long t_beg = (long long)split * a.tiles_per_split;
  long long t_end = t_beg + a.tiles_per_split;
  if (t_end > a.num_tiles) t_end = a.num_tiles;
  const uint32_t stage_tx = (uint32_t)((a.pair ? 1 : 2) * C::X_BOX + C::G_BYTES);

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (long long t = t_beg; t < t_end; ++t) {
      long long r = t;
      const int bx = (int)(r % a.tiles_x); r /= a.tiles_x;
      const int by = (int)(r % a.tiles_y);
      const int img = (int)(r / a.tiles_y);
      const int x0 = bx * TILE_W, y0 = by * TILE_H;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sx = smem + stage * C::STAGE_BYTES;
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], stage_tx);
        tma_load_4d(&tmap_x, &full_bar[stage], sx, cb * 128, x0 - a.pad, y0 - a.pad, img);
        if (!a.pair) tma_load_4d(&tmap_x, &full_bar[stage], sx + C::X_SLAB, cb * 128 + 64, x0 - a.pad, y0 - a.pad, img);
        tma_load_4d(&tmap_g, &full_bar[stage], sx + 2 * C::X_SLAB, nb * 64, x0, y0, img);
      }
      __syncwarp();
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = tcw::umma_idesc_mn(128, BN);
    constexpr uint32_t a_hi = desc_hi(HALO_W * 128), g_hi = desc_hi(1024);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (long long t = t_beg; t < t_end; ++t) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sx = smem_u32(smem + stage * C::STAGE_BYTES);
      const uint32_t sg = sx + 2 * C::X_SLAB;
      if (elect_one()) {
        const uint32_t g_lo = desc_lo(sg, C::G_BYTES);
        for (int slot = slot_beg; slot < slot_end; ++slot) {
          uint32_t a_lo;
          if (a.pair) {
            const int t0 = 2 * slot, t1 = (2 * slot + 1) < 9 ? (2 * slot + 1) : t0;
            const int o0 = (t0 / 3) * HALO_W + (t0 % 3), o1 = (t1 / 3) * HALO_W + (t1 % 3);
            const uint32_t lbo = (uint32_t)(o1 > o0 ? (o1 - o0) * 128 : 128);
            a_lo = desc_lo(sx + o0 * 128, lbo);
          } else {
            const int o0 = (slot / 3) * HALO_W + (slot % 3);
            a_lo = desc_lo(sx + o0 * 128, C::X_SLAB);
          }
          const uint32_t tmem_d = tmem_base + (uint32_t)((slot - slot_beg) * BN);
#pragma unroll
          for (int ks = 0; ks < TILE_M / 16; ++ks)
            umma_bf16_w(tmem_d, a_lo + (uint32_t)(ks * (2 * HALO_W * 8)), a_hi, g_lo + (uint32_t)(ks * 128), g_hi, idesc,
                        accumulate | (uint32_t)(ks > 0));
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const bool has_work = t_end > t_beg;
#pragma unroll 1
    for (int slot = slot_beg; slot < slot_end; ++slot) {
      int tap, ci;
      if (a.pair) { tap = 2 * slot + (row >> 6); ci = row & 63; }
      else { tap = slot; ci = cb * 128 + row; }
      const bool row_ok = tap < 9 && ci < a.cin;
      float* dst = a.part + (((long long)split * 9 + (tap < 9 ? tap : 0)) * a.cin + (row_ok ? ci : 0)) * a.cout;
#pragma unroll 1
      for (int cb32 = 0; cb32 < BN; cb32 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((slot - slot_beg) * BN + cb32), v);
        const int co0 = nb * 64 + cb32;
        if (row_ok && co0 < a.cout) {
          if (a.cout & 3) {                        // thin dY (e.g. 3 channels): scalar stores
            for (int i = 0; i < 32 && co0 + i < a.cout; ++i) dst[co0 + i] = has_work ? __uint_as_float(v[i]) : 0.f;
          } else {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float4 o;
              o.x = has_work ? __uint_as_float(v[4 * g + 0]) : 0.f;
              o.y = has_work ? __uint_as_float(v[4 * g + 1]) : 0.f;
              o.z = has_work ? __uint_as_float(v[4 * g + 2]) : 0.f;
              o.w = has_work ? __uint_as_float(v[4 * g + 3]) : 0.f;
              if (co0 + 4 * g < a.cout) *reinterpret_cast<float4*>(dst + co0 + 4 * g) = o;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static void plan3(const srcgan_conv_params* p, Wg3Args& a) {
  const int bn = p->cout % 64 == 0 ? 64 : 32;
  a.pair = p->cin <= 64 ? 1 : 0;
  a.cblocks = a.pair ? 1 : (p->cin + 127) / 128;
  a.nblocks = (p->cout + 63) / 64;
  a.slots = a.pair ? 5 : 9;
  const int max_slots = 512 / bn;
  a.ngroups = (a.slots + max_slots - 1) / max_slots;
  a.slots_per_group = (a.slots + a.ngroups - 1) / a.ngroups;
  a.tiles_x = (p->wo + TILE_W - 1) / TILE_W;
  a.tiles_y = (p->ho + TILE_H - 1) / TILE_H;
  a.num_tiles = (long long)a.tiles_x * a.tiles_y * p->n;
  const int groups = a.ngroups * a.cblocks * a.nblocks;
  long long s = (kNumSMs + groups - 1) / groups;          // one wave of CTAs: fewer fp32 partials to reduce
  if (s > a.num_tiles) s = a.num_tiles;
  if (s < 1) s = 1;
  a.tiles_per_split = (a.num_tiles + s - 1) / s;
  a.splits = (int)((a.num_tiles + a.tiles_per_split - 1) / a.tiles_per_split);
}

template <int BN>
static int launch3(const CUtensorMap& tx, const CUtensorMap& tg, const Wg3Args& a, cudaStream_t st) {
  using C = W3Cfg<BN>;
  static DeviceOnce attr_set;
  int attr_set_dev;
  if (attr_set.needed(&attr_set_dev)) {
    SRCGAN_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_halo_tc<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
    attr_set.mark(attr_set_dev);
  }
  const unsigned grid = (unsigned)(a.ngroups * a.cblocks * a.nblocks * a.splits);
  conv3x3_wgrad_halo_tc<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(tx, tg, a);
  count_launch();
  return check_launch(BN == 64 ? "conv3x3_wgrad_halo_tc<64>" : "conv3x3_wgrad_halo_tc<32>");
}
}  // namespace tcw3


// ---------------------------------------------------------------------------------------------
// tcgen05 wgrad, kw-stacked (3x3 stride 1 pad 1, cout 32 / 64 - the dense-block and 64->64 layers):
//   dW[kh][kw][ci][co] = sum_q X[q.y + kh - 1][q.x][ci] * dY[q.y][q.x - (kw - 1)][co]
// The horizontal tap moves from X to dY, where it can be stacked into N: the MN-major B operand is an OVERLAPPING view
// of the dY slab [rows][18 px][32 ch] - N atom j (32 channels = one 64-byte SWIZZLE_64B row) starts one pixel after atom
// j-1 (LBO = 64 B < atom size; verified on hardware, scripts/exp/exp_wstack.cu) - so one instruction computes
//   D_kh[ci][(j, co)] += X^T[ci][16 px] * [dY[px-1] | dY[px] | dY[px+1]][co]           (N = 96, kw = 2 - j)
// and a 128-pixel tile costs 3 x 8 MMAs of N = 96 (56 clk each, SMEM-bound) instead of 9 x 8 of N = 32 (40 clk each).
// A = X^T straight from the NHWC slab (MN-major, M = 128 input channels = two 64-channel atoms, vertical tap = slab
// row offset).  Channel blocks of <= 64 channels stack two vertical taps into M instead (second atom = the same
// slab two... one row further down), so M = 128 is always used.  cout = 64 runs as two independent 32-channel halves.
// One CTA per (channel block, cout half, pixel split); fp32 partials [split][tap][ci][co] + the fixed-order reduce.
// ---------------------------------------------------------------------------------------------
namespace tcw4 {
using namespace tc;

constexpr int WS_TH = 8, WS_TW = 16;                                  // pixel tile: 8 rows x 16 px = 8 k-steps of one row each
constexpr int WS_X_ROWS = WS_TH + 2;
constexpr int WS_X_ATOM = WS_X_ROWS * WS_TW * 128;                    // 20480: [10 rows][16 px][64 ch]
constexpr int WS_G_W = WS_TW + 2;
// BN = 32: N = 96, channel blocks of 128 (two X atoms) or <= 64 (one atom, two vertical taps per M).
// BN = 64: N = 192; three accumulators would need 576 TMEM columns, so every channel block is <= 64 channels with two
//          vertical taps per M (2 x 192 columns) - X is then read once for all 64 output channels.
// T22 (BN = 64 only): the third vertical tap without a phantom.  Its accumulator is M = (X | X one pixel to the right) x
// N = 128 = (dY | dY one pixel to the right): the four (X shift, dY shift) pairs are the horizontal taps kw = 1, 0, 2 and a
// duplicate of kw = 1 that is discarded - 96 + 64 clk per k-step for nine taps (90 % useful) instead of 96 + 96 (75 %).  The X
// slab is one pixel wider for it ([10 rows][17 px]); a tile's kw = 2 products then cover X pixels x0+1 .. x0+16, which still
// partitions the image row (X pixel 0 only ever meets dY pixel -1 = padding).
template <int BN, bool T22 = false>
struct W4Cfg {
  static_assert(!T22 || BN == 64, "T22 is the 64-output-channel scheme");
  static constexpr int N = 3 * BN;
  static constexpr int CBLK = BN == 64 ? 64 : 128;                    // input channels per CTA
  static constexpr int NATOM = CBLK / 64;
  static constexpr int PX = BN * 2;                                   // bytes of one dY pixel in the slab
  static constexpr int G_BYTES = WS_TH * WS_G_W * PX;                 // [8 rows][18 px][BN ch]: 9216 / 18432
  static constexpr int XW = T22 ? WS_TW + 1 : WS_TW;                  // X slab width in pixels
  static constexpr int X_BOX = WS_X_ROWS * XW * 128;                  // bytes of one X atom as TMA delivers it: 20480 / 21760
  static constexpr int X_ATOM = (X_BOX + 1023) / 1024 * 1024;         // its slot in the stage (the dY slab stays 1024-byte aligned)
  static constexpr int STAGE = NATOM * X_ATOM + G_BYTES;              // 50176 / 38912 / 40960
  static constexpr int STAGES_RAW = (SMEM_BUDGET - SMEM_AUX - 2048 - 1024) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;      // 4 / 5
  static constexpr int ZROW = WS_TW * 128;                            // one all-zero slab row [16 px][64 ch] (the phantom tap's operand)
  static constexpr int SMEM = STAGES * STAGE + SMEM_AUX + ZROW + 1024;
  static_assert(STAGES >= 3, "wgrad pipeline too shallow");
};

struct Wg4Args {
  int n, ho, wo, cin, cout;
  int cblocks, nblocks, splits;
  int tiles_x, tiles_y;
  long long num_tiles, tiles_per_split;
  float* part;
  float* dbpart;                        // [splits][cout] column sums of dY (bias gradient partials), or nullptr
  int zero_phantom;                     // 1: the phantom fourth tap multiplies a row of zeros (SRCGAN_B200_WGRAD_PHANTOM=data: the slab)
};

__host__ __device__ constexpr uint32_t desc_hi_sw64(uint32_t sbo_bytes) {
  return (sbo_bytes >> 4) | (1u << 14) | (4u << 29);                  // SBO, descriptor version 1, SWIZZLE_64B
}

__host__ __device__ constexpr uint32_t desc_hi_sw(uint32_t sbo_bytes, int bn) {
  return bn == 64 ? desc_hi(sbo_bytes) : desc_hi_sw64(sbo_bytes);     // 128-byte pixels: SWIZZLE_128B, 64-byte: SWIZZLE_64B
}

template <int BN, bool T22 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_wgrad_stack_tc(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_g, const Wg4Args a) {
  using C = W4Cfg<BN, T22>;
  constexpr int WS_STAGES = C::STAGES, WS_STAGE = C::STAGE, WS_G_BYTES = C::G_BYTES;
  constexpr int WS_X_ATOM = C::X_ATOM;
  constexpr int G_OFF = C::NATOM * WS_X_ATOM;                          // dY slab offset inside a stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* aux = smem + WS_STAGES * WS_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + WS_STAGES;
  uint64_t* done_bar = empty_bar + WS_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  uint8_t* zrow = aux + SMEM_AUX;                                      // zeros: see the MMA warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < C::ZROW / 16; i += NUM_THREADS) sts_u4(smem_u32(zrow) + (uint32_t)i * 16u, make_uint4(0u, 0u, 0u, 0u));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  int b = blockIdx.x;
  const int split = b % a.splits; b /= a.splits;
  const int nb = b % a.nblocks; b /= a.nblocks;
  const int cb = b;
  // the CTAs of channel block 0 also sum dY over the pixels (bias gradient): their four epilogue warps read every dY
  // slab out of shared memory while the MMAs run, so a stage is released by the MMA commit AND those four warps
  const bool do_db = a.dbpart != nullptr && cb == 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < WS_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], do_db ? 5 : 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int c_rem = a.cin - cb * C::CBLK;
  const bool paired = BN == 64 || c_rem <= 64;                       // one 64-channel atom: M = (tap kh | tap kh + 1)
  const int ngroups = paired ? 2 : 3;                                // accumulators of N columns
  const long long t_beg = (long long)split * a.tiles_per_split;
  long long t_end = t_beg + a.tiles_per_split;
  if (t_end > a.num_tiles) t_end = a.num_tiles;
  const uint32_t stage_tx = (uint32_t)((paired ? 1 : 2) * C::X_BOX + WS_G_BYTES);

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    pdl_wait();                                                        // X and dY are the previous kernels' outputs
    for (long long t = t_beg; t < t_end; ++t) {
      // tiles run down a 16-pixel-wide strip first: the two halo rows a tile shares with the tile above were read one
      // tile ago and still sit in L2 (x-first order re-read them 16 tiles later: 2.9 GB of DRAM reads for 1.6 GB of data)
      long long r = t;
      const int by = (int)(r % a.tiles_y); r /= a.tiles_y;
      const int bx = (int)(r % a.tiles_x);
      const int img = (int)(r / a.tiles_x);
      const int x0 = bx * WS_TW, y0 = by * WS_TH;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sx = smem + stage * WS_STAGE;
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], stage_tx);
        tma_load_4d(&tmap_x, &full_bar[stage], sx, cb * C::CBLK, x0, y0 - 1, img);
        if (!paired) tma_load_4d(&tmap_x, &full_bar[stage], sx + WS_X_ATOM, cb * C::CBLK + 64, x0, y0 - 1, img);
        tma_load_4d(&tmap_g, &full_bar[stage], sx + G_OFF, nb * BN, x0 - 1, y0, img);
      }
      __syncwarp();
      if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = tcw::umma_idesc_mn(128, C::N);
    constexpr uint32_t a_hi = desc_hi(1024), g_hi = desc_hi_sw(8 * C::PX, BN);   // 8 pixels of a row: 8 x 128 B resp. 8 dY pixels
    constexpr uint32_t ROW_A = C::XW * 128, ROW_G = WS_G_W * C::PX;     // slab row pitches: 2048 (T22: 2176) B, 1152 / 2304 B
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accumulate = 0;
    const uint32_t z_addr = smem_u32(zrow);
    const bool zero_phantom = a.zero_phantom != 0;
    for (long long t = t_beg; t < t_end; ++t) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sx = smem_u32(smem + stage * WS_STAGE);
      const uint32_t sg = sx + G_OFF;
      if (elect_one()) {
        // A: LBO = distance of the second 64-row half of M (the other channel atom, or the next vertical tap's row)
        const uint32_t a_lo = desc_lo(sx, paired ? ROW_A : (uint32_t)WS_X_ATOM);
        const uint32_t g_lo = desc_lo(sg, C::PX);                       // N atoms one pixel apart
        if constexpr (T22) {
          // accumulator 0: taps (kh 0 | kh 1) x kw 2..0 as below; accumulator 1 (columns 192..319): kh = 2, M atoms one PIXEL apart,
          // N atoms = dY one and two slab pixels in (slab pixel 0 is image pixel x0 - 1): X - dY = 0, -1 | +1, 0
          constexpr uint32_t idesc2 = tcw::umma_idesc_mn(128, 128);
          const uint32_t a2_lo = desc_lo(sx, 128), g2_lo = desc_lo(sg + (uint32_t)C::PX, C::PX);
#pragma unroll
          for (int r = 0; r < WS_TH; ++r)
            umma_bf16_w(tmem_base, a_lo + (uint32_t)((r * ROW_A) >> 4), a_hi, g_lo + (uint32_t)((r * ROW_G) >> 4), g_hi, idesc,
                        accumulate | (uint32_t)(r > 0));
#pragma unroll
          for (int r = 0; r < WS_TH; ++r)
            umma_bf16_w(tmem_base + (uint32_t)C::N, a2_lo + (uint32_t)(((r + 2) * ROW_A) >> 4), a_hi,
                        g2_lo + (uint32_t)((r * ROW_G) >> 4), g_hi, idesc2, accumulate | (uint32_t)(r > 0));
        } else
        for (int g = 0; g < ngroups; ++g) {
          const int kh0 = paired ? 2 * g : g;                           // vertical tap of the first half of M
          const uint32_t tmem_d = tmem_base + (uint32_t)(g * C::N);
          if (paired && g == 1 && zero_phantom) {
            // taps (2, phantom 3): the second 64-row half of M has no tap to compute - it would multiply the slab row below
            // the tile's last one.  Point it at a row of zeros instead (LBO = distance from this k-step's row to `zrow`): the
            // result rows are discarded either way, but multipliers fed with zeros do not toggle, and the kernel runs at the
            // board's power limit.
#pragma unroll
            for (int r = 0; r < WS_TH; ++r) {
              const uint32_t row_addr = sx + (uint32_t)((r + 2) * ROW_A);
              umma_bf16_w(tmem_d, desc_lo(row_addr, z_addr - row_addr), a_hi, g_lo + (uint32_t)((r * ROW_G) >> 4), g_hi, idesc,
                          accumulate | (uint32_t)(r > 0));
            }
            continue;
          }
#pragma unroll
          for (int r = 0; r < WS_TH; ++r)
            umma_bf16_w(tmem_d, a_lo + (uint32_t)(((r + kh0) * ROW_A) >> 4), a_hi, g_lo + (uint32_t)((r * ROW_G) >> 4), g_hi, idesc,
                        accumulate | (uint32_t)(r > 0));
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool has_work = t_end > t_beg;
    pdl_wait();                                                        // the partials' workspace is still being reduced by the previous launch
    if (do_db) {
      // thread = (8-channel group cg, pixel lane): 16-byte reads from the swizzled slab [8 rows][18 px][BN ch]
      // (interior pixels only; pixels outside the image were zero-filled by TMA)
      constexpr int CG8 = BN / 8, PL = 128 / CG8;                       // channel groups, pixel lanes
      const int t = (warp - 2) * 32 + lane, cg = t % CG8, pl = t / CG8;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int stage = 0;
      uint32_t phase = 0;
      for (long long tt = t_beg; tt < t_end; ++tt) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t sg = smem_u32(smem + stage * WS_STAGE) + G_OFF;
#pragma unroll
        for (int i = 0; i < 128 / PL; ++i) {
          const int px = pl + PL * i;
          const int idx = (px >> 4) * WS_G_W + (px & 15) + 1;
          const int sw = BN == 64 ? (cg ^ (idx & 7)) : (cg ^ ((idx >> 1) & 3));     // SWIZZLE_128B / SWIZZLE_64B
          uint4 qv;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(qv.x), "=r"(qv.y), "=r"(qv.z), "=r"(qv.w)
                       : "r"(sg + (uint32_t)(idx * C::PX + (sw << 4))));
          float f8[8];
          unpack8(qv, f8);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f8[j];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
      }
      // lanes with equal (lane % CG8) hold the same channel group: fold the other lane bits, then the four warps through smem
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (CG8 == 4) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
      }
      float* red = reinterpret_cast<float*>(aux + 512);                 // [4 warps][BN channels]
      if (lane < CG8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[q * BN + lane * 8 + j] = acc[j];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (t < BN) {
        const float sum = (red[t] + red[BN + t]) + (red[2 * BN + t] + red[3 * BN + t]);
        a.dbpart[(long long)split * a.cout + nb * BN + t] = has_work ? sum : 0.f;
      }
    }
    mbar_wait(done_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
      int kh, ci;
      if (paired) { kh = 2 * g + (row >> 6); ci = cb * C::CBLK + (row & 63); }
      else { kh = g; ci = cb * C::CBLK + row; }
      const bool t22 = T22 && g == 1;                                  // rows = (X, X + 1 px) of tap kh = 2, columns = 2 dY shifts
      if (t22) kh = 2;
      const bool row_ok = kh < 3 && ci < a.cin;
#pragma unroll 1
      for (int j = 0; j < (t22 ? 2 : 3); ++j) {
#pragma unroll 1
        for (int c32 = 0; c32 < BN; c32 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * C::N + j * BN + c32), v);
          const int kw = !t22 ? 2 - j : ((row >> 6) ? 2 : 1 - j);
          if (row_ok && !(t22 && (row >> 6) && j == 1)) {                // (X + 1, dY + 1) repeats kw = 1 on a shifted window
            const int tap = kh * 3 + kw;
            float* dst = a.part + (((long long)split * 9 + tap) * a.cin + ci) * a.cout + nb * BN + c32;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 o;
              o.x = has_work ? __uint_as_float(v[4 * i + 0]) : 0.f;
              o.y = has_work ? __uint_as_float(v[4 * i + 1]) : 0.f;
              o.z = has_work ? __uint_as_float(v[4 * i + 2]) : 0.f;
              o.w = has_work ? __uint_as_float(v[4 * i + 3]) : 0.f;
              *reinterpret_cast<float4*>(dst + 4 * i) = o;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static int bn4(const srcgan_conv_params* p) { return p->cout == 64 && !getenv("SRCGAN_B200_WSTACK32") ? 64 : 32; }
// 64 output channels: the phantom-free third tap (W4Cfg); SRCGAN_B200_WGRAD_T22=0 brings back (kh 2 | zero row) x N = 192
static bool t22_on(const srcgan_conv_params* p) {
  const char* e = getenv("SRCGAN_B200_WGRAD_T22");
  return bn4(p) == 64 && !(e && e[0] == '0');
}
static int x_box_w(const srcgan_conv_params* p) { return t22_on(p) ? WS_TW + 1 : WS_TW; }

static void plan4(const srcgan_conv_params* p, Wg4Args& a) {
  const int bn = bn4(p);
  a.cblocks = bn == 64 ? (p->cin + 63) / 64 : (p->cin + 127) / 128;
  a.nblocks = p->cout / bn;
  a.tiles_x = (p->wo + WS_TW - 1) / WS_TW;
  a.tiles_y = (p->ho + WS_TH - 1) / WS_TH;
  a.num_tiles = (long long)a.tiles_x * a.tiles_y * p->n;
  const int groups = a.cblocks * a.nblocks;
  long long s = kNumSMs / groups;                          // one wave of CTAs, never a second one for a remainder
  if (s > a.num_tiles) s = a.num_tiles;
  if (s < 1) s = 1;
  a.tiles_per_split = (a.num_tiles + s - 1) / s;
  a.splits = (int)((a.num_tiles + a.tiles_per_split - 1) / a.tiles_per_split);
  a.zero_phantom = getenv("SRCGAN_B200_WGRAD_PHANTOM") ? 0 : 1;
}

// NHWC bf16 tensor -> tensor map with an explicit box [boxc ch][boxw px][rows] and swizzle
static int make_tmap_box(CUtensorMap* tm, const void* ptr, int c, int w, int h, int n, int ld, int boxc, int boxw, int rows,
                         CUtensorMapSwizzle sw, const char* what) {
  EncodeTiledFn encode = get_encode_fn();
  SRCGAN_REQUIRE(encode != nullptr, "%s: cuTensorMapEncodeTiled is not available from the driver", what);
  cuuint64_t gdim[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * w, (cuuint64_t)ld * 2 * w * h};
  cuuint32_t box[4] = {(cuuint32_t)boxc, (cuuint32_t)boxw, (cuuint32_t)rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // a 32-channel slice of a wider buffer is 64 useful bytes per pixel: with 128-byte L2 promotion every pixel pulls the whole
  // line from DRAM (ncu: 2.12x the algorithmic bytes on the 32 -> 32 remainders of the paired wgrads, which are HBM-bound)
  const bool narrow = (c < boxc ? c : boxc) * 2 <= 64 && ld > c && !getenv("SRCGAN_B200_NO_L2_64B");
  CUresult cr = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, sw, narrow ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)cr);
    return SRCGAN_E_CUDA;
  }
  return SRCGAN_OK;
}

template <int BN, bool T22 = false>
static int launch4(const CUtensorMap& tx, const CUtensorMap& tg, const Wg4Args& a, cudaStream_t st) {
  using C = W4Cfg<BN, T22>;
  static DeviceOnce attr_set;
  int attr_set_dev;
  if (attr_set.needed(&attr_set_dev)) {
    SRCGAN_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_stack_tc<BN, T22>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set.mark(attr_set_dev);
  }
  const unsigned grid = (unsigned)(a.cblocks * a.nblocks * a.splits);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
  cudaLaunchAttribute at[1];
  cfg.attrs = at; cfg.numAttrs = pdl_attr(&at[0]);
  SRCGAN_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_wgrad_stack_tc<BN, T22>, tx, tg, a));
  count_launch();
  return check_launch(BN == 64 ? "conv3x3_wgrad_stack_tc<64>" : "conv3x3_wgrad_stack_tc<32>");
}
static int launch4_64(const srcgan_conv_params* p, const CUtensorMap& tx, const CUtensorMap& tg, const Wg4Args& a, cudaStream_t st) {
  return t22_on(p) ? launch4<64, true>(tx, tg, a, st) : launch4<64, false>(tx, tg, a, st);
}

// ---------------------------------------------------------------------------------------------
// kw-stacked wgrad for a 32-channel input slice and 32 output channels (the 32 -> 32 remainder launches of the paired dense-block
// weight gradients: x_k slice x dZ_(k+1) slice of two 192-channel concat buffers, 64 useful bytes per pixel and operand).
// In conv3x3_wgrad_stack_tc<32> one 64-channel X atom holds only two vertical taps, so a 128-pixel tile costs 2 x 8 MMAs of N = 96
// (900 clk) beside 304 TMA box rows (160 of X + 144 of dY at ~3.3 clk per row: 1 000 clk); ncu: 71 % SM throughput, 40 % DRAM.
//   * Here X is a SWIZZLE_64B slab of 32-channel atoms: M = 128 = FOUR atoms one slab row apart = the three vertical taps + a
//     phantom, ONE MMA of N = 96 per 16-pixel row (8 per tile, 450 clk).
//   * LSU = false (default): dY through TMA like X (tensor maps with 64-byte L2 promotion: DRAM traffic = the useful bytes).
//   * LSU = true (SRCGAN_B200_WGRAD_R32=lsu, the first version): dY does not go through TMA - the four epilogue warps, idle
//     until the last tile, copy it with 16-byte cp.async into the same swizzled [8 rows][18 px][32 ch] slab (zero fill outside
//     the image), several tiles in flight.  Measured no faster than the stacked kernel: LSU loads of a 64-byte slice pull
//     whole 128-byte lines from DRAM (867 instead of 568 MB per launch, 80 % of the copy bandwidth).
// Same tile order, k-step order and fp32 partial layout as conv3x3_wgrad_stack_tc, so the reduce launch is shared.
// ---------------------------------------------------------------------------------------------
constexpr int R32_X_BYTES = WS_X_ROWS * WS_TW * 64;                   // 10240: [10 rows][16 px][32 ch]
constexpr int R32_G_BYTES = WS_TH * WS_G_W * 64;                      //  9216: [ 8 rows][18 px][32 ch]
constexpr int R32_STAGE = R32_X_BYTES + R32_G_BYTES;                  // 19456 (both slabs 1024-byte aligned)
constexpr int R32_STAGES = 8;
constexpr int R32_LAG = 3;                                            // cp.async groups (tiles) in flight per thread
constexpr int R32_SMEM = R32_STAGES * R32_STAGE + SMEM_AUX + 1024;
constexpr int R32_CHUNKS = WS_TH * WS_G_W * 4;                        // 576 16-byte chunks of a dY slab
constexpr int R32_PER_THREAD = (R32_CHUNKS + 127) / 128;              // 5
static_assert(R32_LAG < R32_STAGES, "the cp.async producers would wait for a stage their own pending arrival frees");

template <bool LSU>      // dY through cp.async by the epilogue warps (true) or through TMA like X (false)
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_wgrad_r32_tc(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_g,
                     const __nv_bfloat16* __restrict__ g, int g_ld, const Wg4Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* aux = smem + R32_STAGES * R32_STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + R32_STAGES;
  uint64_t* done_bar = empty_bar + R32_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const int split = blockIdx.x;
  if (threadIdx.x == 0) {
    // a stage is full when the TMA bytes of X have landed (one arrive.expect_tx) and all 128 copy threads have arrived
    for (int s = 0; s < R32_STAGES; ++s) { mbar_init(&full_bar[s], LSU ? 1 + 128 : 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long t_beg = (long long)split * a.tiles_per_split;
  long long t_end = t_beg + a.tiles_per_split;
  if (t_end > a.num_tiles) t_end = a.num_tiles;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    pdl_wait();
    for (long long t = t_beg; t < t_end; ++t) {
      long long r = t;
      const int by = (int)(r % a.tiles_y); r /= a.tiles_y;
      const int bx = (int)(r % a.tiles_x);
      const int img = (int)(r / a.tiles_x);
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], (uint32_t)(LSU ? R32_X_BYTES : R32_STAGE));
        tma_load_4d(&tmap_x, &full_bar[stage], smem + stage * R32_STAGE, 0, bx * WS_TW, by * WS_TH - 1, img);
        if (!LSU) tma_load_4d(&tmap_g, &full_bar[stage], smem + stage * R32_STAGE + R32_X_BYTES, 0, bx * WS_TW - 1, by * WS_TH, img);
      }
      __syncwarp();
      if (++stage == R32_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = tcw::umma_idesc_mn(128, 96);
    constexpr uint32_t ROW_A = WS_TW * 64, ROW_G = WS_G_W * 64;         // slab row pitches: 1024 B, 1152 B
    constexpr uint32_t a_hi = desc_hi_sw64(8 * 64), g_hi = desc_hi_sw64(8 * 64);   // K groups: 8 pixels of a row
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (long long t = t_beg; t < t_end; ++t) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sx = smem_u32(smem + stage * R32_STAGE);
      const uint32_t sg = sx + R32_X_BYTES;
      if (elect_one()) {
        const uint32_t a_lo = desc_lo(sx, ROW_A);                       // M atoms (32 channels) one slab row apart: taps kh = 0..3
        const uint32_t g_lo = desc_lo(sg, 64);                          // N atoms (32 channels) one pixel apart: taps kw = 2..0
#pragma unroll
        for (int r = 0; r < WS_TH; ++r)                                 // (row 7's phantom atom reads the first row of the dY slab)
          umma_bf16_w(tmem_base, a_lo + (uint32_t)((r * ROW_A) >> 4), a_hi, g_lo + (uint32_t)((r * ROW_G) >> 4), g_hi, idesc,
                      accumulate | (uint32_t)(r > 0));
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == R32_STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // ---- LSU variant only: dY producer, thread -> up to five fixed (slab pixel, 16-byte chunk) slots of every tile
    // (with dY on TMA these warps go straight to the epilogue wait)
    const int tid = threadIdx.x - 64;
    uint32_t dst_off[R32_PER_THREAD], src_off[R32_PER_THREAD];
    int pr[R32_PER_THREAD], pj[R32_PER_THREAD];
#pragma unroll
    for (int i = 0; i < R32_PER_THREAD; ++i) {
      const int id = tid + 128 * i;
      const int idx = id >> 2, ch = id & 3;                             // slab pixel, chunk of its 64 bytes
      pr[i] = idx / WS_G_W;
      pj[i] = idx - pr[i] * WS_G_W;
      dst_off[i] = (uint32_t)(idx * 64 + ((ch ^ ((idx >> 1) & 3)) << 4));            // SWIZZLE_64B (slab 1024-byte aligned)
      src_off[i] = (uint32_t)(((pr[i] * a.wo + pj[i]) * g_ld + ch * 8) * 2);
    }
    const char* gbase = reinterpret_cast<const char*>(g);
    int stage = 0, sig = 0, pend = 0;
    uint32_t phase = 0;
    pdl_wait();                                                         // dY is the previous kernels' output
    for (long long t = t_beg; LSU && t < t_end; ++t) {
      long long r = t;
      const int by = (int)(r % a.tiles_y); r /= a.tiles_y;
      const int bx = (int)(r % a.tiles_x);
      const int img = (int)(r / a.tiles_x);
      const int x0 = bx * WS_TW - 1, y0 = by * WS_TH;
      const char* tile = gbase + (((long long)img * a.ho + y0) * a.wo + x0) * (long long)g_ld * 2;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      const uint32_t sg = smem_u32(smem + stage * R32_STAGE) + R32_X_BYTES;
#pragma unroll
      for (int i = 0; i < R32_PER_THREAD; ++i) {
        if (i < R32_PER_THREAD - 1 || tid + 128 * i < R32_CHUNKS) {
          const int x = x0 + pj[i];
          const bool ok = y0 + pr[i] < a.ho && x >= 0 && x < a.wo;
          const char* src = ok ? tile + src_off[i] : gbase;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sg + dst_off[i]), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (++pend > R32_LAG) {
        asm volatile("cp.async.wait_group %0;" ::"n"(R32_LAG) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> the MMA's async-proxy reads
        mbar_arrive(&full_bar[sig]);
        if (++sig == R32_STAGES) sig = 0;
        --pend;
      }
      if (++stage == R32_STAGES) { stage = 0; phase ^= 1; }
    }
    if (LSU) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (; pend > 0; --pend) {
      mbar_arrive(&full_bar[sig]);
      if (++sig == R32_STAGES) sig = 0;
    }
    // ---- epilogue: TMEM lane = (kh, ci) = (lane quadrant, lane); columns = (j, co), kw = 2 - j
    const int q = warp & 3;
    const bool has_work = t_end > t_beg;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const bool row_ok = q < 3 && lane < a.cin;
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 32), v);
      if (row_ok) {
        const int tap = q * 3 + (2 - j);
        float* dst = a.part + (((long long)split * 9 + tap) * a.cin + lane) * a.cout;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o;
          o.x = has_work ? __uint_as_float(v[4 * i + 0]) : 0.f;
          o.y = has_work ? __uint_as_float(v[4 * i + 1]) : 0.f;
          o.z = has_work ? __uint_as_float(v[4 * i + 2]) : 0.f;
          o.w = has_work ? __uint_as_float(v[4 * i + 3]) : 0.f;
          *reinterpret_cast<float4*>(dst + 4 * i) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// 32-channel (or thinner) input slice, 32 output channels, no in-kernel bias gradient.  SRCGAN_B200_WGRAD_R32 = "0": off (the
// stacked <32> kernel), "lsu": the cp.async variant - measured at 64 x 256 x 256: 0.1754 ms against 0.1712 ms of
// conv3x3_wgrad_stack_tc<32>, because the LSU path reads whole 128-byte lines from DRAM ("What bounds them" 19).
static int r32_mode(const srcgan_conv_params* p, const float* dbpart) {       // 0 off, 1 TMA, 2 cp.async
  if (!(p->cout == 32 && p->cin <= 32 && dbpart == nullptr)) return 0;
  const char* e = getenv("SRCGAN_B200_WGRAD_R32");
  if (!e) return 1;
  return e[0] == '0' ? 0 : (e[0] == 'l' ? 2 : 1);
}

template <bool LSU>
static int launch_r32(const srcgan_conv_params* p, const CUtensorMap& tg, const Wg4Args& a, cudaStream_t st) {
  CUtensorMap tx;
  int rc = make_tmap_box(&tx, p->x, p->cin, p->w, p->h, p->n, p->x_ld, 32, WS_TW, WS_X_ROWS, CU_TENSOR_MAP_SWIZZLE_64B,
                         "conv_wgrad_r32(x)");
  if (rc) return rc;
  static DeviceOnce attr_set;
  int attr_set_dev;
  if (attr_set.needed(&attr_set_dev)) {
    SRCGAN_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_r32_tc<LSU>, cudaFuncAttributeMaxDynamicSharedMemorySize, R32_SMEM));
    attr_set.mark(attr_set_dev);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)a.splits); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = R32_SMEM; cfg.stream = st;
  cudaLaunchAttribute at[1];
  cfg.attrs = at; cfg.numAttrs = pdl_attr(&at[0]);
  SRCGAN_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_wgrad_r32_tc<LSU>, tx, tg, reinterpret_cast<const __nv_bfloat16*>(p->y), (int)p->y_ld, a));
  count_launch();
  return check_launch(LSU ? "conv3x3_wgrad_r32_tc<lsu>" : "conv3x3_wgrad_r32_tc<tma>");
}
}  // namespace tcw4

// shared with the SIMT engine (conv_simt.cu)
int wgrad_reduce_launch(const float* part, int splits, int taps, int cin, int cout, float* dw, int accumulate,
                        float alpha, cudaStream_t st);
int bias_grad_launch(const void* dy, int dy_ld, int dtype, long long M, int cout, float* db, int accumulate,
                     float alpha, void* ws, cudaStream_t st);
int colsum_final_launch(const float* part, int nparts, int c, float* out, int accumulate, float alpha, cudaStream_t st);
int colsum_final_ld_launch(const float* part, int nparts, int c, int ld, float* out, int accumulate, float alpha, cudaStream_t st);
int wgrad_reduce_split_launch(const float* part, int splits, int taps, int cin, int cout, int split, float* d0, int ld0,
                              int ci00, float* d1, int ld1, int ci01, int accumulate, float alpha, cudaStream_t st);
int wgrad_reduce_db_launch(const float* part, int splits, int taps, int cin, int cout, int split, float* d0, int ld0, int ci00,
                           float* d1, int ld1, int ci01, int accumulate, float alpha, const float* dbpart, float* db0, float* db1,
                           cudaStream_t st);

bool conv_wgrad_tc_supported(const srcgan_conv_params* p) {
  if (p->dtype != SRCGAN_DT_BF16 || (p->stride != 1 && p->stride != 2) || p->upsample) return false;
  if (p->kh != p->kw || (p->kh != 1 && p->kh != 3 && p->kh != 4)) return false;
  if (p->kh == 1 && p->pad != 0) return false;
  if (!(p->kh == 3 && p->stride == 1) && (p->cin < 16 || p->cin % 8 || p->cout < 16 || p->cout % 8)) return false;
  if (p->x_ld % 8 || p->y_ld % 8) return false;
  if (((uintptr_t)p->x) % 16 || ((uintptr_t)p->y) % 16) return false;
  return true;
}

static bool wgrad_stack_ok(const srcgan_conv_params* p) {
  // any number of input channels: the X box is 64 channels wide and TMA zero-fills what the slice does not have (an image
  // convolution, 3 -> 64, costs what 64 -> 64 costs: 0.29 ms at 64 x 256^2 against 0.56 ms on the per-tap halo kernel)
  const bool thin_ok = p->cin < 16 && p->cout == 64 && !getenv("SRCGAN_B200_NO_WSTACK_THIN");
  return p->kh == 3 && p->kw == 3 && p->stride == 1 && p->pad == 1 && (p->cout == 32 || p->cout == 64) &&
         ((p->cin >= 16 && p->cin % 8 == 0) || thin_ok) && !getenv("SRCGAN_B200_NO_WSTACK");
}

static bool wgrad_halo_ok(const srcgan_conv_params* p) {
  return p->kh == 3 && p->kw == 3 && p->stride == 1 && (p->cout <= 32 || p->cout % 64 == 0) &&
         !getenv("SRCGAN_B200_NO_HALO");
}

// 3x3 weight gradient of a thin-output layer (64 -> 3 image convolution, 64 -> 1): the stacked kernel with dY's box 32 channels
// wide - TMA zero-fills what the slice does not have - and partials / reduce over 32 output channels of which the first cout are
// kept (the reduce's split destination); 0.45 ms on the per-tap halo kernel at 64 x 256^2
static bool wgrad_stack_thin_out_ok(const srcgan_conv_params* p) {
  return p->kh == 3 && p->kw == 3 && p->stride == 1 && p->pad == 1 && p->cout < 16 && p->cin >= 16 && p->cin % 8 == 0 &&
         !getenv("SRCGAN_B200_NO_WSTACK") && !getenv("SRCGAN_B200_NO_WSTACK_THIN");
}

size_t conv_wgrad_tc_workspace(const srcgan_conv_params* p) {
  if (wgrad_stack_thin_out_ok(p)) {
    srcgan_conv_params q = *p;
    q.cout = 32;
    tcw4::Wg4Args a4;
    tcw4::plan4(&q, a4);
    const size_t wb = (size_t)a4.splits * 9 * q.cin * q.cout * sizeof(float);
    return ((wb + 255) / 256) * 256 + (size_t)1024 * q.cout * sizeof(float) + 256;
  }
  int bn, cblocks, nblocks, splits;
  long long tiles, tps;
  tcw::plan(p, bn, cblocks, nblocks, splits, tiles, tps);
  if (wgrad_halo_ok(p)) {
    tcw3::Wg3Args a3;
    tcw3::plan3(p, a3);
    if (a3.splits > splits) splits = a3.splits;
  }
  if (wgrad_stack_ok(p)) {
    tcw4::Wg4Args a4;
    tcw4::plan4(p, a4);
    if (a4.splits > splits) splits = a4.splits;
  }
  size_t wbytes = (size_t)splits * p->kh * p->kw * p->cin * p->cout * sizeof(float);
  return ((wbytes + 255) / 256) * 256 + (size_t)1024 * p->cout * sizeof(float) + 256;
}

// Weight gradients of TWO layers from one launch of the kw-stacked kernel: output channels [0, split) of dY belong to layer 0,
// [split, cout) to layer 1; each destination is an OIHW gradient tensor with `ld` input channels of which [ci0, ci0 + cin)
// are written.  Dense-block use (nn.py): conv_k and conv_(k+1) read the same input prefix and their dY slices are adjacent in
// the gradient concat buffer, so their common part is ONE 64-output-channel wgrad (N = 192: the tensor pipe is 95-98 % active
// and X is read once) instead of two 32-channel ones (N = 96: 62-68 %, X read twice).
int conv_wgrad_tc_split(const srcgan_conv_params* p, float* dw0, int ld0, int ci00, float* db0, float* dw1, int ld1, int ci01,
                        float* db1, int split, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  SRCGAN_REQUIRE(wgrad_stack_ok(p), "conv_wgrad_split: needs a 3x3 stride-1 pad-1 layer with 32 or 64 output channels");
  SRCGAN_REQUIRE(split > 0 && split <= p->cout && (dw0 || dw1), "conv_wgrad_split: bad split %d", split);
  SRCGAN_REQUIRE((!dw0 || (ld0 >= ci00 + p->cin && ci00 >= 0)) && (!dw1 || (ld1 >= ci01 + p->cin && ci01 >= 0)),
                 "conv_wgrad_split: destination channel window out of range");
  SRCGAN_REQUIRE(ws && ws_bytes >= conv_wgrad_tc_workspace(p), "conv_wgrad_split: workspace too small");
  tcw4::Wg4Args a4;
  tcw4::plan4(p, a4);
  a4.n = p->n; a4.ho = p->ho; a4.wo = p->wo; a4.cin = p->cin; a4.cout = p->cout;
  a4.part = reinterpret_cast<float*>(ws);
  const size_t wbytes = (size_t)a4.splits * 9 * p->cin * p->cout * sizeof(float);
  const bool want_db = db0 != nullptr || db1 != nullptr;
  a4.dbpart = want_db ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ((wbytes + 255) / 256) * 256) : nullptr;
  CUtensorMap tx, tg;
  int rc = tcw4::make_tmap_box(&tx, p->x, p->cin, p->w, p->h, p->n, p->x_ld, 64, tcw4::x_box_w(p), tcw4::WS_X_ROWS,
                               CU_TENSOR_MAP_SWIZZLE_128B, "conv_wgrad_split(x)");
  if (rc) return rc;
  const int bn = tcw4::bn4(p);
  rc = tcw4::make_tmap_box(&tg, p->y, p->cout, p->wo, p->ho, p->n, p->y_ld, bn, tcw4::WS_G_W, tcw4::WS_TH,
                           bn == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, "conv_wgrad_split(dy)");
  if (rc) return rc;
  if (const int m = tcw4::r32_mode(p, a4.dbpart)) rc = m == 2 ? tcw4::launch_r32<true>(p, tg, a4, st) : tcw4::launch_r32<false>(p, tg, a4, st);
  else rc = bn == 64 ? tcw4::launch4_64(p, tx, tg, a4, st) : tcw4::launch4<32>(tx, tg, a4, st);
  if (rc) return rc;
  // split-K reduce of the weight gradients and, in the same launch, of the bias gradients the kernel summed per split
  return wgrad_reduce_db_launch(reinterpret_cast<const float*>(ws), a4.splits, 9, p->cin, p->cout, split, dw0, ld0, ci00, dw1, ld1,
                                ci01, accumulate, p->alpha, a4.dbpart, db0, split < p->cout ? db1 : nullptr, st);
}

int conv_wgrad_tc(const srcgan_conv_params* p, float* dw, float* db, int accumulate, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  SRCGAN_REQUIRE(ws && ws_bytes >= conv_wgrad_tc_workspace(p), "conv_wgrad_tc: workspace too small");
  int bn, cblocks, nblocks, splits;
  long long tiles, tps;
  tcw::plan(p, bn, cblocks, nblocks, splits, tiles, tps);
  size_t wbytes = (size_t)splits * p->kh * p->kw * p->cin * p->cout * sizeof(float);
  if (dw && wgrad_stack_thin_out_ok(p)) {
    srcgan_conv_params q = *p;
    q.cout = 32;
    tcw4::Wg4Args a4;
    tcw4::plan4(&q, a4);
    a4.n = p->n; a4.ho = p->ho; a4.wo = p->wo; a4.cin = p->cin; a4.cout = 32;
    a4.part = reinterpret_cast<float*>(ws);
    wbytes = (size_t)a4.splits * 9 * p->cin * 32 * sizeof(float);
    a4.dbpart = db ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ((wbytes + 255) / 256) * 256) : nullptr;
    CUtensorMap tx, tg;
    int rc = tcw4::make_tmap_box(&tx, p->x, p->cin, p->w, p->h, p->n, p->x_ld, 64, tcw4::WS_TW, tcw4::WS_X_ROWS,
                                 CU_TENSOR_MAP_SWIZZLE_128B, "conv_wgrad_tc(stack x, thin dy)");
    if (rc) return rc;
    rc = tcw4::make_tmap_box(&tg, p->y, p->cout, p->wo, p->ho, p->n, p->y_ld, 32, tcw4::WS_G_W, tcw4::WS_TH,
                             CU_TENSOR_MAP_SWIZZLE_64B, "conv_wgrad_tc(stack thin dy)");
    if (rc) return rc;
    rc = tcw4::launch4<32>(tx, tg, a4, st);
    if (rc) return rc;
    return wgrad_reduce_db_launch(reinterpret_cast<const float*>(ws), a4.splits, 9, p->cin, 32, p->cout, dw, p->cin, 0, nullptr, 0, 0,
                                  accumulate, p->alpha, a4.dbpart, db, nullptr, st);
  }
  if (dw && wgrad_stack_ok(p)) {
    tcw4::Wg4Args a4;
    tcw4::plan4(p, a4);
    a4.n = p->n; a4.ho = p->ho; a4.wo = p->wo; a4.cin = p->cin; a4.cout = p->cout;
    a4.part = reinterpret_cast<float*>(ws);
    wbytes = (size_t)a4.splits * 9 * p->cin * p->cout * sizeof(float);
    a4.dbpart = db ? reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ((wbytes + 255) / 256) * 256) : nullptr;
    CUtensorMap tx, tg;
    int rc = tcw4::make_tmap_box(&tx, p->x, p->cin, p->w, p->h, p->n, p->x_ld, 64, tcw4::x_box_w(p), tcw4::WS_X_ROWS,
                                 CU_TENSOR_MAP_SWIZZLE_128B, "conv_wgrad_tc(stack x)");
    if (rc) return rc;
    const int bn = tcw4::bn4(p);
    rc = tcw4::make_tmap_box(&tg, p->y, p->cout, p->wo, p->ho, p->n, p->y_ld, bn, tcw4::WS_G_W, tcw4::WS_TH,
                             bn == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, "conv_wgrad_tc(stack dy)");
    if (rc) return rc;
    if (const int m = tcw4::r32_mode(p, a4.dbpart)) rc = m == 2 ? tcw4::launch_r32<true>(p, tg, a4, st) : tcw4::launch_r32<false>(p, tg, a4, st);
    else rc = bn == 64 ? tcw4::launch4_64(p, tx, tg, a4, st) : tcw4::launch4<32>(tx, tg, a4, st);
    if (rc) return rc;
    return wgrad_reduce_db_launch(reinterpret_cast<const float*>(ws), a4.splits, 9, p->cin, p->cout, p->cout, dw, p->cin, 0, nullptr,
                                  0, 0, accumulate, p->alpha, a4.dbpart, db, nullptr, st);   // bias gradient: summed in-kernel
  } else if (dw && wgrad_halo_ok(p)) {
    tcw3::Wg3Args a3;
    tcw3::plan3(p, a3);
    a3.n = p->n; a3.ho = p->ho; a3.wo = p->wo; a3.cin = p->cin; a3.cout = p->cout; a3.pad = p->pad;
    a3.part = reinterpret_cast<float*>(ws);
    wbytes = (size_t)a3.splits * 9 * p->cin * p->cout * sizeof(float);
    CUtensorMap tx, tg;
    int rc = tc::make_tmap(&tx, p->x, p->cin, p->w, p->h, p->n, p->x_ld, tc::HALO_H, 1, "conv_wgrad_tc(x)", tc::HALO_W);
    if (rc) return rc;
    rc = tc::make_tmap(&tg, p->y, p->cout, p->wo, p->ho, p->n, p->y_ld, tc::TILE_H, 1, "conv_wgrad_tc(dy)");
    if (rc) return rc;
    rc = p->cout % 64 == 0 ? tcw3::launch3<64>(tx, tg, a3, st) : tcw3::launch3<32>(tx, tg, a3, st);
    if (rc) return rc;
    rc = wgrad_reduce_launch(reinterpret_cast<const float*>(ws), a3.splits, 9, p->cin, p->cout, dw, accumulate,
                             p->alpha, st);
    if (rc) return rc;
  } else if (dw) {
    tc::HostPlan hp = tc::fprop_plan(p->kh, p->stride, p->pad);
    CUtensorMap tx, tg;
    int rc = tc::make_tmap(&tx, p->x, p->cin, p->w, p->h, p->n, p->x_ld, tc::TILE_H + hp.maxt - 1, p->stride,
                           "conv_wgrad_tc(x)");
    if (rc) return rc;
    rc = tc::make_tmap(&tg, p->y, p->cout, p->wo, p->ho, p->n, p->y_ld, tc::TILE_H, 1, "conv_wgrad_tc(dy)");
    if (rc) return rc;
    tcw::WgArgs a;
    a.n = p->n; a.ho = p->ho; a.wo = p->wo; a.cin = p->cin; a.cout = p->cout;
    a.plan = hp.plan;
    for (int s = 0; s < tc::MAX_SLOTS; ++s)
      a.tapidx[s] = s < hp.plan.total_slots ? hp.slot_kh[s] * p->kw + hp.slot_kw[s] : 0;
    a.ktaps = p->kh * p->kw;
    a.cblocks = cblocks; a.nblocks = nblocks; a.splits = splits;
    a.tiles_x = (p->wo + tc::TILE_W - 1) / tc::TILE_W;
    a.tiles_y = (p->ho + tc::TILE_H - 1) / tc::TILE_H;
    a.num_tiles = tiles; a.tiles_per_split = tps;
    a.part = reinterpret_cast<float*>(ws);
    if (hp.maxt == 2) rc = bn == 64 ? tcw::launch<64, 2>(tx, tg, a, st) : tcw::launch<128, 2>(tx, tg, a, st);
    else if (hp.maxt == 3) rc = bn == 64 ? tcw::launch<64, 3>(tx, tg, a, st) : tcw::launch<128, 3>(tx, tg, a, st);
    else rc = bn == 64 ? tcw::launch<64, 4>(tx, tg, a, st) : tcw::launch<128, 4>(tx, tg, a, st);
    if (rc) return rc;
    rc = wgrad_reduce_launch(reinterpret_cast<const float*>(ws), splits, p->kh * p->kw, p->cin, p->cout, dw,
                             accumulate, p->alpha, st);
    if (rc) return rc;
  }
  if (db) {
    void* bws = reinterpret_cast<char*>(ws) + ((wbytes + 255) / 256) * 256;
    return bias_grad_launch(p->y, p->y_ld, p->dtype, (long long)p->n * p->ho * p->wo, p->cout, db, accumulate, p->alpha,
                            bws, st);
  }
  return SRCGAN_OK;
}

}  // namespace srcgan
