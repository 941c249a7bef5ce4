// One launch re-packs EVERY tcgen05-layout weight tensor of a network after an optimizer step.
//
// A packed tensor (SRCGAN_WL_TC: [n_block][chunk][slot][BN][64] bf16, SWIZZLE_128B image, see pack_weights_tc in
// conv_tc.cu) is described as a list of SOURCE BLOCKS: a block copies GEMM-K channels [k0, k0 + k_len) of the packed
// tensor from one fp32 OIHW parameter through arbitrary element strides, an optional 180-degree tap rotation and a scale.
// That covers everything the networks derive from their parameters on the host today:
//   fprop weights            one block, (n, k) = (cout, cin)
//   dgrad (transposed) ones  one block, (n, k) = (cin, cout), rotated taps                                (nn.py::_wT)
//   mirrored dense block     up to five blocks concatenated along K: slices of conv5 (scaled by 0.2 / 0.04) and of
//                            conv4..conv(k+1), each transposed + rotated (nn.py::_dense_wT; model.py:205-211 backward)
// so the ~900 ATen flip / cat / mul / contiguous launches and ~200 pack launches per optimizer step become ONE launch per
// network.  The block table lives in device memory (parameter and destination pointers are stable across steps).
#include "common.cuh"

namespace srcgan {

__global__ void __launch_bounds__(256)
pack_weights_batch_k(const srcgan_pack_block* __restrict__ blocks, int nblocks, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // binary search: last block with elem0 <= i
  int lo = 0, hi = nblocks - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (blocks[mid].elem0 <= i) lo = mid; else hi = mid - 1;
  }
  const srcgan_pack_block& b = blocks[lo];
  long long t = i - b.elem0;
  const int kl = (int)(t % b.k_len); t /= b.k_len;
  const int slot = (int)(t % b.total_slots);
  const int nch = (int)(t / b.total_slots);
  float v = 0.f;
  if (nch < b.n_count) {
    const int off = b.flip ? b.taps - 1 - b.slot_off[slot] : b.slot_off[slot];
    v = b.scale * __ldg(b.src + b.src_off + (long long)nch * b.nstride + (long long)kl * b.kstride + off);
  }
  const int kch = b.k0 + kl, c = kch >> 6, j = kch & 63;
  const int nb = nch / b.bn, r = nch - nb * b.bn;
  const long long tile = ((long long)(nb * b.nchunks + c) * b.total_slots + slot) * b.bn * 64;
  reinterpret_cast<__nv_bfloat16*>(b.out)[tile + (long long)r * 64 + (((j >> 3) ^ (r & 7)) << 3) + (j & 7)] = __float2bfloat16_rn(v);
}

int pack_weights_batch(const srcgan_pack_block* blocks_dev, int nblocks, long long total, cudaStream_t st) {
  SRCGAN_REQUIRE(blocks_dev && nblocks > 0 && total > 0, "pack_weights_batch: empty table");
  pack_weights_batch_k<<<ceil_div(total, 256), 256, 0, st>>>(blocks_dev, nblocks, total);
  count_launch();
  return check_launch("pack_weights_batch");
}

}  // namespace srcgan
