// v6: TWO consecutive dense-block layers in one paired sweep (included by conv_tc.cu inside namespace srcgan::tc).
//
//   layer A :  x_k     = epi_A( conv3x3( P ) )              P = the first cin channels of the concat buffer (cin = 64 | 128)
//   layer B :  x_(k+1) = epi_B( conv3x3( [P | x_k] ) )      (src/model/model.py:205-210: conv1+conv2, conv3+conv4, and the same
//                                                            two pairs of the mirrored dense block of the backward pass)
//
// Run separately the two layers read the prefix P twice and x_k once more ((64 + 96 + 128 + 160 + 4*32) * 2 = 1 152 B per pixel
// over conv1..4 where (64 + 128 + 4*32) * 2 = 640 B do), and each launch is bound by its epilogue: one warp needs ~2 000 clocks
// to drain, activate and store a [32 lanes][32 channels] block, against 590 clocks of MMAs per column of 64->32.  Here one CTA
// pair (cta_group::2, M = 256) covers the WHOLE lane extent of an image (<= 256 pixels), so the only halo of x_k in the lane
// direction is the zero padding of the image plus ONE lane handed between the two CTAs through distributed shared memory; in
// the sweep direction layer B simply trails layer A by a few columns.
//
//   * TMEM: two accumulator rings of 32-column fp32 blocks run in laps exactly like the paired sweep (conv_tc.cu, v5): ring A =
//     6 blocks (lap of 4 input columns, TMEM columns 0..191), ring B = 10 blocks (lap of 8, columns 192..511); input column t
//     accumulates into blocks [k, k+1, k+2] of its lap with one N = 96 instruction per (lane tap, k-step).
//   * two MMA-issuing warps, one per ring.  Step t of warp 1:  P[t] x W_A -> ring A  (commit: A-output t-1 complete);
//     step t of warp 2:  P[t] x W_B[:, P part] -> ring B, then XK[t-3] x W_B[:, x_k part] -> ring B (commit: B-output t-4
//     complete, XK slot free).  A P slab returns to the producer when both have committed it.  The lag of three steps covers
//     the latency of commit -> epilogue A -> shared memory -> MMA (~3 000 clocks); ring B's lap of 8 leaves two more steps for
//     its drain.  One issuer per ring keeps the order of every accumulator's additions fixed (bit-reproducible).
//   * four epilogue groups of four warps: two alternate on the outputs of ring A, two on ring B.  Ring A's groups apply bias /
//     LeakyReLU / packed masks, round to bf16, and write the column (a) into the XK ring in shared memory, in the SWIZZLE_128B
//     K-major layout the MMA reads (two columns share one [130 lanes][128 B] slab: 64 B each), zeros outside the image (the
//     padding layer B must see), (b) to global memory (the x_k slice).  Lanes 127 / 128 also go to the peer CTA's halo rows with
//     st.async (the store completes a transaction count on an mbarrier of the destination CTA: no cluster-scope fence, which
//     would wait for the warp's global stores).  Ring B's groups write x_(k+1) to global memory.
//   * work: the n * w input columns of a launch are dealt to the clusters as equal contiguous ranges (cut at image boundaries into
//     units); every unit sweeps 2 + 2 halo columns.  Static assignment: bit-reproducible.
// Accumulation order differs from the unfused kernel (different laps, P before x_k), so results agree with it to fp32 rounding,
// not bit for bit.

struct SwfLayer {
  const float* bias;
  __nv_bfloat16* y; int y_ld;
  int act; float act_slope;
  uint32_t* signbits;                   // OUT [pixel]: bit c = stored value of channel c > 0
  const uint32_t* maskbits;             // IN  [pixel]: value *= bit ? 1 : mask_slope
  float mask_slope;
};
struct SwfArgs {
  int n, cin, h, w;                     // cin = channels of P (64 | 128); h = lane extent (<= 256), w = sweep extent
  int nchunks, na;                      // K chunks of P; depth of the P slab ring
  int per, total;                       // input columns per cluster; n * w
  int tr;
  long long lane_stride, sweep_stride, img_stride;
  const __nv_bfloat16* wgt_a;           // packed like the sweep's: [chunk][slot = kh*3 + (2-kw)][32][64]
  const __nv_bfloat16* wgt_b;           // cin + 32 input channels: nchunks + 1 chunks, the last one half used
  SwfLayer L[2];
  int pfd;                              // L2 prefetch distance of the producer, in columns (0 = off)
  int issuers;                          // 2 (default): one MMA-issuing warp per accumulator ring | 1: one warp issues everything
  int dbg;                              // 1: no global stores; 32: clock profile of the MMA warp; 128: no halo exchange (wrong seam)
};

constexpr int SWF_BN = 32, SWF_N = 3 * SWF_BN;
constexpr int SWF_RUN_A = 4, SWF_RUN_B = 8;                            // laps; blocks = RUN + 2: 6 + 10 = 16 x 32 TMEM columns
                                                                       // (5 / 7 measures the same, 6 / 6 is 6-8 % slower)
constexpr int SWF_TB0 = (SWF_RUN_A + 2) * SWF_BN, SWF_YB0 = SWF_RUN_A + 2;   // ring B: first TMEM column, first barrier index
constexpr int SWF_W_KH_BYTES = SWF_N * 128 / 2, SWF_W_CHUNK_BYTES = 3 * SWF_W_KH_BYTES;     // per CTA of the pair
constexpr int SWF_NXK = 4;                                                                  // XK column slots (two per slab)
constexpr int SWF_XK_BYTES = (SWF_NXK / 2) * SW_SLAB_STRIDE;
constexpr int SWF_GROUPS = 4;                                          // epilogue groups: (ring, output parity)
constexpr int SWF_THREADS = 96 + SWF_GROUPS * 128;                     // producer, two MMA issuers, 16 epilogue warps

static size_t swf_smem_bytes(int nchunks, int na) {
  return (size_t)(2 * nchunks + 1) * SWF_W_CHUNK_BYTES + SWF_XK_BYTES + (size_t)na * SW_SLAB_STRIDE + SMEM_AUX + 1024;
}

__device__ __forceinline__ bool mbar_test_wait_cluster(uint64_t* bar, uint32_t parity) {   // non-blocking, observes a peer's release
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
// 16 bytes into a peer CTA's shared memory; when they have landed, 16 is subtracted from the pending transaction count of `cbar`,
// an mbarrier of the SAME destination CTA (both addresses in the shared::cluster window)
__device__ __forceinline__ void st_async_u4(uint32_t caddr, const uint4& v, uint32_t cbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(caddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cbar)
               : "memory");
}

__global__ void __launch_bounds__(SWF_THREADS, 1)
conv3x3_pair_sweep_tc(const __grid_constant__ CUtensorMap tmap_x, const SwfArgs a) {
  constexpr int BN = SWF_BN, RUN_A = SWF_RUN_A, RUN_B = SWF_RUN_B, YB0 = SWF_YB0, CG = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_wa = smem;
  uint8_t* smem_wb = smem_wa + (size_t)a.nchunks * SWF_W_CHUNK_BYTES;
  uint8_t* smem_xk = smem_wb + (size_t)(a.nchunks + 1) * SWF_W_CHUNK_BYTES;
  uint8_t* smem_a = smem_xk + SWF_XK_BYTES;
  uint8_t* aux = smem_a + (size_t)a.na * SW_SLAB_STRIDE;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(aux);                 // [SW_MAX_NA]  (leader's copy is the live one)
  uint64_t* a_empty = a_full + SW_MAX_NA;                              // [SW_MAX_NA]  (multicast commits: each CTA's own copy)
  uint64_t* y_full = a_empty + SW_MAX_NA;                              // [16]: ring A = 0..5, ring B = 6..15
  uint64_t* y_empty = y_full + 16;                                     // [16] (leader's copy is the live one)
  uint64_t* w_full = y_empty + 16;
  uint64_t* w_pair = w_full + 1;
  uint64_t* xk_full = w_pair + 1;                                      // [SWF_NXK] (leader's copy is the live one)
  uint64_t* xk_empty = xk_full + SWF_NXK;                              // [SWF_NXK] (multicast commits)
  uint64_t* xk_halo = xk_empty + SWF_NXK;                              // [SWF_NXK] second CTA: the leader's lane 127 has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xk_halo + SWF_NXK);
  float* sbias = reinterpret_cast<float*>(aux + 768);                  // [2][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = (int)(blockIdx.x / CG);
  const bool no_halo = (a.dbg & 128) != 0;
  pdl_launch_dependents();                                             // the next kernel's prologue may overlap this kernel's tail
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.na; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], a.issuers == 2 ? 2 : 1); }
    for (int s = 0; s < 16; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 4 * CG); }
    mbar_init(w_full, 1);
    mbar_init(w_pair, 1);
    for (int s = 0; s < SWF_NXK; ++s) { mbar_init(&xk_full[s], 4 * CG); mbar_init(&xk_empty[s], 1); mbar_init(&xk_halo[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * BN; i += SWF_THREADS) {
    const float* b = i < BN ? a.L[0].bias : a.L[1].bias;
    sbias[i] = b ? b[i % BN] : 0.f;
  }
  // the XK slabs start as zeros: halo row 0 of the first CTA and halo row 129 of the second one are never written again
  for (int i = threadIdx.x; i < SWF_XK_BYTES / 16; i += SWF_THREADS) sts_u4(smem_u32(smem_xk) + (uint32_t)i * 16u, make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async_smem();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 3 && warp < 7) {                                         // every accumulator block starts at zero
    const uint32_t tq = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
    for (int cc = 0; cc < 512; cc += 32) tmem_st32_zero(tq + cc);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                                  // barriers initialised, TMEM and XK zeroed in both CTAs
  tc_fence_after();

  // this cluster's contiguous range of the launch's n * w input columns, cut at image boundaries: unit i = (img, [xs, xe))
  const int g0 = cid * a.per;
  const int g1 = (g0 + a.per) < a.total ? (g0 + a.per) : a.total;
  auto unit = [&](int i, int& img, int& xs, int& xe) -> bool {
    img = g0 / a.w + i;
    const int base = img * a.w;
    if (g0 >= g1 || base >= g1) return false;
    xs = i == 0 ? g0 - base : 0;
    xe = (g1 - base) < a.w ? (g1 - base) : a.w;
    return true;
  };
  const int y0 = (int)rank * SW_ROWS;

  if (warp == 0) {
    // ---- producer: this CTA's half of both layers' stacked weight rows once, then one slab per (P column, K chunk)
    if (elect_one()) {
      mbar_expect_tx(w_full, (uint32_t)((2 * a.nchunks + 1) * SWF_W_CHUNK_BYTES));
      constexpr int PIECE = BN * 128 / CG;
      for (int l = 0; l < 2; ++l) {
        const __nv_bfloat16* wg = l ? a.wgt_b : a.wgt_a;
        uint8_t* dstw = l ? smem_wb : smem_wa;
        const int nch = a.nchunks + l;
        for (int i = 0; i < nch; ++i)
          for (int t = 0; t < 3; ++t)
            for (int h3 = 0; h3 < 3; ++h3) {
              const int hs = (int)rank * 3 + h3;                       // piece index in the stacked tile [s = 2 | 1 | 0] x 32 rows
              const int si = hs >> 1, part = hs & 1;
              const int sw = 2 - si;                                   // sweep-direction tap
              const int kh = a.tr ? sw : t, kw = a.tr ? t : sw;
              const int slot = kh * 3 + (2 - kw);
              bulk_load(wg + ((size_t)(i * 9 + slot) * (BN * 128) + (size_t)part * PIECE) / 2, w_full,
                        dstw + (size_t)(i * 3 + t) * SWF_W_KH_BYTES + (size_t)h3 * PIECE, PIECE);
            }
      }
    }
    __syncwarp();
    if (rank == 1) {
      mbar_wait(w_full, 0);
      if (elect_one()) mbar_arrive_leader(w_pair);
      __syncwarp();
    }
    int as = 0;
    uint32_t aph = 0;
    pdl_wait();                                                        // the prefix is the previous kernels' output
    for (int ui = 0;; ++ui) {
      int img, xs, xe;
      if (!unit(ui, img, xs, xe)) break;
      // the slab ring is short (weights of two layers + the XK ring share the 227 KB): pull the columns ahead into L2
      if (a.pfd > 0 && elect_one())
        for (int c = xs - 2; c < xs - 2 + a.pfd && c <= xe + 1; ++c)
          for (int k = 0; k < a.nchunks; ++k) tma_prefetch_4d(&tmap_x, k * KCH, c, y0 - 1, img);
      __syncwarp();
      for (int c = xs - 2; c <= xe + 1; ++c)
        for (int k = 0; k < a.nchunks; ++k) {
          mbar_wait(&a_empty[as], aph ^ 1);
          if (elect_one()) {
            if (a.pfd > 0 && c + a.pfd <= xe + 1) tma_prefetch_4d(&tmap_x, k * KCH, c + a.pfd, y0 - 1, img);
            if (rank == 0) mbar_expect_tx(&a_full[as], 2 * SW_SLAB_BYTES);
            tma_load_4d_pair(&tmap_x, &a_full[as], smem_a + (size_t)as * SW_SLAB_STRIDE, k * KCH, c, y0 - 1, img);
          }
          __syncwarp();
          if (++as == a.na) { as = 0; aph ^= 1; }
        }
    }
  } else if (warp == 1 || warp == 2) {
    // ---- MMA issuers (leader CTA).  One accumulator ring per issuing warp (default): warp 1 streams P through W_A into ring A,
    // warp 2 streams P through W_B's prefix part and, three columns behind, XK through W_B's x_k part into ring B.  Every ring
    // has ONE issuer, so the order of its fp32 additions is fixed (bit-reproducible), while the waits of one stream (the tensor
    // pipe's instruction queue holds only ~4 MMAs: any wait above ~200 clocks runs it dry) are covered by the other stream's
    // instructions.  A P slab goes back to the producers when both streams have committed it (a_empty counts two).
    // SRCGAN_B200_PAIR_ISSUERS=1: one warp issues everything (the layout of the first version; ~10-25 % slower).
    const bool split = a.issuers == 2;
    const bool do_a = warp == 1, do_b = warp == 2 || !split;
    if (rank == 0 && (warp == 1 || split)) {
      constexpr uint32_t idesc = umma_idesc(TILE_M * CG, SWF_N);
      constexpr uint32_t hi = desc_hi(1024);
      constexpr int LAG = 3;                                           // the x_k part trails the P stream by three columns
                                                                       // (2 and 4, and an XK ring of 8, measure the same)
      mbar_wait(w_full, 0);
      mbar_wait(w_pair, 0);
      const uint32_t wa_lo = desc_lo(smem_u32(smem_wa)), wb_lo = desc_lo(smem_u32(smem_wb));
      const uint32_t xk_lo = desc_lo(smem_u32(smem_xk));
      const uint32_t wbx_lo = wb_lo + (uint32_t)(a.nchunks * (SWF_W_CHUNK_BYTES >> 4));
      int as = 0;
      uint32_t aph = 0;
      int ka = 0, kb = 0;                                              // lap positions of the P stream in ring A / ring B
      uint32_t lapa = 0, lapb = 0;
      int t = 0;
      int kx = 0, xslot = 0;                                           // lap position (ring B) / slot of the XK stream
      uint32_t xph = 0;
      long long tp[6] = {0, 0, 0, 0, 0, 0}, tc0 = 0;
      const bool prof = (a.dbg & 32) != 0;
      const long long t_begin = clock64();
#define SRCGAN_TICK(i) if (prof) { const long long n_ = clock64(); tp[i] += n_ - tc0; tc0 = n_; }
      auto issue_bx = [&]() {                                          // XK[t'] x W_B[:, x_k part] -> ring B
        mbar_wait_cluster(&xk_full[xslot], xph);
        fence_proxy_async_smem();                                      // the peer's halo lane arrived by st.async (generic proxy)
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(SWF_TB0 + kx * BN);
        const uint32_t a_lo = xk_lo + (uint32_t)((xslot >> 1) * (SW_SLAB_STRIDE >> 4) + (xslot & 1) * 4);
        if (elect_one()) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_bf16_w2(d, a_lo + (uint32_t)(kh * 8 + ks * 2), hi, wbx_lo + (uint32_t)(kh * (SWF_W_KH_BYTES >> 4) + ks * 2), hi,
                           idesc, 1u);
          umma_commit_pair(&xk_empty[xslot]);
          umma_commit_pair(&y_full[YB0 + kx]);                         // B-output (kx - 1) of this lap is complete
          if (kx == RUN_B - 1) { umma_commit_pair(&y_full[YB0 + RUN_B]); umma_commit_pair(&y_full[YB0 + RUN_B + 1]); }
        }
        __syncwarp();
        if (++xslot == SWF_NXK) { xslot = 0; xph ^= 1; }
        if (++kx == RUN_B) kx = 0;
      };
      for (int ui = 0;; ++ui) {
        int img, xs, xe;
        if (!unit(ui, img, xs, xe)) break;
        for (int c = xs - 2; c <= xe + 1; ++c) {
          if (prof) tc0 = clock64();
          int s = as;
          uint32_t ph = aph;
          if (do_a) {
            // ring A: blocks touched for the first time in this lap must have been drained (and zeroed)
            if (ka == 0) { mbar_wait(&y_empty[0], lapa ^ 1); mbar_wait(&y_empty[1], lapa ^ 1); }
            mbar_wait(&y_empty[ka + 2], lapa ^ 1);
            tc_fence_after();
            SRCGAN_TICK(0)
            const uint32_t da = tmem_base + (uint32_t)(ka * BN);
            for (int kc = 0; kc < a.nchunks; ++kc) {
              mbar_wait(&a_full[s], ph);
              tc_fence_after();
              const uint32_t a_lo = desc_lo(smem_u32(smem_a + (size_t)s * SW_SLAB_STRIDE));
              const uint32_t b_lo = wa_lo + (uint32_t)(kc * (SWF_W_CHUNK_BYTES >> 4));
              if (elect_one()) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_bf16_w2(da, a_lo + (uint32_t)(kh * 8 + ks * 2), hi, b_lo + (uint32_t)(kh * (SWF_W_KH_BYTES >> 4) + ks * 2), hi,
                                 idesc, 1u);
                if (split) umma_commit_pair(&a_empty[s]);              // this stream is done with the slab
                if (kc == a.nchunks - 1) {
                  umma_commit_pair(&y_full[ka]);                       // A-output (ka - 1) of this lap is complete
                  if (ka == RUN_A - 1) { umma_commit_pair(&y_full[RUN_A]); umma_commit_pair(&y_full[RUN_A + 1]); }
                }
              }
              __syncwarp();
              if (++s == a.na) { s = 0; ph ^= 1; }
            }
            SRCGAN_TICK(1)
          }
          if (do_b) {
            // ring B, P part (same slabs)
            if (kb == 0) { mbar_wait(&y_empty[YB0], lapb ^ 1); mbar_wait(&y_empty[YB0 + 1], lapb ^ 1); }
            mbar_wait(&y_empty[YB0 + kb + 2], lapb ^ 1);
            tc_fence_after();
            SRCGAN_TICK(2)
            const uint32_t db = tmem_base + (uint32_t)(SWF_TB0 + kb * BN);
            s = as;
            ph = aph;
            for (int kc = 0; kc < a.nchunks; ++kc) {
              if (!do_a) { mbar_wait(&a_full[s], ph); tc_fence_after(); }
              const uint32_t a_lo = desc_lo(smem_u32(smem_a + (size_t)s * SW_SLAB_STRIDE));
              const uint32_t b_lo = wb_lo + (uint32_t)(kc * (SWF_W_CHUNK_BYTES >> 4));
              if (elect_one()) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_bf16_w2(db, a_lo + (uint32_t)(kh * 8 + ks * 2), hi, b_lo + (uint32_t)(kh * (SWF_W_KH_BYTES >> 4) + ks * 2), hi,
                                 idesc, 1u);
                umma_commit_pair(&a_empty[s]);                         // slab back to both CTAs' producers
              }
              __syncwarp();
              if (++s == a.na) { s = 0; ph ^= 1; }
            }
            SRCGAN_TICK(3)
            if (t >= LAG) issue_bx();
            SRCGAN_TICK(4)
          }
          as = s;
          aph = ph;
          ++t;
          if (++ka == RUN_A) { ka = 0; lapa ^= 1; }
          if (++kb == RUN_B) { kb = 0; lapb ^= 1; }
        }
      }
      // XK[T - LAG] .. XK[T - 2]: the last one completes B-output T - 3, the stream's last real one
      if (do_b)
        for (int i = 0; i < LAG - 1; ++i)
          if (t >= LAG - i) issue_bx();
      if (prof && (blockIdx.x % 32 == 0) && lane == 0 && t > 0)
        printf("pair mma warp %d (cta %d): cols %d  yA_empty %lld  a_full+issue A %lld  yB_empty %lld  issue B %lld  xk part %lld (clk/col)  total %lld clk\n",
               warp, (int)blockIdx.x, t, tp[0] / t, tp[1] / t, tp[2] / t, tp[3] / t, tp[4] / t, clock64() - t_begin);
#undef SRCGAN_TICK
    }
  } else {
    // ---- epilogues: group g = (ring, parity): groups 0 / 2 take the even / odd outputs of ring A (and feed the XK ring),
    // groups 1 / 3 those of ring B.  Running output index o = running input index of the P column it is centred on; lap
    // r = o / RUN, m = o % RUN: main block m + 1 of lap r, plus block 0 of lap r + 1 when m == RUN - 1, plus block RUN + 1 of
    // lap r - 1 when m == 0.
    const int q = warp & 3;
    const int g = (warp - 3) >> 2;
    const int gb = g & 1, par = g >> 1;
    const int RUN = gb ? RUN_B : RUN_A;
    const int ring = gb ? YB0 : 0;
    __nv_bfloat16* const Ly = gb ? a.L[1].y : a.L[0].y;
    const int Ly_ld = gb ? a.L[1].y_ld : a.L[0].y_ld;
    const int Lact = gb ? a.L[1].act : a.L[0].act;
    const float Lslope = gb ? a.L[1].act_slope : a.L[0].act_slope;
    uint32_t* const Lsign = gb ? a.L[1].signbits : a.L[0].signbits;
    const uint32_t* const Lmask = gb ? a.L[1].maskbits : a.L[0].maskbits;
    const float Lmslope = gb ? a.L[1].mask_slope : a.L[0].mask_slope;
    const uint32_t sbias_addr = smem_u32(sbias) + (uint32_t)(gb * BN * 4);
    const int row = q * 32 + lane;
    const int y = y0 + row;
    const bool row_in = y < a.h;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(gb ? SWF_TB0 : 0);
    const bool wide_st = (Ly_ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(Ly) & 31) == 0);
    auto arrive_empty = [&](int blk) { mbar_arrive_leader(&y_empty[ring + blk]); };
    pdl_wait();                                                        // packed masks in, outputs the previous kernels may still read
    int T = 0;                                                         // input columns of this cluster's stream
    for (int ui = 0;; ++ui) {
      int img, xs, xe;
      if (!unit(ui, img, xs, xe)) break;
      T += xe - xs + 4;
    }
    if (T > 0) {
      if (par == 1) {                                                  // "output -1": block 0 of lap 0 collects only a halo tap
        mbar_wait(&y_full[ring], 0);
        tc_fence_after();
        tmem_st32_zero(lane_base);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_empty(0);
      }
      const int n_out = T - 1 - gb;                                    // A: outputs 0 .. T-2 (XK[T-2] is the last one used); B: 0 .. T-3
      int ui = 0, j = par, img = 0, xs = 0, xe = 0;
      unit(0, img, xs, xe);
      int n_in = xe - xs + 4;
      // this thread's 64 bytes of an XK column: row R = row + 1 of the slab, 16-byte chunks (half * 4 + i) ^ (R & 7)
      const uint32_t xk_base = smem_u32(smem_xk);
      const int R = row + 1;
      const bool halo_lane = !no_halo && ((rank == 0 && row == SW_ROWS - 1) || (rank == 1 && row == 0));
      const int Rp = rank == 0 ? 0 : SW_SLAB_ROWS - 1;                 // the halo row this lane also fills in the peer CTA's slab
      for (int o = par; o < n_out; o += 2) {
        const int x = xs - 2 + j;
        const bool col_in = x >= 0 && x < a.w;
        const bool own = x >= xs && x < xe && row_in && !(a.dbg & 1);  // stored to global memory by this unit
        // ring A also needs the unit's two halo columns of x_k (inside the image) for layer B
        // (warp-uniform: the TMEM loads / stores below are warp-collective; rows below the image only suppress values and stores)
        const bool need = gb == 0 ? (j >= 1 && j <= n_in - 2 && col_in) : (x >= xs && x < xe);
        const int m = o % RUN;
        const uint32_t r = (uint32_t)(o / RUN);
        const int pb = m + 1;
        const uint32_t ppar = r & 1u;
        const int sb = m == RUN - 1 ? 0 : ((m == 0 && o > 0) ? RUN + 1 : -1);
        const uint32_t spar = ppar ^ 1u;
        const long long pix = (long long)img * a.img_stride + (long long)y * a.lane_stride + (long long)x * a.sweep_stride;
        uint32_t pbits = 0xFFFFFFFFu;
        if (need && row_in && Lmask) pbits = __ldg(Lmask + pix);
        mbar_wait(&y_full[ring + pb], ppar);
        if (sb >= 0) mbar_wait(&y_full[ring + sb], spar);
        tc_fence_after();
        const uint32_t taddr = lane_base + (uint32_t)(pb * BN);
        const uint32_t saddr = lane_base + (uint32_t)((sb >= 0 ? sb : 0) * BN);
        float f[32];
        if (!need) {
          tmem_st32_zero(taddr);
          if (sb >= 0) tmem_st32_zero(saddr);
        } else if (sb >= 0) {
          uint32_t z[32], z2[32];
          tmem_ld32_nowait(taddr, z);
          tmem_ld32_nowait(saddr, z2);
          tmem_ld_wait();
          tmem_st32_zero(taddr);
          tmem_st32_zero(saddr);
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(z[i]) + __uint_as_float(z2[i]);
        } else {
          uint32_t z[32];
          tmem_ld32(taddr, z);
          tmem_st32_zero(taddr);
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(z[i]);
        }
        // the blocks go back to the MMA warp before the arithmetic: drained and zeroed (one arrival per warp)
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          arrive_empty(pb);
          if (sb >= 0) arrive_empty(sb);
        }
        uint4 ov[4];
        uint32_t sbits = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) ov[i] = make_uint4(0u, 0u, 0u, 0u);
        if (need && row_in) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = lds_f4(sbias_addr + (uint32_t)i * 4);
            f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
          }
          if (Lact) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], f[i] * Lslope);
          }
          if (Lmask) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] *= ((pbits >> i) & 1u) ? 1.f : Lmslope;
          }
          if (Lsign) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sbits |= (f[i] > 0.f ? 1u : 0u) << i;
          }
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&ov[gq]);
#pragma unroll
            for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(f[gq * 8 + 2 * i], f[gq * 8 + 2 * i + 1]);
          }
        }
        if (gb == 0) {
          // XK[o]: the bf16 values about to be stored (zeros outside the image / in the dummy columns of a unit boundary),
          // BEFORE the global stores - layer B's MMAs wait for this
          const int xslot = o & (SWF_NXK - 1);
          const uint32_t xpar = (uint32_t)(o / SWF_NXK) & 1u;
          mbar_wait(&xk_empty[xslot], xpar ^ 1);                       // XK[o - 4] has been multiplied
          const uint32_t slab = xk_base + (uint32_t)((xslot >> 1) * SW_SLAB_STRIDE);
          const int half4 = (xslot & 1) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) sts_u4(slab + (uint32_t)(R * 128 + (((half4 + i) ^ (R & 7)) << 4)), ov[i]);
          if (halo_lane) {
            // lane 127 of the first CTA is halo row 0 of the second one, lane 128 halo row 129 of the first: st.async into the peer,
            // completing on the barrier its consumer waits on (first CTA: xk_full itself; second CTA: xk_halo, forwarded below)
            const uint32_t peer = mapa_u32(slab, rank ^ 1u);
            const uint32_t pbar = mapa_u32(smem_u32(rank == 0 ? &xk_halo[xslot] : &xk_full[xslot]), rank ^ 1u);
#pragma unroll
            for (int i = 0; i < 4; ++i) st_async_u4(peer + (uint32_t)(Rp * 128 + (((half4 + i) ^ (Rp & 7)) << 4)), ov[i], pbar);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) {
              if (q == 3 && !no_halo) mbar_expect_tx(&xk_full[xslot], 64); else mbar_arrive(&xk_full[xslot]);
            } else {
              if (q == 0 && !no_halo) {                                // the leader's lane 127 must have landed in this CTA's slab
                mbar_expect_tx(&xk_halo[xslot], 64);
                mbar_wait(&xk_halo[xslot], xpar);
                fence_proxy_async_smem();
              }
              mbar_arrive_leader(&xk_full[xslot]);
            }
          }
        }
        if (own) {
          if (Lsign) Lsign[pix] = sbits;
          __nv_bfloat16* dst = Ly + pix * Ly_ld;                       // this pixel's 32 channels: 64 contiguous bytes
          if (wide_st) {
            stg_u8(dst, ov[0], ov[1]);
            stg_u8(dst + 16, ov[2], ov[3]);
          } else {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) *reinterpret_cast<uint4*>(dst + gq * 8) = ov[gq];
          }
        }
        j += 2;
        while (j >= n_in) {
          j -= n_in;
          ++ui;
          if (!unit(ui, img, xs, xe)) break;
          n_in = xe - xs + 4;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                                  // the peer may still signal this CTA's barriers
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}
