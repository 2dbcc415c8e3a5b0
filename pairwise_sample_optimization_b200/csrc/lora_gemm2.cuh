// Two-SM variant of the LoRA projection GEMM: tcgen05.mma.cta_group::2 on a CTA PAIR (cluster of 2).
//
// Why: with cta_group::1 a CTA must ingest 48 KB of operands per 64-wide k-block (A 128 x 64, B 256 x 64) for 512
// cycles of MMA work; measured, the SM takes ~740 cycles to ingest that much, which caps the tensor pipe at ~69 %
// (profiles/r01_lora_gemm.md).  A CTA pair computes a 256 x bn tile: each CTA holds ITS 128 rows of A and HALF of the B
// tile (bn/2 rows), the MMA reads both halves of B across the pair, and each CTA's tensor memory receives its 128 rows
// of the accumulator.  Per-SM ingest drops to 32 KB per k-block for the same 512 cycles of MMA.
//
// Same roles as lora_gemm_kernel; differences:
//   * both CTAs run a TMA producer for their own halves; every copy signals the LEADER's (rank 0) full barrier
//     (cp.async.bulk.tensor ... .cta_group::2 with the barrier address mapped to the even CTA of the pair);
//   * only the leader issues MMAs (M = 256) and commits; commits are MULTICAST to both CTAs' barriers (empty slots,
//     accumulator full);
//   * the peer's epilogue warps release the accumulator on the leader's barrier (remote mbarrier arrive);
//   * tensor memory is allocated / freed with .cta_group::2 by the same warp of both CTAs.
// Supports: problem lists with in-launch dependencies, reduction segments, stacked column groups, bias, alpha, K-major or
// reduction-major B, transposed output copy, PDL waits.  Not: reduction-major A, atomics, split-K (the 1-SM kernel keeps those).
#pragma once

namespace psob200 {

// A stage of the pair kernel holds TWO 64-deep k-blocks (128 reduction elements): the barrier round trip producer -> MMA warp ->
// tcgen05.commit -> producer is paid once per 128 elements.  Measured with loads, MMAs and stores all disabled
// (tools/diag_gemm3.py), the bare hand-shake of 64-deep stages cost 0.21 us per k-block -- 15.7 of the 26.0 us of a
// (8192,1280,1280) launch -- more than the MMAs themselves.
constexpr int kBK2 = 2 * kBK;                                                 // reduction elements per stage
constexpr int kStageA2Bytes = 2 * kStageABytes;                               // 32 KB: two [128 x 64] swizzled sub-tiles
constexpr int kStageB2Bytes = 2 * (kBNMax / 2) * kBK * 2;                     // 32 KB: two halves of the widest B tile's half
constexpr int kStages2 = 3;                                                   // 3 x (32 + 32) KB
constexpr int kGemm2SmemBytes = kStages2 * (kStageA2Bytes + kStageB2Bytes) + 4 * kEpiStageBytesPerWarp + 1024;  // + output staging
constexpr int kMaxAcc2 = 3;                                                   // accumulators in tensor memory (2 x 256 or 3 x 160 columns)
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;                                // shared::cluster address -> even CTA of the pair

namespace ptx {
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_addr(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_addr(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the even CTA's copy of `bar`
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_addr(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
}  // namespace ptx

// Timeline experiment (diag bit 0x40000): per CTA, 8 time stamps (globaltimer ns, SM cycle counter) -- 0 entry, 1 prologue done,
// 2 first operands landed (leader), 3 first accumulator complete, 4 first tile drained, 5 last accumulator complete, 6 last tile
// drained, 7 exit.  Read back with psob200_lora_gemm_timeline (tools/diag_timeline.py).
constexpr int kTimelineCtas = 512, kTimelineSlots = 8;
__device__ unsigned long long g_timeline[kTimelineCtas * kTimelineSlots * 2];
__device__ __forceinline__ void timeline_stamp(int diag, int slot) {
  if ((diag & 0x40000) && blockIdx.x < kTimelineCtas) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g_timeline[(blockIdx.x * kTimelineSlots + slot) * 2] = ns;
    g_timeline[(blockIdx.x * kTimelineSlots + slot) * 2 + 1] = (unsigned long long)clock64();
  }
}

// The order in which a pair's producer and MMA warps walk its tiles: tile after tile -- except when the pair's FIRST tile waits for
// rows of this launch (its adapter segment reads the t / u rows that the t / u tiles of the same first wave are still writing) and
// a second such tile follows.  Then:  base(T0) -> base(T1) -> adapter(T0) -> adapter(T1)  with both tensor-memory accumulators
// open.  Measured (tools/diag_timeline_group.py, (8192,1280,1280), r = 64): a first-wave main tile had its base segments done at
// ~9 us and its accumulator complete at 17 us -- t tile 8.7 us, drain + release 4 us, acquire + load + MMA 4 us -- on 42 of the 74
// pairs, 12 of which then ran two more tiles: the launch took 37 us against 25 us for the frozen pass.
struct TileWalk {
  bool couple;
  __device__ __forceinline__ void init(const GemmLaunch& L, int cluster_id, int n_clusters) {
    couple = false;
    const int t1 = cluster_id + n_clusters;
    if (t1 < L.total_tiles && !(L.diag & 0x80000)) {
      TileInfo a, b;
      decode_tile(L, cluster_id, a);
      decode_tile(L, t1, b);
      couple = L.prob[a.p].wait_seg > 0 && L.prob[b.p].wait_seg > 0;
    }
  }
  // part -> (index of the tile in this pair's sequence, which: 0 whole tile, 1 the segments before wait_seg, 2 the rest)
  __device__ __forceinline__ void part(int part, int& it, int& which) const {
    if (couple && part < 4) { it = part & 1; which = part < 2 ? 1 : 2; }
    else { it = couple ? part - 2 : part; which = 0; }
  }
};
__device__ __forceinline__ int wait_kblock(const GemmProblem& P) {  // first k-block of segment wait_seg
  int kb = 0;
  for (int s = 0; s < P.wait_seg; ++s) kb += P.nk[s];
  return kb;
}

template <typename TD, int kMode>  // one instantiation per output type / bias presence (code size: see lora_gemm_kernel)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
lora_gemm2_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ GemmLaunch L) {
  extern __shared__ unsigned char gemm_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages2], empty_bar[kStages2], tmem_full_bar[kMaxAcc2], tmem_empty_bar[kMaxAcc2];
  __shared__ uint32_t tmem_base_slot;

  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + kStages2 * kStageA2Bytes;
  unsigned char* smem_epi = smem + kStages2 * (kStageA2Bytes + kStageB2Bytes);  // 4 x 4 KB: one staging buffer per epilogue warp

  // the warp index through a shuffle is PROVABLY warp-uniform: ptxas then keeps the role loops (barrier phases, stage counters,
  // UMMA / TMA descriptors) on the uniform datapath instead of moving ~20 per-thread registers to uniform ones (R2UR) per k-block
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) timeline_stamp(L.diag, 0);
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int stage_b_bytes = L.stage_b_bytes;  // two 64-deep sub-tiles of half of the widest problem's B tile
  const int sub_b_bytes = stage_b_bytes / 2;
  const int total_tiles = L.total_tiles;      // m_tiles count 256-row pair tiles here
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  // tensor-memory accumulators: two of 256 columns, or -- when no tile of the launch is wider than 160 -- three of 160: a pair
  // whose first two tiles are interleaved (TileWalk) then starts its third tile while the first two drain
  const int n_acc = L.n_acc, acc_stride = L.acc_stride;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < L.n_prob; ++p)
      for (int s = 0; s < L.prob[p].n_seg; ++s) {
        ptx::prefetch_tensormap(&maps.m[L.prob[p].map_a[s]]);
        ptx::prefetch_tensormap(&maps.m[L.prob[p].map_b[s]]);
      }
  }
  if (L.pdl & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 1 && lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages2; ++s) ptx::mbar_init(&full_bar[s], 1);  // leader's arrive.expect_tx (both CTAs' bytes)
#pragma unroll
    for (int s = 0; s < kStages2; ++s) ptx::mbar_init(&empty_bar[s], 1);  // multicast tcgen05.commit
#pragma unroll
    for (int a = 0; a < kMaxAcc2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);   // multicast tcgen05.commit
      ptx::mbar_init(&tmem_empty_bar[a], 8);  // 4 epilogue warps of each CTA (waited on by the leader only)
    }
    ptx::fence_mbar_init();
  }
  ptx::cluster_sync_all();  // both CTAs are resident and their barriers initialised before anything crosses the pair
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_addr(&tmem_base_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // second cluster barrier (the pair's tensor memory is allocated): the TMA producer only ARRIVES and starts loading -- it never
  // touches tensor memory -- and completes its wait after its loop (the first operands land while the others still wait here)
  ptx::tc_fence_before_sync();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  if (warp != 0) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;  // (not meaningful in warp 0)
  if (threadIdx.x == 0) timeline_stamp(L.diag, 1);

  if (warp == 0) {
    // ================================================================= TMA producer (both CTAs: own A rows, own half of B)
    int stage = 0;
    uint32_t phase = 0;
    bool dep_pending = (L.pdl & 2) != 0;
    if (dep_pending && (L.pdl & 4)) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
      dep_pending = false;
    }
    TileWalk walk;
    walk.init(L, cluster_id, n_clusters);
    for (int part = 0;; ++part) {
      int it, which;
      walk.part(part, it, which);
      const int t = cluster_id + it * n_clusters;
      if (t >= total_tiles) break;
      TileInfo ti;
      decode_tile(L, t, ti);
      const GemmProblem& P = L.prob[ti.p];
      if (which == 1) ti.kb1 = wait_kblock(P);
      if (which == 2) ti.kb0 = wait_kblock(P);
      const int half_bn = P.bn / 2;
      const uint32_t tx_pair = 2u * 2u * ((uint32_t)kStageABytes + (uint32_t)(half_bn * kBK * 2));  // 2 CTAs x 2 sub-tiles
      const int m0 = ti.m_blk * 256 + (int)rank * kBM, n0 = (int)ti.n0 + (int)rank * half_bn;
      int seg = 0, kbs = ti.kb0;
      while (kbs >= P.nk[seg]) { kbs -= P.nk[seg]; ++seg; }
      bool need_wait = P.wait_seg >= 0;
      for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (seg >= 1 && dep_pending) {
          asm volatile("griddepcontrol.wait;" ::: "memory");
          asm volatile("fence.proxy.async;" ::: "memory");
          dep_pending = false;
        }
        if (need_wait && seg >= P.wait_seg) {  // rows produced by other tiles of this launch (8 epilogue warps per pair tile)
          wait_flag(L.flags + ti.m_blk, P.wait_count * 8);
          need_wait = false;
        }
        const CUtensorMap* ma = &maps.m[P.map_a[seg]];
        const CUtensorMap* mb = &maps.m[P.map_b[seg]];
        const int ka = kbs * kBK2 + P.a_koff[seg] + ti.group * P.a_gkoff[seg];
        const int kk = kbs * kBK2;
        const int boff = P.b_off[seg];
        unsigned char* sa = smem_a + stage * kStageA2Bytes;
        unsigned char* sb = smem_b + stage * stage_b_bytes;
        if (ptx::elect_one()) {
          if (L.diag & 12) {  // timing experiment: no operand loads (results are wrong)
            if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 0u);
          } else {
            if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_pair);
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // the two 64-deep sub-tiles (beyond the segment's K the TMA unit fills zeros)
              ptx::tma_load_2d_pair(sa + h * kStageABytes, ma, ka + h * kBK, m0, &full_bar[stage]);
              if (L.b_mn) {
                for (int j = 0; j < half_bn / 64; ++j)
                  ptx::tma_load_2d_pair(sb + h * sub_b_bytes + j * (kBK * 128), mb, n0 + 64 * j, kk + h * kBK + boff, &full_bar[stage]);
              } else {
                ptx::tma_load_2d_pair(sb + h * sub_b_bytes, mb, kk + h * kBK, n0 + boff, &full_bar[stage]);
              }
            }
          }
        }
        __syncwarp();
        if (++stage == kStages2) { stage = 0; phase ^= 1u; }
        if (++kbs == P.nk[seg]) { kbs = 0; ++seg; }
      }
    }
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // the wait of the second cluster barrier (see above)
  } else if (warp == 1 && leader) {
    // ================================================================= MMA issuer (leader CTA, M = 256 across the pair)
    const uint64_t a_hi = ptx::smem_desc_sw128(0, 0, 1024);
    const uint64_t b_hi = ptx::smem_desc_sw128(0, L.b_mn ? kBK * 128 : 0, 1024);
    const uint32_t b_step = L.b_mn ? 2048u >> 4 : 32u >> 4;
    int stage = 0;
    uint32_t phase = 0;
    TileWalk walk;
    walk.init(L, cluster_id, n_clusters);
    for (int part = 0;; ++part) {
      int iter, which;
      walk.part(part, iter, which);
      const int t = cluster_id + iter * n_clusters;
      if (t >= total_tiles) break;
      TileInfo ti;
      decode_tile(L, t, ti);
      const int tile_kb0 = ti.kb0, tile_kb1 = ti.kb1;
      if (which == 1) ti.kb1 = wait_kblock(L.prob[ti.p]);
      if (which == 2) ti.kb0 = wait_kblock(L.prob[ti.p]);
      const uint32_t idesc = (1u << 4) | ((uint32_t)L.ab_format << 7) | ((uint32_t)L.ab_format << 10) |
                             ((uint32_t)L.b_mn << 16) | (((uint32_t)L.prob[ti.p].bn >> 3) << 17) | ((256u >> 4) << 24);
      const int acc = iter % n_acc;
      const uint32_t acc_phase = (uint32_t)((iter / n_acc) & 1);
      if (which != 2) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);  // both CTAs' epilogues have drained this accumulator
        ptx::tc_fence_after_sync();
      }
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
      for (int kb = ti.kb0; kb < ti.kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after_sync();
        if (part == 0 && kb == ti.kb0 && lane == 0) timeline_stamp(L.diag, 2);
        const uint64_t a_desc = a_hi | (uint64_t)((ptx::smem_addr(smem_a + stage * kStageA2Bytes) >> 4) & 0x3FFFu);
        const uint64_t b_desc = b_hi | (uint64_t)((ptx::smem_addr(smem_b + stage * stage_b_bytes) >> 4) & 0x3FFFu);
        const uint64_t a_sub = (uint64_t)(kStageABytes >> 4), b_sub = (uint64_t)(sub_b_bytes >> 4);
        if (ptx::elect_one()) {
          if (!(L.diag & 1)) {  // (timing experiment: bit 0 skips the MMAs)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                ptx::umma_f16_pair(d_tmem, a_desc + h * a_sub + (uint64_t)(k * 2), b_desc + h * b_sub + (uint64_t)(k * b_step), idesc,
                                   (kb > tile_kb0 || h > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit_pair(&empty_bar[stage]);                          // the stage is reusable once these MMAs retire
          if (kb == tile_kb1 - 1) ptx::umma_commit_pair(&tmem_full_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == kStages2) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ================================================================= epilogue (each CTA: its 128 rows)
    const int ew = warp - 4;
    int iter = 0;
    for (int t = cluster_id; t < total_tiles; t += n_clusters, ++iter) {
      TileInfo ti;
      decode_tile(L, t, ti);
      const GemmProblem& P = L.prob[ti.p];
      const int acc = iter % n_acc;
      const uint32_t acc_phase = (uint32_t)((iter / n_acc) & 1);
      const long long row = (long long)ti.m_blk * 256 + (long long)rank * kBM + ew * 32 + lane;
      const bool add_bias = P.bias != nullptr;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after_sync();
      if (threadIdx.x == 128) { if (iter == 0) timeline_stamp(L.diag, 3); timeline_stamp(L.diag, 5); }
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * acc_stride);
      epilogue_tile<TD, false, kMode>(P, L.diag, taddr, row, ti.n0, ti.n_end, add_bias, smem_epi + ew * kEpiStageBytesPerWarp);
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (threadIdx.x == 128) { if (iter == 0) timeline_stamp(L.diag, 4); timeline_stamp(L.diag, 6); }
      if (lane == 0) ptx::mbar_arrive_leader(&tmem_empty_bar[acc]);
      if (P.signal) {  // publish these rows to the tiles of this launch that read them (8 warps = one pair tile)
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(L.flags + ti.m_blk, 1);
      }
    }
  }

  ptx::tc_fence_before_sync();
  ptx::cluster_sync_all();  // both CTAs are done with the pair's tensor memory and with each other's shared memory
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
  if (L.flags != nullptr && threadIdx.x == 0) {  // the last CTA to leave zeroes the flags for the next launch
    __threadfence();
    const int done = atomicAdd(L.flags + L.n_flags, 1);
    if (done == (int)gridDim.x - 1) {
      for (int i = 0; i <= L.n_flags; ++i) L.flags[i] = 0;
      __threadfence();
    }
  }
  if (threadIdx.x == 0) timeline_stamp(L.diag, 7);
}

}  // namespace psob200
