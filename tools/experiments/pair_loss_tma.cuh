// SECOND design of the fused pair-loss kernel (superseded as the fast path by pair_loss_tmem.cuh; NOT BUILT any more: kept as the record of the design measured in profiles/r01_pair_loss.md): PERSISTENT clusters, a continuous TMA ring, and a barrier-free cross-CTA reduction.
// It also defines PairEntry / pair_scalar_function_fast / kTabPairs, which the tensor-memory kernel reuses.
//
// What the profiles said (profiles/r01_pair_loss.md): with one cluster per pair the kernel was bound by
// synchronisation, not memory -- cluster launch/hand-shake, barrier.cluster (which carries a GPU-scope
// MEMBAR that waits for the previous gradient stores), block barriers around the one-thread scalar
// function, fp64 scalar math and dependent table loads on the critical path.  This version removes all of
// them from the per-pair path:
//
//   * the grid is sized to the co-resident clusters and every cluster LOOPS over pairs;
//   * one thread per CTA keeps a shared-memory ring (up to 192 KB) full with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx); the ring runs ahead ACROSS pairs;
//   * every warp pushes its six partial sums straight into the shared memory of all CTAs of the cluster
//     with st.async (mbarrier complete_tx on the destination) -- no barrier.cluster, no __syncthreads;
//   * residuals are double-buffered in registers, so while pair i's partials travel, the CTA already streams
//     pair i+1; when pair i's gather barrier has completed every warp sums the partials itself (fixed order:
//     deterministic, identical in all CTAs), evaluates the pair's scalar function in fp32 from shared memory
//     only, and stores its gradient slice;
//   * per-pair coefficients (fp64, dependent schedule-table loads) are resolved for 64 pairs at a time by the
//     whole CTA in parallel.
//
// Included by pair_loss.cu (needs PairKernelArgs and the shared device helpers defined there).
#pragma once

namespace psob200 {

constexpr int kTmaThreads = 512;  // 16 warps, one 8-element chunk per thread per stage
constexpr int kTmaWarps = kTmaThreads / 32;
constexpr int kTmaStages = 4;     // per pair and CTA: 2 branches x (up to) 2 halves of kTmaThreads chunks
constexpr int kGatherBufs = 4;    // reuse distance that orders a remote write after the local read (see below)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D TMA: global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared::cta address -> the same location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// asynchronous 4-byte store into another CTA's shared memory; counts 4 bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}

template <typename T>
__device__ __forceinline__ void lds8(const unsigned char* slot_sub, int chunk, float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    const float4* p = reinterpret_cast<const float4*>(slot_sub) + chunk * 2;
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 a = reinterpret_cast<const uint4*>(slot_sub)[chunk];
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
      }
    }
  }
}

// A stage is kTmaThreads chunks (one per thread) of the four (three) tensors of one branch.
template <typename TP, typename TL, bool HAS_REF>
struct TmaCfg {
  static constexpr int kNPred = HAS_REF ? 2 : 1;
  static constexpr int kChunkBytes = 8 * (2 * (int)sizeof(TL) + kNPred * (int)sizeof(TP));
  static constexpr int kSlotBytes = kTmaThreads * kChunkBytes;  // bf16/bf16: 32 KB; fp32/fp32: 64 KB
  static constexpr int kDepthRaw = 192 * 1024 / kSlotBytes;     // one CTA per SM: the ring takes 192 KB
  static constexpr int kDepth = kDepthRaw > 6 ? 6 : kDepthRaw;
  static constexpr int kSmemBytes = kDepth * kSlotBytes;
  // sub-buffer offsets inside a slot: x | x' | eps_pol | eps_ref
  static constexpr int kOffX = 0;
  static constexpr int kOffXn = kTmaThreads * 8 * (int)sizeof(TL);
  static constexpr int kOffP = 2 * kTmaThreads * 8 * (int)sizeof(TL);
  static constexpr int kOffR = kOffP + kTmaThreads * 8 * (int)sizeof(TP);
};

// PairEntry / pair_scalar_function_fast / kTabPairs now live in csrc/pair_loss_tmem.cuh (this file is no longer built).

template <typename TP, typename TL, bool HAS_REF>
__global__ void __launch_bounds__(kTmaThreads, 1) pair_loss_grad_tma_kernel(const PairKernelArgs a) {
  using Cfg = TmaCfg<TP, TL, HAS_REF>;
  constexpr int D = Cfg::kDepth;
  constexpr int T = kTmaThreads;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  const long long cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) uint64_t full_bar[6], empty_bar[6], gather_bar[kGatherBufs];
  // partial sums of every warp of every CTA of the cluster, per gather buffer: [buf][cta][warp][6 (+2 pad)]
  __shared__ __align__(16) float s_part[kGatherBufs][kMaxCluster][kTmaWarps][8];
  __shared__ PairEntry s_tab[2][kTabPairs];

  // ---- this CTA's slab of every sample: chunks [cbeg, cend) of each branch, cut into nh halves of <= T chunks
  const int cpc = a.chunks_per_cta;  // <= 2 T
  const long long nchunk = a.N / 8;
  const long long cbeg = (long long)rank * cpc;
  const long long cend = (cbeg + cpc < nchunk) ? cbeg + cpc : nchunk;
  const int mine = cend > cbeg ? (int)(cend - cbeg) : 0;
  const int nh = mine == 0 ? 0 : (mine > T ? 2 : 1);
  const int nstage = 2 * nh;  // non-empty stages per pair; a stage index q maps to (branch, half):
  auto stage_branch = [&](int q) -> int { return nh == 2 ? (q >> 1) : q; };
  auto stage_half = [&](int q) -> int { return nh == 2 ? (q & 1) : 0; };
  auto stage_chunks = [&](int q) -> int {
    const int n = mine - stage_half(q) * T;
    return n > T ? T : n;
  };
  const long long my_pairs = cluster_id < a.B ? (a.B - cluster_id + n_clusters - 1) / n_clusters : 0;
  const long long total_stages = my_pairs * nstage;  // the ring's global stage index g runs over [0, total_stages)

  // thread 0 only: start the bulk copies of global stage g into slot g % D
  auto issue = [&](long long g) {
    const long long it = g >> nh;  // nstage == 1 << nh for nh in {1, 2}
    const int q = (int)(g & (nstage - 1));
    const long long pair = cluster_id + it * n_clusters;
    const int n = stage_chunks(q), k = stage_branch(q);
    const long long first = (cbeg + (long long)stage_half(q) * T) * 8;  // element offset inside the sample
    unsigned char* slot = ring + (size_t)(g % D) * Cfg::kSlotBytes;
    uint64_t* bar = &full_bar[g % D];
    const uint32_t bl = (uint32_t)n * 8u * (uint32_t)sizeof(TL), bp = (uint32_t)n * 8u * (uint32_t)sizeof(TP);
    mbar_expect_tx(bar, 2u * bl + (uint32_t)Cfg::kNPred * bp);
    bulk_g2s(slot + Cfg::kOffX, reinterpret_cast<const TL*>(a.x[k]) + pair * a.stride[2][k] + first, bl, bar);
    bulk_g2s(slot + Cfg::kOffXn, reinterpret_cast<const TL*>(a.xn[k]) + pair * a.stride[3][k] + first, bl, bar);
    bulk_g2s(slot + Cfg::kOffP, reinterpret_cast<const TP*>(a.pred[k]) + pair * a.stride[0][k] + first, bp, bar);
    if constexpr (HAS_REF)
      bulk_g2s(slot + Cfg::kOffR, reinterpret_cast<const TP*>(a.ref[k]) + pair * a.stride[1][k] + first, bp, bar);
  };

  if (tid == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      mbar_init(&full_bar[d], 1);
      mbar_init(&empty_bar[d], kTmaWarps);  // one arrival per warp
    }
#pragma unroll
    for (int b = 0; b < kGatherBufs; ++b) mbar_init(&gather_bar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int g = 0; g < D; ++g)
      if (g < total_stages) issue(g);  // the whole ring is in flight from here on
  }
  // every CTA's barriers are initialised before anyone pushes partial sums into a neighbour (once per kernel)
  cluster.sync();

  const uint32_t gather_tx = C * kTmaWarps * 6u * 4u;  // bytes every CTA receives per pair
  double loss_acc = 0.0;                               // rank 0, thread 0: sum of this cluster's per-pair losses

  // One pipeline step: stream pair `it` into `cur` (if it exists) and push its partial sums to the cluster; then
  // finish pair `it-1`, whose partials have had a whole pass-1 to arrive, from `prev`.
  auto step = [&](float (&cur)[kTmaStages][8], float (&prev)[kTmaStages][8], long long it) {
    if (it < my_pairs) {
      const long long pair = cluster_id + it * n_clusters;
      const int tb = (int)((it / kTabPairs) & 1), slot_t = (int)(it % kTabPairs);
      if (slot_t == 0) {
        // (re)fill one of the two tables of per-pair scalars for the next kTabPairs pairs of this cluster: one
        // (pair, branch) per thread, so fp64 math and dependent global loads of all of them overlap.  The other
        // table still serves the pair being finished.
        __syncthreads();
        if (tid < 2 * kTabPairs && it + (tid >> 1) < my_pairs) {
          const long long p = pair + (long long)(tid >> 1) * n_clusters;
          const int k = tid & 1;
          resolve_pair_coefs(a, p, k, &s_tab[tb][tid >> 1].c[k]);
          s_tab[tb][tid >> 1].h[k] = a.mode == kModeOnline ? a.human_prefer[p * 2 + k] : 0.f;
        }
        __syncthreads();
      }
      const PairEntry& ent = s_tab[tb][slot_t];
      const int buf = (int)(it % kGatherBufs);
      if (tid == 0) mbar_expect_tx(&gather_bar[buf], gather_tx);  // arm this pair's gather (1 arrival + bytes)

      // ---------------------------------------------------------------- pass 1
      float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < kTmaStages; ++q) {
        if (q < nstage) {  // uniform over the CTA
          const long long g = it * nstage + q;
          const int slot_i = (int)(g % D);
          const uint32_t phase = (uint32_t)((g / D) & 1);
          const int n = stage_chunks(q), k = stage_branch(q);
          mbar_wait(&full_bar[slot_i], phase);
          if (tid < n) {
            const unsigned char* slot = ring + (size_t)slot_i * Cfg::kSlotBytes;
            float vx[8], vn[8], vp[8], vr[8];
            lds8<TL>(slot + Cfg::kOffX, tid, vx);
            lds8<TL>(slot + Cfg::kOffXn, tid, vn);
            lds8<TP>(slot + Cfg::kOffP, tid, vp);
            if constexpr (HAS_REF) lds8<TP>(slot + Cfg::kOffR, tid, vr);
            float s_t = 0.f, s_r = 0.f, s_d = 0.f;
            residual8<HAS_REF>(vx, vn, vp, vr, ent.c[k].k, ent.c[k].a, cur[q], s_t, s_r, s_d);
            if (k == 0) { acc[0] += s_t; acc[1] += s_r; acc[2] += s_d; }
            else        { acc[3] += s_t; acc[4] += s_r; acc[5] += s_d; }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[slot_i]);  // this warp has drained the slot
          if (tid == 0 && g + D < total_stages) {          // refill it with the stage D ahead (maybe of a later pair)
            mbar_wait(&empty_bar[slot_i], phase);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before async-proxy writes
            issue(g + D);
          }
        }
      }
      // ---------------------------------------------------------------- warp partials -> every CTA of the cluster
#pragma unroll
      for (int j = 0; j < 6; ++j) acc[j] = warp_sum(acc[j]);  // xor butterflies: every lane holds the warp sums
      const uint32_t my_slot = smem_u32(&s_part[buf][rank][warp][0]);
      const uint32_t my_bar = smem_u32(&gather_bar[buf]);
      for (unsigned idx = lane; idx < C * 6u; idx += 32u) {
        const unsigned r = idx / 6u, j = idx - r * 6u;
        const float v = j == 0 ? acc[0] : j == 1 ? acc[1] : j == 2 ? acc[2] : j == 3 ? acc[3] : j == 4 ? acc[4] : acc[5];
        st_async_f32(map_to_cta(my_slot + j * 4u, r), v, map_to_cta(my_bar, r));
      }
    }
    if (it > 0) {
      // ---------------------------------------------------------------- finish pair it-1
      const long long p = it - 1;
      const long long pair = cluster_id + p * n_clusters;
      const int buf = (int)(p % kGatherBufs);
      const PairEntry& ent = s_tab[(p / kTabPairs) & 1][p % kTabPairs];
      mbar_wait(&gather_bar[buf], (uint32_t)((p / kGatherBufs) & 1));
      // sum the C*16 warp partials: every warp of every CTA uses the same order -> identical S everywhere
      float S[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float4* parts = reinterpret_cast<const float4*>(&s_part[buf][0][0][0]);
      for (unsigned e = lane; e < C * kTmaWarps; e += 32u) {
        const float4 lo4 = parts[2 * e], hi4 = parts[2 * e + 1];
        S[0] += lo4.x; S[1] += lo4.y; S[2] += lo4.z; S[3] += lo4.w; S[4] += hi4.x; S[5] += hi4.y;
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) S[j] = warp_sum(S[j]);
      float g0, g1, st[8];
      const float per = pair_scalar_function_fast(a, S, ent, g0, g1, st);
      // ---------------------------------------------------------------- pass 2: grad = g_k * r from registers
#pragma unroll
      for (int q = 0; q < kTmaStages; ++q) {
        if (q < nstage && tid < stage_chunks(q)) {
          const int k = stage_branch(q);
          const float g = k == 0 ? g0 : g1;
          TP* grad = reinterpret_cast<TP*>(a.grad[k]) + pair * a.N + (cbeg + (long long)stage_half(q) * T + tid) * 8;
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = g * prev[q][i];
          Vec8<TP>::store(grad, o);
        }
      }
      if (rank == 0 && tid == 0) {
        loss_acc += (double)per;
        if (a.stats != nullptr) {
          float4* dst = reinterpret_cast<float4*>(a.stats + pair * 8);
          dst[0] = make_float4(st[0], st[1], st[2], st[3]);
          dst[1] = make_float4(st[4], st[5], st[6], st[7]);
        }
      }
    }
  };

  // Why kGatherBufs = 4: CTA A pushes pair i+4 only after finishing pair i+2, which needs B's partials of pair
  // i+2, which B pushes after it has finished (read) pair i.  So A can never overwrite a buffer B still reads.
  float res_a[kTmaStages][8], res_b[kTmaStages][8];
  for (long long it = 0; it <= my_pairs; it += 2) {
    step(res_a, res_b, it);
    if (it + 1 <= my_pairs) step(res_b, res_a, it + 1);
  }

  // ---- mean over pairs: one ticket per cluster; the last cluster sums the per-cluster partial sums in a fixed order
  if (rank == 0 && warp == 0 && my_pairs > 0) {
    unsigned ticket = 0;
    if (lane == 0) {
      reinterpret_cast<volatile float*>(a.pair_loss)[cluster_id] = (float)loss_acc;
      __threadfence();
      ticket = atomicAdd(a.counter, 1u);
    }
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    const long long active = a.B < n_clusters ? a.B : n_clusters;
    if (ticket == (unsigned)(active - 1)) {
      __threadfence();
      double s = 0.0;
      for (long long i = lane; i < active; i += 32) s += (double)__ldcg(a.pair_loss + i);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) {
        a.loss[0] = (float)((double)a.loss_scale * s / (double)a.B);
        *a.counter = 0u;
      }
    }
  }
  cluster.sync();  // nobody leaves while a neighbour's st.async may still target its shared memory
}

}  // namespace psob200
