// LoRA projection GEMMs on the 5th-generation tensor cores (sm_100a): TMA-fed tcgen05.mma with the
// accumulators in tensor memory.
//
// One persistent, warp-specialised kernel computes
//
//     D[M,N] = alpha * ( A1[M,K1] * B1[N,K1]^T  +  A2[M,K2] * B2[N,K2]^T ) + bias[N]
//
// The second K segment is what makes the LoRA-wrapped projection ONE tensor-core pass over the frozen
// weight instead of the reference stack's three GEMMs + scale + add (peft lora.Linear.forward, reached from
// train_online_pso_sdxl_turbo.py:338-345):   y = x W^T + b + (s x A^T) B^T   is   [x | T] [W | B]^T  with
// T = s x A^T  a skinny first pass of the same kernel.  The backward uses the same two shapes
// (dX = dY W + U A with U = s dY B) plus a split-M reduction with MN-major A operand for dA / dB.
//
// Warp roles (256 threads, 1 CTA per SM, grid = min(tiles, SMs), static round-robin tile schedule):
//   warp 0      TMA producer: 128B-swizzled [128 x 64] A and [bn x 64] B boxes into a 4-stage ring (192 KB)
//   warp 1      MMA issuer: one thread, 4 x tcgen05.mma (128 x bn x 16) per stage, tcgen05.commit frees the stage
//   warp 2      tensor-memory allocator (512 columns = two accumulator buffers of <= 256 columns)
//   warps 4-7   epilogue: tcgen05.ld 32x32b -> alpha, bias, convert -> global (also a transposed copy, or
//               fp32 atomic accumulation for the split reductions); overlaps the next tile's MMAs.
#include <atomic>
#include <mutex>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace psob200 {

constexpr int kBM = 128;            // UMMA M: rows of an output tile
constexpr int kBNMax = 256;         // widest UMMA N
constexpr int kBK = 64;             // reduction elements per stage = one 128-byte swizzle span of a 16-bit type
constexpr int kStages = 4;          // with the widest tile; narrower tiles get more (up to kMaxStages)
constexpr int kMaxStages = 8;
constexpr int kGemmThreads = 256;
constexpr int kStageABytes = kBM * kBK * 2;      // 16 KB
constexpr int kStageBBytes = kBNMax * kBK * 2;   // 32 KB
constexpr int kGemmSmemBytes = kStages * (kStageABytes + kStageBBytes) + 1024;  // + slack for 1024 B alignment
constexpr int kTmemCols = 512;

struct GemmKernelParams {
  long long M, N;         // output extent
  int nk1, nk2;           // 64-wide k-blocks of the two segments
  int bn;                 // tile width, multiple of 16, <= 256
  int m_tiles, n_tiles, splits, kb_per_split;
  void* d; long long ldd;     // row-major output (may be null)
  void* dt; long long lddt;   // transposed output [N, M] (may be null)
  const void* bias;
  float alpha;
  int d_dtype, bias_dtype, ab_format, a_mn, b_mn, atomic, diag;
  int stages;             // even; stage = 16 KB of A + bn*128 B of B
  int pdl;                // programmatic dependent launch role bits (psob200_gemm_args.pdl)
};

template <typename T>
__device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }

// bias[n0 .. n0+32) as fp32, zero beyond N.  One uniform branch per chunk, 128-bit loads for whole chunks.
template <typename TB>
__device__ __forceinline__ void load_bias32(const void* bias, long long n0, long long N, float (&b)[32]) {
  const TB* src = reinterpret_cast<const TB*>(bias) + n0;
  if (n0 + 32 <= N && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float t[8];
      if constexpr (sizeof(TB) == 4) {
        const float4 lo = __ldg(reinterpret_cast<const float4*>(src + j)), hi = __ldg(reinterpret_cast<const float4*>(src + j + 4));
        t[0] = lo.x; t[1] = lo.y; t[2] = lo.z; t[3] = lo.w; t[4] = hi.x; t[5] = hi.y; t[6] = hi.z; t[7] = hi.w;
      } else {
        const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(src + j));
        typename Vec8<TB>::Raw raw{w4};
        Vec8<TB>::decode(raw, t);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) b[j + i] = t[i];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = (n0 + j < N) ? Vec8<TB>::load1(src + j) : 0.f;
  }
}

// Epilogue mode (compile time): bits 0-1 = outputs (1: row-major d only, 2: transposed dt only, 3: whichever pointers are
// set, checked at run time), bit 2 = a bias may be present.  The hot launches of the training step get exact modes so that
// their instruction footprint stays small (see lora_gemm_kernel); mode 7 is the general case.
constexpr int kEpiD = 1, kEpiDt = 2, kEpiAnyOut = 3, kEpiBias = 4, kEpiGeneral = 7;
constexpr int kEpiBiasSameType = 8 | kEpiBias;  // the bias has the output's element type (nn.Linear in one dtype)

// One 32-column chunk of one accumulator row: v[j] belongs to (row, n0 + j).
template <typename TD, bool kAtomic, int kMode>
__device__ __forceinline__ void store_chunk(const GemmKernelParams& p, const float (&v)[32], long long row, long long n0,
                                            long long n_end) {
  if (row >= p.M) return;
  const long long nleft = n_end - n0;  // columns of this chunk that belong to this tile and exist
  if ((kMode & kEpiD) && ((kMode & 3) != kEpiAnyOut || p.d != nullptr)) {
    TD* dst = reinterpret_cast<TD*>(p.d) + row * p.ldd + n0;
    if constexpr (kAtomic) {
      float* acc = reinterpret_cast<float*>(dst);
      if (nleft >= 32 && (reinterpret_cast<uintptr_t>(acc) & 15u) == 0) {
        // a thread owns 32 consecutive floats of its row: 8 vector reductions instead of 32 scalar ones (each lane
        // of a scalar atomic touches a different row = a different sector: a quarter of the L2 operations)
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(acc + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]),
                       "f"(v[j + 3])
                       : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nleft) atomicAdd(acc + j, v[j]);
      }
    } else if (nleft >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
      if constexpr (sizeof(TD) == 4) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float o[8] = {v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]};
          Vec8<TD>::store(dst + j, o);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nleft) dst[j] = cvt_out<TD>(v[j]);
    }
  }
  if ((kMode & kEpiDt) && ((kMode & 3) != kEpiAnyOut || p.dt != nullptr)) {  // lanes hold consecutive rows: coalesced columns
    TD* dst = reinterpret_cast<TD*>(p.dt) + n0 * p.lddt + row;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < nleft) {
        if constexpr (kAtomic) atomicAdd(reinterpret_cast<float*>(dst) + (long long)j * p.lddt, v[j]);
        else dst[(long long)j * p.lddt] = cvt_out<TD>(v[j]);
      }
    }
  }
}

// Drain one accumulator tile: TMEM -> registers (the load of chunk c+1 is in flight while chunk c is converted and
// stored) -> alpha, bias -> global.
template <typename TD, bool kAtomic, int kMode>
__device__ __forceinline__ void epilogue_tile(const GemmKernelParams& p, uint32_t taddr, long long row, long long n_tile0,
                                              bool add_bias) {
  const int chunks = (p.bn + 31) / 32;
  const long long n_end = n_tile0 + p.bn < p.N ? n_tile0 + p.bn : p.N;  // a chunk must not spill into the next tile
  uint32_t raw[2][32];
  ptx::tmem_ld_32x32(taddr, raw[0]);
#pragma unroll 1
  for (int c = 0; c < chunks; c += 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cc = c + h;
      const long long n0 = n_tile0 + (long long)cc * 32;
      if (cc < chunks && n0 < p.N) {  // uniform over the warp
        ptx::tmem_ld_wait();
        if (cc + 1 < chunks) ptx::tmem_ld_32x32(taddr + (uint32_t)(cc + 1) * 32, raw[h ^ 1]);
        float v[32];
        if ((kMode & kEpiBias) && add_bias) {
          float b[32];
          if constexpr ((kMode & 8) != 0) load_bias32<TD>(p.bias, n0, p.N, b);
          else if (p.bias_dtype == PSOB200_F32) load_bias32<float>(p.bias, n0, p.N, b);
          else if (p.bias_dtype == PSOB200_BF16) load_bias32<__nv_bfloat16>(p.bias, n0, p.N, b);
          else load_bias32<__half>(p.bias, n0, p.N, b);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(p.alpha, __uint_as_float(raw[h][j]), b[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = p.alpha * __uint_as_float(raw[h][j]);
        }
        if (!(p.diag & 2)) store_chunk<TD, kAtomic, kMode>(p, v, row, n0, n_end);
      }
    }
  }
  ptx::tmem_ld_wait();
}

// One instantiation per output type / accumulation mode: the four epilogues together made this kernel 259 KB of SASS, and
// a launch that finds its code evicted from the instruction caches (any launch inside the training step: LayerNorm, SDPA,
// cuBLAS kernels run in between) paid ~10 us for it (tools/diag_cold.py: 15.7 us back to back, 26.6 us with three other
// kernels in between, only 18.4 us with the OPERANDS evicted instead).
template <typename TD, bool kAtomic, int kMode>
__global__ void __launch_bounds__(kGemmThreads, 1)
lora_gemm_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1,
                 const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2,
                 const GemmKernelParams p) {
  extern __shared__ unsigned char gemm_smem_raw[];
  // one "full" barrier per stage, one "empty" barrier per PAIR of stages: tcgen05.commit costs several hundred
  // cycles of tensor-pipe command time, so the MMA warp commits every second k-block only
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages / 2], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;
  const int n_stages = p.stages;
  const int stage_b_bytes = p.bn * kBK * 2;
  unsigned char* smem_b = smem + n_stages * kStageABytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = p.nk1 + p.nk2;
  const long long total_tiles = (long long)p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a1);
    ptx::prefetch_tensormap(&map_b1);
    if (p.nk2 > 0) {
      ptx::prefetch_tensormap(&map_a2);
      ptx::prefetch_tensormap(&map_b2);
    }
  }
  if (warp == 1 && lane == 0) {
#pragma unroll
    for (int s = 0; s < kMaxStages; ++s) ptx::mbar_init(&full_bar[s], 1);  // producer's arrive.expect_tx; TMA completes the bytes
#pragma unroll
    for (int s = 0; s < kMaxStages / 2; ++s) ptx::mbar_init(&empty_bar[s], 1);  // tcgen05.commit
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);   // tcgen05.commit after the tile's last MMA
      ptx::mbar_init(&tmem_empty_bar[a], 4);  // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  // primary of a programmatic dependent launch: the dependent grid (the main GEMM that needs this launch's output only
  // for its last k-blocks) may start filling the SMs this small grid leaves idle
  if (p.pdl & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 2) ptx::tmem_alloc<kTmemCols>(&tmem_base_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;

  auto tile_coords = [&](long long t, int& m_blk, int& n_blk, int& kb0, int& kb1) {
    n_blk = (int)(t % p.n_tiles);
    t /= p.n_tiles;
    m_blk = (int)(t % p.m_tiles);
    const int split = (int)(t / p.m_tiles);
    kb0 = split * p.kb_per_split;
    kb1 = kb0 + p.kb_per_split < nk ? kb0 + p.kb_per_split : nk;
  };

  // The producer and MMA loops run WARP-UNIFORMLY (all 32 lanes keep the loop state) and only the issuing
  // instructions are predicated on elect.sync: UTMALDG / UTCHMMA take uniform-register operands, and a loop that
  // only lane 0 executes makes ptxas wrap every one of them in an ELECT / R2UR.BROADCAST waterfall (measured:
  // ~450 cycles per k-block, more than the MMAs themselves).
  if (warp == 0) {
    // ================================================================= TMA producer
    const uint32_t tx = ((p.diag & 8) ? 0u : (uint32_t)kStageABytes) + ((p.diag & 4) ? 0u : (uint32_t)p.bn * kBK * 2u);
    int stage = 0;
    uint32_t phase = 0;
    bool dep_pending = (p.pdl & 2) != 0;  // launched ahead of the grid that produces a2 (or, bit 2, any operand)
    if (dep_pending && (p.pdl & 4)) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
      dep_pending = false;
    }
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int m_blk, n_blk, kb0, kb1;
      tile_coords(t, m_blk, n_blk, kb0, kb1);
      const int m0 = m_blk * kBM, n0 = n_blk * p.bn;
      for (int kb = kb0; kb < kb1; ++kb) {
        if ((stage & 1) == 0) ptx::mbar_wait(&empty_bar[stage >> 1], phase ^ 1u);  // the pair (stage, stage+1) is free
        const bool seg2 = kb >= p.nk1;
        if (seg2 && dep_pending) {  // first read of the previous launch's output: wait for that grid, then fence the
          asm volatile("griddepcontrol.wait;" ::: "memory");  // generic-proxy writes against the TMA (async proxy) reads
          asm volatile("fence.proxy.async;" ::: "memory");
          dep_pending = false;
        }
        const CUtensorMap* ma = seg2 ? &map_a2 : &map_a1;
        const CUtensorMap* mb = seg2 ? &map_b2 : &map_b1;
        const int kk = (seg2 ? kb - p.nk1 : kb) * kBK;
        unsigned char* sa = smem_a + stage * kStageABytes;
        unsigned char* sb = smem_b + stage * stage_b_bytes;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full_bar[stage], tx);
          if (p.diag & 8) {
          } else if (p.a_mn) {  // A given reduction-major: two [64 k x 64 m] boxes, m contiguous
            ptx::tma_load_2d(sa, ma, m0, kk, &full_bar[stage]);
            ptx::tma_load_2d(sa + kStageABytes / 2, ma, m0 + 64, kk, &full_bar[stage]);
          } else {
            ptx::tma_load_2d(sa, ma, kk, m0, &full_bar[stage]);
          }
          if (p.diag & 4) {
          } else if (p.b_mn) {  // B given reduction-major: bn/64 boxes of [64 k x 64 n], n contiguous
            for (int j = 0; j < p.bn / 64; ++j)
              ptx::tma_load_2d(sb + j * (kBK * 128), mb, n0 + 64 * j, kk, &full_bar[stage]);
          } else {
            ptx::tma_load_2d(sb, mb, kk, n0, &full_bar[stage]);
          }
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    const uint32_t idesc = ptx::umma_idesc_f16((uint32_t)p.ab_format, (uint32_t)p.a_mn, (uint32_t)p.b_mn, (uint32_t)p.bn);
    // per-k-step advance of the descriptor start address (>> 4) and the constant upper halves:
    //   K-major, 128B swizzle: 8-row atoms 1024 B apart (SBO); a 16-element k step is 32 B inside the swizzle span.
    //   reduction-major ("MN-major"): 64(mn) x 8(k) atoms, 1024 B between k atoms (SBO), 64 k-rows * 128 B between
    //   mn atoms (LBO); a 16-row k step is 2048 B.
    const uint64_t a_hi = ptx::smem_desc_sw128(0, p.a_mn ? kBK * 128 : 0, 1024);
    const uint64_t b_hi = ptx::smem_desc_sw128(0, p.b_mn ? kBK * 128 : 0, 1024);
    const uint32_t a_step = p.a_mn ? 2048u >> 4 : 32u >> 4, b_step = p.b_mn ? 2048u >> 4 : 32u >> 4;
    int stage = 0;
    uint32_t phase = 0;
    long long iter = 0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x, ++iter) {
      int m_blk, n_blk, kb0, kb1;
      tile_coords(t, m_blk, n_blk, kb0, kb1);
      const int acc = (int)(iter & 1);
      const uint32_t acc_phase = (uint32_t)((iter >> 1) & 1);
      ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);  // epilogue has drained this accumulator
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * kBNMax;
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after_sync();
        const uint64_t a_desc = a_hi | (uint64_t)((ptx::smem_addr(smem_a + stage * kStageABytes) >> 4) & 0x3FFFu);
        const uint64_t b_desc = b_hi | (uint64_t)((ptx::smem_addr(smem_b + stage * stage_b_bytes) >> 4) & 0x3FFFu);
        if (ptx::elect_one()) {
          if (!(p.diag & 1)) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              ptx::umma_f16(d_tmem, a_desc + (uint64_t)(k * a_step), b_desc + (uint64_t)(k * b_step), idesc,
                            (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (stage & 1) ptx::umma_commit(&empty_bar[stage >> 1]);   // the pair is reusable once these MMAs retire
          if (kb == kb1 - 1) ptx::umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ================================================================= epilogue (TMEM lanes 32*(warp%4) ..)
    const int ew = warp - 4;
    long long iter = 0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x, ++iter) {
      int m_blk, n_blk, kb0, kb1;
      tile_coords(t, m_blk, n_blk, kb0, kb1);
      const int acc = (int)(iter & 1);
      const uint32_t acc_phase = (uint32_t)((iter >> 1) & 1);
      const long long row = (long long)m_blk * kBM + ew * 32 + lane;
      const long long n_tile0 = (long long)n_blk * p.bn;
      const bool add_bias = p.bias != nullptr && kb0 == 0;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)acc * kBNMax;
      epilogue_tile<TD, kAtomic, kMode>(p, taddr, row, n_tile0, add_bias);
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
  // an independent dependent launch (bit 3): it never reads the previous grid's output, but it must not be seen as
  // complete before that grid is, so that later launches on the stream stay ordered after both
  if ((p.pdl & 8) && threadIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace psob200

#include "lora_gemm2.cuh"  // cta_group::2 variant (CTA pairs) for the large tensor-bound shapes

namespace psob200 {

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

// Tensor map of a row-major [rows, cols] 16-bit matrix with leading dimension ld (elements); box = 64 contiguous
// elements x box_rows rows, 128-byte swizzle, out-of-bounds elements read as zero.
static int make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                    int ab_dtype) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return PSOB200_ERR_DRIVER;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2u};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, ab_dtype == PSOB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                        const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PSOB200_OK : PSOB200_ERR_DRIVER;
}

static int gemm_sm_count() {
  static PerDevice<int> per_device;
  std::atomic<int>& cached = per_device.here();
  int v = cached.load(std::memory_order_relaxed);
  if (v > 0) return v;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return 148;
  }
  cached.store(n, std::memory_order_relaxed);
  return n;
}

// per-tile fixed cost of the pair kernel in units of one accumulator column: measured per-wave times at K = 1344 are
// 6.7 us (bn 128) and 9.7 us (bn 256) = 3.7 us + 23.4 ns per column (tools/diag_lora_parts.py --bn ...)
constexpr int kTileFixedCols = 160;

static int choose_bn(long long N, bool b_mn) {
  if (b_mn && N < 256) return (int)(((N + 63) / 64) * 64);  // reduction-major B: whole 64-column boxes
  if (N >= 256) {
    // widest tile that wastes the fewest columns: 256, or 128-wide tiles when that divides N better
    const long long w256 = ((N + 255) / 256) * 256 - N, w128 = ((N + 127) / 128) * 128 - N;
    return w128 < w256 ? 128 : 256;
  }
  return (int)(((N + 15) / 16) * 16);
}

// Tile width of the CTA-pair kernel.  The kernel is persistent over `pairs` CTA pairs with a static tile order, so its time
// is (number of waves) x (time of one tile) and a tile costs ~ its width plus a fixed part (pipeline fill, drain of the last
// accumulator): pick the width that minimises waves x (bn + fixed) instead of always 256.  E.g. M = 8192, N = 1280 on 74
// pairs: 256 -> 160 tiles = 3 waves, 224 -> 192 tiles = 3 narrower waves (measured 29.0 -> 26.4 us).  Widths that are not
// multiples of 32 are excluded (bn = 144 measured 30 % slower than the model predicts).
static int choose_bn2(long long M, long long N, bool b_mn, int pairs) {
  const int unit = b_mn ? 128 : 32;
  if (N < 256) return (int)(((N + (b_mn ? 127 : 31)) / (b_mn ? 128 : 32)) * (b_mn ? 128 : 32));
  const long long m_tiles = (M + 255) / 256;
  int best = 256;
  long long best_cost = -1;
  for (int bn = 256; bn >= 128; bn -= unit) {
    const long long tiles = m_tiles * ((N + bn - 1) / bn);
    const long long waves = (tiles + pairs - 1) / pairs;
    const long long cost = waves * (bn + kTileFixedCols);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace psob200

using namespace psob200;

extern "C" int psob200_lora_gemm(const psob200_gemm_args* args, void* stream) {
  if (args == nullptr) return PSOB200_ERR_INVALID_ARG;
  const psob200_gemm_args& g = *args;
  if (g.M <= 0 || g.N <= 0 || g.K1 <= 0 || g.K2 < 0 || !g.a1 || !g.b1 || (!g.d && !g.dt)) return PSOB200_ERR_INVALID_ARG;
  if (g.K2 > 0 && (!g.a2 || !g.b2)) return PSOB200_ERR_INVALID_ARG;
  if (g.ab_dtype != PSOB200_BF16 && g.ab_dtype != PSOB200_F16) return PSOB200_ERR_DTYPE;
  if (!valid_dtype(g.d_dtype) || (g.bias && !valid_dtype(g.bias_dtype))) return PSOB200_ERR_DTYPE;
  if (g.accumulate && g.d_dtype != PSOB200_F32) return PSOB200_ERR_DTYPE;
  if (g.a_reduction_major && g.K2 > 0) return PSOB200_ERR_INVALID_ARG;
  if (g.split_k < 0 || (g.split_k > 1 && !g.accumulate)) return PSOB200_ERR_INVALID_ARG;
  const void* ptrs[4] = {g.a1, g.b1, g.a2, g.b2};
  const long long lds[4] = {g.lda1, g.ldb1, g.lda2, g.ldb2};
  for (int i = 0; i < (g.K2 > 0 ? 4 : 2); ++i) {
    if (!aligned16(ptrs[i])) return PSOB200_ERR_ALIGNMENT;
    if (lds[i] <= 0 || (lds[i] % 8) != 0) return PSOB200_ERR_SHAPE;  // TMA: row pitch is a multiple of 16 bytes
  }
  if ((g.d && g.ldd < g.N) || (g.dt && g.lddt < g.M)) return PSOB200_ERR_SHAPE;
  if (g.M > 0x7fffffffLL - 256 || g.N > 0x7fffffffLL - 256 || g.K1 > 0x7fffffffLL - 256 || g.K2 > 0x7fffffffLL - 256)
    return PSOB200_ERR_SHAPE;

  // ---- CTA-pair kernel (cta_group::2) for the large K-major problems: at least one full wave of 256-row pair tiles
  {
    const int sms = gemm_sm_count();
    const int bn_unit = g.b_reduction_major ? 128 : 16;  // each CTA loads half the tile: whole 64-column boxes / 8-row atoms
    const bool eligible = !g.a_reduction_major && !g.accumulate && g.dt == nullptr && g.split_k <= 1 &&
                          (g.tune_bn == 0 || (g.tune_bn % bn_unit == 0 && g.tune_bn >= 32));
    const int bn2 = g.tune_bn > 0 ? g.tune_bn : choose_bn2(g.M, g.N, g.b_reduction_major != 0, sms / 2);
    const long long pair_tiles = ((g.M + 255) / 256) * ((g.N + bn2 - 1) / bn2);
    const bool want = (g.diag & 0x10000) || (pair_tiles >= sms / 2 && g.N >= 128 && !(g.diag & 0x20000));
    if (eligible && want && bn2 <= kBNMax) {
      GemmKernelParams p = {};
      p.M = g.M; p.N = g.N;
      p.nk1 = (int)((g.K1 + kBK - 1) / kBK);
      p.nk2 = (int)((g.K2 + kBK - 1) / kBK);
      p.bn = bn2;
      p.m_tiles = (int)((g.M + 255) / 256);
      p.n_tiles = (int)((g.N + bn2 - 1) / bn2);
      p.splits = 1; p.kb_per_split = p.nk1 + p.nk2;
      p.d = g.d; p.ldd = g.ldd; p.bias = g.bias;
      p.alpha = g.alpha; p.d_dtype = g.d_dtype; p.bias_dtype = g.bias_dtype;
      p.ab_format = g.ab_dtype == PSOB200_BF16 ? 1 : 0;
      p.b_mn = g.b_reduction_major ? 1 : 0;
      p.diag = g.diag; p.pdl = g.pdl; p.stages = kStages2;
      CUtensorMap ma1, mb1, ma2, mb2;
      int rc;
      if ((rc = make_map(&ma1, g.a1, g.M, g.K1, g.lda1, kBM, g.ab_dtype)) != PSOB200_OK) return rc;
      if (p.b_mn) rc = make_map(&mb1, g.b1, g.K1, g.N, g.ldb1, kBK, g.ab_dtype);
      else rc = make_map(&mb1, g.b1, g.N, g.K1, g.ldb1, bn2 / 2, g.ab_dtype);
      if (rc != PSOB200_OK) return rc;
      if (g.K2 > 0) {
        if ((rc = make_map(&ma2, g.a2, g.M, g.K2, g.lda2, kBM, g.ab_dtype)) != PSOB200_OK) return rc;
        if (p.b_mn) rc = make_map(&mb2, g.b2, g.K2, g.N, g.ldb2, kBK, g.ab_dtype);
        else rc = make_map(&mb2, g.b2, g.N, g.K2, g.ldb2, bn2 / 2, g.ab_dtype);
        if (rc != PSOB200_OK) return rc;
      } else {
        ma2 = ma1;
        mb2 = mb1;
      }
      typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, GemmKernelParams);
      // 0: general (fp32 output, or a bias of another type); then per 16-bit type: bias of the same type, no bias
      const bool bias_other = p.bias != nullptr && p.bias_dtype != p.d_dtype;
      const int variant = (p.d_dtype == PSOB200_F32 || bias_other) ? (p.d_dtype == PSOB200_F32 ? 0 : (p.d_dtype == PSOB200_BF16 ? 1 : 2))
                                                                   : ((p.d_dtype == PSOB200_BF16 ? 3 : 5) + (p.bias != nullptr ? 0 : 1));
      static const KernelFn kernels2[7] = {
          lora_gemm2_kernel<float, kEpiD | kEpiBias>, lora_gemm2_kernel<__nv_bfloat16, kEpiD | kEpiBias>,
          lora_gemm2_kernel<__half, kEpiD | kEpiBias>,
          lora_gemm2_kernel<__nv_bfloat16, kEpiD | kEpiBiasSameType>, lora_gemm2_kernel<__nv_bfloat16, kEpiD>,
          lora_gemm2_kernel<__half, kEpiD | kEpiBiasSameType>, lora_gemm2_kernel<__half, kEpiD>};
      static PerDevice<bool> configured2_dev[7];
      std::atomic<bool>& conf2 = configured2_dev[variant].here();
      if (!conf2.load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(kernels2[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, kGemm2SmemBytes);
        if (e != cudaSuccess) return consume_launch_error("configure lora_gemm2_kernel", e);
        conf2.store(true, std::memory_order_release);
      }
      const long long max_pairs = sms / 2;
      const unsigned grid = 2u * (unsigned)(pair_tiles < max_pairs ? pair_tiles : max_pairs);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(kGemmThreads);
      cfg.dynamicSmemBytes = kGemm2SmemBytes;
      cfg.stream = reinterpret_cast<cudaStream_t>(stream);
      cudaLaunchAttribute attr[1];
      if (g.pdl & (2 | 8)) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
      }
      const cudaError_t e = cudaLaunchKernelEx(&cfg, kernels2[variant], ma1, mb1, ma2, mb2, p);
      return consume_launch_error("launch lora_gemm2_kernel", e);
    }
  }

  GemmKernelParams p = {};
  p.M = g.M; p.N = g.N;
  p.nk1 = (int)((g.K1 + kBK - 1) / kBK);
  p.nk2 = (int)((g.K2 + kBK - 1) / kBK);
  p.bn = g.tune_bn > 0 ? g.tune_bn : choose_bn(g.N, g.b_reduction_major != 0);
  if (p.bn < 16 || p.bn > kBNMax || (p.bn % (g.b_reduction_major ? 64 : 16)) != 0) return PSOB200_ERR_INVALID_ARG;
  p.m_tiles = (int)((g.M + kBM - 1) / kBM);
  p.n_tiles = (int)((g.N + p.bn - 1) / p.bn);
  const int nk = p.nk1 + p.nk2;
  const int sms = gemm_sm_count();
  int splits = g.split_k;
  if (splits == 0) {  // heuristics: split the reduction only for accumulating launches that cannot fill the GPU
    splits = 1;
    if (g.accumulate) {
      // one wave: the largest split count whose tiles all fit on the SMs at once (rounding UP gave e.g. 150 tiles on 148
      // SMs: two CTAs ran two tiles each and the launch took twice as long; measured 12.2 -> 10.1 us for dA at M = 8192)
      const long long tiles = (long long)p.m_tiles * p.n_tiles;
      splits = (int)(sms / tiles);
      if (splits > nk) splits = nk;
      if (splits < 1) splits = 1;
    }
  }
  if (splits > nk) splits = nk;
  p.kb_per_split = (nk + splits - 1) / splits;
  p.splits = (nk + p.kb_per_split - 1) / p.kb_per_split;
  p.d = g.d; p.ldd = g.ldd; p.dt = g.dt; p.lddt = g.lddt; p.bias = g.bias;
  p.alpha = g.alpha; p.d_dtype = g.d_dtype; p.bias_dtype = g.bias_dtype;
  p.ab_format = g.ab_dtype == PSOB200_BF16 ? 1 : 0;
  p.a_mn = g.a_reduction_major ? 1 : 0;
  p.b_mn = g.b_reduction_major ? 1 : 0;
  p.atomic = g.accumulate ? 1 : 0;
  p.diag = g.diag;
  p.pdl = g.pdl;
  p.stages = (kStages * (kStageABytes + kStageBBytes)) / (kStageABytes + p.bn * kBK * 2);
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  p.stages &= ~1;
  if ((g.diag >> 8) & 15) p.stages = (g.diag >> 8) & 14;

  CUtensorMap ma1, mb1, ma2, mb2;
  int rc;
  if (p.a_mn) rc = make_map(&ma1, g.a1, g.K1, g.M, g.lda1, kBK, g.ab_dtype);  // [K, M] row-major: box 64(m) x 64(k)
  else rc = make_map(&ma1, g.a1, g.M, g.K1, g.lda1, kBM, g.ab_dtype);
  if (rc != PSOB200_OK) return rc;
  if (p.b_mn) rc = make_map(&mb1, g.b1, g.K1, g.N, g.ldb1, kBK, g.ab_dtype);  // [K, N] row-major
  else rc = make_map(&mb1, g.b1, g.N, g.K1, g.ldb1, p.bn, g.ab_dtype);
  if (rc != PSOB200_OK) return rc;
  if (g.K2 > 0) {
    if ((rc = make_map(&ma2, g.a2, g.M, g.K2, g.lda2, kBM, g.ab_dtype)) != PSOB200_OK) return rc;
    if (p.b_mn) rc = make_map(&mb2, g.b2, g.K2, g.N, g.ldb2, kBK, g.ab_dtype);
    else rc = make_map(&mb2, g.b2, g.N, g.K2, g.ldb2, p.bn, g.ab_dtype);
    if (rc != PSOB200_OK) return rc;
  } else {
    ma2 = ma1;
    mb2 = mb1;
  }

  typedef void (*KernelFn)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, GemmKernelParams);
  // exact epilogue modes for the launches of the training step, the general kernel for everything else
  const int outs = (p.d != nullptr ? kEpiD : 0) | (p.dt != nullptr ? kEpiDt : 0);
  const bool has_bias = p.bias != nullptr;
  int variant;
  if (p.atomic) variant = has_bias ? 0 : (outs == kEpiD ? 1 : (outs == kEpiDt ? 2 : 0));
  else if (p.d_dtype == PSOB200_F32) variant = 3;
  else {
    const int base = p.d_dtype == PSOB200_BF16 ? 4 : 8;
    variant = base + (has_bias ? (outs == kEpiD && p.bias_dtype == p.d_dtype ? 1 : 0) : (outs == kEpiD ? 2 : 3));
  }
  static const KernelFn kernels[12] = {
      lora_gemm_kernel<float, true, kEpiGeneral>, lora_gemm_kernel<float, true, kEpiD>, lora_gemm_kernel<float, true, kEpiDt>,
      lora_gemm_kernel<float, false, kEpiGeneral>,
      lora_gemm_kernel<__nv_bfloat16, false, kEpiGeneral>, lora_gemm_kernel<__nv_bfloat16, false, kEpiD | kEpiBiasSameType>,
      lora_gemm_kernel<__nv_bfloat16, false, kEpiD>, lora_gemm_kernel<__nv_bfloat16, false, kEpiAnyOut>,
      lora_gemm_kernel<__half, false, kEpiGeneral>, lora_gemm_kernel<__half, false, kEpiD | kEpiBiasSameType>,
      lora_gemm_kernel<__half, false, kEpiD>, lora_gemm_kernel<__half, false, kEpiAnyOut>};
  static PerDevice<bool> configured_dev[12];
  std::atomic<bool>& conf1 = configured_dev[variant].here();
  if (!conf1.load(std::memory_order_acquire)) {
    const cudaError_t e = cudaFuncSetAttribute(kernels[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
    if (e != cudaSuccess) return consume_launch_error("configure lora_gemm_kernel", e);
    conf1.store(true, std::memory_order_release);
  }
  const long long total = (long long)p.m_tiles * p.n_tiles * p.splits;
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = kGemmSmemBytes;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  if (g.pdl & (2 | 8)) {  // may begin while the previous kernel on the stream is still running (it waits on the device)
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernels[variant], ma1, mb1, ma2, mb2, p);
  return consume_launch_error("launch lora_gemm_kernel", e);
}

// ---------------------------------------------------------------------------------------------- LoRA-wrapped Linear
static psob200_gemm_args gemm_defaults(int32_t dtype) {
  psob200_gemm_args g = {};
  g.alpha = 1.0f;
  g.ab_dtype = dtype;
  g.d_dtype = dtype;
  return g;
}

extern "C" int psob200_lora_linear_forward(const psob200_lora_linear_args* args, void* stream) {
  if (args == nullptr) return PSOB200_ERR_INVALID_ARG;
  const psob200_lora_linear_args& a = *args;
  if (!a.x || !a.w || !a.y || a.M <= 0 || a.K <= 0 || a.N <= 0) return PSOB200_ERR_INVALID_ARG;
  const bool lora = a.adapters_enabled != 0;
  if (lora && (!a.lora_a || !a.lora_b || !a.t || a.r <= 0 || a.r > kBNMax)) return PSOB200_ERR_INVALID_ARG;
  int rc;
  const int fph = a.forward_phases == 0 ? (PSOB200_FWD_DOWN | PSOB200_FWD_MAIN) : a.forward_phases;
  if (lora && (fph & PSOB200_FWD_DOWN)) {  // t = scaling * x A^T  (+ its transpose, the K-major operand of the dB reduction)
    psob200_gemm_args g = gemm_defaults(a.dtype);
    g.a1 = a.x; g.lda1 = a.ldx; g.b1 = a.lora_a; g.ldb1 = a.lda;
    g.M = a.M; g.N = a.r; g.K1 = a.K;
    g.alpha = a.scaling;
    g.d = a.t; g.ldd = a.ldt; g.dt = a.tt; g.lddt = a.ldtt;
    g.pdl = 1;  // the main pass below reads t only in its last k-blocks: let it start on the SMs this launch leaves idle
    if ((rc = psob200_lora_gemm(&g, stream)) != PSOB200_OK) return rc;
  }
  if (!(fph & PSOB200_FWD_MAIN)) return PSOB200_OK;
  psob200_gemm_args g = gemm_defaults(a.dtype);
  g.a1 = a.x; g.lda1 = a.ldx; g.b1 = a.w; g.ldb1 = a.ldw;
  g.M = a.M; g.N = a.N; g.K1 = a.K;
  if (lora) { g.a2 = a.t; g.lda2 = a.ldt; g.b2 = a.lora_b; g.ldb2 = a.ldb; g.K2 = a.r; g.pdl = 2; }
  g.bias = a.bias; g.bias_dtype = a.bias_dtype;
  g.d = a.y; g.ldd = a.ldy;
  return psob200_lora_gemm(&g, stream);
}

extern "C" int psob200_lora_linear_backward(const psob200_lora_linear_args* args, void* stream) {
  if (args == nullptr) return PSOB200_ERR_INVALID_ARG;
  const psob200_lora_linear_args& a = *args;
  if (!a.dy || !a.w || a.M <= 0 || a.K <= 0 || a.N <= 0) return PSOB200_ERR_INVALID_ARG;
  const bool lora = a.adapters_enabled != 0;
  if (lora && (!a.lora_a || !a.lora_b || !a.u || a.r <= 0 || a.r > kBNMax)) return PSOB200_ERR_INVALID_ARG;
  if (lora && a.d_lora_a && (!a.ut || !a.x)) return PSOB200_ERR_INVALID_ARG;
  if (lora && a.d_lora_b && !a.tt) return PSOB200_ERR_INVALID_ARG;
  int rc;
  const int ph = a.backward_phases == 0 ? (PSOB200_BWD_INPUT_GRAD | PSOB200_BWD_WEIGHT_GRAD) : a.backward_phases;
  const bool need_u = lora && (ph & PSOB200_BWD_U) && (a.dx != nullptr || a.ut != nullptr || a.d_lora_a != nullptr);
  if (need_u) {  // u = scaling * dy B   (B [N,r] consumed reduction-major: no transposed copy)
    psob200_gemm_args g = gemm_defaults(a.dtype);
    g.a1 = a.dy; g.lda1 = a.lddy; g.b1 = a.lora_b; g.ldb1 = a.ldb; g.b_reduction_major = 1;
    g.M = a.M; g.N = a.r; g.K1 = a.N;
    g.alpha = a.scaling;
    g.d = a.u; g.ldd = a.ldu; g.dt = a.ut; g.lddt = a.ldut;
    g.pdl = a.dx != nullptr ? 1 : 0;
    if ((rc = psob200_lora_gemm(&g, stream)) != PSOB200_OK) return rc;
  }
  if (a.dx != nullptr && (ph & PSOB200_BWD_DX)) {  // dx = dy W + u A   (W [N,K], A [r,K] reduction-major)
    psob200_gemm_args g = gemm_defaults(a.dtype);
    g.a1 = a.dy; g.lda1 = a.lddy; g.b1 = a.w; g.ldb1 = a.ldw; g.b_reduction_major = 1;
    g.M = a.M; g.N = a.K; g.K1 = a.N;
    if (lora) { g.a2 = a.u; g.lda2 = a.ldu; g.b2 = a.lora_a; g.ldb2 = a.lda; g.K2 = a.r; g.pdl = 2; }
    g.d = a.dx; g.ldd = a.lddx;
    if ((rc = psob200_lora_gemm(&g, stream)) != PSOB200_OK) return rc;
  }
  if (lora && (ph & PSOB200_BWD_DA) && a.d_lora_a != nullptr) {  // dA[r,K] += u^T x : D[K,r] = sum_m x[m,:]^T ut[:,m], written transposed
    psob200_gemm_args g = gemm_defaults(a.dtype);
    g.a1 = a.x; g.lda1 = a.ldx; g.a_reduction_major = 1; g.b1 = a.ut; g.ldb1 = a.ldut;
    g.M = a.K; g.N = a.r; g.K1 = a.M;
    g.dt = a.d_lora_a; g.lddt = a.ld_da; g.d_dtype = PSOB200_F32; g.accumulate = 1;
    g.pdl = a.d_lora_b != nullptr ? 1 : 0;  // dB below is independent of this launch: let the two overlap
    if ((rc = psob200_lora_gemm(&g, stream)) != PSOB200_OK) return rc;
  }
  if (lora && (ph & PSOB200_BWD_DB) && a.d_lora_b != nullptr) {  // dB[N,r] += dy^T t
    psob200_gemm_args g = gemm_defaults(a.dtype);
    g.a1 = a.dy; g.lda1 = a.lddy; g.a_reduction_major = 1; g.b1 = a.tt; g.ldb1 = a.ldtt;
    g.M = a.N; g.N = a.r; g.K1 = a.M;
    g.d = a.d_lora_b; g.ldd = a.ld_db; g.d_dtype = PSOB200_F32; g.accumulate = 1;
    g.pdl = a.d_lora_a != nullptr ? 8 : 0;
    if ((rc = psob200_lora_gemm(&g, stream)) != PSOB200_OK) return rc;
  }
  return PSOB200_OK;
}
