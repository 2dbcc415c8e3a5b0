// EXPERIMENT, NOT BUILT INTO THE LIBRARY (kept for the record; see profiles/r01_lora_gemm.md, "A light mma.sync kernel ...").
// Result on B200: correct (all 75 LoRA / GEMM parity tests passed through it) but SLOWER than the tcgen05 kernel it was meant to
// relieve -- 14.2 us vs 12.7 us for t = s x A^T at M = 8192, K = 1280, r = 64, independent of M, pipeline depth (3 / 8 stages) and
// instruction count: the legacy mma.sync path of sm_100 sustains only ~140 TFLOP/s over the whole GPU (1.34 GFLOP in ~10.7 us on
// 128 SMs), so even this "bandwidth-bound" rank-r product is compute-bound on it.
// Rank-r side products of a LoRA-wrapped projection on a LIGHT kernel:  t = s x A^T  and  u = s dy B  (N = r <= 128).
//
// Why not the tcgen05 kernel: these products carry ~5 % of a projection's flops and are bandwidth / latency bound
// (arithmetic intensity <= 84 flop/B), yet on the TMA / TMEM kernel each launch cost 13 us warm and
// 16-18 us inside the training step AT ANY M: one 128-row tile per CTA walks the whole reduction at ~0.3 us per k-block
// (the producer / MMA / commit handshake, measured with no loads and no MMAs) on top of ~7 us of fixed cost (tensor-map
// fetch, TMEM allocation, pipeline fill and drain) -- profiles/r01_lora_gemm.md.  Here: 64-row tiles (twice the CTAs),
// plain cp.async into padded shared-memory rows, ldmatrix + mma.sync.m16n8k16 with fp32 accumulators in registers, no
// barriers other than one __syncthreads per k-block, no tensor maps, no tensor memory: nothing to set up or tear down.
//
//   D[M, N] = alpha * A[M, K] * B^T      A row-major, K contiguous (x or dy)
//     kBT == false:  B given as [N, K] row-major (lora_A [r, K]):  K contiguous
//     kBT == true :  B given as [K, N] row-major (lora_B [N_out, r] consumed reduction-major): N contiguous, ldmatrix.trans
//   outputs: row-major d [M, ldd] and / or the transposed copy dt [N, lddt] (the K-major operand of the dA / dB reductions),
//   both of the operand type.  Requirements (checked by the host, else the tcgen05 path is used): K % 8 == 0, N <= 128,
//   leading dimensions multiples of 8 elements, 16-byte aligned bases.
#pragma once

#include <atomic>

namespace psob200 {

constexpr int kSkBM = 64;        // rows per CTA: 4 row groups of 16
constexpr int kSkBK = 64;        // reduction elements per stage
constexpr int kSkThreads = 256;  // 8 warps: (row group) x (half of the 8-column accumulator tiles): two warps per scheduler
// cp.async stages: the k loop is latency-bound (a CTA's compute per k-block is ~150 cycles, a global -> shared copy ~1 us), so its
// rate is latency / (stages in flight): 3 stages measured 0.58 us per k-block, hence 8 (6 for the widest tile: 221 KB of 227)
template <int kNT> struct SkStages { static constexpr int value = kNT <= 8 ? 8 : 6; };
constexpr int kSkPitchA = kSkBK * 2 + 16;  // bytes per shared-memory row of an [rows, 64 k] tile: 144 = 9 x 16, conflict-free ldmatrix

struct SkinnyParams {
  const void* a;
  const void* b;
  void* d;
  void* dt;
  long long lda, ldb, ldd, lddt;
  long long M;
  int N, K;
  float alpha;
  int pdl;  // bit 0: primary of a programmatic dependent launch (the main pass that consumes d may start its prologue early)
};

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], const void* smem) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* smem) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
template <typename T>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]);
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// kNT = number of 8-column accumulator tiles (N <= 8 kNT).  Shared memory per stage: A 64 x 144 B, then B:
//   kBT == false: [8 kNT n-rows][144 B]          kBT == true: [64 k-rows][16 kNT + 16 B]
template <typename T, int kNT, bool kBT>
__global__ void __launch_bounds__(kSkThreads)
lora_skinny_kernel(const SkinnyParams p) {
  extern __shared__ __align__(16) unsigned char sk_smem[];
  constexpr int kPitchB = kBT ? (kNT * 16 + 16) : kSkPitchA;
  constexpr int kRowsB = kBT ? kSkBK : kNT * 8;
  constexpr int kStageA = kSkBM * kSkPitchA, kStageB = kRowsB * kPitchB, kStage = kStageA + kStageB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long m0 = (long long)blockIdx.x * kSkBM;
  const T* A = reinterpret_cast<const T*>(p.a);
  const T* B = reinterpret_cast<const T*>(p.b);
  constexpr int kSkStages = SkStages<kNT>::value;
  const int nk = (p.K + kSkBK - 1) / kSkBK;
  if (p.pdl & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // Per-thread copy plan, fixed for the whole k loop (the loop is instruction-bound at this occupancy: first version
  // recomputed rows, 64-bit products and bounds for every 16-byte copy of every k-block -- 0.58 us per k-block).
  // A tile: 64 rows x 8 chunks of 16 bytes = 512 copies, two per thread (rows r, r + 32; same chunk).
  constexpr int kCopiesA = kSkBM * 8 / kSkThreads;
  constexpr int kChunksB = kBT ? kSkBK * kNT : kNT * 8 * 8;  // 16-byte chunks of the B tile
  constexpr int kCopiesB = (kChunksB + kSkThreads - 1) / kSkThreads;
  const T* src_a[kCopiesA];
  bool ok_a[kCopiesA];
  int dst_a[kCopiesA];
  const int ch_a = tid & 7;  // k chunk of this thread's A copies
#pragma unroll
  for (int i = 0; i < kCopiesA; ++i) {
    const int r = (tid >> 3) + i * (kSkThreads / 8);
    ok_a[i] = m0 + r < p.M;
    src_a[i] = A + (ok_a[i] ? (m0 + r) * p.lda + ch_a * 8 : 0);
    dst_a[i] = r * kSkPitchA + ch_a * 16;
  }
  const T* src_b[kCopiesB];
  bool ok_b[kCopiesB];
  int dst_b[kCopiesB], kofs_b[kCopiesB];  // kofs_b: offset of the copy inside the k-block (elements), for the K bound
#pragma unroll
  for (int i = 0; i < kCopiesB; ++i) {
    const int c = tid + i * kSkThreads;
    if constexpr (!kBT) {  // [n rows][64 k]: row = n, chunk = 8 k
      const int r = c >> 3, ch = c & 7;
      ok_b[i] = c < kChunksB && r < p.N;
      src_b[i] = B + (ok_b[i] ? (long long)r * p.ldb + ch * 8 : 0);
      dst_b[i] = r * kPitchB + ch * 16;
      kofs_b[i] = ch * 8;
    } else {  // [64 k rows][8 kNT n]: row = k, chunk = 8 n.  A chunk that starts inside N may take in the row's padding
              // (ldb % 8 == 0): those columns are never stored
      const int r = c / kNT, ch = c % kNT;
      ok_b[i] = c < kChunksB && ch * 8 < p.N;
      src_b[i] = B + (ok_b[i] ? (long long)r * p.ldb + ch * 8 : 0);
      dst_b[i] = r * kPitchB + ch * 16;
      kofs_b[i] = r;
    }
  }
  const long long step_b = kBT ? (long long)kSkBK * p.ldb : kSkBK;  // elements between consecutive k-blocks of B

  auto load_stage = [&](int kb, int slot) {
    unsigned char* sa = sk_smem + slot * kStage;
    unsigned char* sb = sa + kStageA;
    const int k0 = kb * kSkBK;
#pragma unroll
    for (int i = 0; i < kCopiesA; ++i) {
      const bool ok = ok_a[i] && k0 + ch_a * 8 < p.K;
      cp_async16_zfill(sa + dst_a[i], ok ? src_a[i] + k0 : A, ok);
    }
#pragma unroll
    for (int i = 0; i < kCopiesB; ++i) {
      if (tid + i * kSkThreads < kChunksB) {
        const bool ok = ok_b[i] && k0 + kofs_b[i] < p.K;
        cp_async16_zfill(sb + dst_b[i], ok ? src_b[i] + (long long)kb * step_b : B, ok);
      }
    }
  };

  // warp (wrow, nhalf): rows 16 wrow .. +16, accumulator tiles j0 .. j0 + kNTW of the kNT
  constexpr int kNTW = (kNT + 1) / 2;
  const int wrow = warp & 3, j0 = (warp >> 2) * kNTW;
  float acc[kNTW][4];
#pragma unroll
  for (int j = 0; j < kNTW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

  // fragments of one 16-deep k step (registers); loaded one step ahead of the MMAs that consume them: with one or two
  // warps per scheduler nothing else hides the ldmatrix -> mma latency (first version: ld, mma, ld, mma ...: 0.9 us per k-block)
  auto load_frags = [&](const unsigned char* sa, const unsigned char* sb, int k16, uint32_t (&af)[4], uint32_t (&bf)[kNTW][2]) {
    // lanes 0-15: rows 0-15 at k 0-7; lanes 16-31: rows 0-15 at k 8-15  ->  a0a1 | a2a3 | a4a5 | a6a7
    ldmatrix_x4(af, sa + (wrow * 16 + (lane & 15)) * kSkPitchA + (k16 * 2 + (lane >> 4)) * 16);
#pragma unroll
    for (int j = 0; j < kNTW; ++j) {
      if (j0 + j < kNT) {  // uniform over the warp
        if constexpr (!kBT)  // rows = n, chunks = k: lanes 0-7 -> k 0-7, lanes 8-15 -> k 8-15
          ldmatrix_x2(bf[j], sb + ((j0 + j) * 8 + (lane & 7)) * kPitchB + (k16 * 2 + ((lane >> 3) & 1)) * 16);
        else                 // rows = k, chunks = n: lanes 0-15 -> the 16 k rows of this step, transposed on load
          ldmatrix_x2_trans(bf[j], sb + (k16 * 16 + (lane & 15)) * kPitchB + (j0 + j) * 16);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < kSkStages - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kb = 0; kb < nk; ++kb) {
    cp_async_wait<kSkStages - 2>();  // stage kb has landed (this thread's copies); the barrier publishes everyone's
    __syncthreads();                 // ... and guarantees the slot refilled below is no longer being read
    if (kb + kSkStages - 1 < nk) load_stage(kb + kSkStages - 1, (kb + kSkStages - 1) % kSkStages);
    cp_async_commit();
    const unsigned char* sa = sk_smem + (kb % kSkStages) * kStage;
    const unsigned char* sb = sa + kStageA;
    uint32_t af[2][4], bf[2][kNTW][2];
    load_frags(sa, sb, 0, af[0], bf[0]);
#pragma unroll
    for (int k16 = 0; k16 < kSkBK / 16; ++k16) {
      if (k16 + 1 < kSkBK / 16) load_frags(sa, sb, k16 + 1, af[(k16 + 1) & 1], bf[(k16 + 1) & 1]);
#pragma unroll
      for (int j = 0; j < kNTW; ++j)
        if (j0 + j < kNT) mma_16816<T>(acc[j], af[k16 & 1], bf[k16 & 1][j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: thread (g = lane / 4, t = lane % 4) holds rows g, g + 8 and columns 2t, 2t + 1 of every 8-column tile
  const int g = lane >> 2, t4 = lane & 3;
  T* D = reinterpret_cast<T*>(p.d);
  T* Dt = reinterpret_cast<T*>(p.dt);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long row = m0 + wrow * 16 + g + 8 * h;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < kNTW; ++j) {
      const int n = (j0 + j) * 8 + t4 * 2;
      const float v0 = p.alpha * acc[j][2 * h], v1 = p.alpha * acc[j][2 * h + 1];
      if (D != nullptr) {
        if (n + 1 < p.N) {  // ldd is even and n is even: 4-byte aligned pair
          T pair[2] = {cvt_out<T>(v0), cvt_out<T>(v1)};
          *reinterpret_cast<uint32_t*>(D + row * p.ldd + n) = *reinterpret_cast<uint32_t*>(pair);
        } else if (n < p.N) {
          D[row * p.ldd + n] = cvt_out<T>(v0);
        }
      }
      if (Dt != nullptr) {
        if (n < p.N) Dt[(long long)n * p.lddt + row] = cvt_out<T>(v0);
        if (n + 1 < p.N) Dt[(long long)(n + 1) * p.lddt + row] = cvt_out<T>(v1);
      }
    }
  }
}

template <typename T, int kNT, bool kBT>
static cudaError_t launch_skinny_one(const SkinnyParams& p, cudaStream_t st) {
  constexpr int kPitchB = kBT ? (kNT * 16 + 16) : kSkPitchA;
  constexpr int kRowsB = kBT ? kSkBK : kNT * 8;
  constexpr int smem = SkStages<kNT>::value * (kSkBM * kSkPitchA + kRowsB * kPitchB);
  static std::atomic<bool> configured{false};
  if (smem > 48 * 1024 && !configured.load(std::memory_order_acquire)) {
    const cudaError_t e = cudaFuncSetAttribute(lora_skinny_kernel<T, kNT, kBT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured.store(true, std::memory_order_release);
  }
  const unsigned grid = (unsigned)((p.M + kSkBM - 1) / kSkBM);
  lora_skinny_kernel<T, kNT, kBT><<<grid, kSkThreads, smem, st>>>(p);
  return cudaSuccess;
}

template <typename T, bool kBT>
static cudaError_t launch_skinny_nt(const SkinnyParams& p, cudaStream_t st) {
  const int nt = (p.N + 7) / 8;
  if (nt <= 1) return launch_skinny_one<T, 1, kBT>(p, st);
  if (nt <= 2) return launch_skinny_one<T, 2, kBT>(p, st);
  if (nt <= 4) return launch_skinny_one<T, 4, kBT>(p, st);
  if (nt <= 8) return launch_skinny_one<T, 8, kBT>(p, st);
  return launch_skinny_one<T, 16, kBT>(p, st);
}

// true when the launch fits the light kernel (everything else stays on the tcgen05 path)
static bool skinny_eligible(const psob200_gemm_args& g) {
  return g.N <= 128 && g.K2 == 0 && g.bias == nullptr && !g.a_reduction_major && !g.accumulate && g.split_k <= 1 &&
         g.tune_bn == 0 && (g.K1 % 8) == 0 && g.d_dtype == g.ab_dtype && (g.ldd % 2) == 0 && !(g.diag & 0x40000) &&
         (g.diag & 0xFFFF) == 0 && g.M < (1LL << 31) * kSkBM && (reinterpret_cast<uintptr_t>(g.d) & 3u) == 0 &&
         (g.pdl & ~1) == 0;  // a launch that must itself WAIT on its predecessor (pdl bits 1-3) keeps the tcgen05 kernel's protocol
}

static int launch_skinny(const psob200_gemm_args& g, void* stream) {
  SkinnyParams p;
  p.a = g.a1; p.b = g.b1; p.d = g.d; p.dt = g.dt;
  p.lda = g.lda1; p.ldb = g.ldb1; p.ldd = g.ldd; p.lddt = g.lddt;
  p.M = g.M; p.N = (int)g.N; p.K = (int)g.K1; p.alpha = g.alpha; p.pdl = g.pdl;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (g.ab_dtype == PSOB200_BF16)
    e = g.b_reduction_major ? launch_skinny_nt<__nv_bfloat16, true>(p, st) : launch_skinny_nt<__nv_bfloat16, false>(p, st);
  else
    e = g.b_reduction_major ? launch_skinny_nt<__half, true>(p, st) : launch_skinny_nt<__half, false>(p, st);
  return consume_launch_error("launch lora_skinny_kernel", e);
}

}  // namespace psob200
