// One scheduler update + Gaussian log-prob, and its backward (sm_100a).
//
// Forward: one thread-block cluster per sample streams (eps, x, x' | noise) once with 128-bit
// loads; in sampling mode it also writes x' = mu + s*noise (and, for the turbo sampler, the next
// UNet input x'/sqrt(sigma_next^2+1)) in the same pass; the squared residual is reduced through
// warp shuffles -> shared memory -> distributed shared memory, so log_prob[b] is written without
// atomics or a zero-init launch (deterministic).
//
// Replaces turbo_inference_with_logprob.py:24-116 and distilled_inference_with_logprob.py:45-137
// (both modes), their autograd backward (train_online_pso_sdxl_turbo.py:857), the x0 recovery of the
// last DMD2 sampler step (distilled :36-42 via sdxl_dmd_with_logprob.py:158-162) and the samplers'
// elementwise scalings (sdxl_turbo_with_logprob.py:99,121).
#include <atomic>

#include "common.cuh"

namespace psob200 {

constexpr int kStepMaxThreads = 512;
constexpr int kStepMaxCluster = 8;
constexpr int kBatch = 4;  // chunks requested per thread before the first is consumed
enum { kScore = 0, kSampleOutPred = 1, kSampleOutLatent = 2 };

struct StepKernelArgs {
  const void* eps;
  const void* x;
  const void* xn;     // scoring
  const void* noise;  // sampling
  const void* ts;
  const void* ts_prev;
  const float* coef;
  void* prev_out;
  void* scaled_out;
  float* log_prob;
  int32_t* status;
  psob200_schedule sched;
  long long B, N, noise_rows, ts_rows;
  long long stride_eps, stride_x, stride_xn;
  int chunks_per_cta;
  int use_philox;  // sampling mode without a noise tensor: N(0,1) draws from the counter-based generator below
  unsigned long long philox_seed, philox_offset;
};

// ---- Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) + Box-Muller: the throughput mode of
// the sampler (hard part 8).  Four N(0,1) draws per counter; counter = (q, offset) with q = index of the
// 4-element group inside the noise tensor (per sample, or shared by the batch: DS:123-124), key = seed: the stream does not
// depend on the launch geometry.  oracle/philox.py restates it; torch's own generator stream is NOT reproduced (parity mode
// takes the noise as an input instead).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox_normal4(unsigned long long q, unsigned long long seed, unsigned long long offset, float (&n)[4]) {
  uint32_t x[4];
  philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), x);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)(x[2 * h] >> 8) + 0.5f) * 5.9604644775390625e-8f;      // (0, 1): 24 bits
    const float u2 = ((float)(x[2 * h + 1] >> 8) + 0.5f) * 5.9604644775390625e-8f;
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    n[2 * h] = rad * cs;
    n[2 * h + 1] = rad * sn;
  }
}
template <typename T>
__device__ __forceinline__ float round_through(float v) {  // the reference draws the noise IN the tensor's dtype (TS:97)
  if constexpr (sizeof(T) == 4) return v;
  else return static_cast<float>(static_cast<T>(v));
}

template <typename TP, typename TL, int MODE, int W>
__global__ void __launch_bounds__(kStepMaxThreads) step_logprob_kernel(const StepKernelArgs a) {
  using TO = typename std::conditional<MODE == kSampleOutLatent, TL, TP>::type;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  const long long b = blockIdx.x / C;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;

  __shared__ float s_warp[kStepMaxThreads / 32];
  __shared__ float s_part[kStepMaxCluster];
  __shared__ StepCoef s_coef;

  cluster.barrier_arrive();
  if (tid == 0)
    s_coef = resolve_coef(a.sched, a.ts, a.ts_prev, a.coef, a.ts_rows == 1 ? 0 : b, b, a.B, a.N, a.status);
  __syncthreads();
  const float kx = s_coef.k, ca = s_coef.a, sd = s_coef.s, nscale = s_coef.next_input_scale;

  const long long nchunk = (a.N + W - 1) / W;
  const long long cbeg = (long long)rank * a.chunks_per_cta;
  const long long cend = (cbeg + a.chunks_per_cta < nchunk) ? cbeg + a.chunks_per_cta : nchunk;
  const long long base = b * a.N;
  const TP* eps = reinterpret_cast<const TP*>(a.eps) + b * a.stride_eps;
  const TL* x = reinterpret_cast<const TL*>(a.x) + b * a.stride_x;
  const TL* xn = MODE == kScore ? reinterpret_cast<const TL*>(a.xn) + b * a.stride_xn : nullptr;
  const TO* noise = MODE != kScore ? reinterpret_cast<const TO*>(a.noise) + (a.noise_rows == 1 ? 0 : base) : nullptr;
  TO* prev_out = MODE != kScore ? reinterpret_cast<TO*>(a.prev_out) + base : nullptr;
  TO* scaled_out = (MODE != kScore && a.scaled_out != nullptr) ? reinterpret_cast<TO*>(a.scaled_out) + base : nullptr;

  float acc = 0.f;
  if constexpr (W == 8) {
    // kBatch chunks per thread are requested back to back (raw 128-bit registers) before any is
    // decoded, so every thread keeps 3*kBatch independent 16-byte loads in flight
    using TN = typename std::conditional<MODE == kScore, TL, TO>::type;
    const TN* third = MODE == kScore ? reinterpret_cast<const TN*>(xn) : reinterpret_cast<const TN*>(noise);
    for (long long c0 = cbeg + tid; c0 < cend; c0 += (long long)T * kBatch) {
      typename Vec8<TP>::Raw re[kBatch];
      typename Vec8<TL>::Raw rx[kBatch];
      typename Vec8<TN>::Raw rn[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const long long c = c0 + (long long)u * T;
        if (c < cend) {
          re[u] = Vec8<TP>::load_raw(eps + c * 8);
          rx[u] = Vec8<TL>::load_raw(x + c * 8);
          if (MODE == kScore || !a.use_philox) rn[u] = Vec8<TN>::load_raw(third + c * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const long long c = c0 + (long long)u * T;
        if (c < cend) {
          float ve[8], vx[8], vn[8], o[8], o2[8];
          Vec8<TP>::decode(re[u], ve);
          Vec8<TL>::decode(rx[u], vx);
          if (MODE != kScore && a.use_philox) {
            const unsigned long long q = (unsigned long long)((a.noise_rows == 1 ? 0 : base) / 4 + 2 * c);  // N % 8 == 0 here
            float lo[4], hi[4];
            philox_normal4(q, a.philox_seed, a.philox_offset, lo);
            philox_normal4(q + 1, a.philox_seed, a.philox_offset, hi);
#pragma unroll
            for (int i = 0; i < 4; ++i) { vn[i] = round_through<TO>(lo[i]); vn[4 + i] = round_through<TO>(hi[i]); }
          } else {
            Vec8<TN>::decode(rn[u], vn);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float mu = fmaf(ca, ve[i], kx * vx[i]);
            float r;
            if constexpr (MODE == kScore) {
              r = vn[i] - mu;
            } else {
              const float nx = fmaf(vn[i], sd, mu);  // TS:99 / DS:126
              r = nx - mu;                           // TS:109 uses the fp32 prev_sample
              o[i] = nx;
              o2[i] = nx * nscale;
            }
            acc = fmaf(r, r, acc);
          }
          if constexpr (MODE != kScore) {
            Vec8<TO>::store(prev_out + c * 8, o);
            if (scaled_out != nullptr) Vec8<TO>::store(scaled_out + c * 8, o2);
          }
        }
      }
    }
  } else {
    for (long long c = cbeg + tid; c < cend; c += T) {
      const float mu = fmaf(ca, Vec8<TP>::load1(eps + c), kx * Vec8<TL>::load1(x + c));
      float r;
      if constexpr (MODE == kScore) {
        r = Vec8<TL>::load1(xn + c) - mu;
      } else {
        float nz;
        if (a.use_philox) {
          const unsigned long long e = (unsigned long long)((a.noise_rows == 1 ? 0 : base) + c);
          float n4[4];
          philox_normal4(e / 4, a.philox_seed, a.philox_offset, n4);
          nz = round_through<TO>(n4[e & 3]);
        } else {
          nz = Vec8<TO>::load1(noise + c);
        }
        const float nx = fmaf(nz, sd, mu);
        r = nx - mu;
        Vec8<TO>::store1(prev_out + c, nx);
        if (scaled_out != nullptr) Vec8<TO>::store1(scaled_out + c, nx * nscale);
      }
      acc = fmaf(r, r, acc);
    }
  }

  const float v = warp_sum(acc);
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  cluster.barrier_wait();
  if (tid == 0) {
    float part = 0.f;
    for (int w = 0; w < nwarps; ++w) part += s_warp[w];
    *cluster.map_shared_rank(&s_part[rank], 0) = part;  // gather on the cluster leader
  }
  cluster.sync();
  if (rank == 0 && tid == 0) {
    double S = 0.0;
    for (unsigned r = 0; r < C; ++r) S += (double)s_part[r];
    const double kHalfLog2Pi = 0.91893853320467274178;
    a.log_prob[b] = (float)(-S * (double)s_coef.inv_2s2n - (double)s_coef.log_s - kHalfLog2Pi);  // TS:108-114
  }
}

// ------------------------------------------------------------------------------- backward
struct StepBwdKernelArgs {
  const void* eps;
  const void* x;
  const void* xn;
  const void* ts;
  const void* ts_prev;
  const float* coef;
  const float* grad_lp;
  void* grad_eps;
  int32_t* status;
  psob200_schedule sched;
  long long B, N, ts_rows;
  long long stride_eps, stride_x, stride_xn;
};

template <typename TP, typename TL, int W>
__global__ void __launch_bounds__(256) step_logprob_bwd_kernel(const StepBwdKernelArgs a) {
  const long long b = blockIdx.y;
  __shared__ StepCoef s_coef;
  if (threadIdx.x == 0)
    s_coef = resolve_coef(a.sched, a.ts, a.ts_prev, a.coef, a.ts_rows == 1 ? 0 : b, b, a.B, a.N, a.status);
  __syncthreads();
  const float kx = s_coef.k, ca = s_coef.a;
  const float g = a.grad_lp[b] * s_coef.a_over_s2n;
  const long long base = b * a.N;
  const TP* eps = reinterpret_cast<const TP*>(a.eps) + b * a.stride_eps;
  const TL* x = reinterpret_cast<const TL*>(a.x) + b * a.stride_x;
  const TL* xn = reinterpret_cast<const TL*>(a.xn) + b * a.stride_xn;
  TP* out = reinterpret_cast<TP*>(a.grad_eps) + base;
  const long long nchunk = (a.N + W - 1) / W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if constexpr (W == 8) {
    for (long long c0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; c0 < nchunk; c0 += stride * kBatch) {
      typename Vec8<TP>::Raw re[kBatch];
      typename Vec8<TL>::Raw rx[kBatch], rn[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const long long c = c0 + (long long)u * stride;
        if (c < nchunk) {
          re[u] = Vec8<TP>::load_raw(eps + c * 8);
          rx[u] = Vec8<TL>::load_raw(x + c * 8);
          rn[u] = Vec8<TL>::load_raw(xn + c * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const long long c = c0 + (long long)u * stride;
        if (c < nchunk) {
          float ve[8], vx[8], vn[8], o[8];
          Vec8<TP>::decode(re[u], ve);
          Vec8<TL>::decode(rx[u], vx);
          Vec8<TL>::decode(rn[u], vn);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = g * (vn[i] - fmaf(ca, ve[i], kx * vx[i]));
          Vec8<TP>::store(out + c * 8, o);
        }
      }
    }
  } else {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += stride) {
      const float mu = fmaf(ca, Vec8<TP>::load1(eps + c), kx * Vec8<TL>::load1(x + c));
      Vec8<TP>::store1(out + c, g * (Vec8<TL>::load1(xn + c) - mu));
    }
  }
}

// ------------------------------------------------------------------------------- x0 from noise
template <typename TP, typename TL, typename TO, int W>
__global__ void __launch_bounds__(256) x0_from_noise_kernel(const float* __restrict__ alphas_cumprod, int n_table,
                                                            const void* eps_, const void* x_, const void* ts,
                                                            int ts_dtype, long long ts_rows, void* out_, long long N,
                                                            int32_t* status) {
  const long long b = blockIdx.y;
  long long it = (long long)load_timestep(ts, ts_dtype, ts_rows == 1 ? 0 : b);
  if (it < 0) it += n_table;
  float c_x, c_e;
  if (it < 0 || it >= n_table) {
    if (status != nullptr && threadIdx.x == 0 && blockIdx.x == 0) atomicOr(status, PSOB200_STATUS_TIMESTEP_NOT_IN_SCHEDULE);
    c_x = c_e = __int_as_float(0x7fc00000);
  } else {
    const double a_t = (double)alphas_cumprod[it];  // DS:38-41
    c_x = (float)(1.0 / sqrt(a_t));
    c_e = (float)(-sqrt(1.0 - a_t) / sqrt(a_t));
  }
  const long long base = b * N;
  const TP* eps = reinterpret_cast<const TP*>(eps_) + base;
  const TL* x = reinterpret_cast<const TL*>(x_) + base;
  TO* out = reinterpret_cast<TO*>(out_) + base;
  const long long nchunk = (N + W - 1) / W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += stride) {
    if constexpr (W == 8) {
      float ve[8], vx[8], o[8];
      Vec8<TP>::load(eps + c * 8, ve);
      Vec8<TL>::load(x + c * 8, vx);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(c_e, ve[i], c_x * vx[i]);
      Vec8<TO>::store(out + c * 8, o);
    } else {
      Vec8<TO>::store1(out + c, fmaf(c_e, Vec8<TP>::load1(eps + c), c_x * Vec8<TL>::load1(x + c)));
    }
  }
}

template <typename TI, typename TO, int W>
__global__ void __launch_bounds__(256) scale_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long count,
                                                    float scale) {
  const long long nchunk = (count + W - 1) / W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += stride) {
    if constexpr (W == 8) {
      float v[8];
      Vec8<TI>::load(in + c * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= scale;
      Vec8<TO>::store(out + c * 8, v);
    } else {
      Vec8<TO>::store1(out + c, Vec8<TI>::load1(in + c) * scale);
    }
  }
}

template <typename T, int W>
__global__ void __launch_bounds__(256) scale_inplace_dev_kernel(T* __restrict__ data, long long count,
                                                                const float* __restrict__ scale_dev) {
  const float scale = *scale_dev;
  if (scale == 1.0f) return;  // the common case: loss.backward() seeds d loss = 1
  const long long nchunk = (count + W - 1) / W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += stride) {
    if constexpr (W == 8) {
      float v[8];
      Vec8<T>::load(data + c * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= scale;
      Vec8<T>::store(data + c * 8, v);
    } else {
      Vec8<T>::store1(data + c, Vec8<T>::load1(data + c) * scale);
    }
  }
}

static inline int check_launch() { return consume_launch_error("elementwise kernel launch", cudaSuccess); }

static inline unsigned elementwise_blocks(long long nchunk, long long rows) {
  long long per_row = (nchunk + 255) / 256;
  const long long want = (148LL * 8 + rows - 1) / rows;  // ~8 CTAs of 256 threads per SM over all rows
  if (per_row > want) per_row = want;
  if (per_row < 1) per_row = 1;
  return (unsigned)per_row;
}

static int validate_sched(const psob200_schedule* s, const void* ts, const void* ts_prev, const float* coef) {
  if (s == nullptr) return PSOB200_ERR_INVALID_ARG;
  if (s->kind == PSOB200_SCHED_AFFINE) return coef ? PSOB200_OK : PSOB200_ERR_INVALID_ARG;
  if (s->kind != PSOB200_SCHED_TURBO && s->kind != PSOB200_SCHED_DMD) return PSOB200_ERR_INVALID_ARG;
  if (!ts || !s->table || s->n_table <= 0) return PSOB200_ERR_INVALID_ARG;
  if (s->kind == PSOB200_SCHED_TURBO && !s->sched_timesteps) return PSOB200_ERR_INVALID_ARG;
  if (s->kind == PSOB200_SCHED_DMD && !ts_prev) return PSOB200_ERR_INVALID_ARG;
  if (s->ts_dtype != PSOB200_TS_I64 && s->ts_dtype != PSOB200_TS_F32 && s->ts_dtype != PSOB200_TS_I32)
    return PSOB200_ERR_DTYPE;
  return PSOB200_OK;
}

template <typename TP, typename TL, int MODE>
static int launch_step(const StepKernelArgs& ka, bool vec_ok, int threads, int cluster, cudaStream_t stream) {
  cudaError_t e;
  if (vec_ok)
    e = launch_cluster(step_logprob_kernel<TP, TL, MODE, 8>, dim3((unsigned)(ka.B * cluster)), dim3(threads), 0,
                       stream, (unsigned)cluster, ka);
  else
    e = launch_cluster(step_logprob_kernel<TP, TL, MODE, 1>, dim3((unsigned)(ka.B * cluster)), dim3(threads), 0,
                       stream, (unsigned)cluster, ka);
  return consume_launch_error("launch step_logprob_kernel", e);
}

}  // namespace psob200

using namespace psob200;

extern "C" int psob200_step_logprob(const psob200_schedule* sched, const psob200_step_args* args, void* stream) {
  if (args == nullptr) return PSOB200_ERR_INVALID_ARG;
  const psob200_step_args& p = *args;
  int rc = validate_sched(sched, p.ts, p.ts_prev, p.coef);
  if (rc != PSOB200_OK) return rc;
  if (p.B <= 0 || p.N <= 0 || !p.model_output || !p.sample || !p.log_prob) return PSOB200_ERR_INVALID_ARG;
  if (p.ts_rows != 1 && p.ts_rows != p.B) return PSOB200_ERR_INVALID_ARG;
  if (!valid_dtype(p.pred_dtype) || !valid_dtype(p.latent_dtype)) return PSOB200_ERR_DTYPE;
  const bool scoring = p.prev_sample != nullptr;
  const bool philox = p.use_philox != 0;
  if (scoring && (p.noise != nullptr || philox)) return PSOB200_ERR_INVALID_ARG;  // exactly one of the two (DS:115-119)
  if (!scoring && ((p.noise != nullptr) == philox)) return PSOB200_ERR_INVALID_ARG;
  int mode = kScore;
  bool vec_ok = (p.N % 8) == 0 && aligned16(p.model_output) && aligned16(p.sample);
  if (scoring) {
    vec_ok = vec_ok && aligned16(p.prev_sample);
  } else {
    if (!p.prev_out) return PSOB200_ERR_INVALID_ARG;
    if (p.noise_rows != 1 && p.noise_rows != p.B) return PSOB200_ERR_INVALID_ARG;
    if (p.out_dtype == p.pred_dtype) mode = kSampleOutPred;
    else if (p.out_dtype == p.latent_dtype) mode = kSampleOutLatent;
    else return PSOB200_ERR_DTYPE;
    if (philox && p.noise_rows == p.B && (p.N % 4) != 0) return PSOB200_ERR_SHAPE;  // a 4-draw group never straddles two samples
    vec_ok = vec_ok && (philox || aligned16(p.noise)) && aligned16(p.prev_out) && (!p.scaled_next_out || aligned16(p.scaled_next_out));
  }
  if (p.stride_model_output < 0 || p.stride_sample < 0 || p.stride_prev_sample < 0) return PSOB200_ERR_INVALID_ARG;
  const long long st_e = p.stride_model_output ? p.stride_model_output : p.N;
  const long long st_x = p.stride_sample ? p.stride_sample : p.N;
  const long long st_n = p.stride_prev_sample ? p.stride_prev_sample : p.N;
  vec_ok = vec_ok && (st_e % 8) == 0 && (st_x % 8) == 0 && (st_n % 8) == 0;
  const int W = vec_ok ? 8 : 1;
  const long long nchunk = (p.N + W - 1) / W;
  int threads = p.tune_threads > 0 ? p.tune_threads : 256;
  if (threads > kStepMaxThreads || threads < 32 || (threads & 31)) return PSOB200_ERR_INVALID_ARG;
  int cluster = p.tune_cluster;
  if (cluster <= 0) {
    // Clusters are costly on B200 (measured: 4.9 TB/s at cluster size 1, 2.5 TB/s at 8 for this kernel), so a
    // sample is split over a cluster only while there are fewer samples than SMs.
    cluster = 1;
    while (cluster < kStepMaxCluster && p.B * cluster < 148 && nchunk / (cluster * 2) >= threads) cluster <<= 1;
  }
  if (cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8) return PSOB200_ERR_INVALID_ARG;
  StepKernelArgs ka = {};
  ka.eps = p.model_output; ka.x = p.sample; ka.xn = p.prev_sample; ka.noise = p.noise;
  ka.ts = p.ts; ka.ts_prev = p.ts_prev; ka.coef = p.coef;
  ka.prev_out = p.prev_out; ka.scaled_out = p.scaled_next_out; ka.log_prob = p.log_prob; ka.status = p.status;
  ka.sched = *sched;
  ka.B = p.B; ka.N = p.N; ka.noise_rows = p.noise_rows; ka.ts_rows = p.ts_rows;
  ka.stride_eps = st_e; ka.stride_x = st_x; ka.stride_xn = st_n;
  ka.chunks_per_cta = (int)((nchunk + cluster - 1) / cluster);
  ka.use_philox = philox ? 1 : 0; ka.philox_seed = p.philox_seed; ka.philox_offset = p.philox_offset;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dispatch2(p.pred_dtype, p.latent_dtype, [&](auto tp, auto tl) -> int {
    using TP = decltype(tp);
    using TL = decltype(tl);
    if (mode == kScore) return launch_step<TP, TL, kScore>(ka, vec_ok, threads, cluster, st);
    if (mode == kSampleOutPred) return launch_step<TP, TL, kSampleOutPred>(ka, vec_ok, threads, cluster, st);
    return launch_step<TP, TL, kSampleOutLatent>(ka, vec_ok, threads, cluster, st);
  });
}

extern "C" int psob200_step_logprob_backward(const psob200_schedule* sched, const psob200_step_bwd_args* args,
                                             void* stream) {
  if (args == nullptr) return PSOB200_ERR_INVALID_ARG;
  const psob200_step_bwd_args& p = *args;
  int rc = validate_sched(sched, p.ts, p.ts_prev, p.coef);
  if (rc != PSOB200_OK) return rc;
  if (p.B <= 0 || p.N <= 0 || !p.model_output || !p.sample || !p.prev_sample || !p.grad_log_prob ||
      !p.grad_model_output)
    return PSOB200_ERR_INVALID_ARG;
  if (p.ts_rows != 1 && p.ts_rows != p.B) return PSOB200_ERR_INVALID_ARG;
  if (p.B > 65535) return PSOB200_ERR_SHAPE;
  if (!valid_dtype(p.pred_dtype) || !valid_dtype(p.latent_dtype)) return PSOB200_ERR_DTYPE;
  if (p.stride_model_output < 0 || p.stride_sample < 0 || p.stride_prev_sample < 0) return PSOB200_ERR_INVALID_ARG;
  const long long st_e = p.stride_model_output ? p.stride_model_output : p.N;
  const long long st_x = p.stride_sample ? p.stride_sample : p.N;
  const long long st_n = p.stride_prev_sample ? p.stride_prev_sample : p.N;
  const bool vec_ok = (p.N % 8) == 0 && aligned16(p.model_output) && aligned16(p.sample) &&
                      aligned16(p.prev_sample) && aligned16(p.grad_model_output) && (st_e % 8) == 0 &&
                      (st_x % 8) == 0 && (st_n % 8) == 0;
  StepBwdKernelArgs ka = {};
  ka.eps = p.model_output; ka.x = p.sample; ka.xn = p.prev_sample; ka.ts = p.ts; ka.ts_prev = p.ts_prev;
  ka.coef = p.coef; ka.grad_lp = p.grad_log_prob; ka.grad_eps = p.grad_model_output; ka.status = p.status;
  ka.sched = *sched;
  ka.B = p.B; ka.N = p.N; ka.ts_rows = p.ts_rows;
  ka.stride_eps = st_e; ka.stride_x = st_x; ka.stride_xn = st_n;
  const long long nchunk = vec_ok ? p.N / 8 : p.N;
  const dim3 grid(elementwise_blocks(nchunk, p.B), (unsigned)p.B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dispatch2(p.pred_dtype, p.latent_dtype, [&](auto tp, auto tl) -> int {
    using TP = decltype(tp);
    using TL = decltype(tl);
    if (vec_ok) step_logprob_bwd_kernel<TP, TL, 8><<<grid, 256, 0, st>>>(ka);
    else step_logprob_bwd_kernel<TP, TL, 1><<<grid, 256, 0, st>>>(ka);
    return check_launch();
  });
}

extern "C" int psob200_dmd_x0_from_noise(const float* alphas_cumprod, int32_t n_table, const void* model_output,
                                         const void* sample, const void* ts, int32_t ts_dtype, int64_t ts_rows,
                                         void* x0_out, int64_t B, int64_t N, int32_t pred_dtype, int32_t latent_dtype,
                                         int32_t out_dtype, int32_t* status, void* stream) {
  if (!alphas_cumprod || n_table <= 0 || !model_output || !sample || !ts || !x0_out || B <= 0 || N <= 0)
    return PSOB200_ERR_INVALID_ARG;
  if (ts_rows != 1 && ts_rows != B) return PSOB200_ERR_INVALID_ARG;
  if (B > 65535) return PSOB200_ERR_SHAPE;
  if (!valid_dtype(pred_dtype) || !valid_dtype(latent_dtype)) return PSOB200_ERR_DTYPE;
  // the inputs' types, or fp32 (the reference's promotion with the fp32 alphas_cumprod gather, DS:36-42)
  if (out_dtype != pred_dtype && out_dtype != latent_dtype && out_dtype != PSOB200_F32) return PSOB200_ERR_DTYPE;
  const bool vec_ok = (N % 8) == 0 && aligned16(model_output) && aligned16(sample) && aligned16(x0_out);
  const long long nchunk = vec_ok ? N / 8 : N;
  const dim3 grid(elementwise_blocks(nchunk, B), (unsigned)B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool out_is_latent = out_dtype == latent_dtype;
  return dispatch2(pred_dtype, latent_dtype, [&](auto tp, auto tl) -> int {
    using TP = decltype(tp);
    using TL = decltype(tl);
#define PSOB200_X0(TO)                                                                                              \
  do {                                                                                                              \
    if (vec_ok)                                                                                                     \
      x0_from_noise_kernel<TP, TL, TO, 8><<<grid, 256, 0, st>>>(alphas_cumprod, n_table, model_output, sample, ts,  \
                                                                ts_dtype, ts_rows, x0_out, N, status);              \
    else                                                                                                            \
      x0_from_noise_kernel<TP, TL, TO, 1><<<grid, 256, 0, st>>>(alphas_cumprod, n_table, model_output, sample, ts,  \
                                                                ts_dtype, ts_rows, x0_out, N, status);              \
  } while (0)
    if (out_is_latent) PSOB200_X0(TL); else if (out_dtype == pred_dtype) PSOB200_X0(TP); else PSOB200_X0(float);
#undef PSOB200_X0
    return check_launch();
  });
}

extern "C" int psob200_scale(const void* in, void* out, int64_t count, float scale, int32_t in_dtype,
                             int32_t out_dtype, void* stream) {
  if (!in || !out || count <= 0) return PSOB200_ERR_INVALID_ARG;
  if (!valid_dtype(in_dtype) || !valid_dtype(out_dtype)) return PSOB200_ERR_DTYPE;
  const bool vec_ok = (count % 8) == 0 && aligned16(in) && aligned16(out);
  const long long nchunk = vec_ok ? count / 8 : count;
  const unsigned blocks = elementwise_blocks(nchunk, 1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dispatch2(in_dtype, out_dtype, [&](auto ti, auto to) -> int {
    using TI = decltype(ti);
    using TO = decltype(to);
    if (vec_ok) scale_kernel<TI, TO, 8><<<blocks, 256, 0, st>>>(reinterpret_cast<const TI*>(in), reinterpret_cast<TO*>(out), count, scale);
    else scale_kernel<TI, TO, 1><<<blocks, 256, 0, st>>>(reinterpret_cast<const TI*>(in), reinterpret_cast<TO*>(out), count, scale);
    return check_launch();
  });
}

extern "C" int psob200_scale_inplace_by_device_scalar(void* data, int64_t count, int32_t dtype, const float* scale_dev,
                                                      void* stream) {
  if (!data || !scale_dev || count <= 0) return PSOB200_ERR_INVALID_ARG;
  if (!valid_dtype(dtype)) return PSOB200_ERR_DTYPE;
  const bool vec_ok = (count % 8) == 0 && aligned16(data);
  const long long nchunk = vec_ok ? count / 8 : count;
  const unsigned blocks = elementwise_blocks(nchunk, 1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dispatch2(dtype, dtype, [&](auto t, auto) -> int {
    using T = decltype(t);
    if (vec_ok) scale_inplace_dev_kernel<T, 8><<<blocks, 256, 0, st>>>(reinterpret_cast<T*>(data), count, scale_dev);
    else scale_inplace_dev_kernel<T, 1><<<blocks, 256, 0, st>>>(reinterpret_cast<T*>(data), count, scale_dev);
    return check_launch();
  });
}
