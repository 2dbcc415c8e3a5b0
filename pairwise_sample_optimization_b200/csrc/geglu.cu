// Gated GELU of the transformer feed-forward, forward and backward, one launch each.
//
// diffusers==0.27.0 `GEGLU.forward`: `hidden, gate = self.proj(x).chunk(2, dim=-1); return hidden * F.gelu(gate)`.
// The feed-forward is NOT one of the PSO hot-path rows of SURVEY.md section 8; it sits between two LoRA-wrapped attention
// blocks of every transformer layer, and the stock lowering of that one line (strided gelu + strided mul forward;
// gelu_backward + two strided muls + the concatenation of the two halves backward) was the largest single item of the
// measured training micro-step (22 % of the kernel time at 128x128 latents, profiles/r01_step_breakdown.md).  Pure
// streaming work: forward reads 2 I and writes I elements per row, backward reads 3 I and writes 2 I.
//
// Arithmetic in fp32 (exact erf GELU, torch's default `approximate="none"`), one rounding to the storage type.
#include "common.cuh"

namespace psob200 {

constexpr int kGegluThreads = 256;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// d/dx [x Phi(x)] = Phi(x) + x phi(x)
__device__ __forceinline__ void gelu_erf_grad(float x, float& y, float& dy) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  y = x * cdf;
  dy = fmaf(x, pdf, cdf);
}

template <typename T>
__global__ void __launch_bounds__(kGegluThreads)
geglu_fwd_kernel(const T* __restrict__ proj, T* __restrict__ out, long long M, long long I, long long ld_proj, long long ld_out) {
  const long long col = ((long long)blockIdx.x * kGegluThreads + threadIdx.x) * 8;
  if (col >= I) return;
  for (long long row = blockIdx.y; row < M; row += gridDim.y) {
    const T* p = proj + row * ld_proj + col;
    float h[8], g[8], o[8];
    if (col + 8 <= I) {
      const typename Vec8<T>::Raw rh = Vec8<T>::load_raw(p), rg = Vec8<T>::load_raw(p + I);
      Vec8<T>::decode(rh, h);
      Vec8<T>::decode(rg, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = h[j] * gelu_erf(g[j]);
      Vec8<T>::store(out + row * ld_out + col, o);
    } else {
      for (long long j = col; j < I; ++j)
        Vec8<T>::store1(out + row * ld_out + j, Vec8<T>::load1(proj + row * ld_proj + j) *
                                                    gelu_erf(Vec8<T>::load1(proj + row * ld_proj + I + j)));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kGegluThreads)
geglu_bwd_kernel(const T* __restrict__ proj, const T* __restrict__ dout, T* __restrict__ dproj, long long M, long long I,
                 long long ld_proj, long long ld_dout, long long ld_dproj) {
  const long long col = ((long long)blockIdx.x * kGegluThreads + threadIdx.x) * 8;
  if (col >= I) return;
  for (long long row = blockIdx.y; row < M; row += gridDim.y) {
    const T* p = proj + row * ld_proj + col;
    T* d = dproj + row * ld_dproj + col;
    if (col + 8 <= I) {
      const typename Vec8<T>::Raw rh = Vec8<T>::load_raw(p), rg = Vec8<T>::load_raw(p + I),
                                  rd = Vec8<T>::load_raw(dout + row * ld_dout + col);
      float h[8], g[8], dy[8], dh[8], dg[8];
      Vec8<T>::decode(rh, h);
      Vec8<T>::decode(rg, g);
      Vec8<T>::decode(rd, dy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y, yp;
        gelu_erf_grad(g[j], y, yp);
        dh[j] = dy[j] * y;
        dg[j] = dy[j] * h[j] * yp;
      }
      Vec8<T>::store(d, dh);
      Vec8<T>::store(d + I, dg);
    } else {
      for (long long j = col; j < I; ++j) {
        const float hv = Vec8<T>::load1(proj + row * ld_proj + j), gv = Vec8<T>::load1(proj + row * ld_proj + I + j);
        const float dv = Vec8<T>::load1(dout + row * ld_dout + j);
        float y, yp;
        gelu_erf_grad(gv, y, yp);
        Vec8<T>::store1(dproj + row * ld_dproj + j, dv * y);
        Vec8<T>::store1(dproj + row * ld_dproj + I + j, dv * hv * yp);
      }
    }
  }
}

static int geglu_check(const psob200_geglu_args& a, bool bwd) {
  if (a.M <= 0 || a.I <= 0 || !a.proj) return PSOB200_ERR_INVALID_ARG;
  if (bwd ? (!a.dout || !a.dproj) : !a.out) return PSOB200_ERR_INVALID_ARG;
  if (!valid_dtype(a.dtype)) return PSOB200_ERR_DTYPE;
  const long long vec = a.dtype == PSOB200_F32 ? 4 : 8;  // elements per 16 bytes
  // the vector path needs every row start and the column split on a 16-byte boundary
  if ((a.I % vec) || (a.ld_proj % vec) || a.ld_proj < 2 * a.I) return PSOB200_ERR_SHAPE;
  if (bwd ? ((a.ld_dout % vec) || (a.ld_dproj % vec) || a.ld_dout < a.I || a.ld_dproj < 2 * a.I) : ((a.ld_out % vec) || a.ld_out < a.I))
    return PSOB200_ERR_SHAPE;
  if (!aligned16(a.proj) || (bwd ? (!aligned16(a.dout) || !aligned16(a.dproj)) : !aligned16(a.out))) return PSOB200_ERR_ALIGNMENT;
  return PSOB200_OK;
}

static dim3 geglu_grid(const psob200_geglu_args& a) {
  const long long bx = (a.I + kGegluThreads * 8 - 1) / (kGegluThreads * 8);
  int sms = psob200_device_sm_count();
  if (sms <= 0) sms = 148;
  long long by = ((long long)sms * 8 + bx - 1) / bx;  // 8 resident 256-thread CTAs per SM, one wave
  if (by > a.M) by = a.M;
  if (by > 65535) by = 65535;
  return dim3((unsigned)bx, (unsigned)by);
}

}  // namespace psob200

using namespace psob200;

extern "C" int psob200_geglu_forward(const psob200_geglu_args* args, void* stream) {
  if (args == nullptr) return PSOB200_ERR_INVALID_ARG;
  const psob200_geglu_args& a = *args;
  const int rc = geglu_check(a, false);
  if (rc != PSOB200_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid = geglu_grid(a);
  if (a.dtype == PSOB200_BF16)
    geglu_fwd_kernel<__nv_bfloat16><<<grid, kGegluThreads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(a.proj),
                                                                     reinterpret_cast<__nv_bfloat16*>(a.out), a.M, a.I, a.ld_proj, a.ld_out);
  else if (a.dtype == PSOB200_F16)
    geglu_fwd_kernel<__half><<<grid, kGegluThreads, 0, st>>>(reinterpret_cast<const __half*>(a.proj), reinterpret_cast<__half*>(a.out),
                                                              a.M, a.I, a.ld_proj, a.ld_out);
  else
    geglu_fwd_kernel<float><<<grid, kGegluThreads, 0, st>>>(reinterpret_cast<const float*>(a.proj), reinterpret_cast<float*>(a.out),
                                                             a.M, a.I, a.ld_proj, a.ld_out);
  return consume_launch_error("launch geglu_fwd_kernel", cudaSuccess);
}

extern "C" int psob200_geglu_backward(const psob200_geglu_args* args, void* stream) {
  if (args == nullptr) return PSOB200_ERR_INVALID_ARG;
  const psob200_geglu_args& a = *args;
  const int rc = geglu_check(a, true);
  if (rc != PSOB200_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid = geglu_grid(a);
  if (a.dtype == PSOB200_BF16)
    geglu_bwd_kernel<__nv_bfloat16><<<grid, kGegluThreads, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(a.proj), reinterpret_cast<const __nv_bfloat16*>(a.dout),
        reinterpret_cast<__nv_bfloat16*>(a.dproj), a.M, a.I, a.ld_proj, a.ld_dout, a.ld_dproj);
  else if (a.dtype == PSOB200_F16)
    geglu_bwd_kernel<__half><<<grid, kGegluThreads, 0, st>>>(reinterpret_cast<const __half*>(a.proj),
                                                              reinterpret_cast<const __half*>(a.dout), reinterpret_cast<__half*>(a.dproj),
                                                              a.M, a.I, a.ld_proj, a.ld_dout, a.ld_dproj);
  else
    geglu_bwd_kernel<float><<<grid, kGegluThreads, 0, st>>>(reinterpret_cast<const float*>(a.proj), reinterpret_cast<const float*>(a.dout),
                                                             reinterpret_cast<float*>(a.dproj), a.M, a.I, a.ld_proj, a.ld_dout, a.ld_dproj);
  return consume_launch_error("launch geglu_bwd_kernel", cudaSuccess);
}
