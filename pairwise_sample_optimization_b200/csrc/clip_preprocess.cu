// Reward-side image preprocessing on the device: decoded image -> uint8 -> Pillow-exact bicubic resize -> centre crop ->
// rescale + normalise -> channels-first tensor for the CLIP / PickScore image tower, in ONE launch.
//
// Replaces train_online_pso_sdxl_turbo.py:632-640 (GPU -> uint8 -> host -> numpy -> PIL) + the CLIPImageProcessor call inside
// pso_pytorch/pickscore_utils.py:24-33 (PIL BICUBIC resize, centre crop, * 1/255, normalise, back to the GPU).
//
// The resize is a restatement of Pillow's libImaging/Resample.c for 8-bit channels (a third-party dependency of the reference's
// processor; restated from its published algorithm, pinned by fixtures generated with the Pillow installed in this image):
//   * per axis: scale = in / out, support = 2 * max(scale, 1), taps centred on (j + 0.5) * scale, bicubic a = -0.5, weights
//     normalised to sum 1 in double, then rounded to 22 fractional bits  (host: psob200_resample_plan);
//   * horizontal pass first, result rounded and clamped to uint8 ((acc + 2^21) >> 22), then the vertical pass on those bytes.
// Byte work: B * (3 H W read + 3 h w * sizeof(out) written) of HBM traffic; one CTA per (image, block of output rows) stages the
// input rows it needs in shared memory as bytes (each read from HBM once per CTA, 128-bit loads), runs the horizontal pass
// shared -> shared and the vertical pass shared -> global, so the intermediate image never touches HBM.  The arithmetic (11 + 11
// integer taps per output at 512 -> 224) runs on the CUDA cores by design: this is not a GEMM.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace psob200 {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kPreThreads = 256;

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// A source sample becomes the byte PIL would see: uint8 NHWC as is; float NCHW quantised like
// ((x + 1.0) * 127.5).clamp(0, 255).to(torch.uint8) evaluated in the tensor's own type (truncation toward zero) -- quantize8.
template <typename TD>
__device__ __forceinline__ TD cvt_px(float v);
template <> __device__ __forceinline__ float cvt_px<float>(float v) { return v; }
template <> __device__ __forceinline__ __half cvt_px<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_px<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

struct PreKernelArgs {
  const void* src;
  void* dst;
  const int32_t *bounds_h, *coeffs_h, *bounds_v, *coeffs_v;
  const float* table;
  long long in_h, in_w, out_h, out_w, crop_top, crop_left;
  int taps_h, taps_v, rows_per_cta, max_in_rows;
  int x_lo, x_w;  // input columns [x_lo, x_lo + x_w) are all the horizontal taps of the crop window touch (x_w a multiple of 8)
};

// 8 consecutive samples of one input row as the bytes PIL would see (see SrcPixel), packed little-endian into two words.
template <typename TS>
__device__ __forceinline__ uint2 quantize8(const TS* p, long long valid) {
  uint32_t w[2] = {0u, 0u};
  if constexpr (sizeof(TS) == 2) {
    if (valid >= 8 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
      const TS* e = reinterpret_cast<const TS*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const TS v = __hmul(__hadd(e[j], static_cast<TS>(1.0f)), static_cast<TS>(127.5f));
        const float f = fminf(fmaxf(static_cast<float>(v), 0.f), 255.f);
        w[j >> 2] |= (uint32_t)(int)f << (8 * (j & 3));
      }
      return make_uint2(w[0], w[1]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < valid) {
      int q;
      if constexpr (sizeof(TS) == 4) {
        float v = __fmul_rn(__fadd_rn(p[j], 1.0f), 127.5f);
        q = (int)fminf(fmaxf(v, 0.f), 255.f);
      } else {
        const TS v = __hmul(__hadd(p[j], static_cast<TS>(1.0f)), static_cast<TS>(127.5f));
        q = (int)fminf(fmaxf(static_cast<float>(v), 0.f), 255.f);
      }
      w[j >> 2] |= (uint32_t)q << (8 * (j & 3));
    }
  }
  return make_uint2(w[0], w[1]);
}

// grid = (ceil(out_h / rows_per_cta), B).  Shared memory (planar, so that every pass reads and writes consecutive bytes):
//   src_s [3][max_in_rows][x_w]   the input rows this CTA needs, quantised to bytes, read from HBM exactly once per CTA
//   hrow_s[3][max_in_rows][out_w] the same rows after the horizontal pass (rounded to bytes, as Pillow does)
//   kh_s  [out_w][taps_h], bh_s[out_w][2]  horizontal coefficients of the crop window
// kTH / kTV: compile-time tap counts (the trainers' 512 -> 224 resize has 11 + 11; the coefficient rows are zero-padded to the
// tap count, so the unrolled loops always run all taps: a zero weight makes an unused sample harmless); 0 = run-time loops.
template <typename TS, typename TD, int kTH, int kTV>
__global__ void __launch_bounds__(kPreThreads) clip_preprocess_kernel(const PreKernelArgs a) {
  extern __shared__ __align__(16) unsigned char pre_smem[];
  __shared__ float table[768];
  const int out_w = (int)a.out_w, x_w = a.x_w, rows_cap = a.max_in_rows;
  unsigned char* src_s = pre_smem;
  unsigned char* hrow_s = src_s + (size_t)3 * rows_cap * x_w;
  int32_t* kh_s = reinterpret_cast<int32_t*>(hrow_s + (((size_t)3 * rows_cap * out_w + 15) & ~(size_t)15));
  int32_t* bh_s = kh_s + out_w * a.taps_h;
  for (int i = threadIdx.x; i < 768; i += kPreThreads) table[i] = a.table[i];
  for (int i = threadIdx.x; i < out_w * a.taps_h; i += kPreThreads) kh_s[i] = a.coeffs_h[(long long)a.crop_left * a.taps_h + i];
  for (int i = threadIdx.x; i < out_w * 2; i += kPreThreads) bh_s[i] = a.bounds_h[2 * a.crop_left + i];
  const long long b = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;  // first output row (inside the crop window) of this CTA
  const long long r1 = r0 + a.rows_per_cta < a.out_h ? r0 + a.rows_per_cta : a.out_h;
  // input rows the vertical taps of output rows [r0, r1) touch (bounds are monotone in the output index)
  const long long y_first = a.bounds_v[2 * (a.crop_top + r0)];
  const long long y_last = a.bounds_v[2 * (a.crop_top + r1 - 1)] + a.bounds_v[2 * (a.crop_top + r1 - 1) + 1];  // exclusive
  const int n_in = (int)(y_last - y_first);
  // ---- stage the needed input window as bytes
  if constexpr (sizeof(TS) == 1) {  // uint8 NHWC
    const uint8_t* img = reinterpret_cast<const uint8_t*>(a.src) + b * 3 * a.in_h * a.in_w;
    for (int i = threadIdx.x; i < n_in * x_w * 3; i += kPreThreads) {
      const int yy = i / (x_w * 3), rem = i - yy * (x_w * 3);
      const int x = rem / 3, c = rem - x * 3;
      const long long gx = a.x_lo + x;
      src_s[((size_t)c * rows_cap + yy) * x_w + x] = gx < a.in_w ? __ldg(img + ((y_first + yy) * a.in_w + gx) * 3 + c) : 0;
    }
  } else {  // float NCHW: 8 samples per thread per step, one 8-byte shared store
    const TS* img = reinterpret_cast<const TS*>(a.src) + b * 3 * a.in_h * a.in_w;
    const int vec_per_row = x_w / 8;
    for (int i = threadIdx.x; i < 3 * n_in * vec_per_row; i += kPreThreads) {
      const int c = i / (n_in * vec_per_row), rem = i - c * (n_in * vec_per_row);
      const int yy = rem / vec_per_row, v = rem - yy * vec_per_row;
      const long long gx = a.x_lo + 8 * v;
      const uint2 q = quantize8<TS>(img + ((long long)c * a.in_h + y_first + yy) * a.in_w + gx, a.in_w - gx);
      *reinterpret_cast<uint2*>(src_s + ((size_t)c * rows_cap + yy) * x_w + 8 * v) = q;
    }
  }
  __syncthreads();
  // ---- horizontal pass: a thread owns (plane, output column) pairs and walks the input rows with its weights in registers;
  // consecutive threads = consecutive columns (their taps overlap: the byte loads of a warp hit a few words, broadcast)
  if constexpr (kTH > 0) {
    for (int item = threadIdx.x; item < 3 * out_w; item += kPreThreads) {
      const int c = item / out_w, xx = item - c * out_w;
      const int xmin = bh_s[2 * xx] - a.x_lo;
      int k[kTH];
#pragma unroll
      for (int t = 0; t < kTH; ++t) k[t] = kh_s[xx * kTH + t];
      const unsigned char* row = src_s + (size_t)c * rows_cap * x_w + xmin;
      unsigned char* dst = hrow_s + (size_t)c * rows_cap * out_w + xx;
      for (int yy = 0; yy < n_in; ++yy) {
        int acc = 1 << (kPrecisionBits - 1);
#pragma unroll
        for (int t = 0; t < kTH; ++t) acc += (int)row[t] * k[t];
        *dst = (unsigned char)clip8(acc);
        row += x_w;
        dst += out_w;
      }
    }
  } else {
    for (int i = threadIdx.x; i < 3 * n_in * out_w; i += kPreThreads) {
      const int pr = i / out_w, xx = i - pr * out_w;  // pr = c * n_in + yy
      const int c = pr / n_in, yy = pr - c * n_in;
      const int xmin = bh_s[2 * xx] - a.x_lo, n = bh_s[2 * xx + 1];
      const int32_t* k = kh_s + xx * a.taps_h;
      const unsigned char* row = src_s + ((size_t)c * rows_cap + yy) * x_w + xmin;
      int acc = 1 << (kPrecisionBits - 1);
      for (int t = 0; t < n; ++t) acc += (int)row[t] * k[t];
      hrow_s[((size_t)c * rows_cap + yy) * out_w + xx] = (unsigned char)clip8(acc);
    }
  }
  __syncthreads();
  // ---- vertical pass + rescale / normalise table + channels-first store
  TD* out = reinterpret_cast<TD*>(a.dst) + b * 3 * a.out_h * a.out_w;
  const int n_rows = (int)(r1 - r0);
  for (int i = threadIdx.x; i < 3 * n_rows * out_w; i += kPreThreads) {
    const int pr = i / out_w, xx = i - pr * out_w;
    const int c = pr / n_rows, rr = pr - c * n_rows;
    const long long j = a.crop_top + r0 + rr;
    const int ymin = a.bounds_v[2 * j];
    [[maybe_unused]] const int n = a.bounds_v[2 * j + 1];
    const int32_t* k = a.coeffs_v + j * a.taps_v;
    const unsigned char* col = hrow_s + ((size_t)c * rows_cap + (ymin - y_first)) * out_w + xx;
    int acc = 1 << (kPrecisionBits - 1);
    if constexpr (kTV > 0) {
#pragma unroll
      for (int t = 0; t < kTV; ++t) acc += (int)col[(size_t)t * out_w] * __ldg(k + t);  // k is warp-uniform: one broadcast load
    } else {
      for (int t = 0; t < n; ++t) acc += (int)col[(size_t)t * out_w] * __ldg(k + t);
    }
    out[((long long)c * a.out_h + (r0 + rr)) * a.out_w + xx] = cvt_px<TD>(table[c * 256 + clip8(acc)]);
  }
}

static double bicubic_filter(double x) {  // Pillow: #define a -0.5
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

}  // namespace psob200

using namespace psob200;

extern "C" int psob200_resample_taps(int64_t in_size, int64_t out_size) {
  if (in_size <= 0 || out_size <= 0) return PSOB200_ERR_INVALID_ARG;
  double filterscale = (double)in_size / (double)out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  return (int)std::ceil(support) * 2 + 1;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc (box = the whole axis)
extern "C" int psob200_resample_plan(int64_t in_size, int64_t out_size, int32_t* bounds, int32_t* coeffs) {
  if (in_size <= 0 || out_size <= 0 || !bounds || !coeffs) return PSOB200_ERR_INVALID_ARG;
  const double scale = (double)in_size / (double)out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = (int)std::ceil(support) * 2 + 1;
  std::vector<double> k((size_t)ksize);
  for (int64_t xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > (int)in_size) xmax = (int)in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = xmax; x < ksize; ++x) k[x] = 0.0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
    for (int x = 0; x < ksize; ++x)
      coeffs[xx * ksize + x] = k[x] < 0 ? (int32_t)(-0.5 + k[x] * (1 << kPrecisionBits)) : (int32_t)(0.5 + k[x] * (1 << kPrecisionBits));
  }
  return PSOB200_OK;
}

// numpy: rescaled = (uint8 * python_float) in float64 -> astype(float32); normalised = (rescaled - mean32) / std32 in float32
extern "C" int psob200_clip_norm_table(double rescale, const float* mean3, const float* std3, float* table768) {
  if (!mean3 || !std3 || !table768) return PSOB200_ERR_INVALID_ARG;
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 256; ++v) {
      const volatile float r = (float)((double)v * rescale);
      const volatile float d = r - mean3[c];
      table768[c * 256 + v] = d / std3[c];
    }
  return PSOB200_OK;
}

extern "C" int psob200_clip_preprocess(const psob200_clip_preprocess_args* args, void* stream) {
  if (args == nullptr) return PSOB200_ERR_INVALID_ARG;
  const psob200_clip_preprocess_args& p = *args;
  if (!p.src || !p.dst || !p.bounds_h || !p.coeffs_h || !p.bounds_v || !p.coeffs_v || !p.norm_table) return PSOB200_ERR_INVALID_ARG;
  if (p.B <= 0 || p.in_h <= 0 || p.in_w <= 0 || p.rs_h <= 0 || p.rs_w <= 0 || p.out_h <= 0 || p.out_w <= 0)
    return PSOB200_ERR_INVALID_ARG;
  if (p.crop_top < 0 || p.crop_left < 0 || p.crop_top + p.out_h > p.rs_h || p.crop_left + p.out_w > p.rs_w) return PSOB200_ERR_SHAPE;
  if (p.taps_h != psob200_resample_taps(p.in_w, p.rs_w) || p.taps_v != psob200_resample_taps(p.in_h, p.rs_h)) return PSOB200_ERR_INVALID_ARG;
  if (p.src_dtype != PSOB200_U8 && !valid_dtype(p.src_dtype)) return PSOB200_ERR_DTYPE;
  if (!valid_dtype(p.dst_dtype)) return PSOB200_ERR_DTYPE;
  if (p.B > 65535 || p.in_h > (1 << 20) || p.in_w > (1 << 20) || p.out_w > 4096) return PSOB200_ERR_SHAPE;
  // input columns the crop window's horizontal taps touch (host reads nothing from the device: recompute the two bounds)
  auto h_bound = [&](long long j, int& xmin, int& n) {
    const double scale = (double)p.in_w / (double)p.rs_w;
    const double support = 2.0 * (scale < 1.0 ? 1.0 : scale);
    const double center = (j + 0.5) * scale;
    xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > (int)p.in_w) xmax = (int)p.in_w;
    n = xmax - xmin;
  };
  int xa, na, xb, nb;
  h_bound(p.crop_left, xa, na);
  h_bound(p.crop_left + p.out_w - 1, xb, nb);
  const int x_lo = xa & ~7;  // 8-sample vectors stay 16-byte aligned for 16-bit sources whose rows are
  const int x_w = ((xb + nb - x_lo) + p.taps_h + 7) & ~7;  // + one tap row of slack: the unrolled loops read all taps (zero weights)
  // rows per CTA: as many as keep both shared-memory planes within 96 KB (two CTAs per SM), else within 200 KB
  const double vscale = (double)p.in_h / (double)p.rs_h;
  const size_t coef_bytes = (size_t)p.out_w * (p.taps_h + 2) * 4 + 16;
  int rows = 0;
  long long max_in = 0;
  size_t smem = 0;
  for (const size_t limit : {(size_t)96 * 1024, (size_t)200 * 1024}) {
    for (rows = 16; rows >= 1; rows >>= 1) {
      max_in = (long long)std::ceil(rows * vscale) + p.taps_v + 2;
      smem = (size_t)3 * max_in * x_w + (((size_t)3 * max_in * p.out_w + 15) & ~(size_t)15) + coef_bytes;
      if (smem <= limit) break;
    }
    if (rows >= 4 || (rows >= 1 && limit > 100 * 1024)) break;
  }
  if (rows < 1) return PSOB200_ERR_SHAPE;
  PreKernelArgs a;
  a.src = p.src; a.dst = p.dst;
  a.bounds_h = p.bounds_h; a.coeffs_h = p.coeffs_h; a.bounds_v = p.bounds_v; a.coeffs_v = p.coeffs_v;
  a.table = p.norm_table;
  a.in_h = p.in_h; a.in_w = p.in_w; a.out_h = p.out_h; a.out_w = p.out_w; a.crop_top = p.crop_top; a.crop_left = p.crop_left;
  a.taps_h = p.taps_h; a.taps_v = p.taps_v; a.rows_per_cta = rows; a.max_in_rows = (int)max_in;
  a.x_lo = x_lo; a.x_w = x_w;
  const dim3 grid((unsigned)((p.out_h + rows - 1) / rows), (unsigned)p.B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static PerDevice<int> configured[24];
  cudaError_t e = cudaSuccess;
  const bool taps11 = p.taps_h == 11 && p.taps_v == 11;  // 512 -> 224 (turbo :626-640); other sizes: run-time tap loops
#define PSOB200_PRE_LAUNCH(TS, TD, slot)                                                                              \
  do {                                                                                                                \
    auto kern = taps11 ? clip_preprocess_kernel<TS, TD, 11, 11> : clip_preprocess_kernel<TS, TD, 0, 0>;               \
    std::atomic<int>& conf = configured[(slot) + (taps11 ? 12 : 0)].here();                                           \
    if (smem > 48 * 1024 && !conf.load(std::memory_order_acquire)) {                                                  \
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);                        \
      if (e != cudaSuccess) return consume_launch_error("configure clip_preprocess_kernel", e);                       \
      conf.store(1, std::memory_order_release);                                                                       \
    }                                                                                                                 \
    kern<<<grid, kPreThreads, smem, st>>>(a);                                                                         \
  } while (0)
#define PSOB200_PRE_DST(TS, base)                                                           \
  do {                                                                                      \
    if (p.dst_dtype == PSOB200_F32) PSOB200_PRE_LAUNCH(TS, float, base + 0);                \
    else if (p.dst_dtype == PSOB200_BF16) PSOB200_PRE_LAUNCH(TS, __nv_bfloat16, base + 1);  \
    else PSOB200_PRE_LAUNCH(TS, __half, base + 2);                                          \
  } while (0)
  if (p.src_dtype == PSOB200_U8) PSOB200_PRE_DST(uint8_t, 0);
  else if (p.src_dtype == PSOB200_F32) PSOB200_PRE_DST(float, 3);
  else if (p.src_dtype == PSOB200_BF16) PSOB200_PRE_DST(__nv_bfloat16, 6);
  else PSOB200_PRE_DST(__half, 9);
#undef PSOB200_PRE_DST
#undef PSOB200_PRE_LAUNCH
  return consume_launch_error("launch clip_preprocess_kernel", cudaSuccess);
}
