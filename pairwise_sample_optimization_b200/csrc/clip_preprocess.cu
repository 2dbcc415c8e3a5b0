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
// HBM-bound byte work: B * (3 H W read + 3 h w * sizeof(out) written); one CTA per (image, block of output rows) keeps the
// horizontally resized rows it needs in shared memory, so the intermediate image never touches HBM.
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

// one source sample as the byte PIL would see: uint8 NHWC as is; float NCHW quantised like
// ((x + 1.0) * 127.5).clamp(0, 255).to(torch.uint8) evaluated in the tensor's own type (truncation toward zero)
template <typename TS>
struct SrcPixel;
template <>
struct SrcPixel<uint8_t> {
  static __device__ __forceinline__ int get(const uint8_t* img, long long, long long in_w, long long y, long long x, int c) {
    return img[(y * in_w + x) * 3 + c];
  }
};
template <>
struct SrcPixel<float> {
  static __device__ __forceinline__ int get(const float* img, long long in_h, long long in_w, long long y, long long x, int c) {
    float v = __fmul_rn(__fadd_rn(img[((long long)c * in_h + y) * in_w + x], 1.0f), 127.5f);
    v = fminf(fmaxf(v, 0.f), 255.f);
    return (int)v;
  }
};
template <>
struct SrcPixel<__half> {
  static __device__ __forceinline__ int get(const __half* img, long long in_h, long long in_w, long long y, long long x, int c) {
    __half v = __hmul(__hadd(img[((long long)c * in_h + y) * in_w + x], __float2half(1.0f)), __float2half(127.5f));
    float f = fminf(fmaxf(__half2float(v), 0.f), 255.f);
    return (int)f;
  }
};
template <>
struct SrcPixel<__nv_bfloat16> {
  static __device__ __forceinline__ int get(const __nv_bfloat16* img, long long in_h, long long in_w, long long y, long long x,
                                            int c) {
    __nv_bfloat16 v = __hmul(__hadd(img[((long long)c * in_h + y) * in_w + x], __float2bfloat16(1.0f)), __float2bfloat16(127.5f));
    float f = fminf(fmaxf(__bfloat162float(v), 0.f), 255.f);
    return (int)f;
  }
};

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
};

// grid = (ceil(out_h / rows_per_cta), B).  Shared memory: [max_in_rows][out_w][3] bytes of horizontally resized rows.
template <typename TS, typename TD>
__global__ void __launch_bounds__(kPreThreads) clip_preprocess_kernel(const PreKernelArgs a) {
  extern __shared__ unsigned char hrows[];
  __shared__ float table[768];
  for (int i = threadIdx.x; i < 768; i += kPreThreads) table[i] = a.table[i];
  const long long b = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;  // first output row (inside the crop window) of this CTA
  const long long r1 = r0 + a.rows_per_cta < a.out_h ? r0 + a.rows_per_cta : a.out_h;
  // input rows the vertical taps of output rows [r0, r1) touch (bounds are monotone in the output index)
  const long long y_first = a.bounds_v[2 * (a.crop_top + r0)];
  const long long y_last = a.bounds_v[2 * (a.crop_top + r1 - 1)] + a.bounds_v[2 * (a.crop_top + r1 - 1) + 1];  // exclusive
  const int n_in = (int)(y_last - y_first);
  const TS* img = reinterpret_cast<const TS*>(a.src) + b * 3 * a.in_h * a.in_w;
  const int row_elems = (int)a.out_w * 3;
  // ---- horizontal pass: every needed input row -> out_w x 3 bytes
  for (int i = threadIdx.x; i < n_in * row_elems; i += kPreThreads) {
    const int yy = i / row_elems, rem = i - yy * row_elems;
    const int xx = rem / 3, c = rem - xx * 3;
    const int j = (int)a.crop_left + xx;
    const int xmin = a.bounds_h[2 * j], n = a.bounds_h[2 * j + 1];
    const int32_t* k = a.coeffs_h + (long long)j * a.taps_h;
    int acc = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < n; ++t) acc += SrcPixel<TS>::get(img, a.in_h, a.in_w, y_first + yy, xmin + t, c) * k[t];
    hrows[i] = (unsigned char)clip8(acc);
  }
  __syncthreads();
  // ---- vertical pass + rescale / normalise table + channels-first store (consecutive threads: consecutive x)
  TD* out = reinterpret_cast<TD*>(a.dst) + b * 3 * a.out_h * a.out_w;
  const int n_out = (int)(r1 - r0) * row_elems;
  for (int i = threadIdx.x; i < n_out; i += kPreThreads) {
    const int c = i / ((int)(r1 - r0) * (int)a.out_w);
    const int rem = i - c * (int)(r1 - r0) * (int)a.out_w;
    const int rr = rem / (int)a.out_w, xx = rem - rr * (int)a.out_w;
    const long long j = a.crop_top + r0 + rr;
    const int ymin = a.bounds_v[2 * j], n = a.bounds_v[2 * j + 1];
    const int32_t* k = a.coeffs_v + j * a.taps_v;
    const unsigned char* col = hrows + ((long long)(ymin - y_first) * a.out_w + xx) * 3 + c;
    int acc = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < n; ++t) acc += (int)col[(long long)t * row_elems] * k[t];
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
  // rows per CTA: as many as keep the horizontally resized input rows within ~96 KB of shared memory (and >= 1)
  const double vscale = (double)p.in_h / (double)p.rs_h;
  const long long row_bytes = p.out_w * 3;
  int rows = 16;
  long long max_in = 0;
  for (; rows >= 1; rows >>= 1) {
    max_in = (long long)std::ceil(rows * vscale) + p.taps_v + 2;
    if (max_in * row_bytes <= 96 * 1024) break;
  }
  if (rows < 1) return PSOB200_ERR_SHAPE;
  const size_t smem = (size_t)(max_in * row_bytes);
  PreKernelArgs a;
  a.src = p.src; a.dst = p.dst;
  a.bounds_h = p.bounds_h; a.coeffs_h = p.coeffs_h; a.bounds_v = p.bounds_v; a.coeffs_v = p.coeffs_v;
  a.table = p.norm_table;
  a.in_h = p.in_h; a.in_w = p.in_w; a.out_h = p.out_h; a.out_w = p.out_w; a.crop_top = p.crop_top; a.crop_left = p.crop_left;
  a.taps_h = p.taps_h; a.taps_v = p.taps_v; a.rows_per_cta = rows; a.max_in_rows = (int)max_in;
  const dim3 grid((unsigned)((p.out_h + rows - 1) / rows), (unsigned)p.B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static PerDevice<int> configured[12];
  cudaError_t e = cudaSuccess;
#define PSOB200_PRE_LAUNCH(TS, TD, slot)                                                                              \
  do {                                                                                                                \
    auto kern = clip_preprocess_kernel<TS, TD>;                                                                       \
    std::atomic<int>& conf = configured[slot].here();                                                                 \
    if (smem > 48 * 1024 && !conf.load(std::memory_order_acquire)) {                                                  \
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);                        \
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
