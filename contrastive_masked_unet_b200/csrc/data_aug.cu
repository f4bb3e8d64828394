// GPU data pipeline of CM-UNet pretraining: the per-sample work of
// Pretraining/CM-UNet/cmae/datasets/cmunet_dataset.py:74-88 for a whole batch, with the random parameters drawn on the
// host exactly as the reference draws them and passed in explicitly.
//
//   raw (N,H0,W0) u8|f32 --bicubic--> (N,256,256) --crop box + bicubic--> (N,256,256) --flip, shift--> img (N,224,224)
//                                                                                      --flip, shift, + sigma*noise--> img_t
//
// The resize is Pillow's (`Image.resize(size, BICUBIC)`, libImaging/Resample.c), restated so that results are BIT-EXACT
// with the reference for uint8 ("L": 22-bit fixed-point coefficients, int32 accumulation) and float32 ("F": double
// accumulation in tap order, float32 store after each pass) images: horizontal pass, then vertical pass, coefficients
// evaluated once per CTA in double precision (window [xmin, xmax), support 2 * max(scale, 1), normalised).
// This file is compiled with -fmad=false: Pillow / numpy evaluate these expressions without fused multiply-adds and a
// contracted FMA would change the last bit.  Launch-latency bound and tiny next to the training step (N = 64: 0.3 ms).
#include "common.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

__device__ __forceinline__ double bicubic_w(double x) {   // Resample.c bicubic_filter, a = -0.5
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

struct Window {
  int xmin, n;
  double center, ss, ww;
};

// Resample.c precompute_coeffs for one output index (in0 = 0, in1 = in_size: the crop is treated as the whole image,
// which is what `mmcv.imcrop` followed by `Image.resize` does)
__device__ __forceinline__ Window make_window(int xx, int in_size, int out_size) {
  double scale = (double)in_size / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  Window w;
  w.center = (xx + 0.5) * scale;
  w.ss = 1.0 / filterscale;
  int xmin = (int)(w.center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(w.center + support + 0.5);
  if (xmax > in_size) xmax = in_size;
  w.xmin = xmin;
  w.n = xmax - xmin;
  double ww = 0.0;
  for (int x = 0; x < w.n; ++x) ww += bicubic_w((x + xmin - w.center + 0.5) * w.ss);
  w.ww = ww;
  return w;
}
__device__ __forceinline__ double window_k(const Window& w, int x) {
  double k = bicubic_w((x + w.xmin - w.center + 0.5) * w.ss);
  if (w.ww != 0.0) k /= w.ww;
  return k;
}

constexpr int kPrecisionBits = 32 - 8 - 2;

// Coefficient of tap x in the form the accumulation uses: 22-bit fixed point (uint8) or double (float32)
template <typename T> struct Coef;
template <> struct Coef<unsigned char> {
  typedef int type;
  static __device__ __forceinline__ int make(const Window& w, int x) {
    const double k = window_k(w, x);
    return k < 0 ? (int)(-0.5 + k * (1 << kPrecisionBits)) : (int)(0.5 + k * (1 << kPrecisionBits));
  }
  static __device__ __forceinline__ unsigned char finish(int acc) {
    acc >>= kPrecisionBits;
    return (unsigned char)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
  static __device__ __forceinline__ int init() { return 1 << (kPrecisionBits - 1); }
  static __device__ __forceinline__ int mac(int acc, unsigned char v, int k) { return acc + (int)v * k; }
};
template <> struct Coef<float> {
  typedef double type;
  static __device__ __forceinline__ double make(const Window& w, int x) { return window_k(w, x); }
  static __device__ __forceinline__ float finish(double acc) { return (float)acc; }
  static __device__ __forceinline__ double init() { return 0.0; }
  static __device__ __forceinline__ double mac(double acc, float v, double k) { return acc + (double)v * k; }
};

constexpr int kMaxTaps = 16;     // windows up to 16 taps (down-scaling by <= 3.75) are staged; wider ones are recomputed
constexpr int kFinalizeSplit = 8; // CTAs per sample in aug_finalize_kernel
constexpr int kRowsPerCta = 32;  // horizontal pass: rows that share one coefficient evaluation

// boxes: [n][4] = x0, y0, w, h (crop rectangle inside the src plane) or nullptr = the whole plane.
// Horizontal: tmp[n][y][xx], y in [0, box.h), xx in [0, out_w).  Vertical: dst[n][yy][xx].
// The coefficients depend on the output index along the resampled axis only, so they are evaluated ONCE per CTA --
// per thread (its own xx) for kRowsPerCta rows in the horizontal pass, per CTA (one yy) in the vertical pass -- and the
// per-element work is the tap loop alone.  Same values, same accumulation order as the per-element evaluation.
template <typename T, bool HORIZ>
__global__ void __launch_bounds__(256) pil_pass_kernel(const T* __restrict__ src, int src_h, int src_w,
                                                       const int* __restrict__ boxes, T* __restrict__ dst, int dst_h,
                                                       int dst_w, int tmp_h) {
  typedef typename Coef<T>::type CT;
  __shared__ CT s_k[HORIZ ? kMaxTaps * 256 : kMaxTaps];
  const int n = blockIdx.z;
  int x0 = 0, y0 = 0, bw = src_w, bh = src_h;
  if (boxes != nullptr) {
    x0 = boxes[4 * n + 0];
    y0 = boxes[4 * n + 1];
    bw = boxes[4 * n + 2];
    bh = boxes[4 * n + 3];
  }
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (HORIZ) {
    // src: the raw plane (src_h x src_w); dst: tmp plane (tmp_h x dst_w), rows of the crop
    const int y_lo = blockIdx.y * kRowsPerCta;
    if (xx >= dst_w || y_lo >= bh) return;
    const int y_hi = min(y_lo + kRowsPerCta, bh);
    const Window w = make_window(xx, bw, dst_w);
    const bool staged = w.n <= kMaxTaps;
    if (staged)
      for (int x = 0; x < w.n; ++x) s_k[x * 256 + threadIdx.x] = Coef<T>::make(w, x);
    for (int yy = y_lo; yy < y_hi; ++yy) {
      const T* p = src + ((size_t)n * src_h + (y0 + yy)) * src_w + x0 + w.xmin;
      CT acc = Coef<T>::init();
      if (staged)
        for (int x = 0; x < w.n; ++x) acc = Coef<T>::mac(acc, p[x], s_k[x * 256 + threadIdx.x]);
      else
        for (int x = 0; x < w.n; ++x) acc = Coef<T>::mac(acc, p[x], Coef<T>::make(w, x));
      dst[((size_t)n * tmp_h + yy) * dst_w + xx] = Coef<T>::finish(acc);
    }
  } else {
    // src: tmp plane (tmp_h x dst_w) holding bh valid rows; dst: (dst_h x dst_w)
    const int yy = blockIdx.y;
    if (yy >= dst_h) return;
    const Window w = make_window(yy, bh, dst_h);     // identical in every thread of the CTA
    const bool staged = w.n <= kMaxTaps;
    if (staged && (int)threadIdx.x < w.n) s_k[threadIdx.x] = Coef<T>::make(w, threadIdx.x);
    __syncthreads();
    if (xx >= dst_w) return;
    const T* p = src + ((size_t)n * tmp_h + w.xmin) * dst_w + xx;
    CT acc = Coef<T>::init();
    if (staged)
      for (int x = 0; x < w.n; ++x) acc = Coef<T>::mac(acc, p[(size_t)x * dst_w], s_k[x]);
    else
      for (int x = 0; x < w.n; ++x) acc = Coef<T>::mac(acc, p[(size_t)x * dst_w], Coef<T>::make(w, x));
    dst[((size_t)n * dst_h + yy) * dst_w + xx] = Coef<T>::finish(acc);
  }
}

// ---- Philox4x32-10 (counter-based; Salmon et al. 2011) for the production noise field
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double philox_normal(unsigned long long seed, uint32_t sample, uint32_t pixel) {
  uint32_t r[4];
  philox4x32_10(pixel, sample, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return (double)(sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2));
}

template <typename T>
__device__ __forceinline__ float to_f32(T v) { return (float)v; }

// ShiftPixel (processing.py:109-121) on the horizontally flipped (or not) 256 x 256 image + GaussNoise
// (auto_augment.py:1148-1154).  gridDim.y CTAs per sample: each finds the block maximum itself (the 200 KB crop is read
// from L2) and writes its share of the pixels.  params: [n][4] = flip, ph, pw, unused.
template <typename T>
__global__ void __launch_bounds__(256) aug_finalize_kernel(const T* __restrict__ src, int sh, int sw,
                                                           const int* __restrict__ params,
                                                           const double* __restrict__ noise, unsigned long long seed,
                                                           int crop, float* __restrict__ img, float* __restrict__ img_t) {
  const int n = blockIdx.x;
  const int flip = params[4 * n + 0], ph = params[4 * n + 1], pw = params[4 * n + 2];
  const T* plane = src + (size_t)n * sh * sw;
  const int total = crop * crop;
  __shared__ float s_red[8];
  // sigma = max(img_t crop) / 10  (uint8: float64 division; float32: float32 division, numpy scalar semantics)
  float mx = -3.4e38f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int y = i / crop, x = i - y * crop;
    const int sx = flip ? (sw - 1 - (pw + x)) : (pw + x);
    mx = fmaxf(mx, to_f32(plane[(size_t)(ph + y) * sw + sx]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_red[i]);
  double sigma;
  if (sizeof(T) == 1) sigma = (double)mx / 10.0;
  else sigma = (double)(mx / 10.0f);
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += blockDim.x * gridDim.y) {
    const int y = i / crop, x = i - y * crop;
    const int x_a = flip ? (sw - 1 - x) : x;                 // ShiftPixel(pixel = 0)
    const int x_b = flip ? (sw - 1 - (pw + x)) : (pw + x);   // ShiftPixel(pixel = 31) draw (ph, pw)
    img[(size_t)n * total + i] = to_f32(plane[(size_t)y * sw + x_a]);
    const T v = plane[(size_t)(ph + y) * sw + x_b];
    const double z = noise != nullptr ? noise[(size_t)n * total + i] : philox_normal(seed, (uint32_t)n, (uint32_t)i);
    const double out = (double)v + sigma * z;
    float r;
    if (sizeof(T) == 1) r = (float)(unsigned char)(long long)out;   // np.array(out, dtype=uint8): C cast, wraps mod 256
    else r = (float)out;
    img_t[(size_t)n * total + i] = r;
  }
}

template <typename T>
static int run_resize(const void* src, int n, int src_h, int src_w, const int* boxes, void* tmp, void* dst, int out_h,
                      int out_w, cudaStream_t st) {
  dim3 block(256);
  dim3 gh(ceil_div(out_w, 256), ceil_div(src_h, kRowsPerCta), n);
  pil_pass_kernel<T, true><<<gh, block, 0, st>>>((const T*)src, src_h, src_w, boxes, (T*)tmp, out_h, out_w, src_h);
  CMU_LAUNCH_CHECK();
  dim3 gv(ceil_div(out_w, 256), out_h, n);
  pil_pass_kernel<T, false><<<gv, block, 0, st>>>((const T*)tmp, src_h, src_w, boxes, (T*)dst, out_h, out_w, src_h);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // namespace cmu

using namespace cmu;

extern "C" {

// `Image.fromarray(plane[box]).resize((out_w, out_h), Image.BICUBIC)` for n planes; dtype 0 = uint8, 1 = float32.
// tmp: n * src_h * out_w elements of the same dtype (horizontal-pass result).
int cmu_pil_resize_bicubic(const void* src, int dtype, int n, int src_h, int src_w, const int* d_boxes, void* tmp,
                           void* dst, int out_h, int out_w, void* stream) {
  CMU_REQUIRE(dtype == 0 || dtype == 1, "pil_resize: dtype must be 0 (uint8) or 1 (float32)");
  CMU_REQUIRE(n > 0 && src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0 && n <= 65535 && src_h <= 65535 && out_h <= 65535,
              "pil_resize: bad sizes");
  if (dtype == 0) return run_resize<unsigned char>(src, n, src_h, src_w, d_boxes, tmp, dst, out_h, out_w, (cudaStream_t)stream);
  return run_resize<float>(src, n, src_h, src_w, d_boxes, tmp, dst, out_h, out_w, (cudaStream_t)stream);
}

// flip + ShiftPixel(0) -> img; flip + ShiftPixel(ph, pw) + GaussNoise -> img_t.  d_params: [n][4] = flip, ph, pw, 0.
// noise: explicit standard-normal field [n][crop][crop] (float64, what np.random.randn returned) or NULL = Philox(seed).
int cmu_aug_shift_flip_noise(const void* src, int dtype, int n, int src_h, int src_w, const int* d_params,
                             const double* noise, unsigned long long seed, int crop, float* img, float* img_t,
                             void* stream) {
  CMU_REQUIRE(dtype == 0 || dtype == 1, "aug: dtype must be 0 (uint8) or 1 (float32)");
  CMU_REQUIRE(n > 0 && crop > 0 && crop + 31 < src_h + 1 && crop <= src_w, "aug: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 0)
    aug_finalize_kernel<unsigned char><<<dim3(n, kFinalizeSplit), 256, 0, st>>>((const unsigned char*)src, src_h, src_w, d_params, noise, seed,
                                                          crop, img, img_t);
  else
    aug_finalize_kernel<float><<<dim3(n, kFinalizeSplit), 256, 0, st>>>((const float*)src, src_h, src_w, d_params, noise, seed, crop, img, img_t);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
