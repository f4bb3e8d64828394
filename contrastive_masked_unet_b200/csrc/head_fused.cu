// Decoder tail fused: BatchNorm apply + ReLU of the last 64-channel conv (up_conv1.double_conv[3..5]) folded into the
// prologue of the 1x1 head `conv_last` (64 -> 2), forward and backward (munet_neck.py:48-49,72,81 == FT/model.py:79-81,131).
//
// Unfused, the tail of each decoder moves the full-resolution 64-channel tensor (T = N*H*W*128 B; 2.1 GB at B = 64 @512^2)
// nine times: bn_relu reads y and writes a, the head reads a; backward: the head reads a and writes da, BatchNorm
// backward reads (da, y) twice and writes dy.  Here `a` and `da` never exist in HBM -- both are one multiply-add away
// from what the kernels already hold (a = relu(y*scale + shift), da = W^T dout with a 2-channel dout):
//   forward   y -> pred                                   reads T            (was 3T)
//   backward  reduce: (y, dout) -> sum dz, sum dz*xhat, dW_head, db_head      reads T            (was 2T + T + T)
//             apply : (y, dout) -> dy                     reads T, writes T   (was 3T)
// 9T -> 4T per decoder: ~25 GB of the step's 467 GB.  Values are rounded to bf16 exactly where the unfused path stored
// bf16 (a, da), so results are bit-compatible with it up to summation order.  HBM-bound, roofline = copy bandwidth.
#include "common.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

__device__ __forceinline__ void unpack8h(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 v = __bfloat1622float2(h[k]);
    f[2 * k] = v.x;
    f[2 * k + 1] = v.y;
  }
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

constexpr int kHeadC = 64;
constexpr int kHeadRow = 2 * kHeadC + 2 * kHeadC + 2;   // partial row: sum dz [64], sum dz*y [64], dW [2][64], db [2]

// a warp takes 32 consecutive pixels per iteration: 8 independent 16-byte loads per lane (4 pixels x 8 lanes each),
// 8-lane butterflies, then lane L collects pixel L so that both output planes get one coalesced 128-byte store
__global__ void __launch_bounds__(256) bn_relu_head_fwd_kernel(const __nv_bfloat16* __restrict__ y,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ shift, const float* __restrict__ w,
                                                               const float* __restrict__ b, float* __restrict__ out,
                                                               size_t npix, size_t hw) {
  const int cg = threadIdx.x & 7;
  float w0[8], w1[8], sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    w0[k] = w[cg * 8 + k];
    w1[k] = w[kHeadC + cg * 8 + k];
    sc[k] = scale[cg * 8 + k];
    sh[k] = shift[cg * 8 + k];
  }
  const float b0 = b[0], b1 = b[1];
  const uint32_t lane = threadIdx.x & 31;
  const size_t warp_id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp_id * 32; base < npix; base += n_warps * 32) {
    uint4 raw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t pix = base + j * 4 + (lane >> 3);
      raw[j] = (pix < npix) ? __ldcs(reinterpret_cast<const uint4*>(y) + pix * 8 + cg) : make_uint4(0, 0, 0, 0);
    }
    float m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f[8];
      unpack8h(raw[j], f);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        // the activation the unfused path stored (bf16), two channels per conversion
        const __nv_bfloat162 ab = __floats2bfloat162_rn(fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f),
                                                        fmaxf(fmaf(f[k + 1], sc[k + 1], sh[k + 1]), 0.f));
        const float2 a = __bfloat1622float2(ab);
        d0 = fmaf(a.x, w0[k], d0);
        d1 = fmaf(a.x, w1[k], d1);
        d0 = fmaf(a.y, w0[k + 1], d0);
        d1 = fmaf(a.y, w1[k + 1], d1);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        d0 += __shfl_xor_sync(0xffffffffu, d0, o);
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      }
      const float v0 = __shfl_sync(0xffffffffu, d0, (lane & 3) * 8);
      const float v1 = __shfl_sync(0xffffffffu, d1, (lane & 3) * 8);
      if ((int)(lane >> 2) == j) { m0 = v0; m1 = v1; }
    }
    const size_t pix = base + lane;
    if (pix < npix) {
      const size_t n = pix / hw, p = pix % hw;
      out[(n * 2 + 0) * hw + p] = m0 + b0;
      out[(n * 2 + 1) * hw + p] = m1 + b1;
    }
  }
}

// bf16 rounding of a channel pair (what the unfused kernels stored), back in fp32
__device__ __forceinline__ float2 round_bf16x2(float2 v) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return __bfloat1622float2(h);
}

// kApply = false: per-block partial row {sum dz, sum dz*y, dW_head, db_head}; kApply = true: dy.
// Packed fp32 math (FFMA2): one instruction per channel PAIR.  The BatchNorm backward is evaluated in the algebraically
// flattened form  dy = scale*dz + (A + B*y),  A = scale*(-k1 + k2*rstd*mean),  B = -scale*k2*rstd  (k1 = sum dz / n,
// k2 = sum dz*xhat / n), and the reduce pass accumulates sum dz*y instead of sum dz*xhat (finished per channel in fp64 by
// head_reduce_kernel: sum dz*xhat = rstd * (sum dz*y - mean * sum dz)): ~8 instructions per element instead of ~20 --
// the first version of this kernel was instruction-bound at 2.4 TB/s.
template <bool kApply>
__global__ void __launch_bounds__(256, 2) bn_head_bwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const float* __restrict__ w,
                                                             const float* __restrict__ dout, const float* __restrict__ sums,
                                                             float inv_count, float* __restrict__ partial,
                                                             __nv_bfloat16* __restrict__ dy, size_t npix, size_t hw) {
  __shared__ float sacc[kHeadRow];
  if (!kApply) {
    for (int i = threadIdx.x; i < kHeadRow; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
  }
  const int cg = threadIdx.x & 7;
  float2 w0[4], w1[4], sc[4], sh[4], cA[4], cB[4];
  float2 a1[4], a2[4], g0[4], g1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = cg * 8 + 2 * k;
    w0[k] = make_float2(w[c], w[c + 1]);
    w1[k] = make_float2(w[kHeadC + c], w[kHeadC + c + 1]);
    sc[k] = make_float2(scale[c], scale[c + 1]);
    sh[k] = make_float2(shift[c], shift[c + 1]);
    if (kApply) {
      float A[2], B[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float k1 = sums[c + e] * inv_count, k2 = sums[kHeadC + c + e] * inv_count;
        const float s_ = scale[c + e], r_ = rstd[c + e], m_ = mean[c + e];
        A[e] = s_ * (k2 * r_ * m_ - k1);
        B[e] = -s_ * k2 * r_;
      }
      cA[k] = make_float2(A[0], A[1]);
      cB[k] = make_float2(B[0], B[1]);
    }
    a1[k] = a2[k] = g0[k] = g1[k] = make_float2(0.f, 0.f);
  }
  float s0 = 0.f, s1 = 0.f;
  const uint32_t lane = threadIdx.x & 31;
  const size_t warp_id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp_id * 32; base < npix; base += n_warps * 32) {
    // lane L fetches dout of pixel L (two coalesced 128-byte loads); handed round by shuffle below
    float dl0 = 0.f, dl1 = 0.f;
    {
      const size_t pix = base + lane;
      if (pix < npix) {
        const size_t n = pix / hw, p = pix % hw;
        dl0 = __ldg(dout + (n * 2 + 0) * hw + p);
        dl1 = __ldg(dout + (n * 2 + 1) * hw + p);
      }
    }
    uint4 raw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t pix = base + j * 4 + (lane >> 3);
      raw[j] = (pix < npix) ? __ldcs(reinterpret_cast<const uint4*>(y) + pix * 8 + cg) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t pix = base + j * 4 + (lane >> 3);
      const bool in = pix < npix;
      const float d0 = __shfl_sync(0xffffffffu, dl0, j * 4 + (lane >> 3));   // 0 for out-of-range pixels
      const float d1 = __shfl_sync(0xffffffffu, dl1, j * 4 + (lane >> 3));
      const float2 d0v = make_float2(d0, d0), d1v = make_float2(d1, d1);
      const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw[j]);
      uint4 pk;
      __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(hp[k]);
        const float2 z = __ffma2_rn(f, sc[k], sh[k]);
        const float2 da = round_bf16x2(__ffma2_rn(d0v, w0[k], __fmul2_rn(d1v, w1[k])));   // what head1x1_bwd stored
        float2 dz;
        dz.x = (in && z.x > 0.f) ? da.x : 0.f;
        dz.y = (in && z.y > 0.f) ? da.y : 0.f;
        if (kApply) {
          const float2 o = __ffma2_rn(sc[k], dz, __ffma2_rn(cB[k], f, cA[k]));
          ho[k] = __floats2bfloat162_rn(o.x, o.y);
        } else {
          const float2 a = round_bf16x2(make_float2(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f)));   // the activation bn_relu stored
          a1[k] = __fadd2_rn(a1[k], dz);
          a2[k] = __ffma2_rn(dz, f, a2[k]);
          g0[k] = __ffma2_rn(d0v, a, g0[k]);      // out-of-range pixels carry d = 0
          g1[k] = __ffma2_rn(d1v, a, g1[k]);
        }
      }
      if (kApply && in) reinterpret_cast<uint4*>(dy)[pix * 8 + cg] = pk;
    }
    s0 += dl0;    // every pixel's dout is held by exactly one lane
    s1 += dl1;
  }
  if (!kApply) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8] = {a1[k].x, a1[k].y, a2[k].x, a2[k].y, g0[k].x, g0[k].y, g1[k].x, g1[k].y};
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        v[t] += __shfl_xor_sync(0xffffffffu, v[t], 8);
        v[t] += __shfl_xor_sync(0xffffffffu, v[t], 16);
      }
      if (lane < 8) {
        const int c = cg * 8 + 2 * k;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          atomicAdd(&sacc[c + e], v[0 + e]);
          atomicAdd(&sacc[kHeadC + c + e], v[2 + e]);
          atomicAdd(&sacc[2 * kHeadC + c + e], v[4 + e]);
          atomicAdd(&sacc[3 * kHeadC + c + e], v[6 + e]);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
      atomicAdd(&sacc[4 * kHeadC], s0);
      atomicAdd(&sacc[4 * kHeadC + 1], s1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHeadRow; i += blockDim.x) partial[(size_t)blockIdx.x * kHeadRow + i] = sacc[i];
  }
}

// out[c] = sum_r partial[r][c] in fp64, fixed order; columns [0,128) -> sums, [128,258) -> acc
__global__ void __launch_bounds__(256) head_reduce_kernel(const float* __restrict__ partial, int rows,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          float* __restrict__ sums, float* __restrict__ acc) {
  __shared__ double sred[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double a = 0.0;
  if (c < kHeadRow)
    for (int r = ry; r < rows; r += 8) a += partial[(size_t)r * kHeadRow + c];
  sred[ry][cx] = a;
  __syncthreads();
  if (ry == 0 && c < kHeadRow) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sred[k][cx];
    if (c < kHeadC) {
      sums[c] = (float)t;                                         // sum dz
    } else if (c < 2 * kHeadC) {
      // sum dz*xhat = rstd * (sum dz*y - mean * sum dz): the sum dz of the same channel, reduced here again in fp64
      double dz = 0.0;
      for (int r = 0; r < rows; ++r) dz += partial[(size_t)r * kHeadRow + (c - kHeadC)];
      sums[c] = (float)((double)rstd[c - kHeadC] * (t - (double)mean[c - kHeadC] * dz));
    } else {
      acc[c - 2 * kHeadC] = (float)t;
    }
  }
}

// one resident wave per launch (grid-stride loops): SMs x blocks that fit per SM
template <typename K>
static int resident_grid(K kernel) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, 0) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 1;
  }
  return num_sms() * (n > 8 ? 8 : n);
}
static int head_grid_max() { return num_sms() * 8; }

}  // namespace cmu

using namespace cmu;

extern "C" {

int cmu_bn_relu_head_grid(void) { return head_grid_max(); }

// pred (N,2,H,W) fp32 = conv_last(relu(y * scale + shift)) for a 64-channel act tensor y (raw conv output, NHWC bf16)
int cmu_bn_relu_head_fwd(const void* y, const float* scale, const float* shift, const float* w, const float* b, float* out,
                         int n, int h, int wd, int cin, int cout, void* stream) {
  CMU_REQUIRE(cin == 64 && cout == 2, "bn_relu_head: only 64 -> 2 is supported (got %d -> %d)", cin, cout);
  const size_t npix = (size_t)n * h * wd;
  static int grid_fwd = 0;
  if (grid_fwd == 0) grid_fwd = resident_grid(bn_relu_head_fwd_kernel);
  bn_relu_head_fwd_kernel<<<grid_fwd, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)y, scale, shift, w, b, out, npix,
                                                                          (size_t)h * wd);
  CMU_LAUNCH_CHECK();
  return 0;
}

// dout (N,2,H,W) fp32 -> dy (act, gradient of the raw conv output), sums[2][64] = (sum dz, sum dz*xhat) = (dbeta, dgamma),
// acc[130] = dW_head (2,64) followed by db_head (2).  partial: float[cmu_bn_relu_head_grid()][258].
int cmu_bn_relu_head_bwd(const void* y, const float* scale, const float* shift, const float* mean, const float* rstd,
                         const float* w, const float* dout, float* partial, float* sums, float* acc, void* dy, int n, int h,
                         int wd, int cin, int cout, int training, void* stream) {
  CMU_REQUIRE(cin == 64 && cout == 2, "bn_relu_head: only 64 -> 2 is supported (got %d -> %d)", cin, cout);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t npix = (size_t)n * h * wd;
  const size_t hw = (size_t)h * wd;
  static int grid = 0, grid_apply = 0;
  if (grid == 0) {
    grid = resident_grid(bn_head_bwd_kernel<false>);
    grid_apply = resident_grid(bn_head_bwd_kernel<true>);
  }
  // eval-mode BN is a fixed affine map: the batch-statistics terms of the backward vanish
  const float inv_count = training ? 1.f / (float)npix : 0.f;
  bn_head_bwd_kernel<false><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y, scale, shift, mean, rstd, w, dout, nullptr, 0.f,
                                                  partial, nullptr, npix, hw);
  CMU_LAUNCH_CHECK();
  head_reduce_kernel<<<ceil_div(kHeadRow, 32), 256, 0, st>>>(partial, grid, mean, rstd, sums, acc);
  CMU_LAUNCH_CHECK();
  bn_head_bwd_kernel<true><<<grid_apply, 256, 0, st>>>((const __nv_bfloat16*)y, scale, shift, mean, rstd, w, dout, sums, inv_count,
                                                 nullptr, (__nv_bfloat16*)dy, npix, hw);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
