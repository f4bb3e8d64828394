// HBM-bound kernels around the tensor-core convolutions: weight packing, the Cin=1 first convolution (fused with
// the patch-mask multiply), BatchNorm statistics finalisation, BN-apply + ReLU (+ 2x2 max-pool), and the
// BatchNorm/ReLU/max-pool backward passes.  All activations are NHWC bf16; every thread moves 16-byte vectors
// (8 channels) so that warps read/write whole 128-byte lines.
#include <algorithm>
#include "common.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

struct bf16x8 {
  uint4 u;
};
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// ---------------------------------------------------------------------------------- weight packing
// torch Conv2d weight (Cout, Cin, 3, 3) fp32 ->
//   wf[t = s*3 + r][co][ci]   (fprop B operand, K-major)         t indexes kernel column s first (see tc_conv.cu)
//   wd[t' = s'*3 + r'][ci][co] = w[co][ci][2-r'][2-s']            (dgrad = conv of dy with the rotated kernel)
__global__ void pack_conv3_kernel(const float* __restrict__ w, int cout, int cin, __nv_bfloat16* __restrict__ wf,
                                  __nv_bfloat16* __restrict__ wd) {
  const size_t total = (size_t)cout * cin * 9;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int s = i % 3, r = (i / 3) % 3;
    const int ci = (i / 9) % cin;
    const int co = i / (9 * (size_t)cin);
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    if (wf) wf[((size_t)(s * 3 + r) * cout + co) * cin + ci] = v;
    if (wd) wd[((size_t)((2 - s) * 3 + (2 - r)) * cin + ci) * cout + co] = v;
  }
}
// torch ConvTranspose2d weight (Cin, Cout, 2, 2) fp32 ->
//   wf[(r*2+s)*Cout + co][ci]     (fprop: GEMM N = 4*Cout)
//   wd[ci][(r*2+s)*Cout + co]     (dgrad: GEMM K = 4*Cout)
__global__ void pack_convT_kernel(const float* __restrict__ w, int cin, int cout, __nv_bfloat16* __restrict__ wf,
                                  __nv_bfloat16* __restrict__ wd) {
  const size_t total = (size_t)cin * cout * 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int rs = i % 4;
    const int co = (i / 4) % cout;
    const int ci = i / (4 * (size_t)cout);
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    if (wf) wf[((size_t)rs * cout + co) * cin + ci] = v;
    if (wd) wd[(size_t)ci * 4 * cout + (size_t)rs * cout + co] = v;
  }
}
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 o;
    *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(y)[i] = o;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}

// fp32 (R, C) row-major -> bf16 (C, R) row-major through a 64x64 shared-memory tile (both sides coalesced);
// optionally also writes the untransposed bf16 copy (same read of x).
__global__ void __launch_bounds__(256) transpose_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ yt,
                                                             __nv_bfloat16* __restrict__ y, long long R, long long C) {
  __shared__ float tile[64][65];
  const long long r0 = (long long)blockIdx.y * 64, c0 = (long long)blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  for (int i = ty; i < 64; i += 4) {
    const long long r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      v = x[r * C + c];
      if (y != nullptr) y[r * C + c] = __float2bfloat16_rn(v);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 4) {
    const long long c = c0 + i, r = r0 + tx;
    if (r < R && c < C) yt[c * R + r] = __float2bfloat16_rn(tile[tx][i]);
  }
}
// same, 16-byte accesses on both sides (needs R % 8 == 0, C % 4 == 0 and 16-byte aligned pointers): float4 loads,
// 8-byte stores of the plain copy, 16-byte (8 x bf16) stores of the transposed one
__global__ void __launch_bounds__(256) transpose_cast_vec_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ yt,
                                                                 __nv_bfloat16* __restrict__ y, long long R, long long C) {
  __shared__ float tile[64][65];
  const long long r0 = (long long)blockIdx.y * 64, c0 = (long long)blockIdx.x * 64;
  const int q = threadIdx.x & 15, rr = threadIdx.x >> 4;   // 16 float4 per row, 16 rows per pass
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int i = pass * 16 + rr;
    const long long r = r0 + i, c = c0 + q * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < R && c < C) {
      v = __ldcs(reinterpret_cast<const float4*>(x + r * C + c));
      if (y != nullptr) {
        uint2 pk;
        pk.x = pack_bf16x2(v.x, v.y);
        pk.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(y + r * C + c) = pk;
      }
    }
    tile[i][q * 4 + 0] = v.x;
    tile[i][q * 4 + 1] = v.y;
    tile[i][q * 4 + 2] = v.z;
    tile[i][q * 4 + 3] = v.w;
  }
  __syncthreads();
  const int rb = threadIdx.x & 7, cc = threadIdx.x >> 3;   // 8 row blocks of 8, 32 columns per pass
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int ci = pass * 32 + cc;
    const long long c = c0 + ci, r = r0 + rb * 8;
    if (c < C && r < R) {
      uint4 pk;
      pk.x = pack_bf16x2(tile[rb * 8 + 0][ci], tile[rb * 8 + 1][ci]);
      pk.y = pack_bf16x2(tile[rb * 8 + 2][ci], tile[rb * 8 + 3][ci]);
      pk.z = pack_bf16x2(tile[rb * 8 + 4][ci], tile[rb * 8 + 5][ci]);
      pk.w = pack_bf16x2(tile[rb * 8 + 6][ci], tile[rb * 8 + 7][ci]);
      *reinterpret_cast<uint4*>(yt + c * R + r) = pk;
    }
  }
}

// ---------------------------------------------------------------------------------- first conv (Cin = 1)
// y[n,h,w,co] = sum_{r,s} w[co][r][s] * (x[n,h+r-1,w+s-1] * (1 - mask0[h+r-1,w+s-1]))     (Q1: image-0 mask)
// Persistent blocks walk image rows (n,h); the three masked input rows are staged once in shared memory (with a
// zero halo column on both sides), then 8 threads per pixel produce 8 channels each (Cout = 64) and write one
// coalesced 128-byte NHWC row.  Per-channel sum / sum-of-squares stay in registers across rows and are reduced per
// block -> stats_partial[block][2][64].  No 64-bit div/mod in the inner loop.
constexpr int kC1Cout = 64;
constexpr int kC1MaxW = 2048;

constexpr int kC1Rows = 4;   // output rows per staging round: kC1Rows + 2 masked input rows are staged once
__device__ __forceinline__ void c1_stage_rows(float* srow, int pitch, const float* __restrict__ x,
                                              const uint8_t* __restrict__ mask0, int nb, int h, int H, int W) {
  for (int i = threadIdx.x; i < (kC1Rows + 2) * pitch; i += blockDim.x) {
    const int r = i / pitch, c = i - r * pitch;
    const int hh = h + r - 1, ww = c - 1;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
      v = __ldg(x + ((size_t)nb * H + hh) * W + ww);
      if (mask0 != nullptr && mask0[hh * W + ww]) v = 0.f;
    }
    srow[i] = v;
  }
}

__global__ void __launch_bounds__(256) conv_c1_fprop_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask0,
                                                            const float* __restrict__ w, __nv_bfloat16* __restrict__ y,
                                                            float* __restrict__ stats_partial, int N, int H, int W) {
  extern __shared__ float c1_smem[];
  float* srow = c1_smem;
  const int pitch = W + 2;
  __shared__ float sw[kC1Cout * 9];
  __shared__ float sred[2][kC1Cout];
  for (int i = threadIdx.x; i < kC1Cout * 9; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < 2 * kC1Cout; i += blockDim.x) (&sred[0][0])[i] = 0.f;
  __syncthreads();
  const int cg = threadIdx.x & 7;   // channel group: channels cg*8 .. cg*8+7
  const int px = threadIdx.x >> 3;  // 32 pixels per pass
  // packed fp32 math (sm_100 FFMA2): one instruction per channel PAIR; each lane is an ordinary IEEE fma, so the
  // results are bit-identical to the scalar loop
  float2 wr[4][9];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[c][t] = make_float2(sw[(cg * 8 + 2 * c) * 9 + t], sw[(cg * 8 + 2 * c + 1) * 9 + t]);
  float2 s1p[4], s2p[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) s1p[c] = s2p[c] = make_float2(0.f, 0.f);
  const int hgroups = (H + kC1Rows - 1) / kC1Rows;
  const int groups = N * hgroups;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int nb = grp / hgroups, h0 = (grp - nb * hgroups) * kC1Rows;
    __syncthreads();
    c1_stage_rows(srow, pitch, x, mask0, nb, h0, H, W);
    __syncthreads();
    for (int rr = 0; rr < kC1Rows && h0 + rr < H; ++rr) {
      uint4* yrow = reinterpret_cast<uint4*>(y) + ((size_t)nb * H + h0 + rr) * W * 8;
      const float* sr = srow + rr * pitch;
      for (int w0 = 0; w0 < W; w0 += 32) {
        const int wq = w0 + px;
        if (wq < W) {
          float2 xin[9];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const float v = sr[r * pitch + wq + s];
              xin[r * 3 + s] = make_float2(v, v);
            }
          float o[8];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float2 a = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < 9; ++t) a = __ffma2_rn(wr[c][t], xin[t], a);
            o[2 * c] = a.x;
            o[2 * c + 1] = a.y;
            s1p[c] = __fadd2_rn(s1p[c], a);
            s2p[c] = __ffma2_rn(a, a, s2p[c]);
          }
          yrow[wq * 8 + cg] = pack8(o);
        }
      }
    }
  }
  float s1[8], s2[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    s1[2 * c] = s1p[c].x;
    s1[2 * c + 1] = s1p[c].y;
    s2[2 * c] = s2p[c].x;
    s2[2 * c + 1] = s2p[c].y;
  }
  if (stats_partial != nullptr) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a = s1[c], b = s2[c];   // lanes with equal cg: xor-reduce over lane bits 3,4
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      b += __shfl_xor_sync(0xffffffffu, b, 8);
      b += __shfl_xor_sync(0xffffffffu, b, 16);
      if ((threadIdx.x & 31) < 8) {
        atomicAdd(&sred[0][cg * 8 + c], a);
        atomicAdd(&sred[1][cg * 8 + c], b);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * kC1Cout; i += blockDim.x)
      stats_partial[(size_t)blockIdx.x * 2 * kC1Cout + i] = (&sred[0][0])[i];
  }
}

// dW[co][r][s] = sum_p dy[p,co] * xm[p + (r-1, s-1)]; partial[block][64*9], reduced by a second tiny kernel.
__global__ void __launch_bounds__(256) conv_c1_wgrad_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask0,
                                                            const __nv_bfloat16* __restrict__ dy, float* __restrict__ partial,
                                                            int N, int H, int W) {
  extern __shared__ float c1_smem[];
  float* srow = c1_smem;
  const int pitch = W + 2;
  __shared__ float sred[kC1Cout * 9];
  for (int i = threadIdx.x; i < kC1Cout * 9; i += blockDim.x) sred[i] = 0.f;
  const int cg = threadIdx.x & 7;
  const int px = threadIdx.x >> 3;
  float2 acc2[4][9];   // channel pairs: packed fp32 FMA (FFMA2), bit-identical to the scalar accumulation
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc2[c][t] = make_float2(0.f, 0.f);
  const int hgroups = (H + kC1Rows - 1) / kC1Rows;
  const int groups = N * hgroups;
  for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int nb = grp / hgroups, h0 = (grp - nb * hgroups) * kC1Rows;
    __syncthreads();
    c1_stage_rows(srow, pitch, x, mask0, nb, h0, H, W);
    __syncthreads();
    for (int rr = 0; rr < kC1Rows && h0 + rr < H; ++rr) {
      const uint4* grow = reinterpret_cast<const uint4*>(dy) + ((size_t)nb * H + h0 + rr) * W * 8;
      const float* sr = srow + rr * pitch;
      for (int w0 = 0; w0 < W; w0 += 32) {
        const int wq = w0 + px;
        if (wq < W) {
          float2 xin[9];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const float v = sr[r * pitch + wq + s];
              xin[r * 3 + s] = make_float2(v, v);
            }
          float g[8];
          unpack8(__ldcs(grow + wq * 8 + cg), g);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float2 gp = make_float2(g[2 * c], g[2 * c + 1]);
#pragma unroll
            for (int t = 0; t < 9; ++t) acc2[c][t] = __ffma2_rn(gp, xin[t], acc2[c][t]);
          }
        }
      }
    }
  }
  float acc[8][9];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      acc[2 * c][t] = acc2[c][t].x;
      acc[2 * c + 1][t] = acc2[c][t].y;
    }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float a = acc[c][t];
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      if ((threadIdx.x & 31) < 8) atomicAdd(&sred[(cg * 8 + c) * 9 + t], a);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < kC1Cout * 9; i += blockDim.x) partial[(size_t)blockIdx.x * kC1Cout * 9 + i] = sred[i];
}
// out[c] (+)= sum_r partial[r][c].  256 threads = 32 columns x 8 row lanes (coalesced 128-byte row segments, 8 rows in
// flight per column), fp64 accumulation, fixed summation order.
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                          int rows, int cols, int accumulate) {
  __shared__ double sred[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double a = 0.0;
  if (c < cols)
    for (int r = ry; r < rows; r += 8) a += partial[(size_t)r * cols + c];
  sred[ry][cx] = a;
  __syncthreads();
  if (ry == 0 && c < cols) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sred[k][cx];
    out[c] = accumulate ? out[c] + (float)t : (float)t;
  }
}

// out[c] = sum over CTAs of the SUM half of conv-epilogue statistics: partial[grid][2][bn_tile], the CTA with index b owns
// channel tile (b % n_tiles).  Used for the ConvTranspose bias gradient (= per-channel sum of the dgrad output).
__global__ void __launch_bounds__(256) stats_colsum_kernel(const float* __restrict__ partial, int grid, int bn, int n_tiles,
                                                           int c_count, float* __restrict__ out) {
  __shared__ double sred[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double a = 0.0;
  if (c < c_count) {
    const int nt = c / bn, cc = c - nt * bn;
    for (int r = nt + ry * n_tiles; r < grid; r += 8 * n_tiles) a += partial[((size_t)r * 2) * bn + cc];
  }
  sred[ry][cx] = a;
  __syncthreads();
  if (ry == 0 && c < c_count) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sred[k][cx];
    out[c] = (float)t;
  }
}

// ---------------------------------------------------------------------------------- BN finalize
// partial layout: [grid][2][bn_tile]; the CTA with index b owns channel tile (b % n_tiles).
// Train: mean/var from the batch (biased var for normalisation, unbiased for running_var); the conv bias is not
// applied in the data path (it cancels inside train-mode BN) but it shifts the batch mean that running_mean tracks.
// Eval: scale/shift from the running statistics (bias folded into the shift).
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partial, int grid, int bn_tile, int C,
                                                          double count, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                                          float* running_mean, float* running_var, float momentum, float eps,
                                                          int training, float* __restrict__ scale, float* __restrict__ shift,
                                                          float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  // 256 threads = 32 channels x 8 row lanes; the per-CTA partial rows are summed in a fixed order in fp64
  __shared__ double sred[2][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double s1 = 0.0, s2 = 0.0;
  if (training && c < C) {
    const int n_tiles = C / bn_tile;
    const int nt = c / bn_tile, j = c % bn_tile;
    for (int b = nt + ry * n_tiles; b < grid; b += 8 * n_tiles) {
      s1 += partial[((size_t)b * 2 + 0) * bn_tile + j];
      s2 += partial[((size_t)b * 2 + 1) * bn_tile + j];
    }
  }
  sred[0][ry][cx] = s1;
  sred[1][ry][cx] = s2;
  __syncthreads();
  if (ry != 0 || c >= C) return;
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  const float cb = conv_bias ? conv_bias[c] : 0.f;
  float mean, rstd;
  if (training) {
    s1 = s2 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s1 += sred[0][k][cx];
      s2 += sred[1][k][cx];
    }
    const double m = s1 / count;
    double var = s2 / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mean + cb);
      const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c] - cb;   // data path carries y without the conv bias
    rstd = rsqrtf(running_var[c] + eps);
  }
  scale[c] = g * rstd;
  shift[c] = bt - mean * g * rstd;
  if (mean_out) mean_out[c] = mean;
  if (rstd_out) rstd_out[c] = rstd;
}

// ---------------------------------------------------------------------------------- BN apply + ReLU (+ pool)
__global__ void __launch_bounds__(256) bn_relu_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, __nv_bfloat16* __restrict__ a,
                                                      size_t npix, int C) {
  // the grid stride is a multiple of C/8, so a thread keeps ONE channel group: scale/shift live in registers and four
  // independent 16-byte loads are in flight per thread
  const int cgs = C >> 3;
  const size_t total = npix * cgs;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = (int)(i % cgs);
  float sc[8], sh[8];
  *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale) + cg * 2);
  *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale) + cg * 2 + 1);
  *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift) + cg * 2);
  *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift) + cg * 2 + 1);
  const uint4* yv = reinterpret_cast<const uint4*>(y);
  uint4* av = reinterpret_cast<uint4*>(a);
  for (; i + 3 * stride < total; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldcs(yv + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f);
      av[i + u * stride] = pack8(f);
    }
  }
  for (; i < total; i += stride) {
    float f[8];
    unpack8(__ldcs(yv + i), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f);
    av[i] = pack8(f);
  }
}
// one thread: a 2x2 pixel quad x 8 channels -> 4 activated vectors + 1 pooled vector
__global__ void __launch_bounds__(256) bn_relu_pool_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, __nv_bfloat16* __restrict__ a,
                                                           __nv_bfloat16* __restrict__ pooled, int N, int H, int W, int C) {
  const int cgs = C >> 3;
  const int Hp = H >> 1, Wp = W >> 1;
  const size_t total = (size_t)N * Hp * Wp * cgs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cg = i % cgs;
    const size_t q = i / cgs;
    const int wp = q % Wp;
    const int hp = (q / Wp) % Hp;
    const size_t nb = q / ((size_t)Wp * Hp);
    float sc[8], sh[8], mx[8];
    *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(scale) + cg * 2);
    *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(scale) + cg * 2 + 1);
    *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(shift) + cg * 2);
    *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(shift) + cg * 2 + 1);
#pragma unroll
    for (int k = 0; k < 8; ++k) mx[k] = 0.f;  // ReLU outputs are >= 0
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const size_t pix = (nb * H + (hp * 2 + (d >> 1))) * W + (wp * 2 + (d & 1));
      float f[8];
      unpack8(reinterpret_cast<const uint4*>(y)[pix * cgs + cg], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f);
      }
      const uint4 pk = pack8(f);
      if (a != nullptr) reinterpret_cast<uint4*>(a)[pix * cgs + cg] = pk;   // a == nullptr: only the pooled map is wanted
      float fr[8];
      unpack8(pk, fr);  // pool the bf16-rounded values (what consumers of `a` see)
#pragma unroll
      for (int k = 0; k < 8; ++k) mx[k] = fmaxf(mx[k], fr[k]);
    }
    reinterpret_cast<uint4*>(pooled)[q * cgs + cg] = pack8(mx);
  }
}

// ---------------------------------------------------------------------------------- BN/ReLU/pool backward
// g  = da (+ dpool routed to the first maximum of its 2x2 window, ATen max_pool2d semantics)
// dz = g * [z > 0],  z = y*scale + shift,  xhat = (y - mean) * rstd
// pass 1: per-channel sum(dz), sum(dz * xhat)  -> partial[block][2][C]
// pass 2: dy = scale * (dz - sum_dz/n - xhat * sum_dzx/n)
template <typename K>
static int resident_blocks(K kernel, int threads, size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 1;
  }
  return n;
}

template <bool kPool, bool kApply>
__global__ void __launch_bounds__(256, kPool ? 2 : 3) bn_bwd_kernel(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ dpool,
                                                     const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const float* __restrict__ sums,
                                                     float inv_count, float* __restrict__ partial,
                                                     __nv_bfloat16* __restrict__ dy, int N, int H, int W, int C) {
  extern __shared__ float sacc[];  // [2][C] (reduce pass), then [6][C] per-channel constants (pooled variants)
  const int cgs = C >> 3;
  // the pooled variants keep 4 pixels x 8 channels live per thread: their per-channel constants stay in shared memory
  // (read as needed) instead of 48 registers, which otherwise spill under the 128-register cap of 2 blocks/SM
  float* consts = sacc + (kApply ? 0 : 2 * C);
  if (kPool) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      consts[i] = scale[i];
      consts[C + i] = shift[i];
      consts[2 * C + i] = mean[i];
      consts[3 * C + i] = rstd[i];
      consts[4 * C + i] = kApply ? sums[i] * inv_count : 0.f;
      consts[5 * C + i] = kApply ? sums[C + i] * inv_count : 0.f;
    }
  }
  if (!kApply) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
  }
  if (kPool || !kApply) __syncthreads();
  // a block handles a fixed channel group per thread: threads are laid out [pixel-slot][cg] with cg fastest
  const int tpb = blockDim.x;
  const int slots = tpb / cgs > 0 ? tpb / cgs : 1;
  const int cg = threadIdx.x % cgs;
  const int slot = threadIdx.x / cgs;
  const bool active = slot < slots && cgs <= tpb;
  float sc[8], sh[8], mu[8], rs[8], k1[8], k2[8];
  if (active && !kPool) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = scale[cg * 8 + k];
      sh[k] = shift[cg * 8 + k];
      mu[k] = mean[cg * 8 + k];
      rs[k] = rstd[cg * 8 + k];
      if (kApply) {
        k1[k] = sums[cg * 8 + k] * inv_count;
        k2[k] = sums[C + cg * 8 + k] * inv_count;
      }
    }
  }
  float a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a1[k] = a2[k] = 0.f;
  const int Hq = kPool ? H >> 1 : H, Wq = kPool ? W >> 1 : W;
  const size_t units = (size_t)N * Hq * Wq;  // quads (pool) or pixels
  if (active) {
    for (size_t u = (size_t)blockIdx.x * slots + slot; u < units; u += (size_t)gridDim.x * slots) {
      if (kPool) {
        const float* cc = consts + cg * 8;
        const int wp = u % Wq;
        const int hp = (u / Wq) % Hq;
        const size_t nb = u / ((size_t)Wq * Hq);
        float gp[8];
        unpack8(reinterpret_cast<const uint4*>(dpool)[u * cgs + cg], gp);
        float yv[4][8];
        size_t pixs[4];
        uint4 rg[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          pixs[d] = (nb * H + (hp * 2 + (d >> 1))) * W + (wp * 2 + (d & 1));
          unpack8(__ldcs(reinterpret_cast<const uint4*>(y) + pixs[d] * cgs + cg), yv[d]);
          rg[d] = (da != nullptr) ? __ldcs(reinterpret_cast<const uint4*>(da) + pixs[d] * cgs + cg) : make_uint4(0, 0, 0, 0);
        }
        // per channel: which window element the forward pooled (first maximum in row-major order of the bf16-rounded
        // activations, exactly what bn_relu_pool compared) and which elements passed the ReLU
        int win[8];
        bool pos[4][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float zb[4];
#pragma unroll
          for (int d = 0; d < 4; ++d)
            zb[d] = __bfloat162float(__float2bfloat16_rn(fmaxf(fmaf(yv[d][k], cc[k], cc[C + k]), 0.f)));
          float m = zb[0];
          int w = 0;
#pragma unroll
          for (int d = 1; d < 4; ++d) {
            if (zb[d] > m) { m = zb[d]; w = d; }
          }
          win[k] = w;
#pragma unroll
          for (int d = 0; d < 4; ++d) pos[d][k] = zb[d] > 0.f;
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          float g[8], o[8];
          unpack8(rg[d], g);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float gg = g[k] + (win[k] == d ? gp[k] : 0.f);
            const float dz = pos[d][k] ? gg : 0.f;
            const float xh = (yv[d][k] - cc[2 * C + k]) * cc[3 * C + k];
            if (kApply) o[k] = cc[k] * (dz - cc[4 * C + k] - xh * cc[5 * C + k]);
            else { a1[k] += dz; a2[k] += dz * xh; }
          }
          if (kApply) reinterpret_cast<uint4*>(dy)[pixs[d] * cgs + cg] = pack8(o);
        }
      } else {
        // two independent pixels per iteration (4 x 16-byte loads in flight per thread)
        const size_t ustride = (size_t)gridDim.x * slots;
        const bool two = !kApply && (u + ustride) < units;   // (the apply pass is store-bound: no gain, more registers)
        const uint4 ry0 = __ldcs(reinterpret_cast<const uint4*>(y) + u * cgs + cg);
        const uint4 rg0 = __ldcs(reinterpret_cast<const uint4*>(da) + u * cgs + cg);
        uint4 ry1 = ry0, rg1 = rg0;
        if (two) {
          ry1 = __ldcs(reinterpret_cast<const uint4*>(y) + (u + ustride) * cgs + cg);
          rg1 = __ldcs(reinterpret_cast<const uint4*>(da) + (u + ustride) * cgs + cg);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && !two) break;
          float yv[8], g[8], o[8];
          unpack8(half ? ry1 : ry0, yv);
          unpack8(half ? rg1 : rg0, g);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float z = fmaf(yv[k], sc[k], sh[k]);
            const float dz = (z > 0.f) ? g[k] : 0.f;
            const float xh = (yv[k] - mu[k]) * rs[k];
            if (kApply) o[k] = sc[k] * (dz - k1[k] - xh * k2[k]);
            else { a1[k] += dz; a2[k] += dz * xh; }
          }
          if (kApply) reinterpret_cast<uint4*>(dy)[(u + (half ? ustride : 0)) * cgs + cg] = pack8(o);
        }
        if (two) u += ustride;   // consumed two units this iteration (the loop increment adds the other stride)
      }
    }
  }
  if (!kApply) {
    if (active) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        atomicAdd(&sacc[cg * 8 + k], a1[k]);
        atomicAdd(&sacc[C + cg * 8 + k], a2[k]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) partial[(size_t)blockIdx.x * 2 * C + i] = sacc[i];
  }
}

// per-channel sums of a bf16 (rows, C) tensor -> partial[block][C]   (ConvTranspose2d bias gradient)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, size_t rows, int C,
                                                          float* __restrict__ partial) {
  extern __shared__ float sacc[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cgs = C >> 3;
  const int slots = blockDim.x / cgs;
  const int cg = threadIdx.x % cgs, slot = threadIdx.x / cgs;
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.f;
  if (slot < slots) {
    const size_t stride = (size_t)gridDim.x * slots;
    size_t r = (size_t)blockIdx.x * slots + slot;
    for (; r + 3 * stride < rows; r += 4 * stride) {   // four independent 16-byte loads in flight per thread
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __ldcs(reinterpret_cast<const uint4*>(x) + (r + j * stride) * cgs + cg);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
        unpack8(v[j], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += f[k];
      }
    }
    for (; r < rows; r += stride) {
      float f[8];
      unpack8(reinterpret_cast<const uint4*>(x)[r * cgs + cg], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += f[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&sacc[cg * 8 + k], a[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) partial[(size_t)blockIdx.x * C + i] = sacc[i];
}

static int ew_grid(size_t work_items, int threads, int per_sm = 8) {
  size_t blocks = (work_items + threads - 1) / threads;
  const size_t cap = (size_t)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace cmu

using namespace cmu;

extern "C" {

int cmu_pack_conv3x3_weights(const float* w, int cout, int cin, void* wf, void* wd, void* stream) {
  const size_t total = (size_t)cout * cin * 9;
  pack_conv3_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(w, cout, cin, (__nv_bfloat16*)wf,
                                                                          (__nv_bfloat16*)wd);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_pack_convT2x2_weights(const float* w, int cin, int cout, void* wf, void* wd, void* stream) {
  const size_t total = (size_t)cout * cin * 4;
  pack_convT_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(w, cin, cout, (__nv_bfloat16*)wf,
                                                                          (__nv_bfloat16*)wd);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream) {
  CMU_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0,
              "cast: pointers must be 16/8-byte aligned");
  cast_bf16_kernel<<<ew_grid((size_t)n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y, (size_t)n);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_transpose_cast_bf16(const float* x, void* yt, void* y, long long rows, long long cols, void* stream) {
  dim3 grid((unsigned)((cols + 63) / 64), (unsigned)((rows + 63) / 64));
  CMU_REQUIRE(grid.y <= 65535, "transpose_cast: too many rows (%lld)", rows);
  const bool vec = rows % 8 == 0 && cols % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)yt % 16 == 0) &&
                   ((uintptr_t)y % 8 == 0);
  if (vec)
    transpose_cast_vec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)yt, (__nv_bfloat16*)y, rows, cols);
  else
    transpose_cast_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)yt, (__nv_bfloat16*)y, rows, cols);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_conv3x3_c1_grid(void) { return num_sms() * 6; }

int cmu_conv3x3_c1_fprop(const float* x, const unsigned char* mask0, const float* w, int cout, void* y,
                         float* stats_partial, int n, int h, int wd, void* stream) {
  CMU_REQUIRE(cout == kC1Cout, "conv3x3_c1: Cout must be 64 (got %d)", cout);
  CMU_REQUIRE(wd <= kC1MaxW, "conv3x3_c1: image width %d exceeds %d", wd, kC1MaxW);
  const int c1_shmem = (kC1Rows + 2) * (wd + 2) * (int)sizeof(float);
  static bool attr_f = false;
  if (!attr_f) {
    CMU_CHECK_CUDA(cudaFuncSetAttribute(conv_c1_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (kC1Rows + 2) * (kC1MaxW + 2) * (int)sizeof(float)));
    attr_f = true;
  }
  conv_c1_fprop_kernel<<<cmu_conv3x3_c1_grid(), 256, c1_shmem, (cudaStream_t)stream>>>(x, mask0, w, (__nv_bfloat16*)y,
                                                                                      stats_partial, n, h, wd);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_conv3x3_c1_wgrad(const float* x, const unsigned char* mask0, const void* dy, int cout, float* partial, float* dw,
                         int accumulate, int n, int h, int wd, void* stream) {
  CMU_REQUIRE(cout == kC1Cout, "conv3x3_c1: Cout must be 64 (got %d)", cout);
  const int grid = cmu_conv3x3_c1_grid();
  CMU_REQUIRE(wd <= kC1MaxW, "conv3x3_c1: image width %d exceeds %d", wd, kC1MaxW);
  const int c1_shmem = (kC1Rows + 2) * (wd + 2) * (int)sizeof(float);
  static bool attr_w = false;
  if (!attr_w) {
    CMU_CHECK_CUDA(cudaFuncSetAttribute(conv_c1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (kC1Rows + 2) * (kC1MaxW + 2) * (int)sizeof(float)));
    attr_w = true;
  }
  conv_c1_wgrad_kernel<<<grid, 256, c1_shmem, (cudaStream_t)stream>>>(x, mask0, (const __nv_bfloat16*)dy, partial, n, h, wd);
  CMU_LAUNCH_CHECK();
  reduce_rows_kernel<<<ceil_div(kC1Cout * 9, 32), 256, 0, (cudaStream_t)stream>>>(partial, dw, grid, kC1Cout * 9,
                                                                                   accumulate);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_bn_finalize(const float* partial, int grid, int bn_tile, int c, double count, const float* gamma,
                    const float* beta, const float* conv_bias, float* running_mean, float* running_var, float momentum,
                    float eps, int training, float* scale, float* shift, float* mean, float* rstd, void* stream) {
  CMU_REQUIRE(!training || (partial != nullptr && grid > 0 && bn_tile > 0 && c % bn_tile == 0), "bn_finalize: bad partial layout");
  CMU_REQUIRE(training || (running_mean && running_var), "bn_finalize: eval mode needs running statistics");
  bn_finalize_kernel<<<ceil_div(c, 32), 256, 0, (cudaStream_t)stream>>>(partial, grid, bn_tile, c, count, gamma, beta,
                                                                        conv_bias, running_mean, running_var, momentum,
                                                                        eps, training, scale, shift, mean, rstd);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_bn_relu_apply(const void* y, const float* scale, const float* shift, void* a, void* pooled, int n, int h, int w,
                      int c, void* stream) {
  CMU_REQUIRE(c % 8 == 0 && 256 % (c / 8) == 0, "bn_relu_apply: C/8 must divide 256 (C = 8, 16, ..., 2048; got %d)", c);
  if (pooled != nullptr) {
    CMU_REQUIRE(h % 2 == 0 && w % 2 == 0, "bn_relu_apply: pooling needs even H, W");
    const size_t total = (size_t)n * (h / 2) * (w / 2) * (c / 8);
    bn_relu_pool_kernel<<<ew_grid(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)y, scale, shift, (__nv_bfloat16*)a, (__nv_bfloat16*)pooled, n, h, w, c);
  } else {
    const size_t npix = (size_t)n * h * w;
    bn_relu_kernel<<<ew_grid(npix * (c / 8), 256, 16), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)y, scale, shift, (__nv_bfloat16*)a, npix, c);
  }
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_colsum_bf16(const void* x, long long rows, int c, float* partial, float* out, void* stream) {
  CMU_REQUIRE(c % 8 == 0 && c / 8 <= 256, "colsum_bf16: C must be a multiple of 8 and <= 2048");
  const int grid = cmu_bn_bwd_grid();
  colsum_bf16_kernel<<<grid, 256, c * sizeof(float), (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (size_t)rows, c,
                                                                            partial);
  CMU_LAUNCH_CHECK();
  reduce_rows_kernel<<<ceil_div(c, 32), 256, 0, (cudaStream_t)stream>>>(partial, out, grid, c, 0);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_stats_colsum(const float* partial, int grid, int bn_tile, int c_total, int c_count, float* out, void* stream) {
  CMU_REQUIRE(partial != nullptr && grid > 0 && bn_tile > 0 && c_total % bn_tile == 0 && c_count <= c_total,
              "stats_colsum: bad partial layout");
  stats_colsum_kernel<<<ceil_div(c_count, 32), 256, 0, (cudaStream_t)stream>>>(partial, grid, bn_tile, c_total / bn_tile,
                                                                              c_count, out);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_bn_bwd_grid(void) { return num_sms() * 4; }

// sums: [2][C] device buffer receiving (sum dz, sum dz*xhat) == (dbeta, dgamma); partial: [cmu_bn_bwd_grid()][2][C]
int cmu_bn_relu_bwd(const void* da, const void* dpool, const void* y, const float* scale, const float* shift,
                    const float* mean, const float* rstd, float* partial, float* sums, void* dy, int n, int h, int w,
                    int c, int training, void* stream) {
  CMU_REQUIRE(c % 8 == 0 && c / 8 <= 256, "bn_relu_bwd: C must be a multiple of 8 and <= 2048");
  CMU_REQUIRE(da != nullptr || dpool != nullptr, "bn_relu_bwd: no incoming gradient");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = cmu_bn_bwd_grid();
  const size_t shmem = 2 * (size_t)c * sizeof(float);
  // eval-mode BN is a fixed affine map: the batch-statistics terms of the backward vanish
  const float inv_count = training ? 1.f / ((float)n * h * w) : 0.f;
  const __nv_bfloat16 *pda = (const __nv_bfloat16*)da, *pdp = (const __nv_bfloat16*)dpool, *py = (const __nv_bfloat16*)y;
  // One resident wave per launch: grid = SMs x blocks that fit per SM (a fixed 4 blocks/SM left the 80-register reduce
  // kernel with a second wave at one-third occupancy: 3.8 TB/s instead of ~5.5).
  const int sms = num_sms();
  if (dpool != nullptr) {
    CMU_REQUIRE(h % 2 == 0 && w % 2 == 0, "bn_relu_bwd: pooling needs even H, W");
    const size_t shmem_p = 8 * (size_t)c * sizeof(float);
    const int g = std::min(grid, sms * resident_blocks(bn_bwd_kernel<true, false>, 256, shmem_p));
    bn_bwd_kernel<true, false><<<g, 256, shmem_p, st>>>(pda, pdp, py, scale, shift, mean, rstd, nullptr, inv_count,
                                                      partial, nullptr, n, h, w, c);
    CMU_LAUNCH_CHECK();
    reduce_rows_kernel<<<ceil_div(2 * c, 32), 256, 0, st>>>(partial, sums, g, 2 * c, 0);
  } else {
    const int g = std::min(grid, sms * resident_blocks(bn_bwd_kernel<false, false>, 256, shmem));
    bn_bwd_kernel<false, false><<<g, 256, shmem, st>>>(pda, pdp, py, scale, shift, mean, rstd, nullptr, inv_count,
                                                       partial, nullptr, n, h, w, c);
    CMU_LAUNCH_CHECK();
    reduce_rows_kernel<<<ceil_div(2 * c, 32), 256, 0, st>>>(partial, sums, g, 2 * c, 0);
  }
  CMU_LAUNCH_CHECK();
  if (dpool != nullptr)
    bn_bwd_kernel<true, true><<<sms * resident_blocks(bn_bwd_kernel<true, true>, 256, 6 * (size_t)c * sizeof(float)), 256,
                                6 * (size_t)c * sizeof(float), st>>>(
        pda, pdp, py, scale, shift, mean, rstd, sums, inv_count, nullptr, (__nv_bfloat16*)dy, n, h, w, c);
  else
    bn_bwd_kernel<false, true><<<sms * resident_blocks(bn_bwd_kernel<false, true>, 256, 0), 256, 0, st>>>(
        pda, pdp, py, scale, shift, mean, rstd, sums, inv_count, nullptr, (__nv_bfloat16*)dy, n, h, w, c);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
