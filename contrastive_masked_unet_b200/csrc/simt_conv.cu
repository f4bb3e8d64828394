// Straightforward CUDA-core versions of the K1 / K2 math, used ONLY as an on-device cross-check for the tcgen05
// kernels (tests) and as a debugging switch (cmu_debug_set(0, 1)).  Same operand layouts, same semantics, no
// tensor cores, no tiling -- slow by design.
#include "common.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

enum { S_CONV3 = 0, S_PLAIN = 1, S_CONVT_FPROP = 2, S_CONVT_DGRAD = 3 };

__device__ __forceinline__ float bf(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }

// out[p, n] = bias + sum_{tap,k} A(p,tap,k) * Wp[tap][n][k]
__global__ void simt_k1_kernel(int mode, const __nv_bfloat16* a0, int c0, const __nv_bfloat16* a1, int c1, int N, int H,
                               int W, const __nv_bfloat16* wp, int n_total, __nv_bfloat16* out0, int oc0,
                               __nv_bfloat16* out1, int oc1, const float* bias, int bias_mod) {
  const size_t total = (size_t)N * H * W * n_total;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int n = idx % n_total;
    const size_t pix = idx / n_total;
    const int w = pix % W, h = (pix / W) % H;
    const size_t img = pix / ((size_t)W * H);
    float acc = 0.f;
    if (mode == S_CONV3) {
      const int ktot = c0 + c1;
      for (int s = 0; s < 3; ++s)
        for (int r = 0; r < 3; ++r) {
          const int hh = h + r - 1, ww = w + s - 1;
          if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
          const size_t sp = (img * H + hh) * W + ww;
          const __nv_bfloat16* wr = wp + ((size_t)(s * 3 + r) * n_total + n) * ktot;
          for (int k = 0; k < c0; ++k) acc = fmaf(bf(a0, sp * c0 + k), bf(wr, k), acc);
          for (int k = 0; k < c1; ++k) acc = fmaf(bf(a1, sp * c1 + k), bf(wr, c0 + k), acc);
        }
    } else if (mode == S_PLAIN || mode == S_CONVT_FPROP) {
      const int ktot = c0 + c1;
      const __nv_bfloat16* wr = wp + (size_t)n * ktot;
      for (int k = 0; k < c0; ++k) acc = fmaf(bf(a0, pix * c0 + k), bf(wr, k), acc);
      for (int k = 0; k < c1; ++k) acc = fmaf(bf(a1, pix * c1 + k), bf(wr, c0 + k), acc);
    } else {  // CONVT_DGRAD: A gathered from dy at (2h+r, 2w+s); K = (r*2+s)*c0 + co
      const __nv_bfloat16* wr = wp + (size_t)n * 4 * c0;
      for (int rs = 0; rs < 4; ++rs) {
        const size_t sp = (img * 2 * H + (2 * h + (rs >> 1))) * 2 * W + (2 * w + (rs & 1));
        for (int k = 0; k < c0; ++k) acc = fmaf(bf(a0, sp * c0 + k), bf(wr, rs * c0 + k), acc);
      }
    }
    if (bias) acc += bias[n % bias_mod];
    const __nv_bfloat16 o = __float2bfloat16_rn(acc);
    if (mode == S_CONVT_FPROP) {
      const int rs = n / oc0, co = n % oc0;
      out0[((img * 2 * H + (2 * h + (rs >> 1))) * 2 * W + (2 * w + (rs & 1))) * oc0 + co] = o;
    } else if (n < oc0) {
      out0[pix * oc0 + n] = o;
    } else {
      out1[pix * oc1 + (n - oc0)] = o;
    }
  }
}

// per-channel (sum, sumsq) of a bf16 (npix, C) tensor -> partial[block][2][C]   (bn_tile = C layout)
__global__ void simt_stats_kernel(const __nv_bfloat16* y, size_t npix, int C, float* partial) {
  extern __shared__ float sacc[];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const size_t total = npix * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float v = __bfloat162float(y[i]);
    atomicAdd(&sacc[i % C], v);
    atomicAdd(&sacc[C + i % C], v * v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) partial[(size_t)blockIdx.x * 2 * C + i] = sacc[i];
}

// conv3 wgrad: dw[co][ci][r][s] (+)= sum_p dy[p,co] * x[p+(r-1,s-1), ci];  convT wgrad: dw[ci][co][r][s] (+)= sum_p
// x[p,ci] * dy[(2h+r,2w+s), co].  One thread per weight element, pixels split over blockIdx.y with float atomics.
__global__ void simt_wgrad_kernel(int mode, const __nv_bfloat16* x0, int c0, const __nv_bfloat16* x1, int c1,
                                  const __nv_bfloat16* dy, int cout, int N, int H, int W, float* dw) {
  const int cin = c0 + c1;
  const int taps = mode == 0 ? 9 : 4;
  const size_t total = (size_t)cin * cout * taps;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int co, ci, r, s;
  if (mode == 0) { s = idx % 3; r = (idx / 3) % 3; ci = (idx / 9) % cin; co = idx / (9 * (size_t)cin); }
  else { s = idx % 2; r = (idx / 2) % 2; co = (idx / 4) % cout; ci = idx / (4 * (size_t)cout); }
  const size_t npix = (size_t)N * H * W;
  const size_t per = (npix + gridDim.y - 1) / gridDim.y;
  const size_t pb = blockIdx.y * per, pe = min(npix, pb + per);
  float acc = 0.f;
  for (size_t pix = pb; pix < pe; ++pix) {
    const int w = pix % W, h = (pix / W) % H;
    const size_t img = pix / ((size_t)W * H);
    if (mode == 0) {
      const int hh = h + r - 1, ww = w + s - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const size_t sp = (img * H + hh) * W + ww;
      const float xv = ci < c0 ? bf(x0, sp * c0 + ci) : bf(x1, sp * c1 + (ci - c0));
      acc = fmaf(bf(dy, pix * cout + co), xv, acc);
    } else {
      const size_t sp = (img * 2 * H + (2 * h + r)) * 2 * W + (2 * w + s);
      acc = fmaf(bf(x0, pix * c0 + ci), bf(dy, sp * cout + co), acc);
    }
  }
  atomicAdd(&dw[idx], acc);
}

int simt_k1(int mode, const void* a0, int c0, const void* a1, int c1, int N, int H, int W, const void* wpk, int n_total,
            void* out0, int oc0, void* out1, int oc1, const float* bias, int bias_mod, float* stats_partial,
            int* stats_grid, cudaStream_t st) {
  const size_t total = (size_t)N * H * W * n_total;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  simt_k1_kernel<<<(int)blocks, 256, 0, st>>>(mode, (const __nv_bfloat16*)a0, c0, (const __nv_bfloat16*)a1, c1, N, H, W,
                                             (const __nv_bfloat16*)wpk, n_total, (__nv_bfloat16*)out0, oc0,
                                             (__nv_bfloat16*)out1, oc1, bias, bias_mod > 0 ? bias_mod : n_total);
  CMU_LAUNCH_CHECK();
  if (stats_partial != nullptr) {
    const int grid = 64;
    simt_stats_kernel<<<grid, 256, 2 * n_total * sizeof(float), st>>>((const __nv_bfloat16*)out0, (size_t)N * H * W,
                                                                     n_total, stats_partial);
    CMU_LAUNCH_CHECK();
    if (stats_grid) *stats_grid = grid;
  }
  return 0;
}

int simt_wgrad(int mode, const void* x0, int c0, const void* x1, int c1, const void* dy, int cout, int N, int H, int W,
               float* dw, int accumulate, cudaStream_t st) {
  const int cin = c0 + c1;
  const size_t total = (size_t)cin * cout * (mode == 0 ? 9 : 4);
  if (!accumulate) CMU_CHECK_CUDA(cudaMemsetAsync(dw, 0, total * sizeof(float), st));
  dim3 grid((unsigned)((total + 127) / 128), 32);
  simt_wgrad_kernel<<<grid, 128, 0, st>>>(mode, (const __nv_bfloat16*)x0, c0, (const __nv_bfloat16*)x1, c1,
                                         (const __nv_bfloat16*)dy, cout, N, H, W, dw);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // namespace cmu
