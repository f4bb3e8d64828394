// Loss / head kernels (HBM- or latency-bound): 1x1 output head (64 -> 2), masked-pixel MSE with the per-row target
// normalisation, fused L2-normalise + logits + online-softmax cross-entropy (InfoNCE) with its gradient, and the
// fine-tuning Dice / IoU / soft-target cross-entropy reductions.  Warp-shuffle reductions, one atomic per block.
#include "common.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// ------------------------------------------------------------------------------------------ 1x1 head
// a: (npix, 64) bf16 NHWC (post BN+ReLU);  out: (N, 2, H, W) fp32 NCHW;  w: (2, 64) fp32, b: (2)
// 8 lanes per pixel (16 bytes each) -> fully coalesced 128-byte rows.
__global__ void __launch_bounds__(256) head1x1_fwd_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ w,
                                                          const float* __restrict__ b, float* __restrict__ out,
                                                          size_t npix, size_t hw) {
  const int cg = threadIdx.x & 7;
  float w0[8], w1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    w0[k] = w[cg * 8 + k];
    w1[k] = w[64 + cg * 8 + k];
  }
  const float b0 = b[0], b1 = b[1];
  // a warp takes 32 consecutive pixels per iteration: 8 independent 16-byte loads per lane (4 pixels x 8 lanes each),
  // 8-lane butterflies, then lane L collects pixel L so that both output planes get one coalesced 128-byte store
  const uint32_t lane = threadIdx.x & 31;
  const size_t warp_id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp_id * 32; base < npix; base += n_warps * 32) {
    uint4 raw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t pix = base + j * 4 + (lane >> 3);
      raw[j] = (pix < npix) ? __ldcs(reinterpret_cast<const uint4*>(a) + pix * 8 + cg) : make_uint4(0, 0, 0, 0);
    }
    float m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f[8];
      unpack8f(raw[j], f);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        d0 = fmaf(f[k], w0[k], d0);
        d1 = fmaf(f[k], w1[k], d1);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        d0 += __shfl_xor_sync(0xffffffffu, d0, o);
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      }
      // pixel base + j*4 + q is complete in every lane of group q; lane L = j*4 + q fetches it
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
// dout: (N,2,H,W) fp32.  da[p,k] = d0*w0[k] + d1*w1[k] (bf16 NHWC);  dw[c][k] += sum_p d_c[p]*a[p,k];  db[c] += sum d_c
// acc: float[130] = dw (2*64) followed by db (2); must be zero-initialised by the caller.
__global__ void __launch_bounds__(256) head1x1_bwd_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ w,
                                                          const float* __restrict__ dout, __nv_bfloat16* __restrict__ da,
                                                          float* __restrict__ acc, size_t npix, size_t hw) {
  __shared__ float sacc[130];
  for (int i = threadIdx.x; i < 130; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cg = threadIdx.x & 7;
  float w0[8], w1[8], g0[8], g1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    w0[k] = w[cg * 8 + k];
    w1[k] = w[64 + cg * 8 + k];
    g0[k] = g1[k] = 0.f;
  }
  float s0 = 0.f, s1 = 0.f;
  // a warp takes 32 consecutive pixels per iteration: lane L fetches dout of pixel L (two coalesced 128-byte loads),
  // every lane issues its 8 independent 16-byte loads of `a` (4 pixels x 8 lanes per step), dout is handed round by shuffle
  const uint32_t lane = threadIdx.x & 31;
  const size_t warp_id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp_id * 32; base < npix; base += n_warps * 32) {
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
      raw[j] = (pix < npix) ? __ldcs(reinterpret_cast<const uint4*>(a) + pix * 8 + cg) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const size_t pix = base + j * 4 + (lane >> 3);
      const float d0 = __shfl_sync(0xffffffffu, dl0, j * 4 + (lane >> 3));
      const float d1 = __shfl_sync(0xffffffffu, dl1, j * 4 + (lane >> 3));
      float f[8], o[8];
      unpack8f(raw[j], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        o[k] = d0 * w0[k] + d1 * w1[k];
        g0[k] = fmaf(d0, f[k], g0[k]);     // out-of-range pixels carry d = 0 and f = 0
        g1[k] = fmaf(d1, f[k], g1[k]);
      }
      if (pix < npix) {
        uint4 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(o[2 * k], o[2 * k + 1]);
        reinterpret_cast<uint4*>(da)[pix * 8 + cg] = pk;
      }
    }
    s0 += dl0;    // every pixel's dout is held by exactly one lane
    s1 += dl1;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float x0 = g0[k], x1 = g1[k];
    x0 += __shfl_xor_sync(0xffffffffu, x0, 8);
    x0 += __shfl_xor_sync(0xffffffffu, x0, 16);
    x1 += __shfl_xor_sync(0xffffffffu, x1, 8);
    x1 += __shfl_xor_sync(0xffffffffu, x1, 16);
    if ((threadIdx.x & 31) < 8) {
      atomicAdd(&sacc[cg * 8 + k], x0);
      atomicAdd(&sacc[64 + cg * 8 + k], x1);
    }
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sacc[128], s0);
    atomicAdd(&sacc[129], s1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 130; i += blockDim.x) atomicAdd(&acc[i], sacc[i]);
}

// ------------------------------------------------------------------------------------------ masked MSE
// cmunet_head.py:62-70.  One warp per image row (b,h): t = (x - mean_w) / sqrt(var_w_unbiased + 1e-6);
// acc[0] += sum m*(p-t)^2, acc[1] += sum m.   pred is addressed as pred + b*pred_bstride + h*W + w (channel-1 view).
template <bool kBwd>
__global__ void __launch_bounds__(256) masked_mse_kernel(const float* __restrict__ x, const float* __restrict__ pred,
                                                         long long pred_bstride, const uint8_t* __restrict__ mask,
                                                         double* __restrict__ acc, float* __restrict__ dpred,
                                                         long long dpred_bstride, const float* __restrict__ gscale, int B,
                                                         int H, int W) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const size_t rows = (size_t)B * H;
  double num = 0.0, den = 0.0;
  float coef = 0.f;
  if (kBwd) coef = gscale[0] * 2.f / (float)acc[1];   // rc_weight * upstream grad * 2 / sum(mask)
  for (size_t row = (size_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows;
       row += (size_t)gridDim.x * warps_per_block) {
    const size_t b = row / H, h = row % H;
    const float* xr = x + row * W;
    float s = 0.f;
    for (int w = lane; w < W; w += 32) s += xr[w];
    const float mean = warp_sum(s) / (float)W;
    float v = 0.f;
    for (int w = lane; w < W; w += 32) {
      const float d = xr[w] - mean;
      v += d * d;
    }
    const float var = warp_sum(v) / (float)(W - 1);
    const float inv = rsqrtf(var + 1e-6f);
    const float* pr = pred + b * pred_bstride + h * W;
    const uint8_t* mr = mask + row * W;
    float ln = 0.f, ld = 0.f;
    for (int w = lane; w < W; w += 32) {
      const float t = (xr[w] - mean) * inv;
      const float m = (float)mr[w];
      const float d = pr[w] - t;
      if (kBwd) dpred[b * dpred_bstride + h * W + w] = coef * m * d;
      else {
        ln += m * d * d;
        ld += m;
      }
    }
    if (!kBwd) {
      num += (double)warp_sum(ln);
      den += (double)warp_sum(ld);
    }
  }
  if (!kBwd) {
    __shared__ double sn[8], sd[8];
    if (lane == 0) {
      sn[threadIdx.x >> 5] = num;
      sd[threadIdx.x >> 5] = den;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, c = 0;
      for (int i = 0; i < warps_per_block; ++i) {
        a += sn[i];
        c += sd[i];
      }
      atomicAdd(&acc[0], a);
      atomicAdd(&acc[1], c);
    }
  }
}
__global__ void masked_mse_finish_kernel(const double* acc, float rc_weight, float* loss) {
  loss[0] = rc_weight * (float)(acc[0] / acc[1]);
}

// ------------------------------------------------------------------------------------------ InfoNCE
// cmunet_head.py:74-88.  One block per query row i:
//   p = q_i / max(|q_i|, 1e-12);  s_j = <p, z_j> / tau  (z rows already L2-normalised, all-gathered);
//   loss_i = logsumexp_j s_j - s_label,  label = i + label_offset;
//   dq_i = (I - p p^T)/|q_i| * [ sum_j (softmax_j - onehot_j) z_j ] * coef / tau,  coef = ct_weight*2*tau/B
// The logits never leave the block (online softmax over column chunks); loss_rows[i] and dq are outputs.
constexpr int kNceDim = 256;
__global__ void __launch_bounds__(256) infonce_kernel(const float* __restrict__ q, const float* __restrict__ z, int n_keys,
                                                      int label_offset, float inv_tau, float coef,
                                                      float* __restrict__ loss_rows, float* __restrict__ dq) {
  __shared__ float sp[kNceDim];
  __shared__ float sred[8];
  __shared__ float s_m, s_l;
  extern __shared__ float slog[];  // [n_keys] logits of this row
  const int i = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float qv = q[(size_t)i * kNceDim + tid];
  float ss = warp_sum(qv * qv);
  if (lane == 0) sred[wid] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int k = 0; k < 8; ++k) tot += sred[k];
  const float nrm = fmaxf(sqrtf(tot), 1e-12f);
  sp[tid] = qv / nrm;
  __syncthreads();
  // logits: one warp per key, lanes stride the 256-dim dot product (coalesced reads of z rows)
  for (int j = wid; j < n_keys; j += 8) {
    const float* zr = z + (size_t)j * kNceDim;
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < kNceDim / 32; ++k) d = fmaf(sp[lane + 32 * k], zr[lane + 32 * k], d);
    d = warp_sum(d);
    if (lane == 0) slog[j] = d * inv_tau;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int j = tid; j < n_keys; j += 256) m = fmaxf(m, slog[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __syncthreads();
  if (lane == 0) sred[wid] = m;
  __syncthreads();
  m = sred[0];
  for (int k = 1; k < 8; ++k) m = fmaxf(m, sred[k]);
  float l = 0.f;
  for (int j = tid; j < n_keys; j += 256) l += __expf(slog[j] - m);
  l = warp_sum(l);
  __syncthreads();
  if (lane == 0) sred[wid] = l;
  __syncthreads();
  l = 0.f;
  for (int k = 0; k < 8; ++k) l += sred[k];
  const int label = i + label_offset;
  if (tid == 0) {
    loss_rows[i] = (m + logf(l)) - slog[label];
    s_m = m;
    s_l = l;
  }
  __syncthreads();
  if (dq != nullptr) {
    // g[d] = sum_j (softmax_j - onehot_j) z[j][d]; thread tid owns dimension d = tid (coalesced over j rows)
    float g = 0.f;
    const float invl = 1.f / s_l;
    for (int j = 0; j < n_keys; ++j) {
      float wj = __expf(slog[j] - s_m) * invl;
      if (j == label) wj -= 1.f;
      g = fmaf(wj, z[(size_t)j * kNceDim + tid], g);
    }
    g *= coef * inv_tau;
    // through the normalisation: dq = (g - p <p, g>) / |q|
    float pg = warp_sum(sp[tid] * g);
    __syncthreads();
    if (lane == 0) sred[wid] = pg;
    __syncthreads();
    pg = 0.f;
    for (int k = 0; k < 8; ++k) pg += sred[k];
    dq[(size_t)i * kNceDim + tid] = (g - sp[tid] * pg) / nrm;
  }
}
__global__ void mean_scale_kernel(const float* __restrict__ rows, int n, float scale, float* __restrict__ out) {
  // deterministic single-warp mean
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += rows[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) out[0] = scale * s / (float)n;
}
__global__ void l2_normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int dim) {
  __shared__ float sred[32];
  const float* xr = x + (size_t)blockIdx.x * dim;
  float s = 0.f;
  for (int k = threadIdx.x; k < dim; k += blockDim.x) s += xr[k] * xr[k];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
  for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += sred[k];
  const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);
  for (int k = threadIdx.x; k < dim; k += blockDim.x) y[(size_t)blockIdx.x * dim + k] = xr[k] * inv;
}

// ------------------------------------------------------------------------------------------ Dice / IoU / CE (fine-tune)
// FT/metrics.py:135-220,503.  logits (N,2,H,W) fp32; gt (N,2,H,W) float64 one-hot/probabilities (Q7).
// acc[0]=tp=sum gt1*pr1  acc[1]=sum pr1  acc[2]=sum gt1  acc[3]=sum_c -y_c*log_softmax_c ; pr1 = [logit1 > logit0]
// (softmax_1 > 0.5 <=> logit1 > logit0).  Optional dlogits = (softmax * sum_c y_c - y) * gscale / (N*H*W).
__global__ void __launch_bounds__(256) seg_losses_kernel(const float* __restrict__ logits, const double* __restrict__ gt,
                                                         double* __restrict__ acc, float* __restrict__ dlogits,
                                                         const float* __restrict__ gscale, size_t npix, size_t hw) {
  double tp = 0, spr = 0, sgt = 0, ce = 0;
  const float gs = (dlogits != nullptr) ? gscale[0] / (float)npix : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / hw, p = i % hw;
    const float l0 = logits[(n * 2) * hw + p], l1 = logits[(n * 2 + 1) * hw + p];
    const double y0 = gt[(n * 2) * hw + p], y1 = gt[(n * 2 + 1) * hw + p];
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    const float lse = mx + logf(e0 + e1);
    const double pr1 = (l1 > l0) ? 1.0 : 0.0;
    tp += y1 * pr1;
    spr += pr1;
    sgt += y1;
    ce += -(y0 * (double)(l0 - lse) + y1 * (double)(l1 - lse));
    if (dlogits != nullptr) {
      const float inv = 1.f / (e0 + e1);
      const float ys = (float)(y0 + y1);
      dlogits[(n * 2) * hw + p] = (e0 * inv * ys - (float)y0) * gs;
      dlogits[(n * 2 + 1) * hw + p] = (e1 * inv * ys - (float)y1) * gs;
    }
  }
  __shared__ double sred[4][8];
  tp = warp_sum_d(tp); spr = warp_sum_d(spr); sgt = warp_sum_d(sgt); ce = warp_sum_d(ce);
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    sred[0][w] = tp; sred[1][w] = spr; sred[2][w] = sgt; sred[3][w] = ce;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) a += sred[threadIdx.x][k];
    atomicAdd(&acc[threadIdx.x], a);
  }
}
// out[0]=dice_loss out[1]=iou_loss out[2]=ce_loss   (float64, as the reference's float64 targets promote them)
__global__ void seg_losses_finish_kernel(const double* acc, double npix, double dice_eps, double beta, double iou_eps,
                                         double* out) {
  const double tp = acc[0], fp = acc[1] - acc[0], fn = acc[2] - acc[0];
  const double b2 = beta * beta;
  out[0] = 1.0 - ((1.0 + b2) * tp + dice_eps) / ((1.0 + b2) * tp + b2 * fn + fp + dice_eps);
  out[1] = 1.0 - (tp + iou_eps) / (acc[2] + acc[1] - tp + iou_eps);
  out[2] = acc[3] / npix;
}

// ------------------------------------------------------------------------------------------ small layout helpers
// mean over the 2 channels of (N,2,H,W) fp32 -> (N, H*W) fp32  (cmunet.py:126 feeding projector.fc0)
__global__ void chmean2_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n_img, size_t hw) {
  const size_t total = n_img * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / hw, p = i % hw;
    y[i] = 0.5f * (x[(n * 2) * hw + p] + x[(n * 2 + 1) * hw + p]);
  }
}
// d(N,2,H,W)[:,c] = 0.5 * dx(N,H*W)  (fp32 in, fp32 out), optionally accumulating
__global__ void chmean2_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dout, size_t n_img, size_t hw) {
  const size_t total = n_img * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / hw, p = i % hw;
    const float g = 0.5f * dx[i];
    dout[(n * 2) * hw + p] = g;
    dout[(n * 2 + 1) * hw + p] = g;
  }
}
// (N, HW, C) bf16 -> (N, C, HW) bf16 / fp32 through a 32x32 smem tile  (cmunet.py:130 flattens NCHW order)
__device__ __forceinline__ void store_out(__nv_bfloat16* p, __nv_bfloat16 v) { *p = v; }
__device__ __forceinline__ void store_out(float* p, __nv_bfloat16 v) { *p = __bfloat162float(v); }
template <typename OutT>
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, OutT* __restrict__ y, int hw, int c) {
  __shared__ __nv_bfloat16 tile[32][33];
  const size_t n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, cc = c0 + threadIdx.x;
    if (p < hw && cc < c) tile[r][threadIdx.x] = x[(n * hw + p) * c + cc];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int cc = c0 + r, p = p0 + threadIdx.x;
    if (p < hw && cc < c) store_out(&y[(n * c + cc) * hw + p], tile[threadIdx.x][r]);
  }
}

static int grid_for(size_t items, int threads, int per_sm = 8) {
  size_t b = (items + threads - 1) / threads;
  const size_t cap = (size_t)num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace cmu

using namespace cmu;

extern "C" {

int cmu_head1x1_fprop(const void* a, const float* w, const float* b, float* out, int n, int h, int wd, int cin, int cout,
                      void* stream) {
  CMU_REQUIRE(cin == 64 && cout == 2, "head1x1: only 64 -> 2 is supported (got %d -> %d)", cin, cout);
  const size_t npix = (size_t)n * h * wd;
  head1x1_fwd_kernel<<<grid_for(npix * 8, 256, 16), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, w, b, out,
                                                                                   npix, (size_t)h * wd);
  CMU_LAUNCH_CHECK();
  return 0;
}

// acc: float[130], zeroed here; on return acc[0:128] = dW (2,64), acc[128:130] = db.
int cmu_head1x1_bwd(const void* a, const float* w, const float* dout, void* da, float* acc, int n, int h, int wd, int cin,
                    int cout, void* stream) {
  CMU_REQUIRE(cin == 64 && cout == 2, "head1x1: only 64 -> 2 is supported (got %d -> %d)", cin, cout);
  const size_t npix = (size_t)n * h * wd;
  CMU_CHECK_CUDA(cudaMemsetAsync(acc, 0, 130 * sizeof(float), (cudaStream_t)stream));
  head1x1_bwd_kernel<<<grid_for(npix * 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)a, w, dout, (__nv_bfloat16*)da, acc, npix, (size_t)h * wd);
  CMU_LAUNCH_CHECK();
  return 0;
}

// acc: double[2] scratch (kept for the backward); loss: float[1]
int cmu_masked_mse_fwd(const float* x, const float* pred, long long pred_bstride, const unsigned char* mask, double* acc,
                       float rc_weight, float* loss, int b, int h, int w, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CMU_REQUIRE(w >= 2, "masked_mse: W must be >= 2 (unbiased variance)");
  CMU_CHECK_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  masked_mse_kernel<false><<<grid_for((size_t)b * h * 32, 256, 8), 256, 0, st>>>(x, pred, pred_bstride, mask, acc, nullptr,
                                                                               0, nullptr, b, h, w);
  CMU_LAUNCH_CHECK();
  masked_mse_finish_kernel<<<1, 1, 0, st>>>(acc, rc_weight, loss);
  CMU_LAUNCH_CHECK();
  return 0;
}

// gscale: device float[1] = rc_weight * upstream gradient; dpred addressed like pred (channel-1 view of a (B,2,H,W) grad)
int cmu_masked_mse_bwd(const float* x, const float* pred, long long pred_bstride, const unsigned char* mask,
                       const double* acc, const float* gscale, float* dpred, long long dpred_bstride, int b, int h, int w,
                       void* stream) {
  masked_mse_kernel<true><<<grid_for((size_t)b * h * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      x, pred, pred_bstride, mask, const_cast<double*>(acc), dpred, dpred_bstride, gscale, b, h, w);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_l2_normalize_rows(const float* x, float* y, int rows, int dim, void* stream) {
  l2_normalize_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(x, y, dim);
  CMU_LAUNCH_CHECK();
  return 0;
}

// q: (B,256) raw predictor output; z: (n_keys,256) L2-normalised keys; loss[0] = ct_weight*2*tau*mean_i CE_i;
// dq (optional, (B,256)) = d loss / d q.  loss_rows: float[B] scratch.
int cmu_infonce_fwd_bwd(const float* q, const float* z, int batch, int n_keys, int dim, int label_offset, float tau,
                        float ct_weight, float* loss_rows, float* loss, float* dq, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CMU_REQUIRE(dim == kNceDim, "infonce: embedding dim must be 256 (got %d)", dim);
  CMU_REQUIRE(label_offset >= 0 && label_offset + batch <= n_keys, "infonce: labels out of range");
  const size_t shmem = (size_t)n_keys * sizeof(float);
  CMU_REQUIRE(shmem <= 160 * 1024, "infonce: %d keys exceed the in-smem logits row; use the queue kernel", n_keys);
  if (shmem > 40 * 1024)
    CMU_CHECK_CUDA(cudaFuncSetAttribute(infonce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
  const float coef = ct_weight * 2.f * tau / (float)batch;
  infonce_kernel<<<batch, 256, shmem, st>>>(q, z, n_keys, label_offset, 1.f / tau, coef, loss_rows, dq);
  CMU_LAUNCH_CHECK();
  mean_scale_kernel<<<1, 32, 0, st>>>(loss_rows, batch, ct_weight * 2.f * tau, loss);
  CMU_LAUNCH_CHECK();
  return 0;
}

// acc: double[4] scratch; out: double[3] = (dice_loss, iou_loss, ce_loss); dlogits optional (CE gradient * gscale[0])
int cmu_seg_losses(const float* logits, const double* gt, double* acc, double* out, float* dlogits, const float* gscale,
                   int n, int h, int w, double dice_eps, double beta, double iou_eps, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const size_t npix = (size_t)n * h * w;
  CMU_CHECK_CUDA(cudaMemsetAsync(acc, 0, 4 * sizeof(double), st));
  seg_losses_kernel<<<grid_for(npix, 256, 8), 256, 0, st>>>(logits, gt, acc, dlogits, gscale, npix, (size_t)h * w);
  CMU_LAUNCH_CHECK();
  seg_losses_finish_kernel<<<1, 1, 0, st>>>(acc, (double)npix, dice_eps, beta, iou_eps, out);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_channel_mean2(const float* x, float* y, int n, long long hw, void* stream) {
  chmean2_kernel<<<grid_for((size_t)n * hw, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, y, (size_t)n, (size_t)hw);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_channel_mean2_bwd(const float* dx, float* dout, int n, long long hw, void* stream) {
  chmean2_bwd_kernel<<<grid_for((size_t)n * hw, 256, 8), 256, 0, (cudaStream_t)stream>>>(dx, dout, (size_t)n, (size_t)hw);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_nhwc_to_nchw_bf16(const void* x, void* y, int n, int hw, int c, void* stream) {
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), n), block(32, 8);
  nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, hw, c);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_nhwc_to_nchw_f32(const void* x, float* y, int n, int hw, int c, void* stream) {
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), n), block(32, 8);
  nhwc_to_nchw_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, y, hw, c);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// =========================================================================================== MoCo-v2 queue head
// (BASELINE.json configs[3]; Pretraining/MoCo/pl_bolts/models/self_supervised/moco/moco2_module.py:224-270,160-175)
// logits = [q.k, q.Queue] / T, CE(label 0).  The N x K negative logits come from the tcgen05 1x1 kernel
// (Queue rows as "pixels", the normalised queries as the weight matrix) as lt[K][N] bf16; this kernel does the
// row-wise online softmax over the K+1 logits of each query, writes the probabilities P[K][N] (bf16, the operand of
// the tensor-core dq GEMM  dq_neg = P^T Queue) and the positive-logit terms.
namespace cmu {

// spatial mean of an NHWC bf16 tensor: out[n][c] = mean_p x[n][p][c]   (moco_data_module.py:65 torch.mean(x,[2,3]))
__global__ void __launch_bounds__(256) spatial_mean_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                                           int hw, int C) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const __nv_bfloat16* p = x + (size_t)n * hw * C + c;
  float a = 0.f;
  for (int i = 0; i < hw; ++i) a += __bfloat162float(p[(size_t)i * C]);
  out[(size_t)n * C + c] = a / (float)hw;
}
__global__ void __launch_bounds__(256) spatial_mean_bwd_kernel(const float* __restrict__ dout, __nv_bfloat16* __restrict__ dx,
                                                               int hw, int C, size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C;
    const size_t n = i / ((size_t)hw * C);
    dx[i] = __float2bfloat16_rn(dout[n * C + c] / (float)hw);
  }
}

// one block per query n: online softmax over {l_pos} U {lt[j][n]}.
// lt: [K][N] bf16 = q_hat . queue_j (NOT yet divided by T); out: loss_rows[n], ppos[n] (probability of the positive),
// P[j][n] bf16 = softmax probability of negative j scaled by `pscale` (= 1 / (N*T), so that P^T Queue is dq_hat directly).
__global__ void __launch_bounds__(256) moco_softmax_kernel(const __nv_bfloat16* __restrict__ lt, const float* __restrict__ lpos,
                                                           int K, int N, float inv_t, float pscale,
                                                           float* __restrict__ loss_rows, float* __restrict__ ppos,
                                                           __nv_bfloat16* __restrict__ P) {
  __shared__ float sred[8];
  __shared__ float s_m, s_l;
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float lp = lpos[n] * inv_t;
  float m = lp;
  for (int j = tid; j < K; j += 256) m = fmaxf(m, __bfloat162float(lt[(size_t)j * N + n]) * inv_t);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) sred[wid] = m;
  __syncthreads();
  m = sred[0];
  for (int k = 1; k < 8; ++k) m = fmaxf(m, sred[k]);
  __syncthreads();
  float l = 0.f;
  for (int j = tid; j < K; j += 256) l += __expf(__bfloat162float(lt[(size_t)j * N + n]) * inv_t - m);
  l = warp_sum(l);
  if (lane == 0) sred[wid] = l;
  __syncthreads();
  if (tid == 0) {
    float t = __expf(lp - m);
    for (int k = 0; k < 8; ++k) t += sred[k];
    s_m = m;
    s_l = t;
    loss_rows[n] = (m + logf(t)) - lp;
    ppos[n] = __expf(lp - m) / t;
  }
  __syncthreads();
  if (P != nullptr) {
    const float inv_l = pscale / s_l, mm = s_m;
    for (int j = tid; j < K; j += 256)
      P[(size_t)j * N + n] = __float2bfloat16_rn(__expf(__bfloat162float(lt[(size_t)j * N + n]) * inv_t - mm) * inv_l);
  }
}
// q_hat = q/|q| (bf16 copy for the GEMM), l_pos[n] = <q_hat, k_n>
// Finishes the fused logits / softmax-numerator kernel: Z_n = sum_k E[k][n] + exp((lpos_n - 1)/T) from the per-CTA partial
// column sums (layout of the convolution statistics: partial[grid][2][bn], channel c in rows r = c / bn (mod n_tiles)),
// loss_n = log Z_n - (lpos_n - 1)/T, ppos_n = exp((lpos_n - 1)/T) / Z_n, row_scale_n = 1 / (Z_n * N * T).
__global__ void moco_finish_kernel(const float* __restrict__ partial, int grid, int bn, int N, const float* __restrict__ lpos,
                                   float inv_t, float* __restrict__ loss_rows, float* __restrict__ ppos,
                                   float* __restrict__ row_scale) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int n_tiles = N / bn, nt = n / bn, c = n - nt * bn;
  double z = 0.0;
  for (int r = nt; r < grid; r += n_tiles) z += (double)partial[((size_t)r * 2) * bn + c];
  const float lp = (lpos[n] - 1.f) * inv_t;
  const float ep = __expf(lp);
  const float zz = (float)z + ep;
  loss_rows[n] = logf(zz) - lp;
  ppos[n] = ep / zz;
  row_scale[n] = inv_t / (zz * (float)N);
}

__global__ void __launch_bounds__(256) moco_prep_kernel(const float* __restrict__ q, const float* __restrict__ k, int D,
                                                        __nv_bfloat16* __restrict__ qh16, float* __restrict__ qh,
                                                        float* __restrict__ qnorm, float* __restrict__ lpos) {
  __shared__ float sred[8], sred2[8];
  const int n = blockIdx.x, tid = threadIdx.x;
  float ss = 0.f;
  for (int d = tid; d < D; d += 256) { const float v = q[(size_t)n * D + d]; ss += v * v; }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) sred[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += sred[i];
  const float nrm = fmaxf(sqrtf(tot), 1e-12f);
  float dp = 0.f;
  for (int d = tid; d < D; d += 256) {
    const float v = q[(size_t)n * D + d] / nrm;
    qh[(size_t)n * D + d] = v;
    qh16[(size_t)n * D + d] = __float2bfloat16_rn(v);
    dp += v * k[(size_t)n * D + d];
  }
  dp = warp_sum(dp);
  if ((tid & 31) == 0) sred2[tid >> 5] = dp;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sred2[i];
    lpos[n] = t;
    qnorm[n] = nrm;
  }
}
// dq = (I - q_hat q_hat^T)/|q| * g,  g = dq_neg + (ppos - 1)/(N*T) * k      (all per row)
__global__ void __launch_bounds__(256) moco_dq_kernel(const float* __restrict__ dq_neg, const float* __restrict__ k,
                                                      const float* __restrict__ qh, const float* __restrict__ qnorm,
                                                      const float* __restrict__ ppos, int D, float coef,
                                                      float* __restrict__ dq, const float* __restrict__ row_scale) {
  __shared__ float sred[8];
  const int n = blockIdx.x, tid = threadIdx.x;
  const float cp = (ppos[n] - 1.f) * coef;
  const float rs = row_scale != nullptr ? row_scale[n] : 1.f;   // dq_neg holds E^T Queue: normalise by 1 / (Z N T)
  float pg = 0.f;
  for (int d = tid; d < D; d += 256) {
    const float g = rs * dq_neg[(size_t)n * D + d] + cp * k[(size_t)n * D + d];
    pg += qh[(size_t)n * D + d] * g;
  }
  pg = warp_sum(pg);
  if ((tid & 31) == 0) sred[tid >> 5] = pg;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += sred[i];
  const float inv = 1.f / qnorm[n];
  for (int d = tid; d < D; d += 256) {
    const float g = rs * dq_neg[(size_t)n * D + d] + cp * k[(size_t)n * D + d];
    dq[(size_t)n * D + d] = (g - qh[(size_t)n * D + d] * tot) * inv;
  }
}
// queue rows [ptr, ptr+n) <- keys (bf16 row copy) and queue_ref[:, ptr+i] <- keys[i] (reference (D,K) fp32 layout)
__global__ void queue_enqueue_kernel(const float* __restrict__ keys, int n, int D, int K, int ptr,
                                     __nv_bfloat16* __restrict__ qrows, float* __restrict__ qref) {
  const size_t total = (size_t)n * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int d = i % D;
    const int r = (ptr + (int)(i / D)) % K;
    const float v = keys[i];
    qrows[(size_t)r * D + d] = __float2bfloat16_rn(v);
    if (qref != nullptr) qref[(size_t)d * K + r] = v;
  }
}

}  // namespace cmu

extern "C" {

int cmu_spatial_mean(const void* x, float* out, int n, int hw, int c, void* stream) {
  dim3 grid(ceil_div(c, 256), n);
  spatial_mean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, hw, c);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_spatial_mean_bwd(const float* dout, void* dx, int n, int hw, int c, void* stream) {
  const size_t total = (size_t)n * hw * c;
  spatial_mean_bwd_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(dout, (__nv_bfloat16*)dx, hw, c, total);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_moco_prep(const float* q, const float* k, int n, int d, void* qh16, float* qh, float* qnorm, float* lpos,
                  void* stream) {
  moco_prep_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(q, k, d, (__nv_bfloat16*)qh16, qh, qnorm, lpos);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_moco_softmax(const void* lt, const float* lpos, int k, int n, float temperature, float* loss_rows, float* loss,
                     float* ppos, void* p, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  moco_softmax_kernel<<<n, 256, 0, st>>>((const __nv_bfloat16*)lt, lpos, k, n, 1.f / temperature,
                                         1.f / ((float)n * temperature), loss_rows, ppos, (__nv_bfloat16*)p);
  CMU_LAUNCH_CHECK();
  mean_scale_kernel<<<1, 32, 0, st>>>(loss_rows, n, 1.f, loss);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_moco_dq(const float* dq_neg, const float* k, const float* qh, const float* qnorm, const float* ppos, int n, int d,
                float temperature, float* dq, void* stream) {
  moco_dq_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(dq_neg, k, qh, qnorm, ppos, d, 1.f / ((float)n * temperature), dq,
                                                      nullptr);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_moco_finish(const float* sums_partial, int sums_grid, int sums_bn, int n, const float* lpos, float temperature,
                    float* loss_rows, float* loss, float* ppos, float* row_scale, void* stream) {
  CMU_REQUIRE(sums_bn > 0 && n % sums_bn == 0 && sums_grid >= n / sums_bn, "moco_finish: bad partial layout");
  cudaStream_t st = (cudaStream_t)stream;
  moco_finish_kernel<<<ceil_div(n, 128), 128, 0, st>>>(sums_partial, sums_grid, sums_bn, n, lpos, 1.f / temperature,
                                                       loss_rows, ppos, row_scale);
  CMU_LAUNCH_CHECK();
  mean_scale_kernel<<<1, 32, 0, st>>>(loss_rows, n, 1.f, loss);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_moco_dq_scaled(const float* dq_neg_unnorm, const float* row_scale, const float* k, const float* qh,
                       const float* qnorm, const float* ppos, int n, int d, float temperature, float* dq, void* stream) {
  moco_dq_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(dq_neg_unnorm, k, qh, qnorm, ppos, d,
                                                      1.f / ((float)n * temperature), dq, row_scale);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_queue_enqueue(const float* keys, int n, int d, int k, int ptr, void* queue_rows, float* queue_ref, void* stream) {
  CMU_REQUIRE(ptr >= 0 && ptr < k, "queue_enqueue: bad pointer %d", ptr);
  queue_enqueue_kernel<<<grid_for((size_t)n * d, 256, 4), 256, 0, (cudaStream_t)stream>>>(keys, n, d, k, ptr,
                                                                                         (__nv_bfloat16*)queue_rows, queue_ref);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// =========================================================================================== soft-clDice metric
// Finetuning/metrics.py:401-492 (eval metric of Finetuning/train.py:464): soft skeletons of the thresholded prediction and
// of the target by min/max-pool morphology (10 iterations), then the clDice ratio.  float64 planes like the reference's
// float64 targets; HBM-bound, one 32x32 tile per CTA with a 3-pixel halo staged in shared memory.
namespace cmu {

__global__ void cldice_prep_kernel(const float* __restrict__ logits, const double* __restrict__ gt,
                                   double* __restrict__ planes, size_t n, size_t hw) {
  const size_t total = n * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t img = i / hw, p = i % hw;
    planes[i] = (logits[(img * 2 + 1) * hw + p] > logits[(img * 2) * hw + p]) ? 1.0 : 0.0;   // softmax_1 > 0.5
    planes[total + i] = gt[(img * 2 + 1) * hw + p];
  }
}

constexpr int kSkT = 32;
// one soft_skel step: e = kFirst ? in : erode(in); o = dilate(erode(e)); delta = relu(e - o);
// skel = kFirst ? delta : skel + relu(delta - skel * delta); out = e
template <bool kFirst>
__global__ void __launch_bounds__(256) skel_step_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                        double* __restrict__ skel, int H, int W) {
  __shared__ double s_in[kSkT + 6][kSkT + 7];
  __shared__ double s_e[kSkT + 4][kSkT + 5];
  __shared__ double s_ee[kSkT + 2][kSkT + 3];
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const size_t plane = (size_t)blockIdx.z * H * W;
  const int y0 = blockIdx.y * kSkT, x0 = blockIdx.x * kSkT;
  for (int i = threadIdx.x; i < (kSkT + 6) * (kSkT + 6); i += blockDim.x) {
    const int ly = i / (kSkT + 6), lx = i % (kSkT + 6);
    const int gy = y0 - 3 + ly, gx = x0 - 3 + lx;
    s_in[ly][lx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? in[plane + (size_t)gy * W + gx] : inf;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (kSkT + 4) * (kSkT + 4); i += blockDim.x) {
    const int ly = i / (kSkT + 4), lx = i % (kSkT + 4);
    const int gy = y0 - 2 + ly, gx = x0 - 2 + lx;
    double v = s_in[ly + 1][lx + 1];
    if (!kFirst) v = fmin(fmin(fmin(v, s_in[ly][lx + 1]), s_in[ly + 2][lx + 1]), fmin(s_in[ly + 1][lx], s_in[ly + 1][lx + 2]));
    s_e[ly][lx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? v : inf;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (kSkT + 2) * (kSkT + 2); i += blockDim.x) {
    const int ly = i / (kSkT + 2), lx = i % (kSkT + 2);
    const int gy = y0 - 1 + ly, gx = x0 - 1 + lx;
    const double v = fmin(fmin(fmin(s_e[ly + 1][lx + 1], s_e[ly][lx + 1]), s_e[ly + 2][lx + 1]),
                          fmin(s_e[ly + 1][lx], s_e[ly + 1][lx + 2]));
    s_ee[ly][lx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? v : -inf;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSkT * kSkT; i += blockDim.x) {
    const int ly = i / kSkT, lx = i % kSkT;
    const int gy = y0 + ly, gx = x0 + lx;
    if (gy >= H || gx >= W) continue;
    double o = -inf;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) o = fmax(o, s_ee[ly + dy][lx + dx]);
    const double e = s_e[ly + 2][lx + 2];
    const double delta = fmax(e - o, 0.0);
    const size_t idx = plane + (size_t)gy * W + gx;
    if (kFirst) {
      skel[idx] = delta;
    } else {
      const double sk = skel[idx];
      skel[idx] = sk + fmax(delta - sk * delta, 0.0);
      out[idx] = e;
    }
  }
}

// acc[0] = sum skel_pred * gt   acc[1] = sum skel_pred   acc[2] = sum skel_true * pred   acc[3] = sum skel_true
__global__ void __launch_bounds__(256) cldice_sums_kernel(const double* __restrict__ planes, const double* __restrict__ skel,
                                                          double* __restrict__ acc, size_t total) {
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const double sp = skel[i], st = skel[total + i];
    a0 += sp * planes[total + i];
    a1 += sp;
    a2 += st * planes[i];
    a3 += st;
  }
  __shared__ double sred[4][8];
  a0 = warp_sum_d(a0); a1 = warp_sum_d(a1); a2 = warp_sum_d(a2); a3 = warp_sum_d(a3);
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    sred[0][w] = a0; sred[1][w] = a1; sred[2][w] = a2; sred[3][w] = a3;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) a += sred[threadIdx.x][k];
    atomicAdd(&acc[threadIdx.x], a);
  }
}
__global__ void cldice_finish_kernel(const double* acc, double smooth, double* out) {
  const double tprec = (acc[0] + smooth) / (acc[1] + smooth);
  const double tsens = (acc[2] + smooth) / (acc[3] + smooth);
  out[0] = 1.0 - 2.0 * (tprec * tsens) / (tprec + tsens);
}

}  // namespace cmu

extern "C" {

long long cmu_soft_cldice_workspace_bytes(int n, int h, int w) {
  return (long long)(4 * 2 * (size_t)n * h * w + 8) * (long long)sizeof(double);
}

int cmu_soft_cldice(const float* logits, const double* gt, int n, int h, int w, int num_iter, double smooth, void* ws,
                    long long ws_bytes, double* out, void* stream) {
  using namespace cmu;
  CMU_REQUIRE(n > 0 && h > 0 && w > 0 && num_iter >= 0, "soft_cldice: bad shape");
  CMU_REQUIRE(ws != nullptr && ws_bytes >= cmu_soft_cldice_workspace_bytes(n, h, w), "soft_cldice: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)n * h * w;          // per set (prediction planes, then target planes)
  double* planes = (double*)ws;
  double* buf_a = planes + 2 * total;
  double* buf_b = buf_a + 2 * total;
  double* skel = buf_b + 2 * total;
  double* acc = skel + 2 * total;
  CMU_CHECK_CUDA(cudaMemsetAsync(acc, 0, 4 * sizeof(double), st));
  cldice_prep_kernel<<<grid_for(total, 256, 8), 256, 0, st>>>(logits, gt, planes, (size_t)n, (size_t)h * w);
  CMU_LAUNCH_CHECK();
  dim3 grid(ceil_div(w, kSkT), ceil_div(h, kSkT), 2 * n);
  skel_step_kernel<true><<<grid, 256, 0, st>>>(planes, nullptr, skel, h, w);
  CMU_LAUNCH_CHECK();
  const double* cur = planes;
  for (int j = 0; j < num_iter; ++j) {
    double* nxt = (j & 1) ? buf_b : buf_a;
    skel_step_kernel<false><<<grid, 256, 0, st>>>(cur, nxt, skel, h, w);
    CMU_LAUNCH_CHECK();
    cur = nxt;
  }
  cldice_sums_kernel<<<grid_for(total, 256, 8), 256, 0, st>>>(planes, skel, acc, total);
  CMU_LAUNCH_CHECK();
  cldice_finish_kernel<<<1, 1, 0, st>>>(acc, smooth, out);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
