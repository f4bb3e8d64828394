// Multi-tensor parameter kernels (HBM-bound): EMA of the target networks (cmunet.py:78-92) and AdamW
// (cmunet_config.py:76-91).  The host passes a device table of chunks so that one launch covers every tensor.
//   EMA   row: {dst, src, count}                       theta_t = theta_t*m + theta_o*(1-m)   (that order, fp32)
//   AdamW row: {param, grad, exp_avg, exp_avg_sq, count, decay_flag}
#include "common.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

__global__ void __launch_bounds__(256) ema_chunks_kernel(const long long* __restrict__ table, float m) {
  const long long* row = table + (size_t)blockIdx.x * 3;
  float* dst = reinterpret_cast<float*>(row[0]);
  const float* src = reinterpret_cast<const float*>(row[1]);
  const int n = (int)row[2];
  const float om = 1.f - m;
  const bool vec = ((row[0] | row[1]) & 15) == 0;
  if (vec) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 d = reinterpret_cast<float4*>(dst)[i];
      const float4 s = reinterpret_cast<const float4*>(src)[i];
      d.x = d.x * m + s.x * om; d.y = d.y * m + s.y * om; d.z = d.z * m + s.z * om; d.w = d.w * m + s.w * om;
      reinterpret_cast<float4*>(dst)[i] = d;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) dst[i] = dst[i] * m + src[i] * om;
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = dst[i] * m + src[i] * om;
  }
}

__global__ void __launch_bounds__(256) adamw_chunks_kernel(const long long* __restrict__ table, float lr, float beta1,
                                                           float beta2, float eps, float weight_decay, float bc1, float bc2,
                                                           float grad_scale) {
  const long long* row = table + (size_t)blockIdx.x * 6;
  float* p = reinterpret_cast<float*>(row[0]);
  const float* g = reinterpret_cast<const float*>(row[1]);
  float* m = reinterpret_cast<float*>(row[2]);
  float* v = reinterpret_cast<float*>(row[3]);
  const int n = (int)row[4];
  const float wd = row[5] ? weight_decay : 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);               // decoupled weight decay (torch.optim.AdamW)
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

}  // namespace cmu

using namespace cmu;

extern "C" {

int cmu_ema_chunks(const long long* d_table, int n_chunks, float momentum, void* stream) {
  if (n_chunks <= 0) return 0;
  ema_chunks_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(d_table, momentum);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_adamw_chunks(const long long* d_table, int n_chunks, float lr, float beta1, float beta2, float eps,
                     float weight_decay, int step, float grad_scale, void* stream) {
  if (n_chunks <= 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adamw_chunks_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(d_table, lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                                                                 grad_scale);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
