// Multi-tensor parameter kernels (HBM-bound): EMA of the target networks (cmunet.py:78-92) and AdamW
// (cmunet_config.py:76-91).  The host passes a device table of chunks so that one launch covers every tensor.
//   EMA   row: {dst, src, count}                       theta_t = theta_t*m + theta_o*(1-m)   (that order, fp32)
//   AdamW row: {param, grad, exp_avg, exp_avg_sq, count, decay_flag}
#include <string.h>

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

// ---- dynamic loss scaling (mmengine AmpOptimWrapper(loss_scale='dynamic') == torch.amp.GradScaler; cmunet_config.py:76-78)
// Device-resident state so that a step never synchronises with the host:
//   amp[0] float  loss scale            amp[1] int  found_inf of this step     amp[2] int  successful optimizer steps
//   amp[3] int    growth tracker        amp[4] float 1 / (scale the gradients of this step were produced with)
__global__ void __launch_bounds__(256) found_inf_chunks_kernel(const long long* __restrict__ table, int* __restrict__ amp) {
  const long long* row = table + (size_t)blockIdx.x * 6;
  const float* g = reinterpret_cast<const float*>(row[1]);
  const int n = (int)row[4];
  bool bad = false;
  for (int i = threadIdx.x; i < n; i += blockDim.x) bad |= !isfinite(g[i]);
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicExch(&amp[1], 1);
}

__global__ void amp_advance_kernel(int* __restrict__ amp, float growth, float backoff, int interval) {
  float scale = __int_as_float(amp[0]);
  amp[4] = __float_as_int(1.f / scale);
  if (amp[1]) {                  // overflow: skip the step, shrink the scale (GradScaler.update)
    scale *= backoff;
    amp[3] = 0;
  } else {
    amp[2] += 1;
    if (++amp[3] >= interval) {
      scale *= growth;
      amp[3] = 0;
    }
  }
  amp[0] = __float_as_int(scale);
}

__global__ void __launch_bounds__(256) adamw_chunks_amp_kernel(const long long* __restrict__ table, float lr, float beta1,
                                                               float beta2, float eps, float weight_decay,
                                                               const int* __restrict__ amp) {
  if (amp[1]) return;            // found_inf: parameters and moments stay untouched
  const float grad_scale = __int_as_float(amp[4]);
  const float step = (float)amp[2];
  const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
  const long long* row = table + (size_t)blockIdx.x * 6;
  float* p = reinterpret_cast<float*>(row[0]);
  const float* g = reinterpret_cast<const float*>(row[1]);
  float* m = reinterpret_cast<float*>(row[2]);
  float* v = reinterpret_cast<float*>(row[3]);
  const int n = (int)row[4];
  const float wd = row[5] ? weight_decay : 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

// SGD with momentum, torch.optim.SGD semantics (MoCo-v2: moco2_module.py configure_optimizers, lr 0.03, momentum 0.9,
// weight decay 1e-4).  row: {param, grad, momentum_buffer, count}
__global__ void __launch_bounds__(256) sgd_chunks_kernel(const long long* __restrict__ table, float lr, float momentum,
                                                         float weight_decay, int first_step) {
  const long long* row = table + (size_t)blockIdx.x * 4;
  float* p = reinterpret_cast<float*>(row[0]);
  const float* g = reinterpret_cast<const float*>(row[1]);
  float* buf = reinterpret_cast<float*>(row[2]);
  const int n = (int)row[3];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float pi = p[i];
    const float gi = g[i] + weight_decay * pi;
    const float bi = first_step ? gi : momentum * buf[i] + gi;
    buf[i] = bi;
    p[i] = pi - lr * bi;
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

int cmu_amp_init(int* d_amp, float init_scale, void* stream) {
  const float inv = 1.f / init_scale;
  int h[5];
  memcpy(&h[0], &init_scale, 4);
  h[1] = 0;
  h[2] = 0;
  h[3] = 0;
  memcpy(&h[4], &inv, 4);
  CMU_CHECK_CUDA(cudaMemcpyAsync(d_amp, h, sizeof(h), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  CMU_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));   // `h` is a stack buffer (one-off initialisation)
  return 0;
}

// One optimizer step under dynamic loss scaling, no host synchronisation: found-inf reduction over every gradient,
// scale / step-counter update, AdamW that unscales by the old scale and is skipped entirely after an overflow.
int cmu_adamw_chunks_amp(const long long* d_table, int n_chunks, float lr, float beta1, float beta2, float eps,
                         float weight_decay, int* d_amp, float growth, float backoff, int growth_interval, void* stream) {
  if (n_chunks <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CMU_CHECK_CUDA(cudaMemsetAsync(d_amp + 1, 0, sizeof(int), st));
  found_inf_chunks_kernel<<<n_chunks, 256, 0, st>>>(d_table, d_amp);
  CMU_LAUNCH_CHECK();
  amp_advance_kernel<<<1, 1, 0, st>>>(d_amp, growth, backoff, growth_interval);
  CMU_LAUNCH_CHECK();
  adamw_chunks_amp_kernel<<<n_chunks, 256, 0, st>>>(d_table, lr, beta1, beta2, eps, weight_decay, d_amp);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_sgd_chunks(const long long* d_table, int n_chunks, float lr, float momentum, float weight_decay, int first_step,
                   void* stream) {
  if (n_chunks <= 0) return 0;
  sgd_chunks_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(d_table, lr, momentum, weight_decay, first_step);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
