// Projection-head kernels (NonLinearNeck: fc0 -> (Sync)BN -> ReLU -> fc1).  The batch is tiny (64 rows per GPU), so
// fc0 (in = S*S up to 262144) is bound by streaming its fp32 weight once per pass; a split-K SIMT SGEMM with
// 64x64x16 tiles keeps that stream coalesced.  (A tcgen05 version would not change the HBM bound.)
// BatchNorm1d statistics are split in two kernels so that the host can all-reduce the (sum, sumsq) pair between them
// (SyncBatchNorm semantics across data-parallel ranks).
#include "common.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

// C[m][n] (+)= sum_k A(m,k) * B(n,k)   with A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]
// grid = (ceil(N/64), ceil(M/64), splits); split z handles k in [z*kper, min(K,(z+1)*kper)) and writes/accumulates into
// C + z*c_split_stride (c_split_stride = 0 with atomic=1 accumulates all splits into C with float atomics).
constexpr int TM = 64, TN = 64, TK = 16;
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, long long sam, long long sak,
                                                    const float* __restrict__ B, long long sbn, long long sbk,
                                                    float* __restrict__ C, long long ldc, long long c_split_stride, int M,
                                                    int N, int K, int kper, int accumulate, int atomic) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kbeg = blockIdx.z * kper;
  const int kend = min(K, kbeg + kper);
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // load mapping: 1024 elements per tile, 4 per thread.  If k is the contiguous dim, let consecutive threads walk k.
  const bool a_kfast = (sak == 1), b_kfast = (sbk == 1);
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int mm, kk;
      if (a_kfast) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < kend) ? A[gm * sam + gk * sak] : 0.f;
      int nn, kb;
      if (b_kfast) { kb = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kb = idx >> 6; }
      const int gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < N && gkb < kend) ? B[gn * sbn + gkb * sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* Cz = C + (size_t)blockIdx.z * c_split_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* o = Cz + (size_t)gm * ldc + gn;
      if (atomic) atomicAdd(o, acc[i][j]);
      else *o = accumulate ? *o + acc[i][j] : acc[i][j];
    }
  }
}
// out[m][n] = sum_z part[z][m][n] + bias[n]
__global__ void splitk_reduce_bias_kernel(const float* __restrict__ part, int splits, size_t mn, int N,
                                          const float* __restrict__ bias, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < mn; i += (size_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int z = 0; z < splits; ++z) a += part[(size_t)z * mn + i];
    out[i] = a + (bias ? bias[i % N] : 0.f);
  }
}
// column sums: out[n] (+)= sum_m x[m][n]
__global__ void colsum_kernel(const float* __restrict__ x, int M, int N, float* __restrict__ out, int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float a = 0.f;
  for (int m = 0; m < M; ++m) a += x[(size_t)m * N + n];
  out[n] = accumulate ? out[n] + a : a;
}

// ------------------------------------------------------------------------------------------ BatchNorm1d (+ReLU)
// stats[0][c] = sum_m x, stats[1][c] = sum_m x^2   (local rows; all-reduced by the host when world > 1)
__global__ void bn1d_stats_kernel(const float* __restrict__ x, int M, int C, float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s1 = 0.f, s2 = 0.f;
  for (int m = 0; m < M; ++m) {
    const float v = x[(size_t)m * C + c];
    s1 += v;
    s2 += v * v;
  }
  stats[c] = s1;
  stats[C + c] = s2;
}
// y = relu?(gamma*(x-mean)*rstd + beta); saves xhat-ready (mean, rstd); updates running stats with the GLOBAL count
__global__ void bn1d_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, double count, int M, int C,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                                  float* running_var, float momentum, float eps, int training, int relu,
                                  float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, rstd;
  if (training) {
    const double m = stats[c] / count;
    double var = stats[C + c] / count - m * m;
    if (var < 0) var = 0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(count > 1 ? var * count / (count - 1) : var);
    }
  } else {
    mean = running_mean[c];
    rstd = rsqrtf(running_var[c] + eps);
  }
  const float g = gamma[c] * rstd, b = beta[c] - mean * gamma[c] * rstd;
  for (int m = 0; m < M; ++m) {
    float v = fmaf(x[(size_t)m * C + c], g, b);
    if (relu) v = fmaxf(v, 0.f);
    y[(size_t)m * C + c] = v;
  }
  if (mean_out) { mean_out[c] = mean; rstd_out[c] = rstd; }
}
// dz = dy * [y > 0] (if relu);  sums[0][c] = sum dz, sums[1][c] = sum dz*xhat (local; all-reduced by the host)
__global__ void bn1d_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                                      const float* __restrict__ mean, const float* __restrict__ rstd, int M, int C, int relu,
                                      float* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s1 = 0.f, s2 = 0.f;
  const float mu = mean[c], rs = rstd[c];
  for (int m = 0; m < M; ++m) {
    const size_t i = (size_t)m * C + c;
    const float dz = (relu && y[i] <= 0.f) ? 0.f : dy[i];
    s1 += dz;
    s2 += dz * (x[i] - mu) * rs;
  }
  sums[c] = s1;
  sums[C + c] = s2;
}
__global__ void bn1d_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                      const float* __restrict__ gamma, const float* __restrict__ sums, double count, int M,
                                      int C, int relu, float* __restrict__ dx) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mu = mean[c], rs = rstd[c], g = gamma[c] * rs;
  const float k1 = (float)(sums[c] / count), k2 = (float)(sums[C + c] / count);
  for (int m = 0; m < M; ++m) {
    const size_t i = (size_t)m * C + c;
    const float dz = (relu && y[i] <= 0.f) ? 0.f : dy[i];
    dx[i] = g * (dz - k1 - (x[i] - mu) * rs * k2);
  }
}

}  // namespace cmu

using namespace cmu;

extern "C" {

// split-K factor: the projector / predictor GEMMs have 1-24 output tiles, so K is spread over the idle SMs (>= 64 K per
// split); with 4 CTAs the 64 x 1536 x 256 layer took 0.23 ms of pure latency
static int sgemm_splits(int m, int n, int k) {
  const int tiles = ceil_div(m, TM) * ceil_div(n, TN);
  if (k < 512 || tiles >= num_sms()) return 1;
  int splits = num_sms() / tiles;
  if (splits > k / (4 * TK)) splits = k / (4 * TK);
  return splits < 1 ? 1 : splits;
}

long long cmu_sgemm_workspace_bytes(int m, int n, int k) {
  return (long long)sgemm_splits(m, n, k) * m * n * 4;
}

// C[m][n] = sum_k A(m,k) B(n,k) (+ bias[n]) with arbitrary element strides; split-K when the output is small and K large.
int cmu_sgemm(const float* a, long long sam, long long sak, const float* b, long long sbn, long long sbk, float* c,
              long long ldc, const float* bias, int m, int n, int k, int accumulate, float* workspace,
              long long workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int splits = (accumulate || ldc != n) ? 1 : sgemm_splits(m, n, k);
  if (splits > 1 && !(workspace && workspace_bytes >= (long long)splits * m * n * 4)) {
    CMU_REQUIRE(k < 4096, "sgemm: workspace too small");   // large K needs the split; small K may run unsplit
    splits = 1;
  }
  if (splits > 1) {
    int kper = ceil_div(ceil_div(k, splits), TK) * TK;
    splits = ceil_div(k, kper);
    dim3 grid(ceil_div(n, TN), ceil_div(m, TM), splits);
    sgemm_kernel<<<grid, 256, 0, st>>>(a, sam, sak, b, sbn, sbk, workspace, n, (long long)m * n, m, n, k, kper, 0, 0);
    CMU_LAUNCH_CHECK();
    const size_t mn = (size_t)m * n;
    splitk_reduce_bias_kernel<<<(int)std::min<size_t>((mn + 255) / 256, 1024), 256, 0, st>>>(workspace, splits, mn, n, bias, c);
    CMU_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid(ceil_div(n, TN), ceil_div(m, TM), 1);
  sgemm_kernel<<<grid, 256, 0, st>>>(a, sam, sak, b, sbn, sbk, c, ldc, 0, m, n, k, k, accumulate, 0);
  CMU_LAUNCH_CHECK();
  if (bias != nullptr) {
    CMU_REQUIRE(ldc == n && !accumulate, "sgemm: bias needs a dense, non-accumulating C");
    const size_t mn = (size_t)m * n;
    splitk_reduce_bias_kernel<<<(int)std::min<size_t>((mn + 255) / 256, 1024), 256, 0, st>>>(c, 1, mn, n, bias, c);
    CMU_LAUNCH_CHECK();
  }
  return 0;
}

int cmu_colsum(const float* x, int m, int n, float* out, int accumulate, void* stream) {
  colsum_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(x, m, n, out, accumulate);
  CMU_LAUNCH_CHECK();
  return 0;
}

int cmu_bn1d_stats(const float* x, int m, int c, float* stats, void* stream) {
  bn1d_stats_kernel<<<ceil_div(c, 128), 128, 0, (cudaStream_t)stream>>>(x, m, c, stats);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_bn1d_apply(const float* x, const float* stats, double count, int m, int c, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float momentum, float eps, int training, int relu, float* y,
                   float* mean, float* rstd, void* stream) {
  bn1d_apply_kernel<<<ceil_div(c, 128), 128, 0, (cudaStream_t)stream>>>(x, stats, count, m, c, gamma, beta, running_mean,
                                                                       running_var, momentum, eps, training, relu, y, mean,
                                                                       rstd);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_bn1d_bwd_stats(const float* dy, const float* y, const float* x, const float* mean, const float* rstd, int m, int c,
                       int relu, float* sums, void* stream) {
  bn1d_bwd_stats_kernel<<<ceil_div(c, 128), 128, 0, (cudaStream_t)stream>>>(dy, y, x, mean, rstd, m, c, relu, sums);
  CMU_LAUNCH_CHECK();
  return 0;
}
int cmu_bn1d_bwd_apply(const float* dy, const float* y, const float* x, const float* mean, const float* rstd,
                       const float* gamma, const float* sums, double count, int m, int c, int relu, float* dx,
                       void* stream) {
  bn1d_bwd_apply_kernel<<<ceil_div(c, 128), 128, 0, (cudaStream_t)stream>>>(dy, y, x, mean, rstd, gamma, sums, count, m, c,
                                                                           relu, dx);
  CMU_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
