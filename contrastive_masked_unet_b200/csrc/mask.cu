// Seeded integer patch-mask generator, bit-exact with the reference's host code
// (UNet_encoder.py:106-139 -> numpy legacy RandomState: MT19937 + masked-rejection Fisher-Yates shuffle).
//
// The random stream is inherently sequential (the number of 32-bit draws per shuffle is data dependent), so one
// warp owns the stream: its 32 lanes regenerate the 624-word MT19937 state cooperatively (three dependent spans of
// <= 227 words), lane 0 consumes tempered words through the rejection loop and performs the swaps in shared
// memory.  The permutation prefixes (K masked patches per image) are written to global memory and a second,
// grid-wide kernel rasterises them into the (B,S,S) uint8 mask with 16-byte stores.  Meant to run on a side
// stream one step ahead of the training step.
//
// Device state layout (uint32[625]): words [0,624) = MT19937 key, word 624 = position (624 = regenerate first),
// identical to numpy's `get_state()[1:3]`, so a stream can be handed over from / to numpy.
#include "common.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

constexpr int MT_N = 624, MT_M = 397;

__global__ void mt_seed_kernel(uint32_t* state, uint32_t seed) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint32_t prev = seed;
    state[0] = prev;
    for (int i = 1; i < MT_N; ++i) {
      prev = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i;
      state[i] = prev;
    }
    state[MT_N] = MT_N;
  }
}

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
  return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}
// warp-cooperative in-place regeneration of mt[624] in shared memory
__device__ void mt_regen(uint32_t* mt, uint32_t lane) {
  // new[k] = f(old[k], old[k+1], X[k+397 mod 624]) where X is old for k < 227 and new (index k-227) afterwards.
  // Spans [0,227), [227,454), [454,623) only depend on earlier spans; inside a span element k also reads old[k+1],
  // which lane-parallel execution could clobber -> read all inputs of a 32-wide batch before writing it.
  for (int base = 0; base < MT_N - 1; base += 32) {
    const int k = base + lane;
    uint32_t v = 0;
    const bool ok = k < MT_N - 1;
    if (ok) {
      const int src = (k < MT_N - MT_M) ? k + MT_M : k + MT_M - MT_N;
      v = mt_twist(mt[k], mt[k + 1], mt[src]);
    }
    __syncwarp();
    if (ok) mt[k] = v;
    __syncwarp();
  }
  if (lane == 0) mt[MT_N - 1] = mt_twist(mt[MT_N - 1], mt[0], mt[MT_M - 1]);
  __syncwarp();
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9D2C5680u;
  y ^= (y << 15) & 0xEFC60000u;
  y ^= y >> 18;
  return y;
}

// One warp.  n_shuffles consecutive shuffles of arange(P); for shuffle i < n_keep the first K entries of the
// permutation are written to perm_prefix[i][K]; the remaining shuffles only advance the stream (quirk Q2: the
// target encoder draws B shuffles although its mask is empty).
__global__ void __launch_bounds__(32) mask_shuffle_kernel(uint32_t* state, int* __restrict__ perm_prefix, int P, int K,
                                                          int n_shuffles, int n_keep) {
  extern __shared__ uint32_t sm[];
  uint32_t* mt = sm;                              // [624]
  int* perm = reinterpret_cast<int*>(sm + MT_N);  // [P]
  __shared__ int s_need_regen;
  const uint32_t lane = threadIdx.x;
  for (int i = lane; i < MT_N; i += 32) mt[i] = state[i];
  int pos = (int)state[MT_N];
  __syncwarp();
  for (int sh = 0; sh < n_shuffles; ++sh) {
    const bool keep = sh < n_keep;
    if (keep) {
      for (int i = lane; i < P; i += 32) perm[i] = i;
    }
    __syncwarp();
    int i = P - 1;
    // lane 0 consumes words until the state block is exhausted, then the warp regenerates and lane 0 resumes
    while (true) {
      if (lane == 0) {
        s_need_regen = 0;
        while (i >= 1) {
          uint32_t mask = (uint32_t)i;
          mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
          if (pos >= MT_N) { s_need_regen = 1; break; }
          const uint32_t v = mt_temper(mt[pos++]) & mask;
          if (v <= (uint32_t)i) {
            if (keep) {
              const int t = perm[i];
              perm[i] = perm[v];
              perm[v] = t;
            }
            --i;
          }
        }
      }
      __syncwarp();
      const int need = s_need_regen;
      __syncwarp();
      if (!need) break;
      mt_regen(mt, lane);
      pos = 0;
      i = __shfl_sync(0xffffffffu, i, 0);
    }
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (keep) {
      for (int k = lane; k < K; k += 32) perm_prefix[(size_t)sh * K + k] = perm[k];
    }
    __syncwarp();
  }
  for (int i = lane; i < MT_N; i += 32) state[i] = mt[i];
  if (lane == 0) state[MT_N] = (uint32_t)pos;
}

// mask[b, r*ps .. , c*ps ..] = 1 for the K selected patches; the buffer must be zero-filled by `mask_clear`.
__global__ void mask_raster_kernel(const int* __restrict__ perm_prefix, uint8_t* __restrict__ mask, int B, int S, int ps,
                                   int K) {
  const int g = S / ps;
  const int vec_per_row = ps / 16;  // 16-byte stores when ps % 16 == 0
  const size_t total = (size_t)B * K * ps * (vec_per_row > 0 ? vec_per_row : ps);
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    if (vec_per_row > 0) {
      const int v = t % vec_per_row;
      const int y = (t / vec_per_row) % ps;
      const size_t bk = t / ((size_t)vec_per_row * ps);
      const int idx = perm_prefix[bk];
      const size_t b = bk / K;
      const int row = (idx / g) * ps + y, col = (idx % g) * ps + v * 16;
      *reinterpret_cast<uint4*>(mask + (b * S + row) * S + col) = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    } else {
      const int x = t % ps;
      const int y = (t / ps) % ps;
      const size_t bk = t / ((size_t)ps * ps);
      const int idx = perm_prefix[bk];
      const size_t b = bk / K;
      mask[(b * S + (idx / g) * ps + y) * S + (idx % g) * ps + x] = 1;
    }
  }
}

}  // namespace cmu

using namespace cmu;

extern "C" {

int cmu_mask_state_words(void) { return MT_N + 1; }

int cmu_mask_seed(unsigned int* d_state, unsigned int seed, void* stream) {
  mt_seed_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_state, seed);
  CMU_LAUNCH_CHECK();
  return 0;
}

// Advances the stream by `n_shuffles` shuffles of arange((S/ps)^2); the first `batch` of them define `mask`.
// perm_ws: int32[batch * K] workspace, K = masked patches per image (0 -> mask stays all-zero).
int cmu_mask_generate(unsigned int* d_state, unsigned char* mask, int* perm_ws, int batch, int img_size, int patch_size,
                      int k_masked, int n_shuffles, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CMU_REQUIRE(img_size % patch_size == 0, "mask: img_size %% patch_size != 0");
  const int g = img_size / patch_size;
  const int P = g * g;
  CMU_REQUIRE(k_masked >= 0 && k_masked <= P, "mask: bad K");
  CMU_REQUIRE(n_shuffles >= 0 && batch >= 0, "mask: bad counts");
  const size_t shmem = (MT_N + (size_t)P) * 4;
  CMU_REQUIRE(shmem <= 200 * 1024, "mask: too many patches (%d)", P);
  const int n_keep = (k_masked > 0 && mask != nullptr) ? (batch < n_shuffles ? batch : n_shuffles) : 0;
  if (shmem > 48 * 1024) {
    CMU_CHECK_CUDA(cudaFuncSetAttribute(mask_shuffle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
  }
  if (n_shuffles > 0) {
    mask_shuffle_kernel<<<1, 32, shmem, st>>>(d_state, perm_ws, P, k_masked, n_shuffles, n_keep);
    CMU_LAUNCH_CHECK();
  }
  if (mask != nullptr) {
    CMU_CHECK_CUDA(cudaMemsetAsync(mask, 0, (size_t)batch * img_size * img_size, st));
    if (n_keep > 0) {
      CMU_REQUIRE(patch_size % 16 != 0 || ((reinterpret_cast<uintptr_t>(mask) & 15) == 0 && img_size % 16 == 0),
                  "mask: buffer must be 16-byte aligned");
      const size_t total = (size_t)n_keep * k_masked * patch_size * (patch_size % 16 == 0 ? patch_size / 16 : patch_size);
      size_t blocks = (total + 255) / 256;
      if (blocks > (size_t)num_sms() * 8) blocks = (size_t)num_sms() * 8;
      mask_raster_kernel<<<(int)blocks, 256, 0, st>>>(perm_ws, mask, n_keep, img_size, patch_size, k_masked);
      CMU_LAUNCH_CHECK();
    }
  }
  return 0;
}

}  // extern "C"
