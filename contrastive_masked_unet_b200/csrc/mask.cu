// Seeded integer patch-mask generator, bit-exact with the reference's host code
// (UNet_encoder.py:106-139 -> numpy legacy RandomState: MT19937 + masked-rejection Fisher-Yates shuffle).
//
// The random stream is sequential (the number of 32-bit draws per Fisher-Yates step is data dependent), but the work
// splits into two chains that need not run in lock-step:
//   1. draw resolution (mask_draw_kernel, ONE warp): which word of the MT19937 stream does step i of which shuffle
//      consume, and what index v_i does it yield?  The warp looks at 32 tempered words at a time.  Lane l's word is
//      accepted iff (w_l & mask(i_l)) <= i_l with i_l = i - (#accepted among lanes < l): a recurrence over lanes that
//      is solved by fixed-point iteration with ballots -- lane 0 is exact after the first pass and every pass fixes at
//      least one more lane, in practice 2-3 passes resolve all 32 words (= ~23 Fisher-Yates steps).  The 624-word
//      state is regenerated cooperatively (three dependent spans).  v_i goes to global memory as uint16.
//   2. swap application (mask_apply_kernel, one THREAD per kept shuffle, shuffles are independent once the v_i are
//      known): perm[i] <-> perm[v_i] for i = P-1..1 in shared memory, then the K-entry prefix is written out.
// A grid-wide kernel rasterises the prefixes into the (B,S,S) uint8 mask with 16-byte stores.
// 128 shuffles of arange(1024) (one B = 64 step: online + discarded target draws, quirk Q2): ~0.5 ms, down from 14 ms
// for the round-1 kernel that walked the rejection loop and the swaps on one lane.
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

// Chain 1.  One warp; n_shuffles consecutive shuffles of arange(P).  For shuffle sh < n_keep the accepted index of
// step i is written to vs[sh][i]; the remaining shuffles only advance the stream (quirk Q2: the target encoder draws
// B shuffles although its mask is empty).
__global__ void __launch_bounds__(32) mask_draw_kernel(uint32_t* state, uint16_t* __restrict__ vs, int P, int n_shuffles,
                                                       int n_keep) {
  __shared__ uint32_t mt[MT_N];
  const uint32_t lane = threadIdx.x;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (int i = lane; i < MT_N; i += 32) mt[i] = state[i];
  int pos = (int)state[MT_N];   // warp-uniform
  __syncwarp();
  for (int sh = 0; sh < n_shuffles; ++sh) {
    const bool keep = sh < n_keep;
    uint16_t* v = vs + (size_t)sh * P;
    int i = P - 1;              // warp-uniform: the Fisher-Yates step the next accepted word serves
    while (i >= 1) {
      if (pos >= MT_N) {
        mt_regen(mt, lane);
        pos = 0;
      }
      const int idx = pos + (int)lane;
      const bool valid = idx < MT_N;
      const uint32_t w = valid ? mt_temper(mt[idx]) : 0u;
      // fixed point of: ok_l = valid_l && i_l >= 1 && (w_l & mask(i_l)) <= i_l,  i_l = i - popc(ok & lanes < l)
      int a = 0, il = i;
      uint32_t x = 0, acc = 0, prev = 0;
      bool first = true;
      while (true) {
        il = i - a;
        const uint32_t m = il >= 1 ? (0xffffffffu >> __clz(il)) : 0u;   // smallest 2^k - 1 >= il (rk_interval)
        x = w & m;
        const bool ok = valid && il >= 1 && x <= (uint32_t)il;
        acc = __ballot_sync(0xffffffffu, ok);
        a = __popc(acc & lt_mask);
        if (!first && acc == prev) break;
        prev = acc;
        first = false;
      }
      const bool mine = (acc >> lane) & 1u;
      if (mine && keep) v[il] = (uint16_t)x;
      // the shuffle ends at the lane that served step 1: words behind it belong to the next shuffle
      const uint32_t fin = __ballot_sync(0xffffffffu, mine && il == 1);
      const int nvalid = (MT_N - pos) < 32 ? (MT_N - pos) : 32;
      pos += fin ? __ffs((int)fin) : nvalid;
      i -= __popc(acc);
    }
  }
  __syncwarp();
  for (int i = lane; i < MT_N; i += 32) state[i] = mt[i];
  if (lane == 0) state[MT_N] = (uint32_t)pos;
}

// Chain 2.  Thread t of block b applies the swaps of shuffle b*T + t to its own column of perm[P][T] (uint16, shared
// memory) and writes the first K entries of the permutation.
__global__ void mask_apply_kernel(const uint16_t* __restrict__ vs, int* __restrict__ perm_prefix, int P, int K,
                                  int n_keep, int T) {
  extern __shared__ uint16_t perm[];
  const int t = threadIdx.x;
  const int sh = blockIdx.x * T + t;
  if (sh >= n_keep) return;
  for (int i = 0; i < P; ++i) perm[i * T + t] = (uint16_t)i;
  const uint16_t* v = vs + (size_t)sh * P;
  constexpr int kChunk = 16;
  uint16_t cur[kChunk], nxt[kChunk];
#pragma unroll
  for (int u = 0; u < kChunk; ++u) {
    const int i = P - 1 - u;
    cur[u] = i >= 1 ? v[i] : (uint16_t)0;
  }
  for (int base = P - 1; base >= 1; base -= kChunk) {
#pragma unroll
    for (int u = 0; u < kChunk; ++u) {   // indices of the next chunk are in flight while this one is applied
      const int i = base - kChunk - u;
      nxt[u] = i >= 1 ? v[i] : (uint16_t)0;
    }
#pragma unroll
    for (int u = 0; u < kChunk; ++u) {
      const int i = base - u;
      if (i >= 1) {
        const int j = cur[u];
        const uint16_t pi = perm[i * T + t];
        perm[i * T + t] = perm[j * T + t];
        perm[j * T + t] = pi;
      }
    }
#pragma unroll
    for (int u = 0; u < kChunk; ++u) cur[u] = nxt[u];
  }
  for (int k = 0; k < K; ++k) perm_prefix[(size_t)sh * K + k] = perm[k * T + t];
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

// Workspace of cmu_mask_generate: int32 perm_prefix[batch][K] followed by uint16 vs[batch][P] (16-byte aligned parts).
long long cmu_mask_workspace_bytes(int batch, int n_patches, int k_masked) {
  const long long a = (((long long)batch * k_masked * 4) + 15) & ~15LL;
  const long long b = (((long long)batch * n_patches * 2) + 15) & ~15LL;
  return a + b + 16;
}

// Advances the stream by `n_shuffles` shuffles of arange((S/ps)^2); the first `batch` of them define `mask`.
// perm_ws: cmu_mask_workspace_bytes(batch, P, K) bytes, K = masked patches per image (0 -> mask stays all-zero);
// mask may be NULL (the stream only advances).
int cmu_mask_generate(unsigned int* d_state, unsigned char* mask, int* perm_ws, int batch, int img_size, int patch_size,
                      int k_masked, int n_shuffles, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  CMU_REQUIRE(patch_size > 0 && img_size % patch_size == 0, "mask: img_size %% patch_size != 0");
  const int g = img_size / patch_size;
  const int P = g * g;
  CMU_REQUIRE(P <= 65536, "mask: too many patches (%d)", P);
  CMU_REQUIRE(k_masked >= 0 && k_masked <= P, "mask: bad K");
  CMU_REQUIRE(n_shuffles >= 0 && batch >= 0, "mask: bad counts");
  const int n_keep = (k_masked > 0 && mask != nullptr) ? (batch < n_shuffles ? batch : n_shuffles) : 0;
  CMU_REQUIRE(n_keep == 0 || perm_ws != nullptr, "mask: workspace missing");
  int* perm_prefix = perm_ws;
  uint16_t* vs = perm_ws == nullptr ? nullptr : reinterpret_cast<uint16_t*>(
      reinterpret_cast<uint8_t*>(perm_ws) + ((((size_t)batch * k_masked * 4) + 15) & ~(size_t)15));
  if (n_shuffles > 0) {
    mask_draw_kernel<<<1, 32, 0, st>>>(d_state, vs, P, n_shuffles, n_keep);
    CMU_LAUNCH_CHECK();
  }
  if (n_keep > 0) {
    int T = 32;
    while (T > 1 && (size_t)P * T * 2 > 160 * 1024) T >>= 1;
    const size_t shmem = (size_t)P * T * 2;
    if (shmem > 48 * 1024) {
      CMU_CHECK_CUDA(cudaFuncSetAttribute(mask_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    }
    mask_apply_kernel<<<(n_keep + T - 1) / T, T, shmem, st>>>(vs, perm_prefix, P, k_masked, n_keep, T);
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
      mask_raster_kernel<<<(int)blocks, 256, 0, st>>>(perm_prefix, mask, n_keep, img_size, patch_size, k_masked);
      CMU_LAUNCH_CHECK();
    }
  }
  return 0;
}

}  // extern "C"
