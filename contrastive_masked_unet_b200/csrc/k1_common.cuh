// Shared definitions of the K1 implicit-GEMM kernels (tc_conv.cu: one CTA per tile; tc_conv2.cu: CTA pairs,
// tcgen05 cta_group::2).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace cmu {

enum { MODE_CONV3 = 0, MODE_PLAIN = 1, MODE_CONVT_FPROP = 2, MODE_CONVT_DGRAD = 3 };

constexpr int kK1Threads = 192;
constexpr int kStagingBytes = 16384;  // one 128 x 64 bf16 output slab
constexpr int kMaxStages = 8;
constexpr int kSchedDepth = 4;
constexpr int kSmemLimit = 232448;    // 227 KB per CTA

struct K1Params {
  CUtensorMap tmA0, tmA1, tmB, tmO0, tmO1;
  int mode;
  int N, H, W;      // pixel space of GEMM-M (output pixels; for convT: the low-resolution grid)
  int c0, c1;       // channels of A source 0 / 1 (concat along K)
  int n_total;      // GEMM N
  int oc0;          // channels of output 0 (dual-output split; convT fprop: Cout)
  int TW, TH, tw_shift;
  int tiles_w, tiles_h, m_tiles, n_tiles;
  int kc;           // number of 64-wide K chunks per tap
  // shared-memory plan (host-computed): [n_stages x stage_bytes][stg_bufs x 16 KB staging][barriers][stats]
  int a_bytes, b_bytes, b_off, stage_bytes, n_stages, stg_bufs;
  int w_resident;   // 1: this CTA's whole weight slab (all taps x K chunks of its n-tile) is loaded ONCE into shared
  int w_bytes;      //    memory and the pipeline streams activations only (small-weight layers)
  unsigned int* sched;  // [n_tiles] m-tile counters, zeroed before the launch (dynamic tile scheduler)
  const float* bias;
  int bias_mod;
  float* stats;     // [gridDim.x][2][BN] partial (sum, sum of squares) or nullptr
};

template <int OFF>
__device__ __forceinline__ void bfly(float (&v)[32], uint32_t lane) {
  const bool up = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < OFF; ++i) {
    const float send = up ? v[i] : v[i + OFF];
    const float keep = up ? v[i + OFF] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}
// After the call, lane L holds in v[0] the sum over the 32 lanes of their v[L].
__device__ __forceinline__ void column_sums(float (&v)[32], uint32_t lane) {
  bfly<16>(v, lane);
  bfly<8>(v, lane);
  bfly<4>(v, lane);
  bfly<2>(v, lane);
  bfly<1>(v, lane);
}


// pair kernel (tc_conv2.cu): returns 0 on success; *used = 1 if the launch was taken by the pair kernel
int run_k1_pair(K1Params& p, const void* wpk, int ktot, cudaStream_t stream, int* used_grid, int* used_bn);

}  // namespace cmu
