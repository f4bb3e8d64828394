// Shared definitions of the K1 implicit-GEMM kernels (tc_conv.cu: one CTA per tile; tc_conv2.cu: CTA pairs,
// tcgen05 cta_group::2).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace cmu {

enum { MODE_CONV3 = 0, MODE_PLAIN = 1, MODE_CONVT_FPROP = 2, MODE_CONVT_DGRAD = 3 };

constexpr int kK1Threads = 320;   // TMA warp, MMA warp, 2 groups of 4 epilogue warps (one group per TMEM accumulator)
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
  int a_bytes, b_bytes, b_off, stage_bytes, n_stages, stg_bufs;   // stg_bufs: staging buffers PER epilogue group
  int epi_groups;   // 1: warps 2-5 drain both accumulators; 2: warps 2-5 drain accumulator 0, warps 6-9 accumulator 1
  int w_resident;   // 1: this CTA's whole weight slab (all taps x K chunks of its n-tile) is loaded ONCE into shared
  int w_bytes;      //    memory and the pipeline streams activations only (small-weight layers)
  unsigned int* sched;  // [n_tiles] m-tile counters, zeroed before the launch (dynamic tile scheduler)
  const float* bias;
  int bias_mod;
  float* stats;     // [gridDim.x][2][BN] partial (sum, sum of squares) or nullptr
  int single_patch; // conv3 pair kernel: ONE haloed (TH+2) x (TW+2) patch per K chunk, the 9 taps are descriptor offsets
  float exp_scale;  // != 0: epilogue stores exp((acc - 1) * exp_scale) (MoCo queue logits -> unnormalised softmax terms)
};

// Shared memory the K1 kernels plan with.  Knob 14 = KB left free per SM so that blocks of the HBM-bound BatchNorm /
// element-wise kernels of ANOTHER stream can be co-resident with a persistent tensor-core CTA (A/B).
static inline int smem_budget() { return kSmemLimit - 1024 * debug_knob(14); }

// staging / epilogue-group plan: two groups with two 16 KB buffers each when the pipeline keeps >= 4 stages, then two
// groups with one buffer, else the single-group plans.  (Measured: the second group only helps ConvTranspose fprop,
// +20 %; the other short-K layers are bound by shared-memory bandwidth, profiles/r1_step_breakdown_v5.md.)
static inline void plan_epilogue(int avail, int stage_bytes, int force_single, int* epi_groups, int* stg_bufs) {
  if (!force_single && (avail - 4 * kStagingBytes) / stage_bytes >= 4) { *epi_groups = 2; *stg_bufs = 2; }
  else if (!force_single && (avail - 2 * kStagingBytes) / stage_bytes >= 5) { *epi_groups = 2; *stg_bufs = 1; }
  else if ((avail - 2 * kStagingBytes) / stage_bytes >= 4) { *epi_groups = 1; *stg_bufs = 2; }
  else { *epi_groups = 1; *stg_bufs = 1; }
}

// BatchNorm statistics of one staged output slab (128 pixel rows x 64 channels, bf16, SWIZZLE_128B rows of 128 B):
// each warp sums its own 32 staged rows (warp_stg), lane = channel pair, so one LDS.32 per row touches all 32 banks once and no
// shuffles are needed.  The sums are taken over the bf16-ROUNDED outputs, i.e. exactly the values BatchNorm later
// normalises.  valid_rows: bit r = pixel row 32q + r lies inside the image.
__device__ __forceinline__ void slab_stats(const uint8_t* warp_stg, uint32_t lane, uint32_t valid_rows,
                                           float* sum, float* sumsq) {
  const uint8_t* base = warp_stg + (lane & 3) * 4;
  const uint32_t c16 = lane >> 2;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    uint32_t w = *reinterpret_cast<const uint32_t*>(base + r * 128 + ((c16 ^ (uint32_t)(r & 7)) << 4));
    if (!((valid_rows >> r) & 1u)) w = 0u;
    const float x0 = __uint_as_float(w << 16);
    const float x1 = __uint_as_float(w & 0xffff0000u);
    a0 += x0;
    a1 += x1;
    b0 = fmaf(x0, x0, b0);
    b1 = fmaf(x1, x1, b1);
  }
  red_shared_add(sum + 2 * lane, a0);
  red_shared_add(sum + 2 * lane + 1, a1);
  red_shared_add(sumsq + 2 * lane, b0);
  red_shared_add(sumsq + 2 * lane + 1, b1);
}


// pair kernel (tc_conv2.cu): returns 0 on success; *used = 1 if the launch was taken by the pair kernel
int run_k1_pair(K1Params& p, const void* a0, const void* a1, const void* wpk, int ktot, cudaStream_t stream,
                int* used_grid, int* used_bn);

}  // namespace cmu
