// K2: tcgen05 weight-gradient GEMM with the reduction dimension = pixels (sm_100a).
//
//   D_tap[m, n] = sum_{pixels p} Q[p + tap_offset, m] * P[p, n]        (fp32 accumulate in TMEM)
//
// Both operands are the SAME kind of TMA tile K1 uses (128 pixels x 64 channels, SWIZZLE_128B); here the
// 64 contiguous channels are the M / N dimension and the pixels are K, i.e. both UMMA operands are MN-major.
// conv3x3 wgrad:  Q = layer input (haloed patch, one kernel column s per CTA, the 3 kernel rows are row offsets
//                 into the patch), P = dy;                     D[ci, co] -> dW[co, ci, r, s]
// convT2x2 wgrad: Q = dy gathered at (2h+r, 2w+s) (one (r,s) per CTA), P = x;  D[co, ci] -> dW[ci, co, r, s]
// The pixel range is split across CTAs (split-K); every CTA writes an fp32 partial block to a workspace and
// a small second kernel reduces the splits into the torch weight layout (no float atomics).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

constexpr int kK2Threads = 192;
constexpr int kQAtom = 20480;   // haloed patch atom (max (TH+2)*TW*128)
constexpr int kPAtom = 16384;   // plain 128-pixel atom
constexpr int kK2Stage = 2 * kQAtom + 2 * kPAtom;
constexpr int kK2Stages = 3;
constexpr int kPace = 4;
constexpr int kK2Smem = kK2Stages * kK2Stage + 1024 + 512;

struct K2Params {
  CUtensorMap tmP, tmQ0, tmQ1;
  int mode;            // 0 = conv3x3 wgrad, 1 = convT2x2 wgrad, 2 = plain rows (D = Q^T P over matrix rows)
  int N, H, W;         // pixel space
  int q0, q1;          // channels of Q source 0 / 1
  int pc;              // channels of P
  int TW, TH, tiles_w, tiles_h, pix_tiles;
  int MT, NT, G, splits;
  int n_stages;        // TMA ring depth: kK2Stages, or fewer when a CTA has only 1-2 pixel tiles (then several CTAs share an
                       // SM and overlap their start-up / epilogue latencies: projector dW has 24576 one-tile CTAs)
  int m_atoms;         // 1 (M = 64 duplicated to 128, or two kernel rows stacked when `stack`) or 2
  int paced;           // launched as clusters of G CTAs (the G kernel-column / tap groups of one pixel range): their TMA
                       // producers meet every kPace pixel tiles so the shared Q / P tiles are still in L2 for the others
  int stack;           // conv3x3 with 64 input channels: UMMA rows 0-63 = kernel row r, rows 64-127 = kernel row r+1
                       // of the SAME smem patch (LBO = one patch row), so 3 kernel rows take 2 UMMA sets instead of 3
  int n_cols;          // 64 or 128
  int taps;            // accumulators per CTA: 3 (conv) or 1 (convT)
  float* ws;           // [splits][G*taps][QC][PC]
};

__global__ void __launch_bounds__(kK2Threads, 1) k2_kernel(const __grid_constant__ K2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_stages = p.n_stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + n_stages * kK2Stage);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kK2Stages;
  uint64_t* tfull_bar = bars + 2 * kK2Stages;
  uint64_t* pace_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pace_bar + 1);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const int acc_sets = p.stack ? 2 : p.taps;
  const uint32_t tmem_cols = (acc_sets * p.n_cols <= 64) ? 64 : (acc_sets * p.n_cols <= 128) ? 128
                             : (acc_sets * p.n_cols <= 256) ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(pace_bar, p.G);
    fence_mbar_init();
    tma_prefetch_desc(&p.tmP);
    tma_prefetch_desc(&p.tmQ0);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (p.paced) cluster_sync_all();   // every CTA's pace barrier is initialised before a peer arrives on it
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // CTA -> (m tile, n tile, split, tap group); the tap group is the rank inside the cluster when paced
  int b = blockIdx.x;
  const int g = b % p.G;          b /= p.G;
  const int split = b % p.splits; b /= p.splits;
  const int nt = b % p.NT;
  const int mt = b / p.NT;
  const int t_begin = (int)(((long long)p.pix_tiles * split) / p.splits);
  const int t_end = (int)(((long long)p.pix_tiles * (split + 1)) / p.splits);
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int n_atoms = p.n_cols >> 6;
  const uint32_t q_atom_bytes = (p.mode == 0) ? (p.TH + 2) * p.TW * 128 : kPAtom;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = p.m_atoms * q_atom_bytes + n_atoms * kPAtom;
      uint32_t it = 0, pace_phase = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int img = t / tiles_per_img;
        const int rem = t - img * tiles_per_img;
        const int h0 = (rem / p.tiles_w) * p.TH;
        const int w0 = (rem % p.tiles_w) * p.TW;
        const uint32_t st = it % n_stages;
        const uint32_t ph = (it / n_stages) & 1;
        mbar_wait(&empty_bar[st], ph ^ 1);
        uint8_t* sQ = smem + st * kK2Stage;
        uint8_t* sP = sQ + 2 * kQAtom;
        mbar_arrive_expect_tx(&full_bar[st], tx);
        for (int a = 0; a < p.m_atoms; ++a) {
          const int qc = mt * 128 + a * 64;
          if (p.mode == 0) {
            const bool second = qc >= p.q0;
            tma_load_4d(sQ + a * kQAtom, second ? &p.tmQ1 : &p.tmQ0, &full_bar[st], second ? qc - p.q0 : qc,
                        w0 + g - 1, h0 - 1, img);
          } else if (p.mode == 1) {
            tma_load_5d(sQ + a * kQAtom, &p.tmQ0, &full_bar[st], (g & 1) * p.q0 + qc, w0, g >> 1, h0, img);
          } else {
            tma_load_4d(sQ + a * kQAtom, &p.tmQ0, &full_bar[st], qc, w0, h0, img);
          }
        }
        for (int a = 0; a < n_atoms; ++a)
          tma_load_4d(sP + a * kPAtom, &p.tmP, &full_bar[st], nt * p.n_cols + a * 64, w0, h0, img);
        if (p.paced && ((it % kPace) == kPace - 1 || t == t_end - 1)) {
          // all G CTAs of the cluster walk the same pixel tiles: tell every one of them (and myself) that I am here,
          // then wait until all of them are
          const uint32_t pb = smem_u32(pace_bar);
          for (int c = 0; c < p.G; ++c) mbar_arrive_cluster(mapa_u32(pb, (uint32_t)c));
          mbar_wait(pace_bar, pace_phase);
          pace_phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // whole warp runs the loop (warp-uniform descriptors in uniform registers), one elected lane issues
    const uint32_t idesc = umma_idesc_bf16(128, p.n_cols, 1, 1);
    // M = 64: the second M atom is the same patch one kernel row further down (stack) or an alias of the first
    const uint32_t a_lbo = (p.m_atoms == 2) ? kQAtom : p.stack ? (uint32_t)p.TW * 128 : 0;
    // MN-major SWIZZLE_128B: LBO = byte stride between 64-channel atoms, SBO = stride between 8-pixel K groups
    const uint64_t a_hi = umma_smem_desc(0, a_lbo, 1024);
    const uint64_t b_hi = umma_smem_desc(0, kPAtom, 1024);
    const uint32_t tap_stride16 = ((p.TW * 128) >> 4) * (p.stack ? 2 : 1);
    const uint32_t smem0 = smem_u32(smem);
    const int taps = acc_sets;
    uint32_t it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const uint32_t st = it % n_stages;
      const uint32_t ph = (it / n_stages) & 1;
      mbar_wait(&full_bar[st], ph);
      tc_fence_after();
      const uint32_t sQ = smem0 + st * kK2Stage;
      const uint32_t sP = sQ + 2 * kQAtom;
      const uint64_t a0 = a_hi | (uint64_t)((sQ >> 4) & 0x3FFF);
      const uint64_t b0 = b_hi | (uint64_t)((sP >> 4) & 0x3FFF);
      if (elect_one()) {
        for (int r = 0; r < taps; ++r) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // 128 pixels = 8 x K16; 16 pixel rows = 2048 bytes = 128 x 16 B
            umma_bf16(tmem_base + r * p.n_cols, a0 + (uint64_t)(r * tap_stride16 + k * 128), b0 + (uint64_t)(k * 128),
                      idesc, (it | (uint32_t)k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[st]);
        if (t == t_end - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
    }
    if (t_end <= t_begin && elect_one()) umma_commit(tfull_bar);
    __syncwarp();
  } else {
    const uint32_t q = warp & 3;
    const uint32_t row = q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int QC = p.q0 + p.q1;
    const bool any = (t_end > t_begin);
    for (int a = 0; a < acc_sets; ++a) {
      // accumulator set a, TMEM lane `row` -> (kernel row r, input channel ch)
      const int r = p.stack ? 2 * a + (int)(row >> 6) : a;
      const int ch = (p.m_atoms == 2) ? mt * 128 + (int)row : (int)(row & 63);
      const bool row_ok = (p.m_atoms == 2 || p.stack || row < 64) && r < p.taps && ch < QC;
      float* dst = p.ws + ((((size_t)split * p.G + g) * p.taps + r) * QC + ch) * p.pc + (size_t)nt * p.n_cols;
      for (int j = 0; j < p.n_cols / 32; ++j) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + ((q * 32) << 16) + a * p.n_cols + j * 32, raw);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 o;
            o.x = any ? __uint_as_float(raw[i + 0]) : 0.f;
            o.y = any ? __uint_as_float(raw[i + 1]) : 0.f;
            o.z = any ? __uint_as_float(raw[i + 2]) : 0.f;
            o.w = any ? __uint_as_float(raw[i + 3]) : 0.f;
            *reinterpret_cast<float4*>(dst + j * 32 + i) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.paced) cluster_sync_all();   // no CTA leaves while a peer may still arrive on its pace barrier
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dW[co][ci][r][s] = sum_split ws[split][s*3+r][ci][co]      (conv3x3; torch Conv2d layout)
__global__ void wgrad_reduce_conv3(const float* __restrict__ ws, float* __restrict__ dw, int cin, int cout, int splits,
                                   int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over (t, ci, co) with co fastest -> coalesced reads
  const int total = 9 * cin * cout;
  if (idx >= total) return;
  const int co = idx % cout;
  const int ci = (idx / cout) % cin;
  const int t = idx / (cout * cin);
  float acc = 0.f;
  for (int sp = 0; sp < splits; ++sp) acc += ws[(size_t)sp * total + idx];
  const int s = t / 3, r = t % 3;
  float* o = dw + (((size_t)co * cin + ci) * 3 + r) * 3 + s;
  *o = accumulate ? *o + acc : acc;
}
// dW[ci][co][r][s] = sum_split ws[split][r*2+s][co][ci]       (convT2x2; torch ConvTranspose2d layout)
__global__ void wgrad_reduce_convT(const float* __restrict__ ws, float* __restrict__ dw, int cin, int cout, int splits,
                                   int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over (t, co, ci) with ci fastest
  const int total = 4 * cin * cout;
  if (idx >= total) return;
  const int ci = idx % cin;
  const int co = (idx / cin) % cout;
  const int t = idx / (cout * cin);
  float acc = 0.f;
  for (int sp = 0; sp < splits; ++sp) acc += ws[(size_t)sp * total + idx];
  float* o = dw + (((size_t)ci * cout + co) * 2 + (t >> 1)) * 2 + (t & 1);
  *o = accumulate ? *o + acc : acc;
}

// out[m][n] (+)= sum_split ws[split][m][n] (+ bias[n])        (plain rows mode)
__global__ void rows_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, int total, int pc, int splits,
                                   const float* __restrict__ bias, int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float acc = 0.f;
  for (int sp = 0; sp < splits; ++sp) acc += ws[(size_t)sp * total + idx];
  if (bias != nullptr) acc += bias[idx % pc];
  out[idx] = accumulate ? out[idx] + acc : acc;
}

static int make_act_map4(CUtensorMap* m, const void* base, int N, int H, int W, int C, int box_w, int box_h) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, (uint32_t)box_w, (uint32_t)box_h, 1};
  return encode_tmap_bf16(m, base, 4, dims, str, box);
}
static int make_up_map5(CUtensorMap* m, const void* base, int N, int H, int W, int C, int box_w, int box_h) {
  uint64_t dims[5] = {(uint64_t)2 * C, (uint64_t)W, 2, (uint64_t)H, (uint64_t)N};
  uint64_t str[4] = {(uint64_t)2 * C * 2, (uint64_t)2 * W * C * 2, (uint64_t)4 * W * C * 2, (uint64_t)4 * H * W * C * 2};
  uint32_t box[5] = {64, (uint32_t)box_w, 1, (uint32_t)box_h, 1};
  return encode_tmap_bf16(m, base, 5, dims, str, box);
}

// CTAs that can be co-resident when the kernel is launched as clusters of `g` (clusters must fit inside a GPC)
static int k2_cluster_slots(int g);

struct K2Plan {
  int TW, TH, tiles_w, tiles_h, pix_tiles, MT, NT, G, splits, m_atoms, n_cols, taps;
};

static void plan_k2(int mode, int N, int H, int W, int qc, int pc, K2Plan* pl) {
  if (mode == 0) {
    pl->TW = (W > 8 && H <= 8) ? 16 : 8;
  } else {
    const int cand[5] = {8, 16, 32, 64, 128};
    long best_cost = -1;
    pl->TW = 8;
    for (int i = 0; i < 5; ++i) {
      const int tw = cand[i], th = 128 / tw;
      const long cost = (long)ceil_div(W, tw) * tw * ceil_div(H, th) * th;
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; pl->TW = tw; }
    }
  }
  pl->TH = 128 / pl->TW;
  pl->tiles_w = ceil_div(W, pl->TW);
  pl->tiles_h = ceil_div(H, pl->TH);
  pl->pix_tiles = N * pl->tiles_w * pl->tiles_h;
  pl->m_atoms = (qc >= 128) ? 2 : 1;
  pl->MT = ceil_div(qc, 128);
  pl->n_cols = (pc % 128 == 0) ? 128 : 64;
  pl->NT = pc / pl->n_cols;
  pl->G = (mode == 0) ? 3 : (mode == 1 ? 4 : 1);
  pl->taps = (mode == 0) ? 3 : 1;
  // split-K over pixel tiles.  Cost model: a launch takes waves(s) x (pixel tiles per CTA + fixed CTA cost), the fixed
  // cost (TMEM alloc, pipeline fill, fp32 partial-block epilogue) being worth ~10 pixel-tile stages.  This fills the
  // machine for small layers and repairs the 1.3-wave quantisation of the big ones (192 CTAs on 148 SMs).
  const int base = pl->MT * pl->NT * pl->G;
  const int sms = (pl->G > 1 && debug_knob(10) == 1) ? k2_cluster_slots(pl->G) : num_sms();
  const long long per_split_bytes = (long long)pl->G * pl->taps * qc * pc * 4;
  int splits = 1;
  double best = 1e300;
  for (int sp = 1; sp <= pl->pix_tiles && sp <= 1024; ++sp) {
    const long long ctas = (long long)base * sp;
    if (sp > 1 && (ctas > 4ll * sms || per_split_bytes * sp > (768ll << 20))) break;
    const long long waves = (ctas + sms - 1) / sms;
    const double cost = (double)waves * ((double)pl->pix_tiles / sp + 10.0);
    if (cost < best * 0.999) {
      best = cost;
      splits = sp;
    }
  }
  pl->splits = splits;
}

static int k2_cluster_slots(int g) {
  static int cache[8] = {0};
  if (g < 1 || g > 7) return num_sms();
  if (cache[g] == 0) {
    cudaFuncSetAttribute(k2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2Smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(g * num_sms()), 1, 1);
    cfg.blockDim = dim3(kK2Threads, 1, 1);
    cfg.dynamicSmemBytes = kK2Smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)g;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k2_kernel, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms() / g;
    }
    cache[g] = n * g;
  }
  return cache[g];
}

int simt_wgrad(int mode, const void* x0, int c0, const void* x1, int c1, const void* dy, int cout, int N, int H, int W,
               float* dw, int accumulate, cudaStream_t st);

static int run_k2(int mode, const void* q0, int qc0, const void* q1, int qc1, const void* pten, int pc, int N, int H,
                  int W, float* ws, size_t ws_bytes, float* dw, int accumulate, cudaStream_t stream,
                  const float* bias = nullptr) {
  if (debug_knob(0) == 1 && mode != 2) {  // CUDA-core cross-check path (tests / debugging only)
    if (mode == 0) return simt_wgrad(0, q0, qc0, q1, qc1, pten, pc, N, H, W, dw, accumulate, stream);
    return simt_wgrad(1, pten, pc, nullptr, 0, q0, qc0, N, H, W, dw, accumulate, stream);
  }
  const int qc = qc0 + qc1;
  CMU_REQUIRE(qc0 % 64 == 0 && qc1 % 64 == 0 && pc % 64 == 0, "wgrad: channels must be multiples of 64");
  K2Plan pl;
  plan_k2(mode, N, H, W, qc, pc, &pl);
  const size_t need = (size_t)pl.splits * pl.G * pl.taps * qc * pc * sizeof(float);
  const bool direct = (mode == 2 && pl.splits == 1 && bias == nullptr && !accumulate);   // D already has the output layout
  if (direct) ws = dw;
  CMU_REQUIRE(direct || (ws != nullptr && ws_bytes >= need), "wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
  K2Params p;
  memset(&p, 0, sizeof(p));
  p.mode = mode;
  p.N = N; p.H = H; p.W = W;
  p.q0 = qc0; p.q1 = qc1; p.pc = pc;
  p.TW = pl.TW; p.TH = pl.TH; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.pix_tiles = pl.pix_tiles;
  p.MT = pl.MT; p.NT = pl.NT; p.G = pl.G; p.splits = pl.splits;
  p.m_atoms = pl.m_atoms; p.n_cols = pl.n_cols; p.taps = pl.taps;
  p.stack = (mode == 0 && pl.m_atoms == 1 && debug_knob(8) == 0) ? 1 : 0;
  p.ws = ws;
  if (make_act_map4(&p.tmP, pten, N, H, W, pc, pl.TW, pl.TH)) return 1;
  if (mode == 0) {
    if (make_act_map4(&p.tmQ0, q0, N, H, W, qc0, pl.TW, pl.TH + 2)) return 1;
    if (q1 != nullptr) {
      if (make_act_map4(&p.tmQ1, q1, N, H, W, qc1, pl.TW, pl.TH + 2)) return 1;
    } else {
      p.tmQ1 = p.tmQ0;
    }
  } else if (mode == 1) {
    if (make_up_map5(&p.tmQ0, q0, N, H, W, qc0, pl.TW, pl.TH)) return 1;
    p.tmQ1 = p.tmQ0;
  } else {
    if (make_act_map4(&p.tmQ0, q0, N, H, W, qc0, pl.TW, pl.TH)) return 1;
    p.tmQ1 = p.tmQ0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    CMU_CHECK_CUDA(cudaFuncSetAttribute(k2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kK2Smem));
    attr_set = true;
  }
  const int grid = pl.MT * pl.NT * pl.G * pl.splits;
  const int tiles_per_cta = ceil_div(pl.pix_tiles, pl.splits);
  p.n_stages = (debug_knob(12) != 1 && tiles_per_cta < kK2Stages) ? tiles_per_cta : kK2Stages;
  if (debug_knob(15) == 2 && p.n_stages > 2) p.n_stages = 2;   // A/B: leave 72 KB of the SM to co-resident kernels
  const int smem_bytes = p.n_stages * kK2Stage + 1024 + 512;
  p.paced = (pl.G > 1 && debug_knob(10) == 1) ? 1 : 0;   // A/B: see profiles/r1_k2_dram_traffic.md
  if (p.paced) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(kK2Threads, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)pl.G;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CMU_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k2_kernel, p));
  } else {
    k2_kernel<<<grid, kK2Threads, smem_bytes, stream>>>(p);
  }
  CMU_LAUNCH_CHECK();
  const int total = pl.G * pl.taps * qc * pc;
  if (mode == 0)
    wgrad_reduce_conv3<<<ceil_div(total, 256), 256, 0, stream>>>(ws, dw, qc, pc, pl.splits, accumulate);
  else if (mode == 1)
    wgrad_reduce_convT<<<ceil_div(total, 256), 256, 0, stream>>>(ws, dw, pc, qc, pl.splits, accumulate);
  else if (!direct)
    rows_reduce_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(ws, dw, total, pc, pl.splits, bias, accumulate);
  if (mode != 2 || !direct) CMU_LAUNCH_CHECK();
  return 0;
}

}  // namespace cmu

using namespace cmu;

extern "C" {

long long cmu_conv3x3_wgrad_workspace_bytes(int cin, int cout, int n, int h, int w) {
  K2Plan pl;
  plan_k2(0, n, h, w, cin, cout, &pl);
  return (long long)pl.splits * 9 * cin * cout * 4;
}

int cmu_conv3x3_wgrad(const void* x0, int c0, const void* x1, int c1, const void* dy, int cout, int n, int h, int w,
                      float* workspace, long long workspace_bytes, float* dw, int accumulate, void* stream) {
  return run_k2(0, x0, c0, x1, c1, dy, cout, n, h, w, workspace, (size_t)workspace_bytes, dw, accumulate,
                (cudaStream_t)stream);
}

long long cmu_convT2x2_wgrad_workspace_bytes(int cin, int cout, int n, int h, int w) {
  K2Plan pl;
  plan_k2(1, n, h, w, cout, cin, &pl);
  return (long long)pl.splits * 4 * cin * cout * 4;
}

// D[qc][pc] = Q^T P over `rows` matrix rows: Q (rows, qc) bf16, P (rows, pc) bf16, out (qc, pc) fp32 (+ bias[pc]).
// Used for the projection-head linears (nonlinear_neck.py:94): forward, dgrad and wgrad are all of this form once the
// small operand is transposed.  qc, pc multiples of 64; rows arbitrary (zero-filled to 128-row tiles by TMA).
long long cmu_gemm_tn_workspace_bytes(int qc, int pc, long long rows) {
  K2Plan pl;
  plan_k2(2, 1, 1, (int)rows, qc, pc, &pl);
  return pl.splits > 1 ? (long long)pl.splits * qc * pc * 4 : 0;
}

int cmu_gemm_tn_bf16(const void* q, int qc, const void* p, int pc, long long rows, float* out, const float* bias,
                     int accumulate, float* workspace, long long workspace_bytes, void* stream) {
  CMU_REQUIRE(rows > 0 && rows < (1ll << 31), "gemm_tn: bad row count");
  return run_k2(2, q, qc, nullptr, 0, p, pc, 1, 1, (int)rows, workspace, (size_t)workspace_bytes, out, accumulate,
                (cudaStream_t)stream, bias);
}

int cmu_convT2x2_wgrad(const void* x, int cin, const void* dy, int cout, int n, int h, int w, float* workspace,
                       long long workspace_bytes, float* dw, int accumulate, void* stream) {
  return run_k2(1, dy, cout, nullptr, 0, x, cin, n, h, w, workspace, (size_t)workspace_bytes, dw, accumulate,
                (cudaStream_t)stream);
}

}  // extern "C"
