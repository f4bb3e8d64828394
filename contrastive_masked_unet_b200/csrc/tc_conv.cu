// K1: persistent, warp-specialised tcgen05 implicit-GEMM for NHWC bf16 activations (sm_100a).
//
//   D[pixels, n] = sum_{tap, k} A[pixel + tap_offset, k] * Wp[tap][n][k]        (fp32 accumulate in TMEM)
//
// One kernel covers: conv3x3 pad 1 fprop and dgrad (dgrad = fprop with rotated/transposed weights),
// 1x1 / plain GEMM, ConvTranspose2d(k2,s2) fprop (GEMM + 2x2 scatter epilogue) and its dgrad (gathered A).
//
// Data movement: every operand tile is a TMA box with SWIZZLE_128B (64 bf16 channels = one 128-byte row per
// pixel).  conv3x3 loads, per 64-channel K chunk, THREE horizontally shifted (TH+2) x TW haloed patches (one per
// kernel column s); the three kernel rows r are then plain 1024B-aligned row offsets of the same patch in the
// UMMA shared-memory descriptor, so 3 TMA loads feed 9 taps (zero padding = TMA out-of-bounds fill).
// Roles (320 threads): warp 0 = tile scheduler + TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner,
// warps 2..5 (and 6..9 when shared memory allows a second staging ring: one warp group per TMEM accumulator) =
// epilogue (TMEM -> regs -> [bias] -> bf16 -> the warp's own 32 rows of a swizzled staging slab -> the warp's own TMA
// store; BatchNorm sum / sum-of-squares are taken from the staged bf16 slab).  Accumulators are double buffered in TMEM
// so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "ptx.cuh"
#include "k1_common.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

// Tile scheduling: every CTA owns ONE n-tile (blockIdx % n_tiles; its BatchNorm partial sums live in one smem slot)
// and pulls m-tiles of that n-tile from a global atomic counter.  The producer thread fetches tile indices and
// publishes them to the MMA and epilogue warps through a small smem ring, so a CTA that starts late (another
// kernel holding its SM) or runs slow simply takes fewer tiles -- no static tail.
template <int BN>
__global__ void __launch_bounds__(kK1Threads, 1) k1_kernel(const __grid_constant__ K1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_res = smem;                                   // [w_bytes] resident weights (w_resident mode)
  uint8_t* stage_base = smem + p.w_bytes;
  uint8_t* staging = stage_base + p.n_stages * p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + p.epi_groups * p.stg_bufs * kStagingBytes);
  uint64_t* full_bar = bars;                          // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;            // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;        // [2]
  uint64_t* tempty_bar = tfull_bar + 2;               // [2]
  uint64_t* sfull_bar = tempty_bar + 2;               // [kSchedDepth]
  uint64_t* sempty_bar = sfull_bar + kSchedDepth;     // [kSchedDepth]
  uint64_t* wfull_bar = sempty_bar + kSchedDepth;     // [1]
  int* sched_tile = reinterpret_cast<int*>(wfull_bar + 1);  // [kSchedDepth]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sched_tile + kSchedDepth);
  float* s_stats = reinterpret_cast<float*>(bars + 40);  // [2][BN]
  float* s_bias = s_stats + 2 * BN;                      // [BN] bias of this CTA's n-tile (broadcast reads in the epilogue)
  constexpr uint32_t kTmemCols = 2 * BN;

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    for (int i = 0; i < kSchedDepth; ++i) {
      mbar_init(&sfull_bar[i], 1);
      mbar_init(&sempty_bar[i], 1 + 4 * p.epi_groups);  // MMA lane + one lane of each active epilogue warp
    }
    mbar_init(wfull_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmO0);
  }
  for (int i = threadIdx.x; i < 2 * BN; i += kK1Threads) s_stats[i] = 0.f;
  if (p.bias != nullptr)
    for (int i = threadIdx.x; i < BN; i += kK1Threads) s_bias[i] = p.bias[((blockIdx.x % p.n_tiles) * BN + i) % p.bias_mod];
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nt = blockIdx.x % p.n_tiles;
  const int n0 = nt * BN;
  const int shifts = (p.mode == MODE_CONV3) ? 3 : 1;
  const int kc = p.kc;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    // ------------------------------------------------------------------ scheduler + TMA producer
    if (lane == 0) {
      const uint32_t tx_bytes = p.a_bytes + (p.w_resident ? 0 : p.b_bytes);
      if (p.w_resident) {   // one-time load of this n-tile's weights: kc x shifts boxes
        mbar_arrive_expect_tx(wfull_bar, p.w_bytes);
        for (int c = 0; c < kc; ++c)
          for (int s = 0; s < shifts; ++s)
            tma_load_3d(w_res + (c * shifts + s) * p.b_bytes, &p.tmB, wfull_bar, c << 6, n0,
                        (p.mode == MODE_CONV3) ? 3 * s : 0);
      }
      uint32_t it = 0, sit = 0;
      // dynamic: m-tiles come from a global counter, fetched ONE TILE AHEAD so that the atomic's round trip overlaps
      // the loads of the current tile; static (sched == nullptr, A/B switch): blockIdx-strided like a classic
      // persistent kernel.
      const int mstride = gridDim.x / p.n_tiles;
      int static_next = blockIdx.x / p.n_tiles;
      int t_next = p.sched ? (int)atomicAdd(p.sched + nt, 1u) : static_next;
      while (true) {
        const uint32_t sslot = sit % kSchedDepth;
        mbar_wait(&sempty_bar[sslot], ((sit / kSchedDepth) & 1) ^ 1);
        const int mt = (t_next < p.m_tiles) ? t_next : -1;
        sched_tile[sslot] = mt;
        mbar_arrive(&sfull_bar[sslot]);
        ++sit;
        if (mt < 0) break;
        if (p.sched) t_next = (int)atomicAdd(p.sched + nt, 1u);
        else t_next = (static_next += mstride);
        const int img = mt / tiles_per_img;
        const int rem = mt - img * tiles_per_img;
        const int h0 = (rem / p.tiles_w) * p.TH;
        const int w0 = (rem % p.tiles_w) * p.TW;
        for (int c = 0; c < kc; ++c) {
          for (int s = 0; s < shifts; ++s, ++it) {
            const uint32_t st = it % n_stages;
            const uint32_t ph = (it / n_stages) & 1;
            mbar_wait(&empty_bar[st], ph ^ 1);
            uint8_t* sA = stage_base + st * p.stage_bytes;
            uint8_t* sB = sA + p.b_off;
            mbar_arrive_expect_tx(&full_bar[st], tx_bytes);
            if (p.mode == MODE_CONVT_DGRAD) {
              const int per = p.c0 >> 6;
              const int rs = c / per;
              const int cc = (c - rs * per) << 6;
              tma_load_5d(sA, &p.tmA0, &full_bar[st], (rs & 1) * p.c0 + cc, w0, rs >> 1, h0, img);
            } else {
              const int k0 = c << 6;
              const bool second = k0 >= p.c0;
              const CUtensorMap* src = second ? &p.tmA1 : &p.tmA0;
              const int cc = second ? k0 - p.c0 : k0;
              if (p.mode == MODE_CONV3)
                tma_load_4d(sA, src, &full_bar[st], cc, w0 + s - 1, h0 - 1, img);
              else
                tma_load_4d(sA, src, &full_bar[st], cc, w0, h0, img);
            }
            if (!p.w_resident)
              tma_load_3d(sB, &p.tmB, &full_bar[st], c << 6, n0, (p.mode == MODE_CONV3) ? 3 * s : 0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The WHOLE warp runs this loop so that every address / descriptor is warp-uniform (kept in uniform registers);
    // one elected lane issues the tcgen05 instructions.  (With a single-lane branch around the loop the compiler had
    // to broadcast each descriptor through R2UR + an ELECT loop: ~22 instructions per MMA, which made the kernel
    // MMA-issue bound -- profiles/r1_k1_ncu_full.md.)
    const uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
    const uint32_t a_tap_stride = p.TW * 128;
    const bool conv3 = (p.mode == MODE_CONV3);
    uint32_t it = 0, tile_it = 0, sit = 0;
    const uint32_t w_base = smem_u32(w_res);
    const uint32_t stage0 = smem_u32(stage_base);
    // descriptor high words are constant: LBO = 16 B (unused for swizzled K-major), SBO = 1024 B, version 1, SW128
    const uint64_t desc_hi = umma_smem_desc(0, 16, 1024);
    if (p.w_resident) mbar_wait(wfull_bar, 0);
    while (true) {
      const uint32_t sslot = sit % kSchedDepth;
      mbar_wait(&sfull_bar[sslot], (sit / kSchedDepth) & 1);
      const int mt = sched_tile[sslot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sempty_bar[sslot]);
      ++sit;
      if (mt < 0) break;
      const uint32_t acc = tile_it & 1;
      const uint32_t aph = (tile_it >> 1) & 1;
      ++tile_it;
      mbar_wait(&tempty_bar[acc], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      const int nstages = kc * shifts;
      for (int sidx = 0; sidx < nstages; ++sidx, ++it) {
        const uint32_t st = it % n_stages;
        const uint32_t ph = (it / n_stages) & 1;
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        const uint32_t sA = stage0 + st * p.stage_bytes;
        const uint32_t sB = p.w_resident ? w_base + sidx * p.b_bytes : sA + p.b_off;   // sidx = c * shifts + s
        const uint64_t a0 = desc_hi | (uint64_t)((sA >> 4) & 0x3FFF);
        const uint64_t b0 = desc_hi | (uint64_t)((sB >> 4) & 0x3FFF);
        if (elect_one()) {
          if (conv3) {
            const uint32_t ta = a_tap_stride >> 4;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // start-address field is in 16-byte units: +2 per K16 step (32 B), + tap stride per kernel row
                umma_bf16(d_tmem, a0 + (uint64_t)(r * ta + k * 2), b0 + (uint64_t)(r * (BN * 8) + k * 2), idesc,
                          (sidx | r | k) != 0 ? 1u : 0u);
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, a0 + (uint64_t)(k * 2), b0 + (uint64_t)(k * 2), idesc, (sidx | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[st]);  // smem slot reusable once these MMAs have read it
          if (sidx == nstages - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5 [, 6..9])
    const uint32_t q = warp & 3;  // TMEM lane quadrant this warp may access
    const uint32_t row = q * 32 + lane;
    const int th = row >> p.tw_shift;
    const int tw = row & (p.TW - 1);
    const int wth0 = (int)(q * 32) >> p.tw_shift;   // first pixel row of this warp inside the tile
    const int wtw0 = (int)(q * 32) & (p.TW - 1);
    const int stg_bufs = p.stg_bufs;
    const uint32_t grp = (warp - 2) >> 2;             // epilogue group = the TMEM accumulator it drains
    const bool two_groups = (p.epi_groups == 2);
    uint8_t* const grp_staging = staging + grp * stg_bufs * kStagingBytes;
    uint32_t tile_it = 0, sit = 0, slab_it = 0;
    while (grp < (uint32_t)p.epi_groups) {
      const uint32_t sslot = sit % kSchedDepth;
      mbar_wait(&sfull_bar[sslot], (sit / kSchedDepth) & 1);
      const int mt = sched_tile[sslot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sempty_bar[sslot]);
      ++sit;
      if (mt < 0) break;
      const int img = mt / tiles_per_img;
      const int rem = mt - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * p.TH;
      const int w0 = (rem % p.tiles_w) * p.TW;
      const bool valid = (h0 + th < p.H) && (w0 + tw < p.W);
      const uint32_t acc = tile_it & 1;
      const uint32_t aph = (tile_it >> 1) & 1;
      ++tile_it;
      if (two_groups && acc != grp) continue;
      const uint32_t valid_rows = __ballot_sync(0xffffffffu, valid);
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
#pragma unroll 1
      for (int slab = 0; slab < BN / 64; ++slab, ++slab_it) {
        // warp-private staging (32 rows x 128 B) and warp-private TMA stores: no CTA-wide barrier in the tile loop
        uint8_t* stg = grp_staging + (slab_it % stg_bufs) * kStagingBytes + q * 4096;
        if (lane == 0) {  // the store that last used this staging buffer has finished reading it
          if (stg_bufs == 2) tma_store_wait_read1();
          else tma_store_wait_read0();
        }
        __syncwarp();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int j = slab * 2 + half;
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + ((q * 32) << 16) + acc * BN + j * 32, raw);
          tmem_ld_wait();
          if (j == BN / 32 - 1) {  // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          if (p.bias != nullptr) {
            const float4* sb = reinterpret_cast<const float4*>(s_bias + j * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = sb[i];
              v[4 * i + 0] += b4.x;
              v[4 * i + 1] += b4.y;
              v[4 * i + 2] += b4.z;
              v[4 * i + 3] += b4.w;
            }
          }
          if (p.exp_scale != 0.f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __expf((v[i] - 1.f) * p.exp_scale);
          }
          // bf16 pack + swizzled staging store (16-byte chunk index XOR (row & 7): TMA SWIZZLE_128B pattern)
          uint8_t* rowp = stg + lane * 128;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            uint4 pk;
            pk.x = pack_bf16(v[k4 * 8 + 0], v[k4 * 8 + 1]);
            pk.y = pack_bf16(v[k4 * 8 + 2], v[k4 * 8 + 3]);
            pk.z = pack_bf16(v[k4 * 8 + 4], v[k4 * 8 + 5]);
            pk.w = pack_bf16(v[k4 * 8 + 6], v[k4 * 8 + 7]);
            const int chunk = (half * 4 + k4) ^ (lane & 7);
            *reinterpret_cast<uint4*>(rowp + chunk * 16) = pk;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int nch = n0 + slab * 64;
          if (p.mode == MODE_CONVT_FPROP) {
            const int rs = nch / p.oc0;
            const int co = nch - rs * p.oc0;
            tma_store_5d(&p.tmO0, stg, (rs & 1) * p.oc0 + co, w0 + wtw0, rs >> 1, h0 + wth0, img);
          } else if (nch < p.oc0) {
            tma_store_4d(&p.tmO0, stg, nch, w0 + wtw0, h0 + wth0, img);
          } else {
            tma_store_4d(&p.tmO1, stg, nch - p.oc0, w0 + wtw0, h0 + wth0, img);
          }
          tma_store_commit();
        }
        if (p.stats != nullptr) slab_stats(stg, lane, valid_rows, &s_stats[slab * 64], &s_stats[BN + slab * 64]);
      }
    }
    if (lane == 0) tma_store_wait_all0();
    if (p.stats != nullptr) {
      named_bar_sync(1, 256);
      for (int i = threadIdx.x - 64; i < 2 * BN; i += 256) p.stats[(size_t)blockIdx.x * 2 * BN + i] = s_stats[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Global m-tile counters of the dynamic scheduler: a ring of slots, one per launch, zeroed (stream-ordered) before
// the launch that uses it.  Static device memory -- the library still never allocates.
constexpr int kSchedSlots = 512;
constexpr int kSchedWidth = 32;
__device__ unsigned int g_sched_counters[kSchedSlots * kSchedWidth];

// ---------------------------------------------------------------------------------------------- host side
static bool pick_tile(int mode, int H, int W, int* TW, int* TH) {
  // tile = TH x TW = 128 pixels, TW a multiple of 8 (one SWIZZLE_128B atom = 8 pixel rows)
  if (mode == MODE_CONV3) {
    *TW = (W > 8 && H <= 8) ? 16 : 8;
    *TH = 128 / *TW;
    return true;
  }
  // no halo: prefer the widest tile that does not waste columns
  const int cand[5] = {8, 16, 32, 64, 128};
  int best = 8;
  long best_cost = -1;
  for (int i = 0; i < 5; ++i) {
    const int tw = cand[i], th = 128 / tw;
    const long cost = (long)ceil_div(W, tw) * tw * ceil_div(H, th) * th;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = tw;
    }
  }
  *TW = best;
  *TH = 128 / best;
  return true;
}

static int make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int box_w, int box_h) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, (uint32_t)box_w, (uint32_t)box_h, 1};
  return encode_tmap_bf16(m, base, 4, dims, str, box);
}
// view of a (N, 2H, 2W, C) tensor as (sc = s*C + c, w, r, h, n): pixel (2h+r, 2w+s)
static int make_up_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int box_w, int box_h) {
  uint64_t dims[5] = {(uint64_t)2 * C, (uint64_t)W, 2, (uint64_t)H, (uint64_t)N};
  uint64_t str[4] = {(uint64_t)2 * C * 2, (uint64_t)2 * W * C * 2, (uint64_t)4 * W * C * 2, (uint64_t)4 * H * W * C * 2};
  uint32_t box[5] = {64, (uint32_t)box_w, 1, (uint32_t)box_h, 1};
  return encode_tmap_bf16(m, base, 5, dims, str, box);
}
static int make_w_map(CUtensorMap* m, const void* base, int taps, int n_rows, int k, int box_n, int box_taps) {
  uint64_t dims[3] = {(uint64_t)k, (uint64_t)n_rows, (uint64_t)taps};
  uint64_t str[2] = {(uint64_t)k * 2, (uint64_t)n_rows * k * 2};
  uint32_t box[3] = {64, (uint32_t)box_n, (uint32_t)box_taps};
  return encode_tmap_bf16(m, base, 3, dims, str, box);
}

template <int BN>
static int launch_k1(const K1Params& p, int grid, int smem_bytes, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CMU_CHECK_CUDA(cudaFuncSetAttribute(k1_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  k1_kernel<BN><<<grid, kK1Threads, smem_bytes, stream>>>(p);
  CMU_LAUNCH_CHECK();
  return 0;
}

static int next_sched_slot(unsigned int** out, int n_tiles, cudaStream_t stream) {
  static unsigned int* base = nullptr;
  static unsigned int counter = 0;
  if (base == nullptr) CMU_CHECK_CUDA(cudaGetSymbolAddress((void**)&base, g_sched_counters));
  CMU_REQUIRE(n_tiles <= kSchedWidth, "k1: too many n-tiles (%d)", n_tiles);
  unsigned int* slot = base + (size_t)(counter++ % kSchedSlots) * kSchedWidth;
  CMU_CHECK_CUDA(cudaMemsetAsync(slot, 0, kSchedWidth * sizeof(unsigned int), stream));
  *out = slot;
  return 0;
}

// Generic driver used by the C-ABI entry points below.
int simt_k1(int mode, const void* a0, int c0, const void* a1, int c1, int N, int H, int W, const void* wpk, int n_total,
            void* out0, int oc0, void* out1, int oc1, const float* bias, int bias_mod, float* stats_partial,
            int* stats_grid, cudaStream_t st);

static int run_k1(int mode, const void* a0, int c0, const void* a1, int c1, int N, int H, int W, const void* wpk,
                  int n_total, void* out0, int oc0, void* out1, int oc1, const float* bias, int bias_mod,
                  float* stats_partial, int* stats_grid, int* stats_bn, cudaStream_t stream, float exp_scale = 0.f) {
  if (debug_knob(0) == 1) {  // CUDA-core cross-check path (tests / debugging only)
    CMU_REQUIRE(exp_scale == 0.f, "k1: the CUDA-core cross-check path has no exp epilogue");
    if (stats_bn) *stats_bn = n_total;
    return simt_k1(mode, a0, c0, a1, c1, N, H, W, wpk, n_total, out0, oc0, out1, oc1, bias, bias_mod, stats_partial,
                   stats_grid, stream);
  }
  CMU_REQUIRE(c0 > 0 && c0 % 64 == 0 && c1 % 64 == 0, "k1: channel counts must be multiples of 64 (c0=%d c1=%d)", c0, c1);
  CMU_REQUIRE(n_total % 64 == 0, "k1: GEMM-N must be a multiple of 64 (got %d)", n_total);
  K1Params p;
  memset(&p, 0, sizeof(p));
  p.mode = mode;
  p.N = N; p.H = H; p.W = W;
  p.c0 = c0; p.c1 = c1;
  p.n_total = n_total;
  p.oc0 = oc0;
  pick_tile(mode, H, W, &p.TW, &p.TH);
  p.tw_shift = __builtin_ctz(p.TW);
  p.tiles_w = ceil_div(W, p.TW);
  p.tiles_h = ceil_div(H, p.TH);
  p.m_tiles = N * p.tiles_w * p.tiles_h;
  int BN = (n_total % 128 == 0) ? 128 : 64;
  CMU_REQUIRE(oc0 % 64 == 0 && oc1 % 64 == 0, "k1: output channel counts must be multiples of 64");
  if (debug_knob(1) == 64) BN = 64;
  p.n_tiles = n_total / BN;
  if (stats_bn) *stats_bn = BN;
  p.kc = (mode == MODE_CONVT_DGRAD) ? 4 * (c0 / 64) : (c0 + c1) / 64;
  p.bias = bias;
  p.bias_mod = bias_mod > 0 ? bias_mod : n_total;
  p.stats = stats_partial;
  p.exp_scale = exp_scale;

  const int box_h = (mode == MODE_CONV3) ? p.TH + 2 : p.TH;
  if (mode == MODE_CONVT_DGRAD) {
    if (make_up_map(&p.tmA0, a0, N, H, W, c0, p.TW, p.TH)) return 1;
    p.tmA1 = p.tmA0;
  } else {
    if (make_act_map(&p.tmA0, a0, N, H, W, c0, p.TW, box_h)) return 1;
    if (a1 != nullptr) {
      if (make_act_map(&p.tmA1, a1, N, H, W, c1, p.TW, box_h)) return 1;
    } else {
      p.tmA1 = p.tmA0;
    }
  }
  const int ktot = p.kc * 64;
  if (mode == MODE_CONV3) {
    if (make_w_map(&p.tmB, wpk, 9, n_total, ktot, BN, 3)) return 1;
  } else {
    if (make_w_map(&p.tmB, wpk, 1, n_total, ktot, BN, 1)) return 1;
  }
  // every epilogue warp stores its own 32 pixel rows: output box = 32 pixels (part of a tile row or 32/TW tile rows)
  const int obox_w = p.TW < 32 ? p.TW : 32;
  const int obox_h = 32 / obox_w;
  if (mode == MODE_CONVT_FPROP) {
    if (make_up_map(&p.tmO0, out0, N, H, W, oc0, obox_w, obox_h)) return 1;
    p.tmO1 = p.tmO0;
  } else {
    if (make_act_map(&p.tmO0, out0, N, H, W, oc0, obox_w, obox_h)) return 1;
    if (out1 != nullptr) {
      if (make_act_map(&p.tmO1, out1, N, H, W, oc1, obox_w, obox_h)) return 1;
    } else {
      p.tmO1 = p.tmO0;
    }
  }
  // CTA-pair kernel (tcgen05 cta_group::2) for the large tensor-bound layers
  {
    int pair_grid = 0, pair_bn = 0;
    K1Params pp = p;
    if (run_k1_pair(pp, a0, a1, wpk, ktot, stream, &pair_grid, &pair_bn)) return 1;
    if (pair_grid > 0) {
      if (stats_grid) *stats_grid = pair_grid;
      if (stats_bn) *stats_bn = pair_bn;
      return 0;
    }
  }
  // shared-memory plan
  p.a_bytes = (mode == MODE_CONV3) ? (p.TH + 2) * p.TW * 128 : 128 * 128;
  p.b_bytes = ((mode == MODE_CONV3) ? 3 : 1) * BN * 128;
  p.b_off = (p.a_bytes + 1023) & ~1023;
  p.stage_bytes = p.b_off + p.b_bytes;
  const int fixed = 1024 /*align*/ + 320 /*barriers*/ + 3 * BN * 4 /*stats + bias*/ + 64;
  // small-weight layers: keep the n-tile's whole weight slab resident, stream activations only
  const int w_all = p.kc * ((mode == MODE_CONV3) ? 3 : 1) * p.b_bytes;   // all taps x K chunks of one n-tile
  p.w_resident = 0;
  p.w_bytes = 0;
  if (debug_knob(4) != 1 && mode != MODE_CONVT_DGRAD && p.m_tiles >= 4 * num_sms() &&
      (smem_budget() - fixed - kStagingBytes - w_all) / p.b_off >= 3) {
    p.w_resident = 1;
    p.w_bytes = w_all;
    p.stage_bytes = p.b_off;
  }
  plan_epilogue(smem_budget() - fixed - p.w_bytes, p.stage_bytes, debug_knob(9) == 1, &p.epi_groups, &p.stg_bufs);
  p.n_stages = (smem_budget() - fixed - p.w_bytes - p.epi_groups * p.stg_bufs * kStagingBytes) / p.stage_bytes;
  if (p.n_stages > kMaxStages) p.n_stages = kMaxStages;
  CMU_REQUIRE(p.n_stages >= 2, "k1: shared-memory plan failed (stage %d bytes)", p.stage_bytes);
  const int smem_bytes = p.w_bytes + p.n_stages * p.stage_bytes + p.epi_groups * p.stg_bufs * kStagingBytes + fixed;
  if (debug_knob(3) == 1) p.sched = nullptr;   // A/B switch: static tile schedule
  else if (next_sched_slot(&p.sched, p.n_tiles, stream)) return 1;
  int grid = num_sms();
  const int total_tiles = p.m_tiles * p.n_tiles;
  if (grid > total_tiles) grid = total_tiles;
  grid = (grid / p.n_tiles) * p.n_tiles;
  if (grid < p.n_tiles) grid = p.n_tiles;
  if (stats_grid) *stats_grid = grid;
  return BN == 128 ? launch_k1<128>(p, grid, smem_bytes, stream) : launch_k1<64>(p, grid, smem_bytes, stream);
}

}  // namespace cmu

using namespace cmu;

extern "C" {

// Upper bound on the number of CTAs (= rows of the statistics partial buffer) any K1 launch uses.
int cmu_conv_max_grid(void) { return num_sms(); }

int cmu_conv3x3_fprop(const void* x0, int c0, const void* x1, int c1, int n, int h, int w, const void* w_packed,
                      int cout, void* y, float* stats_partial, int* stats_grid, int* stats_bn, void* stream) {
  return run_k1(MODE_CONV3, x0, c0, x1, c1, n, h, w, w_packed, cout, y, cout, nullptr, 0, nullptr, 0, stats_partial,
                stats_grid, stats_bn, (cudaStream_t)stream);
}

int cmu_conv3x3_dgrad(const void* dy, int cout, int n, int h, int w, const void* w_packed_dgrad, void* dx0, int c0,
                      void* dx1, int c1, void* stream) {
  return run_k1(MODE_CONV3, dy, cout, nullptr, 0, n, h, w, w_packed_dgrad, c0 + c1, dx0, c0, dx1, c1, nullptr, 0,
                nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

// dgrad with the per-CTA column sums of its bf16 outputs (same partial layout as the fprop statistics): the sum over
// the first c0 output channels is the bias gradient of the ConvTranspose2d that produced x0 (munet_neck.py:46-49).
int cmu_conv3x3_dgrad_sums(const void* dy, int cout, int n, int h, int w, const void* w_packed_dgrad, void* dx0, int c0,
                           void* dx1, int c1, float* sums_partial, int* sums_grid, int* sums_bn, void* stream) {
  return run_k1(MODE_CONV3, dy, cout, nullptr, 0, n, h, w, w_packed_dgrad, c0 + c1, dx0, c0, dx1, c1, nullptr, 0,
                sums_partial, sums_grid, sums_bn, (cudaStream_t)stream);
}

int cmu_conv1x1_fprop(const void* x, int cin, int n, int h, int w, const void* w_packed, int cout, const float* bias,
                      void* y, void* stream) {
  return run_k1(MODE_PLAIN, x, cin, nullptr, 0, n, h, w, w_packed, cout, y, cout, nullptr, 0, bias, cout, nullptr,
                nullptr, nullptr, (cudaStream_t)stream);
}

// MoCo queue logits fused with the softmax numerator (moco2_module.py:258-266): y[p][c] = exp((x[p].w[c] - 1) / T) in
// bf16 -- the logits themselves never reach HBM -- plus per-CTA partial column sums of the stored values (the softmax
// denominators), in the layout of the convolution statistics: partial[grid][2][bn].
int cmu_conv1x1_fprop_exp(const void* x, int cin, int n, int h, int w, const void* w_packed, int cout, float inv_temperature,
                          void* y, float* sums_partial, int* sums_grid, int* sums_bn, void* stream) {
  CMU_REQUIRE(inv_temperature > 0.f, "conv1x1_fprop_exp: 1/T must be positive");
  return run_k1(MODE_PLAIN, x, cin, nullptr, 0, n, h, w, w_packed, cout, y, cout, nullptr, 0, nullptr, 0, sums_partial,
                sums_grid, sums_bn, (cudaStream_t)stream, inv_temperature);
}

int cmu_convT2x2_fprop(const void* x, int cin, int n, int h, int w, const void* w_packed, int cout, const float* bias,
                       void* y, void* stream) {
  return run_k1(MODE_CONVT_FPROP, x, cin, nullptr, 0, n, h, w, w_packed, 4 * cout, y, cout, nullptr, 0, bias, cout,
                nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int cmu_convT2x2_dgrad(const void* dy, int cout, int n, int h, int w, const void* w_packed_dgrad, int cin, void* dx,
                       void* stream) {
  return run_k1(MODE_CONVT_DGRAD, dy, cout, nullptr, 0, n, h, w, w_packed_dgrad, cin, dx, cin, nullptr, 0, nullptr, 0,
                nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
