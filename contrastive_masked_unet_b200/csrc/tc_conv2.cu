// K1 pair kernel: the same implicit-GEMM convolution as tc_conv.cu, but two CTAs (a cluster of 2 = one TPC) work on one
// 256-pixel x BN tile with tcgen05.mma.cta_group::2.
//
// Why: with one CTA per 128 x 128 tile every UMMA reads 4 KB of A and 4 KB of B from shared memory for 64 tensor
// cycles (128 B/clk = the whole shared-memory bandwidth of the SM) while TMA is writing the next stage into the same
// shared memory; the 1-CTA kernel therefore saturates at ~50 % tensor-pipe utilisation (profiles/r1_k1_ncu_full.md).
// In pair mode each CTA stages its own 128 pixel rows of A but only HALF of the weight tile (BN/2 rows); the tensor
// cores of the two SMs exchange the B halves, so per SM the shared-memory traffic per MMA drops to 4 KB (A) + 2-4 KB
// (B half) for the same tensor cycles, and the weight tile is fetched from L2 once per pair instead of once per CTA.
//
// Structure per CTA = tc_conv.cu (TMA producer warp, MMA warp, 4 or 8 epilogue warps, double-buffered TMEM accumulators),
// with the pair protocol of the PTX ISA:
//   * both CTAs issue their TMA loads with .cta_group::2; the bytes are accounted on the LEADER's (rank 0) full barrier,
//     which the leader arms with the byte count of both CTAs;
//   * only the leader's MMA warp issues tcgen05.mma.cta_group::2 (M = 256: rows 0-127 in its own TMEM, 128-255 in the
//     peer's); tcgen05.commit.cta_group::2 ... multicast signals the smem-slot-free and accumulator-full barriers of
//     BOTH CTAs;
//   * the epilogue warps of both CTAs hand an accumulator back by arriving (remotely for the peer) on the leader's
//     tmem-empty barrier (count 8).
// Tiles are assigned statically per cluster (both CTAs must walk the same tile sequence).  A cluster-level dynamic
// scheduler (leader pulls super-tiles from a global counter and publishes them to both CTAs through a DSMEM ring) was
// built and measured in round 2: no gain at N = 1 (-0.5 %: the ring costs a GPU-scope release per tile) and sporadic
// launch failures under load, so the static walk stays.
// Also covered: ConvTranspose2d fprop (5-D scatter store) and dgrad (5-D .cta_group::2 gather); N = 64 tiles keep their
// half of the weights resident in shared memory and stream activations only.
#include "k1_common.cuh"
#include "../../include/cmu_b200.h"

namespace cmu {

constexpr int kPairThreads = kK1Threads;

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
    k1_pair_kernel(const __grid_constant__ K1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_res = smem;                        // [w_bytes] this CTA's half of the n-tile's weights (resident mode)
  uint8_t* stage_base = smem + p.w_bytes;
  uint8_t* staging = stage_base + p.n_stages * p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + p.epi_groups * p.stg_bufs * kStagingBytes);
  uint64_t* full_bar = bars;                    // [kMaxStages]  (used in the leader only)
  uint64_t* empty_bar = bars + kMaxStages;      // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]           (used in the leader only, 8 arrivals)
  uint64_t* wfull_bar = tempty_bar + 2;         // [1]           (leader: resident weights of both CTAs have landed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);
  float* s_stats = reinterpret_cast<float*>(bars + 40);  // [2][BN]
  constexpr uint32_t kTmemCols = 2 * BN;
  constexpr int kHalfN = BN / 2;

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int n_stages = p.n_stages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    mbar_init(wfull_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmO0);
  }
  for (int i = threadIdx.x; i < 2 * BN; i += kPairThreads) s_stats[i] = 0.f;
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // static schedule: cluster c owns n-tile (c % n_tiles) and walks 256-pixel super-tiles; CTA `rank` takes m-tile
  // 2*super + rank (a possibly out-of-range last tile is all TMA out-of-bounds: zero loads, clipped stores)
  const int n_clusters = gridDim.x >> 1;
  const int cid = blockIdx.x >> 1;
  const int nt = cid % p.n_tiles;
  const int n0 = nt * BN;
  const int sup0 = cid / p.n_tiles;
  const int sup_stride = n_clusters / p.n_tiles;
  const int n_super = (p.m_tiles + 1) >> 1;
  const int shifts = (p.mode == MODE_CONV3 && !p.single_patch) ? 3 : 1;   // pipeline stages per K chunk
  const int kc = p.kc;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      const uint32_t tx_pair = 2u * (uint32_t)(p.a_bytes + (p.w_resident ? 0 : p.b_bytes));
      if (p.w_resident) {   // one-time load of this CTA's half of the n-tile's weights: kc x shifts boxes
        if (leader) mbar_arrive_expect_tx(wfull_bar, 2u * (uint32_t)p.w_bytes);
        const uint32_t wb = mapa_u32(smem_u32(wfull_bar), 0);
        for (int c = 0; c < kc; ++c)
          for (int s = 0; s < shifts; ++s)
            tma_load_3d_pair(w_res + (c * shifts + s) * p.b_bytes, &p.tmB, wb, c << 6, n0 + (int)rank * kHalfN,
                             (p.mode == MODE_CONV3) ? 3 * s : 0);   // single_patch: one box holds all 9 taps (s = 0)
      }
      uint32_t it = 0;
      for (int sup = sup0; sup < n_super; sup += sup_stride) {
        const int mt = 2 * sup + (int)rank;
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
            if (leader) mbar_arrive_expect_tx(&full_bar[st], tx_pair);
            const uint32_t fb = mapa_u32(smem_u32(&full_bar[st]), 0);  // the leader's full barrier
            const int k0 = c << 6;
            if (p.mode == MODE_CONVT_DGRAD) {   // K = (r, s, co): gather dy at (2h + r, 2w + s) through the 5-D map
              const int per = p.c0 >> 6;
              const int rs = c / per;
              const int cc = (c - rs * per) << 6;
              tma_load_5d_pair(sA, &p.tmA0, fb, (rs & 1) * p.c0 + cc, w0, rs >> 1, h0, img);
            } else {
              const bool second = k0 >= p.c0;
              const CUtensorMap* src = second ? &p.tmA1 : &p.tmA0;
              const int cc = second ? k0 - p.c0 : k0;
              if (p.mode == MODE_CONV3)   // single_patch: box (TW+2) x (TH+2) at (w0-1, h0-1); else column-shifted copies
                tma_load_4d_pair(sA, src, fb, cc, w0 + (p.single_patch ? 0 : s) - 1, h0 - 1, img);
              else
                tma_load_4d_pair(sA, src, fb, cc, w0, h0, img);
            }
            if (!p.w_resident)
              tma_load_3d_pair(sB, &p.tmB, fb, k0, n0 + (int)rank * kHalfN, (p.mode == MODE_CONV3) ? 3 * s : 0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      const uint32_t ta = (p.TW * 128) >> 4;
      const bool conv3 = (p.mode == MODE_CONV3);
      const uint32_t stage0 = smem_u32(stage_base);
      const uint64_t desc_hi = umma_smem_desc(0, 16, 1024);
      // single-patch mode: the tile's 8-pixel row groups sit (TW+2) pixel rows apart inside the haloed patch
      const uint64_t desc_hi_sp = umma_smem_desc(0, 16, (uint32_t)(p.TW + 2) * 128u);
      const bool single_patch = p.single_patch != 0;
      const uint32_t pitch16 = ((uint32_t)(p.TW + 2) * 128u) >> 4;   // patch row pitch in 16-byte units
      uint32_t it = 0, tile_it = 0;
      const uint32_t w_base = smem_u32(w_res);
      if (p.w_resident) mbar_wait(wfull_bar, 0);
      for (int sup = sup0; sup < n_super; sup += sup_stride, ++tile_it) {
        const uint32_t acc = tile_it & 1;
        const uint32_t aph = (tile_it >> 1) & 1;
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
          const uint64_t a0 = (single_patch ? desc_hi_sp : desc_hi) | (uint64_t)((sA >> 4) & 0x3FFF);
          const uint64_t b0 = desc_hi | (uint64_t)((sB >> 4) & 0x3FFF);
          if (elect_one()) {
            if (conv3 && single_patch) {
              // tap (s, r): A starts (r * (TW+2) + s) pixel rows into the patch, B at weight tap 3 s + r
#pragma unroll
              for (int s = 0; s < 3; ++s) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16_pair(d_tmem, a0 + (uint64_t)(r * pitch16 + s * 8 + k * 2),
                                   b0 + (uint64_t)((3 * s + r) * (kHalfN * 8) + k * 2), idesc, (sidx | s | r | k) != 0 ? 1u : 0u);
                }
              }
            } else if (conv3) {
#pragma unroll
              for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_pair(d_tmem, a0 + (uint64_t)(r * ta + k * 2), b0 + (uint64_t)(r * (kHalfN * 8) + k * 2), idesc,
                                 (sidx | r | k) != 0 ? 1u : 0u);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_pair(d_tmem, a0 + (uint64_t)(k * 2), b0 + (uint64_t)(k * 2), idesc, (sidx | k) != 0 ? 1u : 0u);
            }
            umma_commit_pair(&empty_bar[st]);                              // slot free in both CTAs
            if (sidx == nstages - 1) umma_commit_pair(&tfull_bar[acc]);    // accumulators complete in both CTAs
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5 [, 6..9], both CTAs)
    const uint32_t q = warp & 3;
    const uint32_t row = q * 32 + lane;
    const int th = row >> p.tw_shift;
    const int tw = row & (p.TW - 1);
    const int wth0 = (int)(q * 32) >> p.tw_shift;   // first pixel row of this warp inside the tile
    const int wtw0 = (int)(q * 32) & (p.TW - 1);
    const int stg_bufs = p.stg_bufs;
    const uint32_t grp = (warp - 2) >> 2;             // epilogue group = the TMEM accumulator it drains
    const bool two_groups = (p.epi_groups == 2);
    uint8_t* const grp_staging = staging + grp * stg_bufs * kStagingBytes;
    uint32_t tile_it = 0, slab_it = 0;
    for (int sup = sup0; sup < n_super && grp < (uint32_t)p.epi_groups; sup += sup_stride, ++tile_it) {
      if (two_groups && (tile_it & 1) != grp) continue;
      const int mt = 2 * sup + (int)rank;
      const int img = mt / tiles_per_img;
      const int rem = mt - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * p.TH;
      const int w0 = (rem % p.tiles_w) * p.TW;
      const bool valid = (img < p.N) && (h0 + th < p.H) && (w0 + tw < p.W);
      const uint32_t acc = tile_it & 1;
      const uint32_t aph = (tile_it >> 1) & 1;
      const uint32_t valid_rows = __ballot_sync(0xffffffffu, valid);
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
#pragma unroll 1
      for (int slab = 0; slab < BN / 64; ++slab, ++slab_it) {
        // warp-private staging (32 rows x 128 B) and warp-private TMA stores: no CTA-wide barrier in the tile loop
        uint8_t* stg = grp_staging + (slab_it % stg_bufs) * kStagingBytes + q * 4096;
        if (lane == 0) {
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
          if (j == BN / 32 - 1) {  // last read of this accumulator: hand it back to the leader's MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
          }
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          if (p.bias != nullptr) {
            const int cb = (n0 + j * 32) % p.bias_mod;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + cb + i);
          }
          if (p.exp_scale != 0.f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __expf((v[i] - 1.f) * p.exp_scale);
          }
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
          if (p.mode == MODE_CONVT_FPROP) {   // N = (r, s, co): scatter to (2h + r, 2w + s)
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
      // partial row r with r % n_tiles == nt (the layout cmu_bn_finalize expects)
      const size_t srow = ((size_t)(cid / p.n_tiles) * 2 + rank) * p.n_tiles + nt;
      for (int i = threadIdx.x - 64; i < 2 * BN; i += 256) p.stats[srow * 2 * BN + i] = s_stats[i];
    }
  }
  tc_fence_before();
  cluster_sync_all();  // nobody frees TMEM / leaves while the partner may still signal or read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

template <int BN>
static int launch_pair(const K1Params& p, int grid, int smem_bytes, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CMU_CHECK_CUDA(cudaFuncSetAttribute(k1_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  k1_pair_kernel<BN><<<grid, kPairThreads, smem_bytes, stream>>>(p);
  CMU_LAUNCH_CHECK();
  return 0;
}

// Takes the launch when the pair kernel applies (conv3x3 / plain modes, GEMM-N multiple of 128, enough tiles).
// `p` arrives fully prepared for the 1-CTA kernel; tile maps are re-encoded where the box differs (weight half tiles).
int run_k1_pair(K1Params& p, const void* a0, const void* a1, const void* wpk, int ktot, cudaStream_t stream, int* used,
                int* used_bn) {
  *used = 0;
  if (debug_knob(5) == 1) return 0;
  const bool convT = (p.mode == MODE_CONVT_FPROP || p.mode == MODE_CONVT_DGRAD);
  if (convT && debug_knob(11) == 1) return 0;                   // A/B: ConvTranspose on the 1-CTA kernel
  if (p.m_tiles < 2 * num_sms()) return 0;
  if (p.n_total % 128 != 0 && debug_knob(7) == 1) return 0;   // A/B: 64-wide layers on the 1-CTA kernel
  const int BN = (p.n_total % 256 == 0 && debug_knob(6) != 1) ? 256 : (p.n_total % 128 == 0 ? 128 : 64);
  p.n_tiles = p.n_total / BN;
  // weight tensor map with a BN/2-row box
  {
    const int taps = (p.mode == MODE_CONV3) ? 9 : 1;
    uint64_t dims[3] = {(uint64_t)ktot, (uint64_t)p.n_total, (uint64_t)taps};
    uint64_t str[2] = {(uint64_t)ktot * 2, (uint64_t)p.n_total * ktot * 2};
    const bool sp = (p.mode == MODE_CONV3 && p.TW == 8 && BN == 64 && debug_knob(13) != 1);
    uint32_t box[3] = {64, (uint32_t)(BN / 2), (uint32_t)((p.mode == MODE_CONV3) ? (sp ? 9 : 3) : 1)};
    if (encode_tmap_bf16(&p.tmB, wpk, 3, dims, str, box)) return 1;
  }
  // single-patch staging (conv3x3, 8-wide tiles, N = 64 tiles -- the shared-memory-bound layers): one haloed patch per K
  // chunk instead of three column-shifted copies; the nine taps become start-address offsets of the UMMA descriptor
  // (row pitch (TW+2) x 128 B: the hardware applies the 128-byte swizzle to the absolute shared-memory address, so
  // starts that are not 1024-byte aligned read what TMA wrote).  Wider n-tiles keep the three-copy scheme: their nine-
  // tap weight stage (9 x BN/2 x 128 B) would leave no room for a pipeline.  Knob 13 = 1: off (A/B).
  p.single_patch = (p.mode == MODE_CONV3 && p.TW == 8 && BN == 64 && debug_knob(13) != 1) ? 1 : 0;
  if (p.single_patch) {   // activation maps with the haloed (TW+2) x (TH+2) box
    const void* srcs[2] = {a0, a1};
    const int chans[2] = {p.c0, p.c1};
    CUtensorMap* maps[2] = {&p.tmA0, &p.tmA1};
    for (int i = 0; i < 2; ++i) {
      if (srcs[i] == nullptr) continue;
      uint64_t dims[4] = {(uint64_t)chans[i], (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
      uint64_t str[3] = {(uint64_t)chans[i] * 2, (uint64_t)p.W * chans[i] * 2, (uint64_t)p.H * p.W * chans[i] * 2};
      uint32_t box[4] = {64, (uint32_t)(p.TW + 2), (uint32_t)(p.TH + 2), 1};
      if (encode_tmap_bf16(maps[i], srcs[i], 4, dims, str, box)) return 1;
    }
    if (a1 == nullptr) p.tmA1 = p.tmA0;
  }
  p.a_bytes = (p.mode == MODE_CONV3) ? (p.TH + 2) * (p.single_patch ? p.TW + 2 : p.TW) * 128 : 128 * 128;
  p.b_bytes = ((p.mode == MODE_CONV3) ? (p.single_patch ? 9 : 3) : 1) * (BN / 2) * 128;
  p.b_off = (p.a_bytes + 1023) & ~1023;
  p.stage_bytes = p.b_off + p.b_bytes;
  p.w_resident = 0;
  p.w_bytes = 0;
  p.sched = nullptr;
  const int fixed = 1024 + 320 + 2 * BN * 4 + 64;
  // small-weight layers (64-wide n-tiles): keep this CTA's half of the weights resident for the whole persistent loop and
  // stream activations only -- 40 % less TMA fill per tile and a deeper activation pipeline (these layers have 1-2 K
  // chunks per tile, so the TMA latency is hidden by the number of stages in flight, not by the length of a tile)
  const int w_all = p.kc * ((p.mode == MODE_CONV3 && !p.single_patch) ? 3 : 1) * p.b_bytes;
  if (debug_knob(4) != 1 && BN == 64 && w_all <= 80 * 1024 &&
      (smem_budget() - fixed - 2 * kStagingBytes - w_all) / p.b_off >= 5) {
    p.w_resident = 1;
    p.w_bytes = w_all;
    p.stage_bytes = p.b_off;
  }
  plan_epilogue(smem_budget() - fixed - p.w_bytes, p.stage_bytes, debug_knob(9) == 1, &p.epi_groups, &p.stg_bufs);
  p.n_stages = (smem_budget() - fixed - p.w_bytes - p.epi_groups * p.stg_bufs * kStagingBytes) / p.stage_bytes;
  if (p.n_stages > kMaxStages) p.n_stages = kMaxStages;
  CMU_REQUIRE(p.n_stages >= 2, "k1 pair: shared-memory plan failed");
  const int smem_bytes = p.w_bytes + p.n_stages * p.stage_bytes + p.epi_groups * p.stg_bufs * kStagingBytes + fixed;
  int n_clusters = num_sms() / 2;
  const int n_super = (p.m_tiles + 1) / 2;
  if (n_clusters > n_super * p.n_tiles) n_clusters = n_super * p.n_tiles;
  n_clusters = (n_clusters / p.n_tiles) * p.n_tiles;
  if (n_clusters < p.n_tiles) n_clusters = p.n_tiles;
  const int grid = 2 * n_clusters;
  *used = grid;   // number of statistic partial rows
  *used_bn = BN;
  if (BN == 256) return launch_pair<256>(p, grid, smem_bytes, stream);
  if (BN == 128) return launch_pair<128>(p, grid, smem_bytes, stream);
  return launch_pair<64>(p, grid, smem_bytes, stream);
}

}  // namespace cmu
