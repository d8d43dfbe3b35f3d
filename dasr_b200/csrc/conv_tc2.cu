// CTA-pair (tcgen05 cta_group::2) 3x3 convolution for the first launch of a dense block — sm_100a.
//
// Replaces (reference, codes/SRN): the products of ResidualDenseBlock_5C's input x with conv1..conv5
// (models/modules/block.py:262-286) — launch 1 of the N-fused dense-block schedule (engine.SCHED2): K = 64 input
// channels against the 192 stacked output channels [x1 | p2 p3 p4 | p5].
//
// Why a pair: a tcgen05.mma M=128 K=16 costs ~84 cycles back to back for every N <= 128 (the 4 KB A read from shared
// memory is the floor) and 101 cycles at N = 192, but a 192-wide resident filter set (9 taps x 64 x 192 bf16 = 221 KB)
// does not fit one SM, so the single-CTA kernel runs this launch as two Cout tiles of 96 (two CTAs each reading every A
// tile: 2 x 84 cycles per K step and pixel tile).  With cta_group::2 the two SMs of a TPC each keep HALF of the filters
// (96 rows, 110 KB), each loads the A halo tile of its own pixel tile, and one M=256 N=192 instruction issued by the
// leader CTA feeds both tensor cores: ~84-101 cycles per K step for TWO pixel tiles.
//
// Layout per CTA: [filters: 9 taps x nchunks x (N/2 rows x 64 B)] [A stages: halo tile 18x10 px x 64 B, SWIZZLE_64B]
// [output staging ring: 64-channel blocks of 128 px x 128 B, SWIZZLE_128B] [barriers, bias].
// Warp roles (TC2_THREADS = 352): warp 0 TMA producer (filters once, A halo tiles) / warp 1 TMEM allocator + (leader
// only) MMA issuer / warps 2..9 epilogue (TMEM -> registers -> +bias, +pre, activation on the first act_cols columns,
// scale, +residuals -> bf16 block in shared memory) / warp 10 epilogue TMA (pre / residual blocks in, finished blocks out).
// The epilogue works in 64-channel blocks (128 px x 128 B) plus one 32-channel tail block when cout % 64 == 32:
// v = alpha*act(acc + bias + pre) + beta1*res1 + beta2*res2,
// the same contract as dasr_conv_tc's staged epilogue, so every launch of the dense-block schedules can run on a pair.
#include <stdlib.h>
#include "tc_common.cuh"

namespace dasr {

constexpr int TC2_EPI_WARPS = 8;
constexpr int TC2_WARPS = 2 + TC2_EPI_WARPS + 1;
constexpr int TC2_THREADS = 32 * TC2_WARPS;
constexpr int TC2_MAX_STAGES = 6;
constexpr int TC2_MAX_BLOCKS = 6;      // staging ring entries (64-channel output blocks)

struct Tc2Maps {          // TMA descriptors of the epilogue blocks: [out, pre, res1, res2] x [64-channel box, 32-channel tail box]
  CUtensorMap m[8];
};

struct Tc2Args {
  DasrConvTcParams p;
  const float* bias;
  int nchunks;        // cin / 32
  int n_half;         // cout / 2: filter rows resident in one CTA
  int tiles_x, tiles_y;
  long ntiles;
  int stages, a_stage_bytes, w_bytes;
  int nblk;           // staging ring entries
  int nb64;           // 64-channel blocks per tile (cout / 64)
  int nb;             // blocks per tile: nb64 + (cout % 64 == 32 ? one 32-channel tail block : 0)
  int acc_stride;     // TMEM columns between the two accumulators
};

template <bool HAS_PRE, int NRES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ Tc2Maps em, const Tc2Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + a.w_bytes;
  uint8_t* sS = sA + (size_t)a.stages * a.a_stage_bytes;                  // [nblk] 64-channel output blocks (pre addend in place)
  uint8_t* sR1 = sS + (size_t)a.nblk * EPI_BLK64_BYTES;                   // [nblk] res1 blocks
  uint8_t* sR2 = sR1 + (NRES >= 1 ? (size_t)a.nblk * EPI_BLK64_BYTES : 0);   // [nblk] res2 blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR2 + (NRES >= 2 ? (size_t)a.nblk * EPI_BLK64_BYTES : 0));
  constexpr bool HAS_LOADS = HAS_PRE || NRES > 0;
  uint64_t* full_bar = bars;                              // [stages] LEADER: both CTAs' A chunks landed
  uint64_t* empty_bar = bars + TC2_MAX_STAGES;            // [stages] each CTA: A chunk consumed (multicast commit)
  uint64_t* w_bar = bars + 2 * TC2_MAX_STAGES;            // [1]      LEADER: both filter halves landed
  uint64_t* tfull_bar = w_bar + 1;                        // [2]      each CTA: accumulator complete (multicast commit)
  uint64_t* tempty_bar = tfull_bar + 2;                   // [2]      LEADER: accumulator drained by BOTH CTAs' epilogues
  uint64_t* sfull_bar = tempty_bar + 2;                   // [nblk]   block written by the epilogue warps
  uint64_t* sfree_bar = sfull_bar + TC2_MAX_BLOCKS;       // [nblk]   block read by its TMA store
  uint64_t* pre_bar = sfree_bar + TC2_MAX_BLOCKS;         // [nblk]   pre / residual blocks landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pre_bar + TC2_MAX_BLOCKS);
  float* sBias = reinterpret_cast<float*>(bars + 40);     // [cout]  (16-byte aligned: read with ld.shared.v4)

  const DasrConvTcParams& p = a.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const long pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int cout = p.nt;                          // Cout tile of this CTA pair (grid.y walks the tiles; nt == cout: one tile)
  const int co_base = (int)blockIdx.y * p.nt;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_in);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&em.m[0]);
    for (int s = 0; s < a.stages; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(w_bar, 1);
    for (int b = 0; b < 2; b++) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 2 * TC2_EPI_WARPS);
    }
    for (int b = 0; b < a.nblk; b++) {
      mbar_init(&sfull_bar[b], TC2_EPI_WARPS);
      mbar_init(&sfree_bar[b], 1);
      mbar_init(&pre_bar[b], 1);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < cout; i += TC2_THREADS) sBias[i] = a.bias ? a.bias[co_base + i] : 0.f;
  cluster_sync();                       // barriers of both CTAs initialised before any cross-CTA signal
  if (warp == 1) tmem_alloc2(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();              // the next launch may start its prologue on SMs this grid has left

  auto tile_of = [&](long it, int& x0, int& y0, int& n) -> bool {     // tile of THIS CTA in pair-iteration `it`
    uint32_t t = 2u * (uint32_t)(pair + it * npairs) + rank;           // tile counts fit 32 bits (checked on the host)
    const bool live = (long)t < a.ntiles;
    // tile_rev: walk the grid backwards — a launch that re-reads what the previous launch wrote last (partial sums, the
    // activations just produced) then starts with the tiles still resident in the 126 MB L2
    if (p.tile_rev && live) t = (uint32_t)(a.ntiles - 1) - t;
    const uint32_t r = t / (uint32_t)a.tiles_x;
    const uint32_t tx = t - r * (uint32_t)a.tiles_x;
    n = (int)(r / (uint32_t)a.tiles_y); // n >= N for the odd tail tile: TMA zero-fills, nothing is stored
    const uint32_t ty = r - (uint32_t)n * (uint32_t)a.tiles_y;
    x0 = (int)tx * TILE_W;
    y0 = (int)ty * TILE_H;
    return live;
  };
  const long niter = (a.ntiles / 2 + (a.ntiles & 1) - pair + npairs - 1) / npairs;   // pair-iterations of this pair

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      if (rank == 0) mbar_expect_tx(w_bar, 2u * (uint32_t)a.w_bytes);
      for (int tap = 0; tap < 9; tap++)
        for (int c = 0; c < a.nchunks; c++) {
          const int slot = tap * a.nchunks + c;
          const int row = slot * p.cout + co_base + (int)rank * a.n_half;
          tma2_load_2d(sW + (size_t)slot * a.n_half * ROW_B, &tmap_w, w_bar, 0, row);
        }
      pdl_wait();                       // activations are written by the previous launch
      int stage = 0;
      uint32_t phase = 0;
      for (long it = 0; it < niter; it++) {
        int x0, y0, n;
        tile_of(it, x0, y0, n);
        for (int c = 0; c < a.nchunks; c++) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * A_HALO_BYTES);
          tma2_load_4d(sA + (size_t)stage * a.a_stage_bytes, &tmap_in, &full_bar[stage],
                       p.nchunk_list ? p.chunk_off[c] : p.in_coff + c * CHUNK, x0 - 1, y0 - 1, n);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA only) ===========================
    if (rank == 0) {
      const uint32_t idesc = make_idesc_16(256, cout, p.f16);
      const uint32_t a_sbo = (uint32_t)(HALO_W * ROW_B);
      const uint64_t a_hi = ((uint64_t)((a_sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61) | ((uint64_t)1 << 16);
      const uint64_t b_hi = ((uint64_t)(((8 * ROW_B) >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61) | ((uint64_t)1 << 16);
      const uint32_t a_lo0 = smem_u32(sA) >> 4;
      const uint32_t a_stage_lo = (uint32_t)a.a_stage_bytes >> 4;
      const uint32_t b_lo0 = smem_u32(sW) >> 4;
      const uint32_t b_slot_lo = (uint32_t)(a.n_half * ROW_B) >> 4;
      uint32_t tap_lo[9];
#pragma unroll
      for (int t = 0; t < 9; t++) tap_lo[t] = (uint32_t)(((t / 3) * HALO_W + (t % 3)) * ROW_B) >> 4;
      mbar_wait(w_bar, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      for (long it = 0; it < niter; it++) {
        const int acc = (int)(it & 1);
        mbar_wait(&tempty_bar[acc], ((uint32_t)(it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.acc_stride);
        for (int c = 0; c < a.nchunks; c++) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + (uint32_t)stage * a_stage_lo;
            uint32_t b_lo = b_lo0 + (uint32_t)c * b_slot_lo;
            const uint32_t b_tap_step = (uint32_t)a.nchunks * b_slot_lo;
#pragma unroll
            for (int tap = 0; tap < 9; tap++) {
              const uint32_t al = a_lo + tap_lo[tap];
              umma2_bf16(d_tmem, a_hi | (uint64_t)(al & 0x3FFF), b_hi | (uint64_t)(b_lo & 0x3FFF), idesc, (uint32_t)((c | tap) != 0));
              umma2_bf16(d_tmem, a_hi | (uint64_t)((al + 2) & 0x3FFF), b_hi | (uint64_t)((b_lo + 2) & 0x3FFF), idesc, 1u);
              b_lo += b_tap_step;
            }
            umma2_commit_mc(&empty_bar[stage]);                       // both CTAs' A slots reusable
            if (c == a.nchunks - 1) umma2_commit_mc(&tfull_bar[acc]);  // both CTAs' accumulators complete
          }
          __syncwarp();
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == TC2_WARPS - 1) {
    // =========================== epilogue TMA warp ===========================
    // blocks are numbered k = it * nb64 + i over this CTA's tiles; block k uses ring slot k % nblk
    if (lane == 0) {
      pdl_wait();                       // output slots / pre / residual tensors belong to earlier launches until now
      const uint32_t nb = (uint32_t)a.nb, nblk = (uint32_t)a.nblk;
      const uint32_t total = (uint32_t)niter * nb;
      const uint32_t nld = (uint32_t)((HAS_PRE ? 1 : 0) + NRES);
      // mask mode (dgrad): res1 is the ACTIVATION whose sign gates output channels [mask_c0, mask_c1) (LeakyReLU backward of
      // the slot this launch completes); only the blocks that intersect the range are loaded
      const bool mask_mode = (NRES >= 1) && (p.mask_c1 > p.mask_c0);
      auto issue_loads = [&](uint32_t k) {
        const uint32_t b = k % nblk;
        int x0, y0, n;
        const uint32_t kt = k / nb;
        tile_of((long)kt, x0, y0, n);                // tail tile (n >= N): zero-filled boxes still complete the barrier
        const int i = (int)(k - kt * nb);
        const int wide = i < a.nb64 ? 1 : 0;          // 64-channel block | 32-channel tail block
        const int col = co_base + i * 64;
        const bool need_r1 = !mask_mode || (col < p.mask_c1 && col + (wide ? 64 : 32) > p.mask_c0);
        mbar_expect_tx(&pre_bar[b], (nld - (need_r1 ? 0u : 1u)) * (uint32_t)(wide ? EPI_BLK64_BYTES : EPI_BLK32_BYTES));
        if constexpr (HAS_PRE) tma_load_4d(sS + (size_t)b * EPI_BLK64_BYTES, &em.m[wide ? 2 : 3], &pre_bar[b], p.pre_coff + col, x0, y0, n);
        if constexpr (NRES >= 1) {
          if (need_r1) tma_load_4d(sR1 + (size_t)b * EPI_BLK64_BYTES, &em.m[wide ? 4 : 5], &pre_bar[b], p.res1_coff + col, x0, y0, n);
        }
        if constexpr (NRES >= 2) tma_load_4d(sR2 + (size_t)b * EPI_BLK64_BYTES, &em.m[wide ? 6 : 7], &pre_bar[b], p.res2_coff + col, x0, y0, n);
      };
      if constexpr (HAS_LOADS)
        for (uint32_t k = 0; k < nblk && k < total; k++) issue_loads(k);
      auto retire = [&](uint32_t k) {                // block k's store has finished reading its ring slot
        if constexpr (HAS_LOADS) {
          if (k + nblk < total) issue_loads(k + nblk);
        } else {
          mbar_arrive(&sfree_bar[k % nblk]);
        }
      };
      int x0 = 0, y0 = 0, n = 0;
      bool live = false;
      uint32_t i = 0;                                 // block index inside the tile
      for (uint32_t k = 0; k < total; k++) {
        const uint32_t b = k % nblk;
        if (i == 0) live = tile_of((long)(k / nb), x0, y0, n);
        mbar_wait(&sfull_bar[b], (k / nblk) & 1);
        if (live) tma_store_4d(&em.m[(int)i < a.nb64 ? 0 : 1], sS + (size_t)b * EPI_BLK64_BYTES, p.out_coff + co_base + (int)i * 64, x0, y0, n);
        bulk_commit();                                // (an empty group for the tail tile keeps the group count uniform)
        if (k >= 1) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // everything but the newest store has been read
          retire(k - 1);
        }
        if (++i == nb) i = 0;
      }
      if (total > 0) {
        bulk_wait_read0();
        retire(total - 1);
      }
      bulk_wait0();
    }
    __syncwarp();
  } else {
    // =========================== epilogue warps (2..9) ===========================
    // Two warpgroups share every 64-channel block: warpgroup w takes the 16-column groups g with g % 2 == w.
    const int ew = warp - 2;
    const int wg = ew >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int sw128 = m & 7, sw64 = (m >> 1) & 3;
    const int act = p.act, f16 = p.f16;
    const float slope = p.slope, alpha = p.alpha;
    const bool scale = alpha != 1.f;
    const uint32_t sS_u = smem_u32(sS), sR1_u = smem_u32(sR1), sR2_u = smem_u32(sR2), sBias_u = smem_u32(sBias);
    const bool has_bias = a.bias != nullptr;
    uint32_t k = 0;
    for (long it = 0; it < niter; it++) {
      const int acc = (int)(it & 1);
      mbar_wait(&tfull_bar[acc], (uint32_t)(it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * a.acc_stride);
      for (int i = 0; i < a.nb; i++, k++) {
        const int b = (int)(k % (uint32_t)a.nblk);
        if constexpr (HAS_LOADS) {
          mbar_wait(&pre_bar[b], (k / (uint32_t)a.nblk) & 1);                 // pre / residual blocks of this block landed
        } else {
          if (k >= (uint32_t)a.nblk) mbar_wait(&sfree_bar[b], ((k / (uint32_t)a.nblk) & 1) ^ 1);
        }
        const bool wide = i < a.nb64;                         // 64-channel block (128 B rows, SWIZZLE_128B) | 32-channel tail (64 B rows, SWIZZLE_64B)
        const uint32_t rowoff = (uint32_t)m * (wide ? 128u : 64u);
        const uint32_t bS = sS_u + (uint32_t)b * EPI_BLK64_BYTES + rowoff;
        const uint32_t bR1 = sR1_u + (uint32_t)b * EPI_BLK64_BYTES + rowoff;
        const uint32_t bR2 = sR2_u + (uint32_t)b * EPI_BLK64_BYTES + rowoff;
        uint32_t ra[16], rb[16];
        const int c0 = i * 64 + wg * 16, c1 = c0 + 32;        // this warp's 16-column groups of the block (tail: one group)
        tmem_ld16(t_addr + c0, ra);
        if (wide) tmem_ld16(t_addr + c1, rb);
        tmem_ld_wait();
        if (i == a.nb - 1) {            // last TMEM read of this tile: hand the accumulator back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (h == 1 && !wide) break;
          const uint32_t* rr = h ? rb : ra;
          const int cg = h ? c1 : c0;
          const bool do_act = (act != DASR_ACT_NONE) && (co_base + cg + 16 <= p.act_cols);
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j++) v[j] = __uint_as_float(rr[j]);
          if (has_bias) {
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
              const float4 b4 = lds128f(sBias_u + (cg + 4 * j4) * 4);
              v[4 * j4 + 0] += b4.x;
              v[4 * j4 + 1] += b4.y;
              v[4 * j4 + 2] += b4.z;
              v[4 * j4 + 3] += b4.w;
            }
          }
          const int ch = (cg & (wide ? 63 : 31)) >> 3;        // 16-byte chunk index inside the 128 B (64 B) row
          const int sw = wide ? sw128 : sw64;
          const uint32_t o0 = (uint32_t)(((ch) ^ sw) << 4), o1 = (uint32_t)(((ch + 1) ^ sw) << 4);
          if constexpr (HAS_PRE) {
            fma_h16x8(v, lds128(bS + o0), 1.f, f16);
            fma_h16x8(v + 8, lds128(bS + o1), 1.f, f16);
          }
          if (do_act) {
            if (act == DASR_ACT_LRELU) {
#pragma unroll
              for (int j = 0; j < 16; j++) v[j] = fmaxf(v[j], v[j] * slope);
            } else {
#pragma unroll
              for (int j = 0; j < 16; j++) v[j] = fmaxf(v[j], 0.f);
            }
          }
          if (scale) {
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] *= alpha;
          }
          if constexpr (NRES >= 1) {
            if (p.mask_c1 > p.mask_c0) {
              const int co = co_base + cg;
              if (co >= p.mask_c0 && co < p.mask_c1) {      // 16-column groups never straddle the range (multiples of 16)
                const uint4 m0 = lds128(bR1 + o0), m1 = lds128(bR1 + o1);
                const short* ms0 = reinterpret_cast<const short*>(&m0);
                const short* ms1 = reinterpret_cast<const short*>(&m1);
#pragma unroll
                for (int j = 0; j < 8; j++) {               // a positive finite bf16 / half is a positive 16-bit integer
                  if (!(ms0[j] > 0)) v[j] *= p.mask_slope;
                  if (!(ms1[j] > 0)) v[8 + j] *= p.mask_slope;
                }
              }
            } else {
              fma_h16x8(v, lds128(bR1 + o0), p.beta1, f16);
              fma_h16x8(v + 8, lds128(bR1 + o1), p.beta1, f16);
            }
          }
          if constexpr (NRES >= 2) {
            fma_h16x8(v, lds128(bR2 + o0), p.beta2, f16);
            fma_h16x8(v + 8, lds128(bR2 + o1), p.beta2, f16);
          }
          uint4 o[2];
          pack_h16x16(v, o, f16);
          sts128(bS + o0, o[0]);
          sts128(bS + o1, o[1]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfull_bar[b]);
      }
    }
  }

  tc_fence_before();
  cluster_sync();                       // no CTA leaves while its peer may still signal its barriers / read its filters
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// tcgen05.mma cta_group::2 issue-rate probe (selftest only): the leader issues `iters` groups of 18 MMAs
// (M=256, N=n, K=16, bf16; A descriptors walk the 9 tap offsets of a halo tile), back to back, one commit at the end.
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma2_rate_kernel(int n, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  cluster_sync();
  if (threadIdx.x < 32) tmem_alloc2(&tptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tptr;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x < 32 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(256, n);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 16384);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      if (elect_one()) {
#pragma unroll 1
        for (int tap = 0; tap < 9; tap++) {
          uint32_t aa = a0 + (uint32_t)(((tap / 3) * 10 + (tap % 3)) * 64);
#pragma unroll
          for (int k = 0; k < 2; k++)
            umma2_bf16(tb, make_desc_sw64(aa + k * 32, 640), make_desc_sw64(b0 + k * 32, 512), idesc, 1u);
        }
        if (it == iters - 1)
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0);
  }
  tc_fence_before();
  cluster_sync();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc2(tb, 512); }
}

}  // namespace dasr

using namespace dasr;

extern "C" {

static int tc2_plan(const DasrConvTcParams* p, int has_pre, int nres, int* stages_out, int* nblk_out) {
  const int nchunks = p->cin / CHUNK;
  const long w_bytes = (long)9 * nchunks * (p->nt / 2) * ROW_B;
  const int a_stage = (A_HALO_BYTES + 1023) / 1024 * 1024;
  const int bar_bytes = 40 * 8 + 256 * 4 + 64;
  const int per_blk = (1 + nres) * EPI_BLK64_BYTES;
  (void)has_pre;
  static int force_nblk = -1;
  if (force_nblk < 0) {
    const char* e = getenv("DASR_TC2_NBLK");       // experiments: force the epilogue ring depth
    force_nblk = e ? atoi(e) : 0;
  }
  // measured on B200 (selftest fused, K=64 N=192): 4 A stages + 4 blocks 0.137 ms, 3 A stages + 5 blocks 0.171 ms — the A
  // ring matters more than the epilogue ring, so take the deepest epilogue ring that still leaves 4 A stages
  for (int want = 4; want >= 2; want--)
    for (int nblk = TC2_MAX_BLOCKS; nblk >= 2; nblk--) {
      if (force_nblk > 0 && nblk != force_nblk) continue;
      const long avail = (long)SMEM_LIMIT - 1024 - w_bytes - (long)nblk * per_blk - bar_bytes;
      int stages = (int)(avail / a_stage);
      if (stages > TC2_MAX_STAGES) stages = TC2_MAX_STAGES;
      if (stages >= want) {
        *stages_out = stages;
        *nblk_out = nblk;
        return 1;
      }
    }
  return 0;
}

int dasr_conv_tc2_supported(const DasrConvTcParams* p) {
  // pair kernel: plain 3x3 fprop/dgrad geometry, staged bf16 epilogue in 64-channel blocks
  if (!p) return 0;
  if (p->nvar != 1 || p->ntaps != 9 || p->out_mul != 1 || p->epi_mode != 0 || p->a_mode != 0) return 0;
  if (p->nt % 32 != 0 || p->nt < 32 || p->nt > 256 || p->cout % p->nt != 0 || p->cin % CHUNK != 0 || p->cin <= 0) return 0;
  // pre / residual tensors are announced through their channel strides (pre_cs, res1_cs, res2_cs > 0), as dasr_conv_tc2
  // callers fill them; each one costs one more 16 KB block array per ring slot
  int st, nb;
  const int nres = (p->res1_cs > 0 ? 1 : 0) + (p->res2_cs > 0 ? 1 : 0);
  return tc2_plan(p, p->pre_cs > 0 ? 1 : 0, nres, &st, &nb);
}

int dasr_conv_tc2(const void* in, const void* w, const float* bias, const void* pre, const void* res1, const void* res2,
                  void* out, const DasrConvTcParams* p, void* stream) {
  DASR_REQUIRE(p && in && w && out, "conv_tc2: null argument");
  DASR_REQUIRE(p->nvar == 1 && p->ntaps == 9 && p->out_mul == 1 && p->epi_mode == 0 && p->a_mode == 0,
               "conv_tc2: plain 3x3 geometry with the staged epilogue only");
  DASR_REQUIRE(p->nt % 32 == 0 && p->nt >= 32 && p->nt <= 256 && p->cout % p->nt == 0,
               "conv_tc2: the Cout tile nt must be a multiple of 32 in [32, 256] that divides cout (nt=%d cout=%d)", p->nt, p->cout);
  DASR_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0, "conv_tc2: bad dims");
  DASR_REQUIRE((long)p->N * cdiv(p->W, TILE_W) * cdiv(p->H, TILE_H) < (1L << 30), "conv_tc2: too many tiles for 32-bit tile arithmetic");
  DASR_REQUIRE(p->cin > 0 && p->cin % CHUNK == 0, "conv_tc2: cin must be a multiple of 32 (got %d)", p->cin);
  if (p->nchunk_list > 0) {
    DASR_REQUIRE(p->nchunk_list <= 8 && p->nchunk_list * CHUNK == p->cin, "conv_tc2: chunk list must cover cin");
    for (int i = 0; i < p->nchunk_list; i++)
      DASR_REQUIRE(p->chunk_off[i] >= 0 && p->chunk_off[i] % 8 == 0 && p->chunk_off[i] + CHUNK <= p->in_cs, "conv_tc2: chunk_off[%d]", i);
  } else {
    DASR_REQUIRE(p->in_cs % 8 == 0 && p->in_coff % 8 == 0 && p->in_coff + p->cin <= p->in_cs, "conv_tc2: input slice");
  }
  DASR_REQUIRE(p->out_cs % 8 == 0 && p->out_coff % 8 == 0 && p->out_coff + p->cout <= p->out_cs, "conv_tc2: output slice");
  DASR_REQUIRE(p->act_cols % 16 == 0, "conv_tc2: act_cols must be a multiple of 16");
  DASR_REQUIRE(!(res2 && !res1), "conv_tc2: res2 without res1");
  if (p->mask_c1 > p->mask_c0)
    DASR_REQUIRE(res1 && !res2 && p->mask_c0 % 16 == 0 && p->mask_c1 % 16 == 0 && p->mask_c1 <= p->cout,
                 "conv_tc2: mask mode takes the activation as res1 (no res2) and a range of whole 16-channel groups");
  if (pre) DASR_REQUIRE(p->pre_cs % 8 == 0 && p->pre_coff % 8 == 0 && p->pre_coff + p->cout <= p->pre_cs, "conv_tc2: pre slice");
  if (res1) DASR_REQUIRE(p->res1_cs % 8 == 0 && p->res1_coff % 8 == 0 && p->res1_coff + p->cout <= p->res1_cs, "conv_tc2: res1 slice");
  if (res2) DASR_REQUIRE(p->res2_cs % 8 == 0 && p->res2_coff % 8 == 0 && p->res2_coff + p->cout <= p->res2_cs, "conv_tc2: res2 slice");
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("conv_tc2: cuTensorMapEncodeTiled not available");
    return DASR_E_NODRIVER;
  }
  const int has_pre = pre ? 1 : 0, nres = (res1 ? 1 : 0) + (res2 ? 1 : 0);
  Tc2Args a;
  a.p = *p;
  a.bias = bias;
  a.nchunks = p->cin / CHUNK;
  a.n_half = p->nt / 2;
  a.tiles_x = cdiv(p->W, TILE_W);
  a.tiles_y = cdiv(p->H, TILE_H);
  a.ntiles = (long)p->N * a.tiles_x * a.tiles_y;
  a.w_bytes = 9 * a.nchunks * a.n_half * ROW_B;
  a.a_stage_bytes = (A_HALO_BYTES + 1023) / 1024 * 1024;
  a.nb64 = p->nt / 64;
  a.nb = a.nb64 + ((p->nt & 32) ? 1 : 0);
  a.acc_stride = 256;
  const int bar_bytes = 40 * 8 + 256 * 4 + 64;
  int stages = 0, nblk = 0;
  if (!tc2_plan(p, has_pre, nres, &stages, &nblk)) {
    set_error("conv_tc2: filters (%d B per CTA) + %d epilogue block arrays do not fit shared memory", a.w_bytes, 1 + nres);
    return DASR_E_SMEM;
  }
  a.stages = stages;
  a.nblk = nblk;
  size_t smem = 1024 + (size_t)a.w_bytes + (size_t)stages * a.a_stage_bytes + (size_t)nblk * (1 + nres) * EPI_BLK64_BYTES + bar_bytes;
  if (smem < 120 * 1024) smem = 120 * 1024;      // one CTA per SM (the pair allocates all 512 TMEM columns)

  CUtensorMap tm_in, tm_w;
  Tc2Maps em;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)p->in_cs, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
    cuuint64_t gstr[3] = {(cuuint64_t)p->in_cs * 2, (cuuint64_t)p->W * p->in_cs * 2, (cuuint64_t)p->H * p->W * p->in_cs * 2};
    cuuint32_t box[4] = {CHUNK, HALO_W, HALO_H, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     p->in_cs > CHUNK ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc2: cuTensorMapEncodeTiled(input) failed: %d", (int)r); return DASR_E_LAUNCH; }
  }
  {
    cuuint64_t rows = (cuuint64_t)9 * a.nchunks * p->cout;
    cuuint64_t gdim[2] = {CHUNK, rows};
    cuuint64_t gstr[1] = {ROW_B};
    cuuint32_t box[2] = {CHUNK, (cuuint32_t)a.n_half};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc2: cuTensorMapEncodeTiled(filter) failed: %d", (int)r); return DASR_E_LAUNCH; }
  }
  {
    const void* bases[4] = {out, pre, res1, res2};
    const int css[4] = {p->out_cs, p->pre_cs, p->res1_cs, p->res2_cs};
    const char* names[4] = {"output", "pre", "res1", "res2"};
    for (int t = 0; t < 4; t++) {
      em.m[2 * t] = em.m[2 * t + 1] = tm_in;      // placeholders when unused
      if (!bases[t]) continue;
      for (int narrow = 0; narrow < 2; narrow++) {
        if (narrow ? !(p->nt & 32) : (a.nb64 == 0)) continue;
        const int width = narrow ? 32 : 64;
        cuuint64_t gdim[4] = {(cuuint64_t)css[t], (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
        cuuint64_t gstr[3] = {(cuuint64_t)css[t] * 2, (cuuint64_t)p->W * css[t] * 2, (cuuint64_t)p->H * p->W * css[t] * 2};
        cuuint32_t box[4] = {(cuuint32_t)width, TILE_W, TILE_H, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&em.m[2 * t + narrow], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(bases[t]), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, narrow ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         (narrow && css[t] > 32) ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv_tc2: cuTensorMapEncodeTiled(%s) failed: %d", names[t], (int)r); return DASR_E_LAUNCH; }
      }
    }
  }
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const Tc2Maps, const Tc2Args);
  static const KernelFn kernels[6] = {conv_tc2_kernel<false, 0>, conv_tc2_kernel<false, 1>, conv_tc2_kernel<false, 2>,
                                      conv_tc2_kernel<true, 0>,  conv_tc2_kernel<true, 1>,  conv_tc2_kernel<true, 2>};
  const int ki = has_pre * 3 + nres;
  static bool attr_set[6] = {false, false, false, false, false, false};
  if (!attr_set[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kernels[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) { set_error("conv_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return DASR_E_LAUNCH; }
    attr_set[ki] = true;
  }
  long npairs = (a.ntiles + 1) / 2;
  int gx = num_sms() & ~1;
  if ((long)gx > 2 * npairs) gx = (int)(2 * npairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, p->cout / p->nt, 1);
  cfg.blockDim = dim3(TC2_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernels[ki], tm_in, tm_w, em, a);
  if (e != cudaSuccess) {
    set_error("conv_tc2: launch failed: %s", cudaGetErrorString(e));
    return DASR_E_LAUNCH;
  }
  return check_launch("conv_tc2");
}

// selftest-only probe (declared in selftest.cu, not in the public header)
int dasr_probe_mma2_rate(int n, int iters, double* cycles_per_mma) {
  long long* d;
  if (cudaMalloc(&d, 8) != cudaSuccess) return DASR_E_LAUNCH;
  cudaFuncSetAttribute(mma2_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  mma2_rate_kernel<<<num_sms() & ~1, 128, 64 * 1024>>>(n, iters, d);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) { set_error("probe mma2: %s", cudaGetErrorString(e)); return DASR_E_LAUNCH; }
  *cycles_per_mma = (double)h / ((double)iters * 18.0);
  return DASR_OK;
}

}  // extern "C"
