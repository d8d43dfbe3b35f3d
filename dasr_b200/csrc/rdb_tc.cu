// One persistent kernel for a whole ResidualDenseBlock_5C (reference: codes/SRN/models/modules/block.py:254-286) in
// the dense-block N-fused schedule of conv_tc.cu: the five stage "launches" (x -> x1 | partial conv2..5, x1 -> x2 |
// partial conv3..5, ...) run INSIDE one kernel, image chunk by image chunk, separated by grid-wide barriers.
//
// Why: as separate launches over the whole batch every stage streams its 160..64 partial-sum channels through HBM
// (read-modify-write, 400-700 MB per launch, 4.4-4.7 TB/s = DRAM bound with the tensor pipe 30 % busy).  A chunk of
// two 256x256 images is a 50 MB working set: run all five stages on it before moving on and the partial sums never
// leave the 126 MB L2.  Separate launches per chunk cannot do that (fixed launch + pipeline-fill cost of ~10 us
// against ~6 us of work); a grid barrier costs ~2-4 us.
//
// The tile pipeline of a stage is the one of conv_tc_kernel<0, HAS_PRE, NRES> (same MMA order, same epilogue
// arithmetic => bit-identical results): warp 0 TMA producer (resident filters, A halo tiles), warp 1 MMA issuer,
// warps 2..9 epilogue, warp 10 epilogue TMA (pre / residual tiles in, finished tiles out).
// Stage boundary inside a CTA: __syncthreads (everything of the stage drained, shared memory re-carved).
// Grid barrier: warp 10 arrives (release) once its TMA stores are complete; the two warps that read global memory
// produced by other CTAs (warp 0: A tiles, warp 10: pre / residual tiles) wait (acquire + async-proxy fence).
#include "tc_common.cuh"

namespace dasr {
namespace rdb {

constexpr int A_SLOTS = 4;                      // A ring depth (fixed for all stages: the ring state carries over)
constexpr int A_SLOT_BYTES = 12288;             // >= A_HALO_BYTES, 1024-aligned
constexpr int ACC_STRIDE = 256;                 // TMEM columns per accumulator buffer (max nt = 192 < 256)
constexpr int SPIN_LIMIT = 1 << 22;             // bounded grid-barrier spin (x ~100 ns): trap instead of hanging the GPU

struct Stage {
  int nchunks, in_coff;          // input slice: nchunks x 32 channels from in_coff of buffer B
  int nt, n_ntiles;              // Cout columns per CTA tile; 1 or 2 column tiles (stage 1: 2 x 96)
  int out_coff;                  // first destination channel
  int pre_coff;                  // partial sums to extend (same columns as the output); < 0: none
  int act_cols;                  // leading columns that get LeakyReLU
  int nres;                      // residual tiles added after scaling (stage 5): 0..2
  int res1_coff, res2_coff;
  float alpha, beta1, beta2;
  int dst_next;                  // 0: write into buffer B, 1: into the next RDB's buffer
  const float* bias;             // [n_ntiles * nt]
};

struct Args {
  Stage st[5];
  int N, H, W, tiles_x, tiles_y, tiles_per_img, chunk_imgs;
  float slope;
  unsigned int* counter;         // grid-barrier arrivals (zero at launch)
  int* error;                    // set to 1 if a barrier spin ran out
};

struct Maps {
  CUtensorMap in;                // buffer B, box 32 ch x HALO_W x HALO_H, SWIZZLE_64B
  CUtensorMap w[5];              // packed filters of the five stages, box 32 x nt rows
  CUtensorMap b64, b32;          // buffer B, box 64 / 32 ch x TILE_W x TILE_H (SWIZZLE_128B / 64B)
  CUtensorMap n64, n32;          // next buffer
  CUtensorMap r64, r32;          // residual-2 buffer (RRDB input, every third RDB)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// whole warp: wait until `target` CTAs-arrivals are visible, then order the async proxy (TMA) after them
__device__ __forceinline__ void grid_wait(const Args& a, unsigned target) {
  if (target == 0) return;
  int spins = 0;
  while (ld_acquire_u32(a.counter) < target) {
    __nanosleep(64);
    if (++spins > SPIN_LIMIT) {
      *a.error = 1;
      __threadfence_system();
      __trap();
    }
  }
  __syncwarp();
  asm volatile("fence.proxy.async;" ::: "memory");
}

struct Smem {            // fixed part of the carve-up (offsets from the 1024-aligned base)
  uint8_t* sA;           // [A_SLOTS][A_SLOT_BYTES]
  uint64_t* full_bar;    // [A_SLOTS]
  uint64_t* empty_bar;   // [A_SLOTS]
  uint64_t* w_bar;       // [1]
  uint64_t* tfull_bar;   // [2]
  uint64_t* tempty_bar;  // [2]
  uint64_t* pre_bar;     // [2]
  uint64_t* sfull_bar;   // [2]
  uint64_t* sfree_bar;   // [2]
  uint32_t* tmem_ptr;
  float* sBias;          // [256]
  uint8_t* dyn;          // stage-dependent: [W resident][staging x2][res1 x2][res2 x2]
};

// pipeline state that survives stage boundaries (each role keeps the fields it owns)
struct Carry {
  int a_slot;  uint32_t a_phase;      // A ring position (producer and MMA warp walk it identically)
  uint32_t it;                        // tiles processed by this CTA so far (accumulator / staging buffer = it & 1)
  uint32_t n_pre[2], n_free[2];       // completed phases of pre_bar[b] / sfree_bar[b]
  uint32_t n_w;                       // completed phases of w_bar
  unsigned epoch;                     // grid barriers passed
};

template <bool HAS_PRE, int NRES>
__device__ __noinline__ void run_stage(const Maps& maps, const Args& a, const Stage& s, const CUtensorMap* tm_w,
                                       const Smem& sm, Carry& cy, const uint32_t tmem_base, const int img0,
                                       const int nimg, const bool wait_grid, const bool arrive) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = s.nt;
  const int G = (int)gridDim.x / s.n_ntiles;                 // CTAs sharing one column tile
  const int ntile = (s.n_ntiles == 2) ? (int)(blockIdx.x & 1) : 0;
  const int first = (s.n_ntiles == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const long ntiles = (long)nimg * a.tiles_per_img;
  const int w_bytes = 9 * s.nchunks * nt * ROW_B;
  const int epi_bytes = nt * 128 * 2;
  const int nb64 = nt >> 6;
  const bool tail32 = (nt & 32) != 0;
  uint8_t* sW = sm.dyn;
  uint8_t* sS = sW + ((w_bytes + 1023) & ~1023);
  uint8_t* sR1 = sS + 2 * epi_bytes;
  uint8_t* sR2 = sR1 + (NRES >= 1 ? 2 * epi_bytes : 0);
  const CUtensorMap* m_out64 = s.dst_next ? &maps.n64 : &maps.b64;
  const CUtensorMap* m_out32 = s.dst_next ? &maps.n32 : &maps.b32;

  for (int i = threadIdx.x; i < nt; i += TC_THREADS) sm.sBias[i] = s.bias ? s.bias[ntile * nt + i] : 0.f;
  __syncthreads();                       // previous stage drained everywhere in this CTA; bias visible

  auto tile_xyz = [&](long t, int& x0, int& y0, int& n) {
    int tx = (int)(t % a.tiles_x);
    long r = t / a.tiles_x;
    int ty = (int)(r % a.tiles_y);
    n = img0 + (int)(r / a.tiles_y);
    x0 = tx * TILE_W;
    y0 = ty * TILE_H;
  };

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_expect_tx(sm.w_bar, (uint32_t)w_bytes);
      for (int tap = 0; tap < 9; tap++)
        for (int c = 0; c < s.nchunks; c++) {
          const int slot = tap * s.nchunks + c;
          const int row = (tap * s.nchunks + c) * (nt * s.n_ntiles) + ntile * nt;
          tma_load_2d(sW + (size_t)slot * nt * ROW_B, tm_w, sm.w_bar, 0, row);
        }
    }
    if (wait_grid) grid_wait(a, cy.epoch * gridDim.x);
    int slot = cy.a_slot;
    uint32_t phase = cy.a_phase;
    for (long t = first; t < ntiles; t += G) {
      int x0, y0, n;
      tile_xyz(t, x0, y0, n);
      if (lane == 0) {
        for (int c = 0; c < s.nchunks; c++) {
          mbar_wait(&sm.empty_bar[slot], phase ^ 1);
          mbar_expect_tx(&sm.full_bar[slot], A_HALO_BYTES);
          tma_load_4d(sm.sA + (size_t)slot * A_SLOT_BYTES, &maps.in, &sm.full_bar[slot], s.in_coff + c * CHUNK, x0 - 1,
                      y0 - 1, n);
          if (++slot == A_SLOTS) { slot = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_bf16(128, nt);
    const uint32_t a_sbo = (uint32_t)(HALO_W * ROW_B);
    const uint64_t a_hi = ((uint64_t)((a_sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61) | ((uint64_t)1 << 16);
    const uint64_t b_hi = ((uint64_t)(((8 * ROW_B) >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61) | ((uint64_t)1 << 16);
    const uint32_t a_lo0 = smem_u32(sm.sA) >> 4;
    const uint32_t a_slot_lo = (uint32_t)A_SLOT_BYTES >> 4;
    const uint32_t b_lo0 = smem_u32(sW) >> 4;
    const uint32_t b_slot_lo = (uint32_t)(nt * ROW_B) >> 4;
    uint32_t tap_lo[9];
#pragma unroll
    for (int t = 0; t < 9; t++) tap_lo[t] = (uint32_t)(((t / 3) * HALO_W + (t % 3)) * ROW_B) >> 4;
    mbar_wait(sm.w_bar, cy.n_w & 1);
    tc_fence_after();
    int slot = cy.a_slot;
    uint32_t phase = cy.a_phase;
    uint32_t it = cy.it;
    for (long t = first; t < ntiles; t += G, it++) {
      const int acc = it & 1;
      mbar_wait(&sm.tempty_bar[acc], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
      for (int c = 0; c < s.nchunks; c++) {
        mbar_wait(&sm.full_bar[slot], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)slot * a_slot_lo;
          uint32_t b_lo = b_lo0 + (uint32_t)c * b_slot_lo;
          const uint32_t b_tap_step = (uint32_t)s.nchunks * b_slot_lo;
#pragma unroll
          for (int tap = 0; tap < 9; tap++) {
            const uint32_t al = a_lo + tap_lo[tap];
            umma_bf16(d_tmem, a_hi | (uint64_t)(al & 0x3FFF), b_hi | (uint64_t)(b_lo & 0x3FFF), idesc, (uint32_t)((c | tap) != 0));
            umma_bf16(d_tmem, a_hi | (uint64_t)((al + 2) & 0x3FFF), b_hi | (uint64_t)((b_lo + 2) & 0x3FFF), idesc, 1u);
            b_lo += b_tap_step;
          }
          umma_commit(&sm.empty_bar[slot]);
          if (c == s.nchunks - 1) umma_commit(&sm.tfull_bar[acc]);
        }
        __syncwarp();
        if (++slot == A_SLOTS) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 10) {
    // =========================== epilogue TMA warp ===========================
    const int co_base = s.out_coff + ntile * nt;
    const uint32_t load_bytes = (uint32_t)(((HAS_PRE ? 1 : 0) + NRES) * epi_bytes);
    constexpr bool has_loads = HAS_PRE || NRES > 0;
    auto issue_loads = [&](long t, int b) {
      int x0, y0, n;
      tile_xyz(t, x0, y0, n);
      if (lane == 0) {
        mbar_expect_tx(&sm.pre_bar[b], load_bytes);
        const int cb = ntile * nt;
        for (int i = 0; i < nb64 + (tail32 ? 1 : 0); i++) {
          const int wide = i < nb64;
          const int off = wide ? i * EPI_BLK64_BYTES : nb64 * EPI_BLK64_BYTES;
          const int col = cb + (wide ? i * 64 : nb64 * 64);
          if constexpr (HAS_PRE) tma_load_4d(sS + b * epi_bytes + off, wide ? &maps.b64 : &maps.b32, &sm.pre_bar[b], s.pre_coff + col, x0, y0, n);
          if constexpr (NRES >= 1) tma_load_4d(sR1 + b * epi_bytes + off, wide ? &maps.b64 : &maps.b32, &sm.pre_bar[b], s.res1_coff + col, x0, y0, n);
          if constexpr (NRES >= 2) tma_load_4d(sR2 + b * epi_bytes + off, wide ? &maps.r64 : &maps.r32, &sm.pre_bar[b], s.res2_coff + col, x0, y0, n);
        }
      }
      __syncwarp();
    };
    if (has_loads && wait_grid) grid_wait(a, cy.epoch * gridDim.x);
    uint32_t it = cy.it;
    if (has_loads) {
      if (first < ntiles) issue_loads(first, it & 1);
      if (first + G < ntiles) issue_loads(first + G, (it + 1) & 1);
    }
    for (long t = first; t < ntiles; t += G, it++) {
      const int b = it & 1;
      if (lane == 0) {
        int x0, y0, n;
        tile_xyz(t, x0, y0, n);
        mbar_wait(&sm.sfull_bar[b], (it >> 1) & 1);
        for (int i = 0; i < nb64 + (tail32 ? 1 : 0); i++) {
          const int wide = i < nb64;
          const int off = wide ? i * EPI_BLK64_BYTES : nb64 * EPI_BLK64_BYTES;
          tma_store_4d(wide ? m_out64 : m_out32, sS + b * epi_bytes + off, co_base + (wide ? i * 64 : nb64 * 64), x0, y0, n);
        }
        bulk_commit();
        bulk_wait_read0();
        if (!has_loads) mbar_arrive(&sm.sfree_bar[b]);
      }
      __syncwarp();
      if (has_loads && t + 2L * G < ntiles) issue_loads(t + 2L * G, b);
    }
    if (lane == 0) {
      bulk_wait0();                       // this CTA's stores of the stage are complete
      if (arrive) {                       // grid barrier arrive (release): a dependent stage follows
        __threadfence();
        red_release_add(a.counter, 1u);
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue warps (2..9) ===========================
    const int ew = warp - 2;
    const int wg = ew >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ngroups = nt >> 4;
    const float slope = a.slope, alpha = s.alpha;
    const bool scale = alpha != 1.f;
    const int sw64 = (m >> 1) & 3, sw128 = m & 7;
    const uint32_t sS_u = smem_u32(sS), sR1_u = smem_u32(sR1), sR2_u = smem_u32(sR2), sBias_u = smem_u32(sm.sBias);
    const bool has_bias = s.bias != nullptr;
    constexpr bool has_loads = HAS_PRE || NRES > 0;
    uint32_t it = cy.it;
    uint32_t n_pre0 = cy.n_pre[0], n_pre1 = cy.n_pre[1], n_free0 = cy.n_free[0], n_free1 = cy.n_free[1];
    uint32_t used0 = 0, used1 = 0;       // tiles of THIS stage staged through buffer 0 / 1 (sfree only covers this stage)
    for (long t = first; t < ntiles; t += G, it++) {
      const int acc = it & 1;
      const uint32_t bS = sS_u + acc * epi_bytes, bR1 = sR1_u + acc * epi_bytes, bR2 = sR2_u + acc * epi_bytes;
      mbar_wait(&sm.tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
      if (has_loads) {
        uint32_t& np = acc ? n_pre1 : n_pre0;
        mbar_wait(&sm.pre_bar[acc], np & 1);
        np++;
      } else {
        uint32_t& used = acc ? used1 : used0;
        uint32_t& nf = acc ? n_free1 : n_free0;
        if (used > 0) {                   // the stores of this stage's previous tile in this buffer have read it
          mbar_wait(&sm.sfree_bar[acc], nf & 1);
          nf++;
        }
        used++;
      }
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_STRIDE);
      auto process = [&](const uint32_t* rr, int g) {
        const int cg = g << 4;
        const bool do_act = (ntile * nt + cg + 16 <= s.act_cols);
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
        int o0, o1;
        if (cg < (nb64 << 6)) {
          const int base = (cg >> 6) * EPI_BLK64_BYTES + m * 128;
          const int c0 = (cg & 63) >> 3;
          o0 = base + ((c0 ^ sw128) << 4);
          o1 = base + (((c0 + 1) ^ sw128) << 4);
        } else {
          const int base = nb64 * EPI_BLK64_BYTES + m * 64;
          const int c0 = (cg & 31) >> 3;
          o0 = base + ((c0 ^ sw64) << 4);
          o1 = base + (((c0 + 1) ^ sw64) << 4);
        }
        if constexpr (HAS_PRE) {
          fma_bf16x8(v, lds128(bS + o0), 1.f);
          fma_bf16x8(v + 8, lds128(bS + o1), 1.f);
        }
        if (do_act) {
#pragma unroll
          for (int j = 0; j < 16; j++) v[j] = fmaxf(v[j], v[j] * slope);
        }
        if (scale) {
#pragma unroll
          for (int j = 0; j < 16; j++) v[j] *= alpha;
        }
        if constexpr (NRES >= 1) {
          fma_bf16x8(v, lds128(bR1 + o0), s.beta1);
          fma_bf16x8(v + 8, lds128(bR1 + o1), s.beta1);
        }
        if constexpr (NRES >= 2) {
          fma_bf16x8(v, lds128(bR2 + o0), s.beta2);
          fma_bf16x8(v + 8, lds128(bR2 + o1), s.beta2);
        }
        uint4 o[2];
        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
        for (int j = 0; j < 8; j++) ob[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        sts128(bS + o0, o[0]);
        sts128(bS + o1, o[1]);
      };
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.tempty_bar[acc]);
      };
      uint32_t ra[16], rb[16];
      if (wg < ngroups) tmem_ld16(t_addr + wg * 16, ra);
      for (int g = wg; g < ngroups; g += 4) {
        tmem_ld_wait();
        if (g + 2 < ngroups) tmem_ld16(t_addr + (g + 2) * 16, rb); else release_acc();
        process(ra, g);
        if (g + 2 < ngroups) {
          tmem_ld_wait();
          if (g + 4 < ngroups) tmem_ld16(t_addr + (g + 4) * 16, ra); else release_acc();
          process(rb, g + 2);
        }
      }
      if (wg >= ngroups) release_acc();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.sfull_bar[acc]);
    }
  }

  // ---- stage end: the CTA joins; every thread advances the carried pipeline state the same way (all roles walked the
  //      same tile sequence): A ring by tiles x chunks, tile counter, one phase per tile on the barriers a tile uses ----
  __syncthreads();
  {
    uint32_t nmine = 0, cnt[2] = {0, 0};
    for (long t = first; t < ntiles; t += G) {
      cnt[(cy.it + nmine) & 1]++;
      nmine++;
    }
    const uint32_t lin = (uint32_t)cy.a_slot + nmine * (uint32_t)s.nchunks;
    cy.a_phase ^= (lin / A_SLOTS) & 1;
    cy.a_slot = (int)(lin % A_SLOTS);
    constexpr bool has_loads = HAS_PRE || NRES > 0;
    if (has_loads) { cy.n_pre[0] += cnt[0]; cy.n_pre[1] += cnt[1]; }
    else { cy.n_free[0] += cnt[0]; cy.n_free[1] += cnt[1]; }
    cy.it += nmine;
    cy.n_w++;
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
rdb_tc_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Smem sm;
  sm.sA = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + A_SLOTS * A_SLOT_BYTES);
  sm.full_bar = bars;
  sm.empty_bar = bars + A_SLOTS;
  sm.w_bar = bars + 2 * A_SLOTS;
  sm.tfull_bar = bars + 2 * A_SLOTS + 1;
  sm.tempty_bar = bars + 2 * A_SLOTS + 3;
  sm.pre_bar = bars + 2 * A_SLOTS + 5;
  sm.sfull_bar = bars + 2 * A_SLOTS + 7;
  sm.sfree_bar = bars + 2 * A_SLOTS + 9;
  sm.tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * A_SLOTS + 11);
  sm.sBias = reinterpret_cast<float*>(bars + 2 * A_SLOTS + 12);
  sm.dyn = base + A_SLOTS * A_SLOT_BYTES + 2048;          // barriers + bias (256 floats) fit in 2 KB; stays 1024-aligned

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.in);
    for (int i = 0; i < 5; i++) tma_prefetch_desc(&maps.w[i]);
    tma_prefetch_desc(&maps.b64);
    tma_prefetch_desc(&maps.b32);
    for (int i = 0; i < A_SLOTS; i++) {
      mbar_init(&sm.full_bar[i], 1);
      mbar_init(&sm.empty_bar[i], 1);
    }
    mbar_init(sm.w_bar, 1);
    for (int b = 0; b < 2; b++) {
      mbar_init(&sm.tfull_bar[b], 1);
      mbar_init(&sm.tempty_bar[b], 8);
      mbar_init(&sm.pre_bar[b], 1);
      mbar_init(&sm.sfull_bar[b], 8);
      mbar_init(&sm.sfree_bar[b], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(sm.tmem_ptr, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  Carry cy;
  cy.a_slot = 0; cy.a_phase = 0; cy.it = 0; cy.n_w = 0; cy.epoch = 0;
  cy.n_pre[0] = cy.n_pre[1] = cy.n_free[0] = cy.n_free[1] = 0;

  for (int img0 = 0; img0 < a.N; img0 += a.chunk_imgs) {
    const int nimg = min(a.chunk_imgs, a.N - img0);
    for (int si = 0; si < 5; si++) {
      const Stage& s = a.st[si];
      const bool wait_grid = si > 0;                   // stage 1 of a chunk only reads what earlier kernels wrote
      if (s.pre_coff < 0) run_stage<false, 0>(maps, a, s, &maps.w[si], sm, cy, tmem_base, img0, nimg, wait_grid, si < 4);
      else if (s.nres == 0) run_stage<true, 0>(maps, a, s, &maps.w[si], sm, cy, tmem_base, img0, nimg, wait_grid, si < 4);
      else if (s.nres == 1) run_stage<true, 1>(maps, a, s, &maps.w[si], sm, cy, tmem_base, img0, nimg, wait_grid, si < 4);
      else run_stage<true, 2>(maps, a, s, &maps.w[si], sm, cy, tmem_base, img0, nimg, wait_grid, si < 4);
      if (si < 4) cy.epoch++;        // warp 10 arrived at the grid barrier inside run_stage
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace rdb
}  // namespace dasr

using namespace dasr;

extern "C" {

int dasr_rdb_tc(void* buf, void* buf_next, const void* buf_res2, const void* const* w_packed, const float* const* bias,
                const DasrRdbParams* p, unsigned int* counter, int* error_flag, void* stream) {
  DASR_REQUIRE(buf && buf_next && w_packed && bias && p && counter && error_flag, "rdb_tc: null argument");
  DASR_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->nf == 64 && p->gc == 32, "rdb_tc: needs nf=64, gc=32");
  DASR_REQUIRE(p->cs == 256 && p->next_cs >= 64 && p->next_cs % 8 == 0, "rdb_tc: buffer strides");
  DASR_REQUIRE(p->chunk_imgs >= 1, "rdb_tc: chunk_imgs");
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("rdb_tc: cuTensorMapEncodeTiled not available");
    return DASR_E_NODRIVER;
  }
  const int nf = 64, gc = 32, CS = 192;     // buffer: [x 0:64 | x1..x4 64:192 | p5 192:256]
  rdb::Args a;
  // stage j: input chunk, stacked filters of all convs k >= j
  const int nts[5] = {96, 160, 128, 96, 64};
  for (int j = 0; j < 5; j++) {
    rdb::Stage& s = a.st[j];
    s.nchunks = (j == 0) ? 2 : 1;
    s.in_coff = (j == 0) ? 0 : nf + (j - 1) * gc;
    s.nt = nts[j];
    s.n_ntiles = (j == 0) ? 2 : 1;
    s.out_coff = nf + j * gc;
    s.pre_coff = (j == 0) ? -1 : nf + j * gc;
    s.act_cols = (j < 4) ? gc : 0;
    s.nres = 0;
    s.res1_coff = s.res2_coff = 0;
    s.alpha = 1.f;
    s.beta1 = s.beta2 = 0.f;
    s.dst_next = 0;
    s.bias = bias[j];
  }
  {
    rdb::Stage& s = a.st[4];                 // x4 -> conv5 complete: alpha*(acc + b5 + p5) + beta1*x (+ beta2*x_rrdb)
    s.pre_coff = CS;
    s.out_coff = p->next_coff;
    s.dst_next = 1;
    s.alpha = p->alpha;
    s.nres = buf_res2 ? 2 : 1;
    s.res1_coff = 0;
    s.beta1 = p->beta1;
    s.res2_coff = p->res2_coff;
    s.beta2 = p->beta2;
  }
  a.N = p->N; a.H = p->H; a.W = p->W;
  a.tiles_x = cdiv(p->W, TILE_W);
  a.tiles_y = cdiv(p->H, TILE_H);
  a.tiles_per_img = a.tiles_x * a.tiles_y;
  a.chunk_imgs = p->chunk_imgs;
  a.slope = p->slope;
  a.counter = counter;
  a.error = error_flag;

  rdb::Maps mp;
  auto act_map = [&](CUtensorMap* tm, const void* base, int cs, int width, int bw, int bh, bool halo) -> int {
    cuuint64_t gdim[4] = {(cuuint64_t)cs, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
    cuuint64_t gstr[3] = {(cuuint64_t)cs * 2, (cuuint64_t)p->W * cs * 2, (cuuint64_t)p->H * p->W * cs * 2};
    cuuint32_t box[4] = {(cuuint32_t)width, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    (void)halo;
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, width == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     width == 32 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("rdb_tc: cuTensorMapEncodeTiled failed: %d", (int)r);
      return DASR_E_LAUNCH;
    }
    return DASR_OK;
  };
  int rc;
  if ((rc = act_map(&mp.in, buf, p->cs, CHUNK, HALO_W, HALO_H, true))) return rc;
  if ((rc = act_map(&mp.b64, buf, p->cs, 64, TILE_W, TILE_H, false))) return rc;
  if ((rc = act_map(&mp.b32, buf, p->cs, 32, TILE_W, TILE_H, false))) return rc;
  if ((rc = act_map(&mp.n64, buf_next, p->next_cs, 64, TILE_W, TILE_H, false))) return rc;
  if ((rc = act_map(&mp.n32, buf_next, p->next_cs, 32, TILE_W, TILE_H, false))) return rc;
  const void* r2 = buf_res2 ? buf_res2 : buf;
  const int r2cs = buf_res2 ? p->res2_cs : p->cs;
  if ((rc = act_map(&mp.r64, r2, r2cs, 64, TILE_W, TILE_H, false))) return rc;
  if ((rc = act_map(&mp.r32, r2, r2cs, 32, TILE_W, TILE_H, false))) return rc;
  for (int j = 0; j < 5; j++) {
    const rdb::Stage& s = a.st[j];
    DASR_REQUIRE(w_packed[j] && (reinterpret_cast<uintptr_t>(w_packed[j]) & 15) == 0, "rdb_tc: filter pointer %d", j);
    cuuint64_t rows = (cuuint64_t)9 * s.nchunks * s.nt * s.n_ntiles;
    cuuint64_t gdim[2] = {CHUNK, rows};
    cuuint64_t gstr[1] = {ROW_B};
    cuuint32_t box[2] = {CHUNK, (cuuint32_t)s.nt};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&mp.w[j], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed[j]), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("rdb_tc: cuTensorMapEncodeTiled(filter %d) failed: %d", j, (int)r);
      return DASR_E_LAUNCH;
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(rdb::rdb_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) {
      set_error("rdb_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return DASR_E_LAUNCH;
    }
    attr_set = true;
  }
  int grid = num_sms();
  if (grid & 1) grid--;                        // stage 1 pairs CTAs over its two column tiles
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(counter, 0, sizeof(unsigned int), st);
  rdb::rdb_tc_kernel<<<grid, TC_THREADS, SMEM_LIMIT, st>>>(mp, a);
  return check_launch("rdb_tc");
}

}  // extern "C"
