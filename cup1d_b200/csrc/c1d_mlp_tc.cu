// Stage 2 on the 5th-generation tensor cores: the LaCE-style MLP as a chain of
// tcgen05.mma (kind::tf32) GEMMs over a 128-row tile of (walker, z) pairs, in 3xTF32
// (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32 accumulation in TMEM).
//
// One persistent CTA per SM, 4 TC_EPI_PER_QUAD + 2 warps (14 by default):
//   epilogue warps  TC_EPI_PER_QUAD threads per row / TMEM lane (taking 16-column batches in turn): input
//              normalisation, per-layer epilogue TMEM -> registers -> bias + LeakyReLU -> hi/lo split ->
//              next A operand; every layer, including 6->10 and 50->6, is a tensor-core GEMM
//   MMA warp   TMEM allocation; lane 0 issues every tcgen05.mma and tcgen05.commit
//   streamer   lane 0 streams the pre-swizzled weight images with cp.async.bulk (TMA bulk copy) through a
//              byte-granular 96 KB ring with 16 mbarrier pairs; units are (layer, n-block, k-range):
//              32 k in SWIZZLE_128B rows for layers up to 128 wide, 16 k in SWIZZLE_64B rows with
//              n-blocks up to 256 (one tcgen05.mma spans the whole layer) otherwise.  With
//              C1D_TC_CLUSTER > 1 the CTAs of a cluster fetch each unit once and multicast it.
//
// Operand placement (227 KB of shared memory cannot hold hi and lo of a 128 x 256 tile):
//   A_hi   shared memory, K-major SWIZZLE_128B, 8 chunks of 32 k            (SS form)
//   A_lo   tensor memory columns [256, 512), lane = row, column = k         (TS form)
//   B_hi, B_lo  shared memory ring, hi rows followed by lo rows inside a unit
//   D      tensor memory columns [0, 256)
//
// Measured limits (tools/tc_issue_probe.py): one kind::tf32 MMA (M=128, K=8) holds the tensor pipe
// for max(92, 128 N / 256 + 11) cycles with A in TMEM and 108 / 140 / 172 cycles at N = 128 / 192 / 256
// with A in shared memory; a tile needs ~290 of them (41 k cycles) plus a 14-18 k cycle epilogue that
// cannot overlap, because TMEM holds exactly one tile (256 D + 256 A_lo columns).  Measured without gain:
// cluster multicast of the weight stream, and folding a_hi . b_hi and a_hi . b_lo of the narrow layers
// into one MMA over the stacked [b_hi; b_lo] rows (2 instead of 3 MMAs per k-step; layer times unchanged).
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "c1d_internal.h"
#include "c1d_tc_common.cuh"

namespace {

constexpr int TC_ROWS = 128;
constexpr int TC_KCH = 32;                         // tf32 elements per 128-byte swizzle row
constexpr int TC_A_CHUNK_BYTES = TC_ROWS * 128;    // 16 KB
constexpr int TC_MAX_KCHUNKS = 8;                  // K <= 256
constexpr int TC_A_BYTES = TC_MAX_KCHUNKS * TC_A_CHUNK_BYTES;
#ifndef C1D_TC_BK
#define C1D_TC_BK 16
#endif
#ifndef C1D_TC_NBLK
#define C1D_TC_NBLK 256
#endif
constexpr int TC_BK = C1D_TC_BK;                   // k per weight unit: 32 (SWIZZLE_128B rows) or 16 (SWIZZLE_64B)
constexpr int TC_NBLK = C1D_TC_NBLK;               // widest n-block = N of one tcgen05.mma (128 or 256)
constexpr int TC_RING_BYTES = 98304;               // weight ring: units of different size packed back to back
constexpr int TC_NBAR = 16;                        // mbarrier pairs, used round-robin by the units
static_assert(TC_NBLK * TC_BK * 4 * 2 * 2 <= TC_RING_BYTES, "weight ring too small for two of the largest units");
#ifndef TC_EPI_PER_QUAD
#define TC_EPI_PER_QUAD 3  // measured: epilogue 18.3k -> 15.2k -> 13.7k cycles per tile for 2, 3, 4 (4 needs <= 112 registers)
#endif
constexpr int TC_EPI_WARPS = 4 * TC_EPI_PER_QUAD;   // TC_EPI_PER_QUAD warps per TMEM lane quadrant
constexpr int TC_BSTEP = 16 * TC_EPI_PER_QUAD;      // columns between consecutive 16-column batches of one warp
constexpr int TC_THREADS = 32 * (TC_EPI_WARPS + 2);
constexpr int TC_MMA_TID = 32 * TC_EPI_WARPS;      // lane 0 of the MMA warp
constexpr int TC_LOAD_TID = 32 * (TC_EPI_WARPS + 1);
constexpr int TC_MAX_UNITS = 96;
constexpr int TC_MAX_TL = C1D_MAX_LAYERS;
constexpr uint32_t TC_ALO_COL = 256;
constexpr int TC_BAR_BYTES = 16 * TC_NBAR + 64;
constexpr size_t TC_SMEM = 1024 + TC_A_BYTES + TC_RING_BYTES + TC_BAR_BYTES;
static_assert(TC_SMEM <= 232448, "shared memory budget");

struct TcUnit {
  uint32_t w_off;      // byte offset of the unit in the weight image
  uint16_t rows;       // rows of the n-block (multiple of 16, <= 128)
  uint16_t dcol;       // first D column of the n-block
  uint8_t layer;       // tensor-layer index
  uint8_t kstep0;      // first k-step (of 8) of this unit
  uint8_t ksteps;      // k-steps issued in this unit (1 or 2)
  uint8_t flags;       // 1 = first unit of its (layer, n-block), 2 = last unit of its layer,
                       // 4 = 32-k unit in SWIZZLE_128B rows (else 16-k unit in SWIZZLE_64B rows)
  uint16_t ring_kb;    // offset of the unit in the weight ring, in KiB
  uint8_t back;        // the streamer waits for the completion of the unit `back` places earlier (1..TC_NBAR)
  uint8_t pad;
};
static_assert(sizeof(TcUnit) == 16, "TcUnit layout");

struct TcParams {
  const uint8_t* wimg;
  const float* bias;        // tensor layers, padded, concatenated
  const double *emu_lo, *emu_rng_inv;  // input normalisation: (x - lo) * 1/(hi - lo) - 0.5
  int n_units, n_tl;
  int n_in, last_out;
  int tl_npad[TC_MAX_TL];
  int tl_bias_off[TC_MAX_TL];
  int tl_unit_begin[TC_MAX_TL + 1];
  double slope;
  TcUnit units[TC_MAX_UNITS];  // in the kernel-parameter constant bank: uniform loads for the MMA issuer
  int debug;  // C1D_TC_DEBUG=1: block 0 prints per-phase cycle counts of its second tile
};

struct TcState {
  TcParams p;
  void *d_wimg, *d_bias, *d_rng_inv;
  bool ready;
  int cluster;  // CTAs per cluster sharing the weight stream (C1D_TC_CLUSTER: 1, 2 or 4; default 1)
};

using namespace c1d_tc;  // PTX helpers: c1d_tc_common.cuh

// split 16 fp32 activations into tf32 hi (to shared memory, swizzled) and lo (to TMEM)
__device__ __forceinline__ void store_act16(uint32_t a_smem, uint32_t tmem_lane_base, uint32_t row, int col0, const float* v) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint32_t b = __float_as_uint(v[i]);
    hi[i] = b & 0xffffe000u;
    lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
  }
  const uint32_t chunk = (uint32_t)col0 >> 5, g0 = ((uint32_t)col0 & 31u) >> 2;
  const uint32_t cbase = a_smem + chunk * TC_A_CHUNK_BYTES;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t addr = cbase + sw128_off(row, g0 + q);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hi[4 * q]), "r"(hi[4 * q + 1]), "r"(hi[4 * q + 2]),
                 "r"(hi[4 * q + 3])
                 : "memory");
  }
  tmem_st16(tmem_lane_base + TC_ALO_COL + (uint32_t)col0, lo);
}

// ------------------------------------------------------------------------------------
// the MLP kernel
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
k_mlp_tc(TcParams p, const double* __restrict__ emu_in, int64_t nrows, double* __restrict__ out, int out_stride) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base;
  const uint32_t b_smem = base + TC_A_BYTES;
  const uint32_t bar0 = b_smem + TC_RING_BYTES;
  // barriers: b_full[s], b_empty[s], d_full, a_ready, then the TMEM pointer
  const uint32_t bar_bfull = bar0, bar_bempty = bar0 + 8 * TC_NBAR;
  const uint32_t bar_dfull = bar0 + 16 * TC_NBAR, bar_aready = bar_dfull + 8;
  const uint32_t tmem_slot = bar_aready + 8;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

  __shared__ long long dbg_t[2 * TC_MAX_TL];
  // warp index through a shuffle: warp-uniform role branches keep the MMA issuer's descriptors in uniform registers
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  // CTAs of a cluster walk the same weight stream in lockstep: every unit is fetched from L2 once
  // per cluster (each CTA loads 1/csize of it and multicasts the slice to all)
  const uint32_t csize = cluster_nctarank(), crank = cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  if (tid == 0) {
    for (int s = 0; s < TC_NBAR; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, csize);  // one tcgen05.commit arrival from every CTA of the cluster
    }
    mbar_init(bar_dfull, 1);
    mbar_init(bar_aready, 32 * TC_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of every CTA are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_p, 0);

  const int64_t ntiles = (nrows + TC_ROWS - 1) / TC_ROWS;
  // the same trip count in every CTA keeps the cluster in lockstep (surplus tiles hold no rows)
  const int64_t tiles_per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
  const float slope = (float)p.slope;

  if (warp < TC_EPI_WARPS) {
    // ============== row owners / epilogue: TC_EPI_PER_QUAD warps per lane quadrant ==============
    // warp w works on TMEM lanes 32 (w & 3) .. +31 (one row per thread) and on the 16-column
    // batches congruent to (w >> 2) modulo TC_EPI_PER_QUAD
    const uint32_t quad = (uint32_t)warp & 3u, half = (uint32_t)warp >> 2;
    const uint32_t row = quad * 32u + ((uint32_t)tid & 31u);
    const uint32_t tmem_lane = tmem_base + ((quad * 32u) << 16);
    uint32_t d_phase = 0;
    for (int64_t it = 0; it < tiles_per_cta; ++it) {
      const int64_t tile = it * gridDim.x + blockIdx.x;
      const int64_t grow = tile * TC_ROWS + row;
      const bool dbg = p.debug && blockIdx.x == 0 && tid == 0 && it == 1;
      long long tq0 = dbg ? clock64() : 0, t_mma = 0, t_epi = 0;
      // ---- normalised inputs -> A operand of the first layer (k padded to 16 with zeros) ----
      if (half == 0) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          double xn = 0.0;
          if (j < C1D_MAX_EMU && j < p.n_in) {
            const double xv = (grow < nrows) ? emu_in[grow * C1D_MAX_EMU + j] : 0.0;
            xn = (xv - __ldg(p.emu_lo + j)) * __ldg(p.emu_rng_inv + j) - 0.5;
          }
          v[j] = (float)xn;
        }
        store_act16(a_smem, tmem_lane, row, 0, v);
      }
      fence_async_smem();
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bar_aready);

      for (int tl = 0; tl < p.n_tl; ++tl) {
        long long tq1 = dbg ? clock64() : 0;
        mbar_wait(bar_dfull, d_phase);
        d_phase ^= 1;
        tc_fence_after();
        long long tq2 = dbg ? clock64() : 0;
        const int npad = p.tl_npad[tl];
        const float* bias = p.bias + p.tl_bias_off[tl];
        if (tl < p.n_tl - 1) {
          // TMEM -> registers is double buffered: the load of the next 16 columns is in flight
          // while the current 16 are biased, activated, split and stored
          auto process = [&](const uint32_t* r, int col0) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float t = __uint_as_float(r[i]) + __ldg(bias + col0 + i);
              v[i] = (t > 0.f) ? t : slope * t;
            }
            store_act16(a_smem, tmem_lane, row, col0, v);
          };
          uint32_t r0[16], r1[16];
          const int c_first = 16 * (int)half;
          if (c_first < npad) tmem_ld16(tmem_lane + (uint32_t)c_first, r0);
          for (int col0 = c_first; col0 < npad; col0 += 2 * TC_BSTEP) {
            tc_wait_ld();
            const bool more1 = col0 + TC_BSTEP < npad;
            if (more1) tmem_ld16(tmem_lane + (uint32_t)(col0 + TC_BSTEP), r1);
            process(r0, col0);
            if (more1) {
              tc_wait_ld();
              if (col0 + 2 * TC_BSTEP < npad) tmem_ld16(tmem_lane + (uint32_t)(col0 + 2 * TC_BSTEP), r0);
              process(r1, col0 + TC_BSTEP);
            }
          }
          fence_async_smem();
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_aready);
        } else {
          // ---- output layer: coefficients = D + bias (no activation) ----
          if (half == 0) {
            uint32_t r[16];
            tmem_ld16(tmem_lane, r);
            tc_wait_ld();
            if (grow < nrows) {
#pragma unroll
              for (int o = 0; o < C1D_MAX_COEF; ++o)
                if (o < p.last_out) out[grow * out_stride + o] = (double)(__uint_as_float(r[o]) + __ldg(bias + o));
            }
          }
          tc_fence_before();
        }
        if (dbg) { t_mma += tq2 - tq1; t_epi += clock64() - tq2; dbg_t[2 * tl] = tq2 - tq1; dbg_t[2 * tl + 1] = clock64() - tq2; }
      }
      if (dbg) {
        printf("tc tile cycles: mma-wait %lld  epilogue %lld  total %lld |", t_mma, t_epi, clock64() - tq0);
        for (int q = 0; q < p.n_tl; ++q) printf(" L%d %lld/%lld", q, dbg_t[2 * q], dbg_t[2 * q + 1]);
        printf("\n");
      }
    }
  } else if (warp == TC_EPI_WARPS) {
    // ================================ MMA issuer ================================
    // the whole warp runs the loop (warp-uniform values stay in uniform registers, the unit
    // table comes from the kernel-parameter constant bank); one elected lane issues
    const bool leader = elect_one();
    uint32_t a_phase = 0;
    uint32_t g = 0;  // running unit count: barrier pair g % TC_NBAR, phase (g / TC_NBAR) & 1
    for (int64_t it = 0; it < tiles_per_cta; ++it) {
      for (int tl = 0; tl < p.n_tl; ++tl) {
        mbar_wait(bar_aready, a_phase);
        a_phase ^= 1;
        tc_fence_after();
        const int u_end = p.tl_unit_begin[tl + 1];
        for (int u = p.tl_unit_begin[tl]; u < u_end; ++u, ++g) {
          const uint32_t rows = p.units[u].rows, dcol = p.units[u].dcol;
          const uint32_t kstep0 = p.units[u].kstep0, nks = p.units[u].ksteps, uflags = p.units[u].flags;
          const uint32_t stage = g % TC_NBAR, b_phase = (g / TC_NBAR) & 1u;
          mbar_wait(bar_bfull + 8 * stage, b_phase);
          tc_fence_after();
          const uint32_t sb = b_smem + (uint32_t)p.units[u].ring_kb * 1024u;
          const uint32_t idesc = make_idesc_tf32(rows);
          const uint32_t d_tmem = tmem_base + dcol;
          for (uint32_t ks = 0; ks < nks; ++ks) {
            const uint32_t kg = kstep0 + ks;  // global k-step: A chunk kg / 4, 32 B steps inside
            const uint64_t a_hi = make_desc_sw128(a_smem + (kg >> 2) * TC_A_CHUNK_BYTES + (kg & 3u) * 32u);
            const bool wide = (uflags & 4u) != 0;
            const uint64_t b_hi = wide ? make_desc_sw128(sb + ks * 32u) : make_desc_sw64(sb + ks * 32u);
            const uint64_t b_lo = wide ? make_desc_sw128(sb + rows * 128u + ks * 32u) : make_desc_sw64(sb + rows * 64u + ks * 32u);
            const uint32_t a_lo = tmem_base + TC_ALO_COL + kg * 8u;
            const uint32_t acc0 = ((uflags & 1) && ks == 0) ? 0u : 1u;
            if (leader) {
              mma_ts(d_tmem, a_lo, b_hi, idesc, acc0);  // small terms first
              mma_ss(d_tmem, a_hi, b_lo, idesc, 1u);
              mma_ss(d_tmem, a_hi, b_hi, idesc, 1u);
            }
          }
          if (leader) {
            if (csize > 1) tc_commit_mcast(bar_bempty + 8 * stage, cmask);
            else tc_commit(bar_bempty + 8 * stage);
            if (uflags & 2) tc_commit(bar_dfull);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ============================== weight streamer ==============================
    const bool leader = (tid & 31) == 0;
    uint32_t g = 0;
    for (int64_t it = 0; it < tiles_per_cta; ++it) {
      for (int u = 0; u < p.n_units; ++u, ++g) {
        const uint32_t bytes = (uint32_t)p.units[u].rows * ((p.units[u].flags & 4u) ? 256u : 128u);
        const uint32_t back = p.units[u].back;
        if (g >= back) {
          // the bytes this unit overwrites (and its barrier pair) are free once unit g - back is done
          const uint32_t d = g - back;
          mbar_wait(bar_bempty + 8 * (d % TC_NBAR), (d / TC_NBAR) & 1u);
        }
        const uint32_t stage = g % TC_NBAR;
        if (leader) {
          mbar_expect_tx(bar_bfull + 8 * stage, bytes);  // all slices, whoever sends them
          const uint32_t dst = b_smem + (uint32_t)p.units[u].ring_kb * 1024u;
          if (csize > 1) {
            const uint32_t slice = bytes / csize;  // unit sizes are multiples of 2 KB
            bulk_g2s_mcast(dst + crank * slice, p.wimg + p.units[u].w_off + crank * slice, slice, bar_bfull + 8 * stage, cmask);
          } else {
            bulk_g2s(dst, p.wimg + p.units[u].w_off, bytes, bar_bfull + 8 * stage);
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == TC_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------
// self test: D[128, N] = A[128, 32] . B[N, 32]^T with one tcgen05.mma chain (4 k-steps)
//   mode 0: A from shared memory (SS), mode 1: A from tensor memory (TS)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1)
k_tc_selftest(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_smem = base, b_smem = base + TC_A_CHUNK_BYTES, bar = b_smem + 256 * 128, tmem_slot = bar + 8;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  if (warp < 4) {
    const uint32_t row = tid;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    float v[16];
    for (int col0 = 0; col0 < 32; col0 += 16) {
      for (int i = 0; i < 16; ++i) v[i] = A[row * 32 + col0 + i];
      // hi to shared memory, and the same hi values to TMEM columns [256, 288) for the TS test
      uint32_t hi[16];
      for (int i = 0; i < 16; ++i) hi[i] = __float_as_uint(v[i]) & 0xffffe000u;
      for (int q = 0; q < 4; ++q) {
        uint32_t addr = a_smem + sw128_off(row, (col0 >> 2) + q);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hi[4 * q]), "r"(hi[4 * q + 1]),
                     "r"(hi[4 * q + 2]), "r"(hi[4 * q + 3])
                     : "memory");
      }
      tmem_st16(tmem_lane + TC_ALO_COL + col0, hi);
    }
    for (int n = row; n < N; n += 128)
      for (int g = 0; g < 8; ++g) {
        uint32_t w[4];
        for (int i = 0; i < 4; ++i) w[i] = __float_as_uint(B[n * 32 + 4 * g + i]) & 0xffffe000u;
        // mode 2: two SWIZZLE_64B chunks of 16 k, each N rows x 64 B
        uint32_t addr = (mode == 2) ? b_smem + (uint32_t)(g >> 2) * (uint32_t)N * 64u + sw64_off(n, g & 3)
                                    : b_smem + sw128_off(n, g);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      }
    fence_async_smem();
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 128 && mode >= 100) {
    // issue-rate probe: 256 back-to-back MMAs of width N (100: SS, 101: TS, 102: SS,SS,TS);
    // D[0] = cycles per MMA to completion, D[1] = cycles per MMA spent issuing
    const uint32_t idesc = make_idesc_tf32(N);
    const int REP = 256;
    const long long t0 = clock64();
    for (int r = 0; r < REP; ++r) {
      const uint32_t ks = r & 3;
      const uint64_t bd = make_desc_sw128(b_smem + ks * 32u);
      const bool ts = (mode == 101) || (mode == 102 && (r % 3) == 2);
      if (ts) mma_ts(tmem_base, tmem_base + TC_ALO_COL + ks * 8u, bd, idesc, r ? 1u : 0u);
      else mma_ss(tmem_base, make_desc_sw128(a_smem + ks * 32u), bd, idesc, r ? 1u : 0u);
    }
    const long long t1 = clock64();
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    D[0] = (float)(t2 - t0) / REP;
    D[1] = (float)(t1 - t0) / REP;
  } else if (tid == 128) {
    const uint32_t idesc = make_idesc_tf32(N);
    for (uint32_t ks = 0; ks < 4; ++ks) {
      const uint64_t bd = (mode == 2) ? make_desc_sw64(b_smem + (ks >> 1) * (uint32_t)N * 64u + (ks & 1u) * 32u)
                                      : make_desc_sw128(b_smem + ks * 32u);
      if (mode == 0 || mode == 2)
        mma_ss(tmem_base, make_desc_sw128(a_smem + ks * 32u), bd, idesc, ks ? 1u : 0u);
      else
        mma_ts(tmem_base, tmem_base + TC_ALO_COL + ks * 8u, bd, idesc, ks ? 1u : 0u);
    }
    tc_commit(bar);
  }
  if (warp < 4 && mode < 100) {
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int col0 = 0; col0 < N; col0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_lane + col0, r);
      tc_wait_ld();
      for (int i = 0; i < 16; ++i) D[tid * N + col0 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
static inline float tf32_trunc(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  b &= 0xffffe000u;
  memcpy(&x, &b, 4);
  return x;
}

int c1d_tc_prepare(c1d_ctx* ctx) {
  const c1d_config& c = ctx->cfg;
  if (ctx->tc_state) { c1d_tc_destroy(ctx); }
  if (c.emu_kind != 0 || c.n_layers < 1) return C1D_OK;
  TcState* st = new TcState();
  memset(st, 0, sizeof(*st));
  ctx->tc_state = st;
  const int L = c.n_layers;
  size_t nw = 0, nb = 0;
  std::vector<size_t> woff(L), boff(L);
  for (int l = 0; l < L; ++l) {
    woff[l] = nw;
    boff[l] = nb;
    nw += (size_t)c.layer_dims[l] * c.layer_dims[l + 1];
    nb += c.layer_dims[l + 1];
  }
  std::vector<double> W(nw), Bv(nb);
  C1D_CUDA(ctx, cudaMemcpy(W.data(), ctx->dev[C1D_F_NN_W], nw * sizeof(double), cudaMemcpyDeviceToHost));
  C1D_CUDA(ctx, cudaMemcpy(Bv.data(), ctx->dev[C1D_F_NN_B], nb * sizeof(double), cudaMemcpyDeviceToHost));
  // weights are stored in-major: Wt[k * N + n]
  auto wat = [&](int l, int n, int k) { return W[woff[l] + (size_t)k * c.layer_dims[l + 1] + n]; };

  TcParams& p = st->p;
  p.n_in = c.layer_dims[0];
  p.last_out = c.layer_dims[L];
  p.n_tl = L;  // every layer runs on the tensor cores
  p.slope = c.nn_slope;
  p.debug = getenv("C1D_TC_DEBUG") ? 1 : 0;
  st->cluster = getenv("C1D_TC_CLUSTER") ? atoi(getenv("C1D_TC_CLUSTER")) : 1;
  if (st->cluster != 1 && st->cluster != 2 && st->cluster != 4) st->cluster = 1;
  std::vector<TcUnit> units;
  std::vector<uint8_t> img;
  std::vector<float> bias;
  for (int tl = 0; tl < p.n_tl; ++tl) {
    const int l = tl;
    const int K = c.layer_dims[l], N = c.layer_dims[l + 1];
    const int npad = (N + 15) & ~15;
    p.tl_npad[tl] = npad;
    p.tl_bias_off[tl] = (int)bias.size();
    for (int n = 0; n < npad; ++n) bias.push_back(n < N ? (float)Bv[boff[l] + n] : 0.f);
    p.tl_unit_begin[tl] = (int)units.size();
    // narrow layers (N <= 128): 32-k units (fewer, longer units); wide layers: 16-k units so that one
    // tcgen05.mma can span up to 256 columns within a 32 KB unit
    const int bk = (npad <= 128) ? 32 : 16;
    const int nkc = (K + bk - 1) / bk;
    int blocks[2], nblk = 1;
    if (npad <= TC_NBLK) blocks[0] = npad;
    else { blocks[0] = ((npad / 2) + 15) & ~15; blocks[1] = npad - blocks[0]; nblk = 2; }
    int n0 = 0;
    for (int b = 0; b < nblk; ++b) {
      const int rows = blocks[b];
      for (int kc = 0; kc < nkc; ++kc) {
        TcUnit u;
        memset(&u, 0, sizeof(u));
        u.w_off = (uint32_t)img.size();
        u.rows = (uint16_t)rows;
        u.dcol = (uint16_t)n0;
        u.layer = (uint8_t)tl;
        u.kstep0 = (uint8_t)(kc * (bk / 8));
        int kleft = K - kc * bk;
        if (kleft > bk) kleft = bk;
        u.ksteps = (uint8_t)((kleft + 7) / 8);
        u.flags = (uint8_t)((kc == 0 ? 1 : 0) | ((b == nblk - 1 && kc == nkc - 1) ? 2 : 0) | (bk == 32 ? 4 : 0));
        units.push_back(u);
        size_t o = img.size();
        img.resize(o + (size_t)rows * bk * 8, 0);
        for (int n = 0; n < rows; ++n)
          for (int k = 0; k < bk; ++k) {
            int gn = n0 + n, gk = kc * bk + k;
            double w = (gn < N && gk < K) ? wat(l, gn, gk) : 0.0;
            float hi = tf32_trunc((float)w);
            float lo = (float)(w - (double)hi);
            uint32_t off = ((bk == 32) ? sw128_off((uint32_t)n, (uint32_t)k >> 2) : sw64_off((uint32_t)n, (uint32_t)k >> 2)) +
                           ((uint32_t)k & 3u) * 4u;
            memcpy(&img[o + off], &hi, 4);
            memcpy(&img[o + (size_t)rows * bk * 4 + off], &lo, 4);
          }
      }
      n0 += rows;
    }
  }
  p.tl_unit_begin[p.n_tl] = (int)units.size();
  p.n_units = (int)units.size();
  {
    // pack the units of one tile back to back into the ring (every tile restarts at offset 0, so
    // the placement is the same for all tiles) and find, for each unit, the most recent earlier
    // unit of the cyclic sequence whose bytes it overwrites: the streamer waits for that one.
    const int nu = p.n_units;
    std::vector<int> off(nu), len(nu);
    int cur = 0;
    for (int u = 0; u < nu; ++u) {
      len[u] = ((int)units[u].rows * ((units[u].flags & 4) ? 256 : 128) + 1023) & ~1023;
      if (cur + len[u] > TC_RING_BYTES) cur = 0;
      off[u] = cur;
      cur += len[u];
    }
    for (int u = 0; u < nu; ++u) {
      int back = TC_NBAR;  // at least: the previous user of this barrier pair
      for (int b = 1; b < TC_NBAR; ++b) {
        int v = ((u - b) % nu + nu) % nu;
        if (off[v] < off[u] + len[u] && off[u] < off[v] + len[v]) { back = b; break; }
      }
      if (back > nu) back = nu;  // a tile has fewer units than barrier pairs: wait for this unit's previous use
      units[u].ring_kb = (uint16_t)(off[u] / 1024);
      units[u].back = (uint8_t)back;
    }
  }
  if (p.n_units > TC_MAX_UNITS) {
    snprintf(ctx->err, sizeof(ctx->err), "tensor-core MLP: too many weight units (%d)", p.n_units);
    return C1D_ERR_ARG;
  }
  std::vector<double> lo(c.n_emu), hi(c.n_emu), rinv(c.n_emu);
  C1D_CUDA(ctx, cudaMemcpy(lo.data(), ctx->dev[C1D_F_EMU_LO], c.n_emu * sizeof(double), cudaMemcpyDeviceToHost));
  C1D_CUDA(ctx, cudaMemcpy(hi.data(), ctx->dev[C1D_F_EMU_HI], c.n_emu * sizeof(double), cudaMemcpyDeviceToHost));
  for (int j = 0; j < c.n_emu; ++j) rinv[j] = 1.0 / (hi[j] - lo[j]);
#define UP(dst, src, bytes)                                                   \
  do {                                                                        \
    C1D_CUDA(ctx, cudaMalloc(&(dst), (bytes) ? (bytes) : 16));                \
    C1D_CUDA(ctx, cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice)); \
  } while (0)
  UP(st->d_wimg, img.data(), img.size());
  for (size_t q = 0; q < units.size(); ++q) p.units[q] = units[q];
  UP(st->d_bias, bias.data(), bias.size() * sizeof(float));
  UP(st->d_rng_inv, rinv.data(), rinv.size() * sizeof(double));
#undef UP
  p.wimg = (const uint8_t*)st->d_wimg;
  p.bias = (const float*)st->d_bias;
  p.emu_lo = ctx->d.emu_lo;
  p.emu_rng_inv = (const double*)st->d_rng_inv;
  C1D_CUDA(ctx, cudaFuncSetAttribute(k_mlp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  st->ready = true;
  return C1D_OK;
}

int c1d_tc_launch(c1d_ctx* ctx, const double* emu_in, int64_t nrows, double* out, int out_stride, cudaStream_t s) {
  TcState* st = (TcState*)ctx->tc_state;
  if (!st || !st->ready) {
    snprintf(ctx->err, sizeof(ctx->err),
             "tensor-core emulator path unavailable (needs an NN emulator)");
    return C1D_ERR_STATE;
  }
  int64_t ntiles = (nrows + TC_ROWS - 1) / TC_ROWS;
  int grid = (int)(ntiles < ctx->num_sms ? ntiles : ctx->num_sms);
  // optional clusters share the weight stream (multicast).  Measured on B200: no gain, the MMA phase
  // is bound by tensor-pipe occupancy per instruction (tools/tc_issue_probe.py), not by L2 bandwidth
  int csize = st->cluster;
  while (csize > 1 && (grid % csize)) csize >>= 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  C1D_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_mlp_tc, st->p, emu_in, nrows, out, out_stride));
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return C1D_OK;
}

void c1d_tc_destroy(c1d_ctx* ctx) {
  TcState* st = (TcState*)ctx->tc_state;
  if (!st) return;
  cudaFree(st->d_wimg); cudaFree(st->d_bias); cudaFree(st->d_rng_inv);
  delete st;
  ctx->tc_state = nullptr;
}

// D[128, N] = tf32(A[128, 32]) . tf32(B[N, 32])^T through one tcgen05.mma chain; host pointers.
extern "C" int c1d_tc_selftest(int device, const float* A_host, const float* B_host, float* D_host, int N, int mode) {
  if (!A_host || !B_host || !D_host || N < 16 || N > 256 || (N % 16) || (mode < 0 || (mode > 2 && (mode < 100 || mode > 102)))) return C1D_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return C1D_ERR_CUDA;
  float *dA = nullptr, *dB = nullptr, *dD = nullptr;
  int rc = C1D_OK;
  size_t smem = 1024 + TC_A_CHUNK_BYTES + 256 * 128 + 64;
  if (cudaMalloc(&dA, 128 * 32 * 4) || cudaMalloc(&dB, (size_t)N * 32 * 4) || cudaMalloc(&dD, (size_t)128 * N * 4)) rc = C1D_ERR_CUDA;
  if (!rc && (cudaMemcpy(dA, A_host, 128 * 32 * 4, cudaMemcpyHostToDevice) ||
              cudaMemcpy(dB, B_host, (size_t)N * 32 * 4, cudaMemcpyHostToDevice)))
    rc = C1D_ERR_CUDA;
  if (!rc && cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) rc = C1D_ERR_CUDA;
  if (!rc) {
    k_tc_selftest<<<1, 160, smem>>>(dA, dB, dD, N, mode);
    if (cudaDeviceSynchronize() != cudaSuccess) rc = C1D_ERR_CUDA;
  }
  if (!rc && cudaMemcpy(D_host, dD, (size_t)128 * N * 4, cudaMemcpyDeviceToHost)) rc = C1D_ERR_CUDA;
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return rc;
}
