// Stage 4 on the 5th-generation tensor cores: chi2 = d^T C^-1 d (likelihood.py:956-960) with the dense inverse
// covariance as tcgen05.mma kind::i8 GEMMs over balanced base-256 digit planes (an Ozaki-style error-free split, the
// arithmetic of c1d_mlp_i8.cu with five digits).
//
// Arithmetic.  Row j of the symmetric C^-1 is folded to its lower triangle by tiles, A'_jk = 2 C^-1_jk for k before the
// tile of j, C^-1_jk inside the tile, 0 after it, and written as A'_jk = sA_j * qA_jk with five balanced digits
// (|qA| <= 127 * 2^32); the residual row of a walker is d_wk = sd_w * qd_wk likewise, with sd_w from the row's own
// maximum.  The digit-plane products of level l = i + i' <= 4 (15 pairs) accumulate EXACTLY in int32 (|sum| <= 5 * K *
// 2^14) and recombine in float64:  y_jw = sA_j sd_w 2^32 sum_l 256^(4-l) acc_l,  chi2_w = sum_j y_jw d_wj with the
// unquantised d.  The errors are the two quantisations (2^-40 of the row maxima) and the dropped pairs of level >= 5:
// chi2 within 5e-11 relative (2e-7 absolute at chi2 ~ 4e3) of the float64 evaluation on the DESI-DR1-like workload and
// on the reference configurations with a dense covariance (tools/chi2_ozaki_sim.py; tests/test_gpu_chi2_i8.py).
//
// Orientation.  The rows of C^-1 are the M = 128 side (TMEM lanes), the walkers the N side: a CTA quantises the
// residuals of NW = 48 walkers ONCE into shared memory (five planes x 48 rows x K bytes, SWIZZLE_32B, 165 KB at K = 704)
// and then only streams: the pre-swizzled digit planes of C^-1 arrive as 20 KB units (128 rows x 32 k x 5 planes) through
// a three-slot cp.async.bulk ring, in the same order for every CTA.  Because the digit planes of the residuals are
// consecutive 48-row groups of one shared-memory operand, ONE tcgen05.mma of N = 48 (5 - i) multiplies plane i of C^-1
// with planes 0 .. 4 - i of the residuals and lands on the level accumulators i .. 4, which are consecutive 48-column
// groups of tensor memory: 5 instructions per unit instead of 15, each C^-1 plane read once.  Two accumulator sets
// (2 x 240 columns) alternate between row tiles, so the epilogue of one tile (tcgen05.ld, float64 recombination,
// product with d_wj, per-thread partial sums) runs under the MMAs of the next.  The symmetric fold starts with the short
// tile (K = 704: rows 0..63, two units) so that the zero padding of the 128-row tile costs two units, not 22.
//
// One persistent CTA per SM, 14 warps: 12 quantiser / epilogue warps (TMEM lane quadrant = warp % 4, walker columns
// 16 (warp / 4) ..), one MMA-issuing warp, one streamer.
//
// Measured (B200, C4, 65 536 walkers): 0.38 ms against 1.06 ms for the fp64 DMMA kernel; IMMA pipe 39 % active; DRAM
// 394 MB per launch (the residuals once).  The phase clocks (C1D_CHI2_DEBUG=1) show a 48-walker group at ~70 k cycles
// with the MMAs at ~600 cycles per unit against 390 of tensor time: the unit stream (20 KB per unit per SM, 5 KB/clk
// over the chip) runs at the rate L2 delivers bytes to the SMs, which the residual passes share.  Multicasting the
// stream over CTA pairs changed nothing (0.377 against 0.373 ms): delivered bytes, not L2 array reads, are the limit.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "c1d_internal.h"
#include "c1d_tc_common.cuh"

namespace {

using namespace c1d_tc;

constexpr int CQ_NW = 48;                 // walkers per group (N side)
constexpr int CQ_ND = 5;                  // digit planes of both operands = levels kept
constexpr int CQ_EPI_WARPS = 12;
constexpr int CQ_EPI_THREADS = 32 * CQ_EPI_WARPS;
constexpr int CQ_THREADS = CQ_EPI_THREADS + 64;
constexpr uint32_t CQ_BPLANE = CQ_NW * 32;            // residual digits: one plane of one 32-k unit (1536 B)
constexpr uint32_t CQ_BUNIT = CQ_ND * CQ_BPLANE;      // 7680 B
constexpr uint32_t CQ_APLANE = 128 * 32;              // C^-1 digits: one plane of one unit (4096 B)
constexpr uint32_t CQ_AUNIT = CQ_ND * CQ_APLANE;      // 20 480 B
constexpr int CQ_NSLOT = 3;
constexpr int CQ_MAX_TILES = 8;
constexpr uint32_t CQ_ACC_COLS = CQ_ND * CQ_NW;       // 240 TMEM columns per accumulator set
constexpr double CQ_FULL = 127.0 * 4294967296.0;      // largest |q|: the top balanced digit stays within +-127
constexpr size_t CQ_SMEM_MAX = 232448;

struct CQParams {
  const uint8_t* img;     // digit planes of the folded C^-1 in streaming order: [tile][unit][plane][128 rows x 32 B, SWIZZLE_32B]
  const double* rscale;   // [ntile * 128]: sA_j * 2^32 of the row a TMEM lane holds in this tile, 0 for padding rows
  int ntile, nunits_b;    // row tiles; 32-k units of the residual operand (= K / 32)
  int ldd, nd;
  int tile_r0[CQ_MAX_TILES];      // first row of the tile
  int tile_rows[CQ_MAX_TILES];    // rows of the tile (<= 128)
  int tile_units[CQ_MAX_TILES];   // k units of the tile (k < end of the tile)
  int tile_u0[CQ_MAX_TILES];      // first unit of the tile in the image
  uint32_t part_mask[2];          // nsplit = 2: the row tiles of each half of a group's work (balanced by units)
  int part_rs[2];                 // ... and the tile after whose epilogue the next item's scales are found
  uint32_t off_ring, off_sd, off_bar;  // shared-memory offsets (the residual operand starts at 0)
  long long* dbg;  // C1D_CHI2_DEBUG=1: clocks of CTA 0's second group (phases of warp 0 and of the MMA warp)
};

struct CQState {
  CQParams p;
  void *d_img, *d_rscale, *d_dbg;
  size_t smem;
  bool ready;
};

__device__ __forceinline__ void cq_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(CQ_EPI_THREADS) : "memory"); }
__device__ __forceinline__ double cq_i2d(uint32_t x) {  // exact int32 -> float64 without the conversion unit
  return __hiloint2double(0x43300000, (int)(x ^ 0x80000000u)) - 4503601774854144.0;
}

__global__ void __launch_bounds__(CQ_THREADS, 1)
k_chi2_i8(CQParams p, const double* __restrict__ dvec, int64_t W, double* __restrict__ chi2p, int ntj, int nsplit) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t ring = base + p.off_ring;
  const uint32_t bar0 = base + p.off_bar;
  const uint32_t bar_bfull = bar0, bar_bempty = bar0 + 8 * CQ_NSLOT;
  const uint32_t bar_dfull = bar0 + 16 * CQ_NSLOT, bar_dfree = bar_dfull + 16, bar_tready = bar_dfree + 16;
  const uint32_t tmem_slot = bar_tready + 8 * CQ_MAX_TILES;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(smem_raw + p.off_bar + 16 * CQ_NSLOT + 32 + 8 * CQ_MAX_TILES);
  double* sd_s = reinterpret_cast<double*>(smem_raw + p.off_sd);    // [2][NW] scale of the walker's digits (NaN: bad row), this group / the next
  double* red_s = reinterpret_cast<double*>(smem_raw);              // [4 quadrants][NW]: on the operand region, free once the group's MMAs are complete

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;
  if (tid == 0) {
    if (base & 1023u) {
      printf("cup1d_b200: k_chi2_i8 shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    for (int s = 0; s < CQ_NSLOT; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_dfull + 8 * b, 1);
      mbar_init(bar_dfree + 8 * b, CQ_EPI_THREADS);
    }
    for (int t = 0; t < CQ_MAX_TILES; ++t) mbar_init(bar_tready + 8 * t, CQ_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CQ_EPI_WARPS) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_p, 0);

  // Work items: a 48-walker group, or (nsplit = 2: small and medium batches, where whole groups leave SMs idle or
  // fill a last round badly) one of the two halves of its row tiles; both halves quantise the group and write their
  // partial sum to their own slot of the walker's chi2p row.  Items blockIdx.x, + gridDim.x, ..
  const int64_t nitems = ((W + CQ_NW - 1) / CQ_NW) * nsplit;
  const int64_t groups_per_cta = (nitems - blockIdx.x + gridDim.x - 1) / gridDim.x;
  auto item_w0 = [&](int64_t it) { return ((it * gridDim.x + blockIdx.x) / nsplit) * CQ_NW; };
  auto item_part = [&](int64_t it) { return (int)((it * gridDim.x + blockIdx.x) % nsplit); };
  const uint32_t all_tiles = (1u << p.ntile) - 1u;
  const int ldd = p.ldd;

  if (warp < CQ_EPI_WARPS) {
    // ======================= quantiser / epilogue warps =======================
    const uint32_t quad = (uint32_t)warp & 3u, sub = (uint32_t)warp >> 2;
    const uint32_t lane_addr = tmem_base + ((quad * 32u) << 16);
    uint32_t tcount = 0;       // tiles seen so far (accumulator set = tcount & 1)

    // A warp owns four walkers of the group (rows 4 warp .. 4 warp + 3 of the N-side operand): it finds their maxima
    // (lanes along k, coalesced), and later converts them 64 k at a time with lane = (row, 8 k): a 256-bit load touches
    // 16 lines instead of the 32 of a lane-per-row mapping (the L1 tag stage serves a line per cycle), and the four rows
    // land on the four 32-byte bank windows of the swizzled operand, so the 8-byte stores are conflict-free.
    auto row_scales = [&](int64_t w0, int sel) {
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {  // one row at a time, all its loads in flight (44 registers; the rows come from L2)
        const int nn = 4 * warp + h;
        const int64_t w = w0 + nn;
        const double2* row = reinterpret_cast<const double2*>(dvec + (w < W ? w : W - 1) * ldd) + lane;
        double2 v[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) v[i] = (64 * i < ldd) ? row[32 * i] : make_double2(0.0, 0.0);
        uint32_t mhi = 0u;  // |x| orders like its high word; NaN / Inf carry the largest
#pragma unroll
        for (int i = 0; i < 11; ++i)
          mhi = max(mhi, max((uint32_t)__double2hiint(v[i].x) & 0x7fffffffu, (uint32_t)__double2hiint(v[i].y) & 0x7fffffffu));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mhi = max(mhi, __shfl_xor_sync(0xffffffffu, mhi, o));
        if (lane == 0) {
          // upper bound of max|d|: the next value of the high word (2^-20 above at most)
          double sc = 0.0;
          if (w >= W) sc = 0.0;
          else if (mhi >= 0x7ff00000u) sc = __longlong_as_double(0x7ff8000000000000LL);
          else if (mhi != 0u) sc = __hiloint2double((int)(mhi + 1u), 0) * (1.0 / CQ_FULL);
          sd_s[sel * CQ_NW + nn] = sc;
        }
      }
      __syncwarp();
    };
    const int qn = 4 * warp + (lane >> 3);       // the row this lane converts
    const int qk = 8 * (lane & 7);               // its 8 k inside a 64-k step
    auto load8 = [&](const double* src, double (&x)[8]) {
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(src));
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[4]), "=d"(x[5]), "=d"(x[6]), "=d"(x[7]) : "l"(src + 4));
    };

    for (int64_t it = 0; it < groups_per_cta; ++it) {
      const int64_t w0 = item_w0(it);
      const int part = item_part(it);
      const uint32_t tmask = nsplit == 2 ? p.part_mask[part] : all_tiles;
      const int tpre = nsplit == 2 ? p.part_rs[part] : (p.ntile > 3 ? p.ntile - 3 : 0);  // the next item's scales are found under the long tiles of this one
      const int sel = (int)(it & 1);
      const double* sdc = sd_s + sel * CQ_NW;
      const bool dbg = p.dbg && blockIdx.x == 0 && tid == 0 && it == 1;
      if (dbg) p.dbg[0] = clock64();
      if (it == 0) {
        row_scales(w0, 0);
        cq_epi_bar();
      }
      // ---- residuals -> five digit planes of the N-side operand (the MMAs of the previous group are complete: its
      //      last dfull was waited for); the loads of the next step are in flight while one is converted; a row tile's
      //      MMAs start as soon as the k units it needs are in place (bar_tready) ----
      {
        const double s = sdc[qn];
        const double inv_s = (s > 0.0) ? 1.0 / s : 0.0;  // zero rows, bad rows (NaN scale) and padding walkers: digits 0
        const int64_t wq = w0 + qn;
        const double* src = dvec + (wq < W ? wq : W - 1) * ldd + qk;
        const int nho = ldd >> 6;
        double xa[8], xb[8], xc[8];  // two steps of loads in flight (L2 latency exceeds one step of conversions)
        load8(src, xa);
        if (nho > 1) load8(src + 64, xb);
        int tnext = 0;  // next row tile to release
        for (int ho = 0; ho < nho; ++ho) {
          if (ho + 2 < nho) load8(src + 64 * (ho + 2), xc);
          uint32_t pw[CQ_ND][2];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t vlo[4], vhi[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // round to nearest through the mantissa of x / s + 1.5 2^52 (|q| < 2^40): no conversion instruction
              const double r = fma(xa[4 * g + e], inv_s, 6755399441055744.0);
              unsigned long long q = ((unsigned long long)(uint32_t)__double2hiint(r) << 32) | (uint32_t)__double2loint(r);
              q = (q + 0x8080808080ull) ^ 0x8080808080ull;  // two's complement -> balanced digit bytes
              vlo[e] = (uint32_t)q;
              vhi[e] = (uint32_t)(q >> 32);
            }
            const uint32_t a = __byte_perm(vlo[0], vlo[1], 0x5140), b = __byte_perm(vlo[0], vlo[1], 0x7362);
            const uint32_t c = __byte_perm(vlo[2], vlo[3], 0x5140), d = __byte_perm(vlo[2], vlo[3], 0x7362);
            pw[4][g] = __byte_perm(a, c, 0x5410);  // bytes 0: least significant digit = plane 4
            pw[3][g] = __byte_perm(a, c, 0x7632);
            pw[2][g] = __byte_perm(b, d, 0x5410);
            pw[1][g] = __byte_perm(b, d, 0x7632);
            const uint32_t e4 = __byte_perm(vhi[0], vhi[1], 0x0040), f4 = __byte_perm(vhi[2], vhi[3], 0x0040);
            pw[0][g] = __byte_perm(e4, f4, 0x5410);  // most significant digit = plane 0
          }
          const int k0 = 64 * ho + qk;
          const uint32_t dst = base + (uint32_t)(k0 >> 5) * CQ_BUNIT + sw32_off((uint32_t)qn, (uint32_t)(k0 >> 4) & 1u) + ((uint32_t)k0 & 8u);
#pragma unroll
          for (int pl = 0; pl < CQ_ND; ++pl)
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst + (uint32_t)pl * CQ_BPLANE), "r"(pw[pl][0]), "r"(pw[pl][1]) : "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i) { xa[i] = xb[i]; xb[i] = xc[i]; }
          // units [0, 2 ho + 2) are written by this thread: release the row tiles they complete
          while (tnext < p.ntile && p.tile_units[tnext] <= 2 * ho + 2) {
            fence_async_smem();
            mbar_arrive(bar_tready + 8 * tnext);
            ++tnext;
          }
        }
      }

      if (dbg) p.dbg[1] = clock64();
      // ---- epilogues: partial sums over the rows this thread holds, 16 walkers ----
      double acc[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = 0.0;
      for (int t = 0; t < p.ntile; ++t) {
        if (!((tmask >> t) & 1u)) continue;
        const uint32_t buf = tcount & 1u;
        const int r = (int)(quad * 32u) + lane;             // row of the tile this lane holds
        const bool rvalid = r < p.tile_rows[t] && p.tile_r0[t] + r < p.nd;
        const int j = rvalid ? p.tile_r0[t] + r : 0;
        const double cs = __ldg(p.rscale + t * 128 + r);    // 0 for padding rows
        double cd[16];  // (row scale) x (residual of this row) per walker: the products wait for the accumulators
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int64_t w = w0 + 16 * sub + c;
          cd[c] = cs * __ldg(dvec + (w < W ? w : W - 1) * ldd + j);
        }
        if (dbg) p.dbg[8 + 4 * t] = clock64();
        mbar_wait(bar_dfull + 8 * buf, (tcount >> 1) & 1u);
        tc_fence_after();
        if (dbg) p.dbg[9 + 4 * t] = clock64();
        // level by level, eight walkers at a time: eight independent float64 chains per thread, the next level's
        // tcgen05.ld in flight under the arithmetic of this one, few registers (no spills at the 128 cap)
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          const uint32_t ta = lane_addr + buf * CQ_ACC_COLS + 16u * sub + 8u * hb;
          uint32_t ra[8], rb[8];
          double T[8];
          tmem_ld8(ta, ra);
          tc_wait_ld();
          tmem_ld8(ta + CQ_NW, rb);
#pragma unroll
          for (int c = 0; c < 8; ++c) T[c] = cq_i2d(ra[c]);
#pragma unroll
          for (int l = 1; l < CQ_ND; ++l) {
            tc_wait_ld();
            if (l & 1) {
              if (l + 1 < CQ_ND) tmem_ld8(ta + (uint32_t)(l + 1) * CQ_NW, ra);
#pragma unroll
              for (int c = 0; c < 8; ++c) T[c] = fma(T[c], 256.0, cq_i2d(rb[c]));
            } else {
              if (l + 1 < CQ_ND) tmem_ld8(ta + (uint32_t)(l + 1) * CQ_NW, rb);
#pragma unroll
              for (int c = 0; c < 8; ++c) T[c] = fma(T[c], 256.0, cq_i2d(ra[c]));
            }
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[8 * hb + c] = fma(T[c], cd[8 * hb + c], acc[8 * hb + c]);
        }
        tc_fence_before();
        mbar_arrive(bar_dfree + 8 * buf);
        if (dbg) p.dbg[10 + 4 * t] = clock64();
        // the next group's scales under this group's long tiles (the loads also bring its residuals into L2)
        if (t == tpre && it + 1 < groups_per_cta) row_scales(item_w0(it + 1), sel ^ 1);
        if (dbg) p.dbg[11 + 4 * t] = clock64();
        ++tcount;
      }
      // ---- sum over the lanes (rows) in a fixed order, then over the four quadrants ----
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        double v = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[c] = v;
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 16; ++c) red_s[quad * CQ_NW + 16 * sub + c] = acc[c];
      }
      cq_epi_bar();
      if (tid < CQ_NW && w0 + tid < W) {
        const double v = ((red_s[tid] + red_s[CQ_NW + tid]) + (red_s[2 * CQ_NW + tid] + red_s[3 * CQ_NW + tid])) * sdc[tid];
        double* o = chi2p + (w0 + tid) * ntj;  // k_finalize adds the ntj partial sums of the walker
        o[part] = v;
        if (part == 0)
          for (int q = nsplit; q < ntj; ++q) o[q] = 0.0;
      }
      if (dbg) p.dbg[2] = clock64();
      cq_epi_bar();  // red_s lies in the operand region the next group's digits overwrite; the scale buffers alternate
    }
  } else if (warp == CQ_EPI_WARPS) {
    // ================================ MMA issuer ================================
    const bool leader = elect_one();
    const uint64_t desc0 = make_desc_sw32(0);
    const uint32_t base16 = base >> 4, ring16 = ring >> 4;
    uint32_t g = 0, tcount = 0;
    for (int64_t it = 0; it < groups_per_cta; ++it) {
      const bool dbg = p.dbg && blockIdx.x == 0 && leader && it == 1;
      const uint32_t tmask = nsplit == 2 ? p.part_mask[item_part(it)] : all_tiles;
      for (int t = 0; t < p.ntile; ++t) {
        if (!((tmask >> t) & 1u)) continue;
        const uint32_t buf = tcount & 1u;
        if (dbg) p.dbg[48 + 4 * t] = clock64();
        mbar_wait(bar_tready + 8 * t, (uint32_t)it & 1u);  // the k units of this tile are in shared memory
        tc_fence_after();
        if (dbg) p.dbg[49 + 4 * t] = clock64();
        if (tcount >= 2u) {
          mbar_wait(bar_dfree + 8 * buf, ((tcount >> 1) - 1u) & 1u);
          tc_fence_after();
        }
        if (dbg) p.dbg[50 + 4 * t] = clock64();
        const uint32_t d0 = tmem_base + buf * CQ_ACC_COLS;
        const int nu = p.tile_units[t];
        for (int u = 0; u < nu; ++u, ++g) {
          const uint32_t slot = g % CQ_NSLOT;
          mbar_wait_fast(bar_bfull + 8 * slot, (g / CQ_NSLOT) & 1u);  // (complete_tx orders the bulk copy before the MMAs)
          const uint64_t ad0 = desc0 | (uint64_t)(ring16 + slot * (CQ_AUNIT >> 4));
          const uint64_t bd0 = desc0 | (uint64_t)(base16 + (uint32_t)u * (CQ_BUNIT >> 4));
          const uint32_t keep = u ? 1u : 0u;
          if (leader) {
            // plane i of C^-1 x planes 0 .. 4 - i of the residuals -> levels i .. 4 (consecutive 48-column groups)
            mma_i8_ss(d0, ad0, bd0, make_idesc_s8(5 * CQ_NW), keep);
            mma_i8_ss(d0 + 1 * CQ_NW, ad0 + 1 * (CQ_APLANE >> 4), bd0, make_idesc_s8(4 * CQ_NW), 1u);
            mma_i8_ss(d0 + 2 * CQ_NW, ad0 + 2 * (CQ_APLANE >> 4), bd0, make_idesc_s8(3 * CQ_NW), 1u);
            mma_i8_ss(d0 + 3 * CQ_NW, ad0 + 3 * (CQ_APLANE >> 4), bd0, make_idesc_s8(2 * CQ_NW), 1u);
            mma_i8_ss(d0 + 4 * CQ_NW, ad0 + 4 * (CQ_APLANE >> 4), bd0, make_idesc_s8(1 * CQ_NW), 1u);
            tc_commit(bar_bempty + 8 * slot);
            if (u == nu - 1) tc_commit(bar_dfull + 8 * buf);
          }
          __syncwarp();
        }
        if (dbg) p.dbg[51 + 4 * t] = clock64();
        ++tcount;
      }
    }
  } else {
    // ============================== streamer of C^-1 ==============================
    const bool leader = lane == 0;
    uint32_t g = 0;
    for (int64_t it = 0; it < groups_per_cta; ++it) {
      const uint32_t tmask = nsplit == 2 ? p.part_mask[item_part(it)] : all_tiles;
      for (int t = 0; t < p.ntile; ++t) {
        if (!((tmask >> t) & 1u)) continue;
        for (int u = p.tile_u0[t]; u < p.tile_u0[t] + p.tile_units[t]; ++u, ++g) {
          const uint32_t slot = g % CQ_NSLOT;
          if (g >= (uint32_t)CQ_NSLOT) mbar_wait(bar_bempty + 8 * slot, ((g / CQ_NSLOT) - 1u) & 1u);
          if (leader) {
            mbar_expect_tx(bar_bfull + 8 * slot, CQ_AUNIT);
            bulk_g2s(ring + slot * CQ_AUNIT, p.img + (size_t)u * CQ_AUNIT, CQ_AUNIT, bar_bfull + 8 * slot);
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CQ_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
int c1d_chi2i8_prepare(c1d_ctx* ctx) {
  const c1d_config& c = ctx->cfg;
  c1d_chi2i8_destroy(ctx);
  if (!c.has_full || !ctx->dev[C1D_F_ICOV_FULL]) return C1D_OK;
  const int nd = c.nd, ldd = ctx->d.ldd;
  const int ntile = (ldd + 127) / 128;
  if (ntile > CQ_MAX_TILES) return C1D_OK;  // unavailable: the DMMA kernel serves the context
  const int nunits_b = ldd / 32;
  const size_t off_ring = ((size_t)nunits_b * CQ_BUNIT + 1023) & ~(size_t)1023;
  const size_t off_sd = off_ring + (size_t)CQ_NSLOT * CQ_AUNIT;
  const size_t off_bar = off_sd + 2 * CQ_NW * sizeof(double);
  const size_t smem = off_bar + 192;
  if (smem > CQ_SMEM_MAX) return C1D_OK;  // more than 704 data points: unavailable
  CQState* st = new CQState();
  memset(st, 0, sizeof(*st));
  ctx->chi2i8_state = st;
  CQParams& p = st->p;
  p.ntile = ntile; p.nunits_b = nunits_b; p.ldd = ldd; p.nd = nd;
  p.off_ring = (uint32_t)off_ring; p.off_sd = (uint32_t)off_sd; p.off_bar = (uint32_t)off_bar;
  st->smem = smem;
  const int h0 = ldd - 128 * (ntile - 1);  // the short tile first
  int total_units = 0;
  for (int t = 0; t < ntile; ++t) {
    p.tile_r0[t] = t ? h0 + 128 * (t - 1) : 0;
    p.tile_rows[t] = t ? 128 : h0;
    p.tile_units[t] = (p.tile_r0[t] + p.tile_rows[t]) / 32;
    p.tile_u0[t] = total_units;
    total_units += p.tile_units[t];
  }
  {
    // two halves of a group's row tiles with equal units: the largest tile first, each to the lighter half
    int load[2] = {0, 0};
    p.part_mask[0] = p.part_mask[1] = 0;
    for (int t = ntile - 1; t >= 0; --t) {
      const int h = load[1] < load[0] ? 1 : 0;
      p.part_mask[h] |= 1u << t;
      load[h] += p.tile_units[t];
    }
    for (int h = 0; h < 2; ++h) {
      // the next item's scales after the second-last tile of the half (the last one's MMAs run meanwhile)
      int last = -1, prev = -1;
      for (int t = 0; t < ntile; ++t)
        if ((p.part_mask[h] >> t) & 1u) { prev = last; last = t; }
      p.part_rs[h] = prev >= 0 ? prev : last;
    }
  }
  std::vector<double> A((size_t)nd * nd);
  C1D_CUDA(ctx, cudaMemcpy(A.data(), ctx->dev[C1D_F_ICOV_FULL], A.size() * sizeof(double), cudaMemcpyDeviceToHost));
  std::vector<uint8_t> img((size_t)total_units * CQ_AUNIT, 0);
  std::vector<double> rscale((size_t)ntile * 128, 0.0);
  std::vector<long long> q(ldd);
  size_t unit0 = 0;
  for (int t = 0; t < ntile; ++t) {
    const int r0 = p.tile_r0[t], r1 = r0 + p.tile_rows[t], kend = 32 * p.tile_units[t];
    for (int r = 0; r < p.tile_rows[t]; ++r) {
      const int j = r0 + r;
      if (j >= nd) continue;
      // folded row: twice the entries before the tile, the tile's own block as is (C^-1 symmetric: uploaded symmetrised)
      double m = 0.0;
      for (int k = 0; k < r1 && k < nd; ++k) {
        const double a = A[(size_t)j * nd + k] * (k < r0 ? 2.0 : 1.0);
        m = fmax(m, fabs(a));
      }
      if (!(m > 0.0) || !(m < INFINITY)) {
        if (m > 0.0 || m != m) { c1d_chi2i8_destroy(ctx); return C1D_OK; }  // non-finite matrix: leave it to the float64 kernel
        continue;
      }
      const double s = m / CQ_FULL;
      rscale[(size_t)t * 128 + r] = s * 4294967296.0;
      for (int k = 0; k < kend; ++k) {
        q[k] = 0;
        if (k < r1 && k < nd) q[k] = llrint(A[(size_t)j * nd + k] * (k < r0 ? 2.0 : 1.0) / s);
      }
      for (int k = 0; k < kend; ++k) {
        const unsigned long long v = ((unsigned long long)q[k] + 0x8080808080ull) ^ 0x8080808080ull;
        const int u = k >> 5, kk = k & 31;
        for (int pl = 0; pl < CQ_ND; ++pl)  // plane 0 = most significant digit
          img[(unit0 + u) * CQ_AUNIT + (size_t)pl * CQ_APLANE + c1d_tc::sw32_off((uint32_t)r, (uint32_t)kk >> 4) + (kk & 15)] =
              (uint8_t)(v >> (8 * (CQ_ND - 1 - pl)));
      }
    }
    unit0 += p.tile_units[t];
  }
  C1D_CUDA(ctx, cudaMalloc(&st->d_img, img.size()));
  C1D_CUDA(ctx, cudaMemcpy(st->d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
  C1D_CUDA(ctx, cudaMalloc(&st->d_rscale, rscale.size() * sizeof(double)));
  C1D_CUDA(ctx, cudaMemcpy(st->d_rscale, rscale.data(), rscale.size() * sizeof(double), cudaMemcpyHostToDevice));
  p.img = (const uint8_t*)st->d_img;
  p.rscale = (const double*)st->d_rscale;
  C1D_CUDA(ctx, cudaFuncSetAttribute(k_chi2_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p.dbg = nullptr;
  if (getenv("C1D_CHI2_DEBUG")) {
    C1D_CUDA(ctx, cudaMalloc(&st->d_dbg, 96 * sizeof(long long)));
    C1D_CUDA(ctx, cudaMemset(st->d_dbg, 0, 96 * sizeof(long long)));
    p.dbg = (long long*)st->d_dbg;
  }
  st->ready = true;
  return C1D_OK;
}

int c1d_chi2i8_available(const c1d_ctx* ctx) {
  const CQState* st = (const CQState*)ctx->chi2i8_state;
  return st && st->ready;
}

int c1d_chi2i8_launch(c1d_ctx* ctx, int64_t W, int ntj, cudaStream_t s) {
  CQState* st = (CQState*)ctx->chi2i8_state;
  if (!st || !st->ready) {
    snprintf(ctx->err, sizeof(ctx->err), "int8 digit-slice chi2 unavailable (needs a dense inverse covariance of at most 704 data points)");
    return C1D_ERR_STATE;
  }
  const int64_t ngroups = (W + CQ_NW - 1) / CQ_NW;
  // halves of groups where whole groups would leave at least half of the SMs idle (small batches: the kernel's latency
  // drops by the share of the MMAs); a half costs more than half a group (it quantises the whole group), so larger
  // batches keep whole groups.  (Halves between one and one and a half rounds of groups gained 2 % at 8192 walkers and
  // are not used: the partial sums of a walker add in another order, and a shard of a large ensemble - 8192 walkers
  // per GPU of 65 536 - would no longer equal the unsharded evaluation bit for bit.)
  const int64_t sms = ctx->num_sms;
  int nsplit = (st->p.ntile >= 2 && ntj >= 2 && 2 * ngroups <= sms) ? 2 : 1;
  if (const char* e = getenv("C1D_CHI2_SPLIT")) nsplit = (atoi(e) == 2 && st->p.ntile >= 2 && ntj >= 2) ? 2 : 1;
  const int64_t nitems = ngroups * nsplit;
  const int grid = (int)(nitems < sms ? nitems : sms);
  k_chi2_i8<<<grid, CQ_THREADS, st->smem, s>>>(st->p, ctx->ws.dvec, W, ctx->ws.chi2p, ntj, nsplit);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  if (st->p.dbg && nitems > 2 * grid) {
    // diagnostic: the second group of CTA 0, cycles relative to its start; synchronises
    long long h[96];
    C1D_CUDA(ctx, cudaStreamSynchronize(s));
    C1D_CUDA(ctx, cudaMemcpy(h, st->d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
    const long long t0 = h[0];
    fprintf(stderr, "chi2 i8 group: digits done %lld  epilogues done %lld\n", h[1] - t0, h[2] - t0);
    for (int t = 0; t < st->p.ntile; ++t)
      fprintf(stderr, "  tile %d (%2d units): epilogue warp waits from %6lld to %6lld, loads done %6lld, after scales %6lld | MMA warp: at %6lld, digits ready %6lld, accumulators free %6lld, issued %6lld\n",
              t, st->p.tile_units[t], h[8 + 4 * t] - t0, h[9 + 4 * t] - t0, h[10 + 4 * t] - t0, h[11 + 4 * t] - t0, h[48 + 4 * t] - t0,
              h[49 + 4 * t] - t0, h[50 + 4 * t] - t0, h[51 + 4 * t] - t0);
  }
  return C1D_OK;
}

void c1d_chi2i8_destroy(c1d_ctx* ctx) {
  CQState* st = (CQState*)ctx->chi2i8_state;
  if (!st) return;
  cudaFree(st->d_img); cudaFree(st->d_rscale); cudaFree(st->d_dbg);
  delete st;
  ctx->chi2i8_state = nullptr;
}
