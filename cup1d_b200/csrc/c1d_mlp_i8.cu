// Stage 2 on the 5th-generation tensor cores with float64-class accuracy: the LaCE-style MLP as a
// chain of tcgen05.mma kind::i8 GEMMs over a 128-row tile of (walker, z) pairs, by digit slicing
// (an Ozaki-style error-free split).
//
// Arithmetic.  Every activation row and every weight row (one output neuron) is written as
//     x = s * q,   q = sum_i d_i 256^(3-i),   d_i in [-128, 127]  (four balanced base-256 digits)
// with a per-row scale s = max|x| / (127 * 2^24).  The product of two digit planes is an s8 x s8 GEMM
// whose int32 accumulation is EXACT (|sum| <= 4 * 256 * 127^2 < 2^24 per level); the ten digit pairs
// (i, j) with i + j <= 3 are accumulated into four level accumulators acc_l (l = i + j) and recombined in
// float64:  y_n = (((acc_0 * 256 + acc_1) * 256 + acc_2) * 256 + acc_3) * (2^24 s_a s_w,n) + b_n.
// The only errors are the two quantisations (2^-32 of the row maxima) and the dropped pairs of level >= 4:
// model P1D within 2e-9, log-likelihood within 4e-5 of the float64 evaluation on the DESI-DR1-like
// workload (tools/ozaki_sim.py; tests/test_gpu_i8.py), against 5e-3 for 3xTF32.
//
// One persistent CTA per SM, 4 I8_EPQ + 2 warps:
//   epilogue warps  I8_EPQ threads per row / TMEM lane (16-column batches in turn).  Pass 1: tcgen05.ld of
//              the four level accumulators -> float64 recombination -> scale, bias, LeakyReLU -> h written back
//              to tensor memory in place (two words), running row maximum.  Pass 2 (once the row maximum of the
//              layer is known): h -> q -> four digit bytes -> the A operand of the next layer (st.shared).
//   MMA warp   lane 0 issues ten tcgen05.mma per weight unit from descriptors held in registers
//              (tools/i8_probe.cu: an MMA holds the pipe for max(45, N / 2) cycles once the issue loop is lean)
//   streamer   lane 0 streams the pre-swizzled weight digit planes with cp.async.bulk through a ring of
//              four 16 KB slots; a unit is (layer, n-block, 32 k): 4 planes x nb rows x 32 B, SWIZZLE_32B.
//
// Memory plan (227 KB shared, 512 TMEM columns): the four level accumulators of a 128-column block fill the
// tensor memory, so layers wider than 128 run as two n-blocks.
//   width <= 128          one block
//   width <= 192          blocks (128, rest <= 64): h of block 0 stays in TMEM columns [0, 256) while block 1
//                         accumulates in [256, 512); one row scale for the layer
//   width <= 256          blocks (128, rest <= 128) with a row scale EACH: block 0 is quantised into region Y
//                         while the MMAs of block 1 still read the layer input in region X; the next layer sees
//                         two K-groups with their own scales and accumulates them separately (2 x 4 levels),
//                         which restricts it to 64 columns - the shape of LaCE's networks, whose widest hidden
//                         layer is followed by the 50-wide head.  Other shapes are reported as unsupported and
//                         the caller keeps the float64 kernel.
// A operand: digit planes in SWIZZLE_64B chunks of 128 rows x 64 k (8 KB); region X = 4 planes x 3 chunks,
// region Y = 4 planes x 2 chunks.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "c1d_internal.h"
#include "c1d_tc_common.cuh"

namespace {

using namespace c1d_tc;

constexpr int I8_ROWS = 128;
#ifndef I8_EPQ
#define I8_EPQ 3
#endif
constexpr int I8_EPI_WARPS = 4 * I8_EPQ;
constexpr int I8_EPI_THREADS = 32 * I8_EPI_WARPS;
constexpr int I8_THREADS = I8_EPI_THREADS + 64;
constexpr uint32_t I8_CHUNK = 8192;                   // 128 rows x 64 B
constexpr uint32_t I8_X_PSTRIDE = 3 * I8_CHUNK;
constexpr uint32_t I8_X_BYTES = 4 * I8_X_PSTRIDE;     // 96 KB
constexpr uint32_t I8_Y_PSTRIDE = 2 * I8_CHUNK;
constexpr uint32_t I8_Y_BYTES = 4 * I8_Y_PSTRIDE;     // 64 KB
constexpr uint32_t I8_SLOT = 16384;
constexpr int I8_NSLOT = 4;
constexpr uint32_t I8_RING = I8_SLOT * I8_NSLOT;      // 64 KB
constexpr uint32_t I8_RMAX_BYTES = I8_EPQ * I8_ROWS * 4;
constexpr uint32_t I8_OFF_X = 0, I8_OFF_Y = I8_X_BYTES, I8_OFF_RING = I8_OFF_Y + I8_Y_BYTES;
constexpr uint32_t I8_OFF_RMAX = I8_OFF_RING + I8_RING, I8_OFF_BAR = I8_OFF_RMAX + I8_RMAX_BYTES;
constexpr size_t I8_SMEM = I8_OFF_BAR + 128;
static_assert(I8_SMEM <= 232448, "shared memory budget");
constexpr int I8_MAX_UNITS = 96;
constexpr int I8_MAX_BLOCKS = 2 * C1D_MAX_LAYERS;
constexpr int I8_MAX_COLS = 1024;  // padded output columns of all layers together (the per-column constants ride in the kernel parameters)
constexpr double I8_FULL = 127.0 * 16777216.0;        // largest |q|: the top balanced digit stays within +-127

struct I8Unit {        // one weight unit = 32 k of one (layer, n-block, K-group): 16 bytes, ready-made descriptor fields
  uint32_t w_off;      // byte offset in the weight image
  uint32_t idesc;      // instruction descriptor of its MMAs (N = rows of the n-block, a multiple of 16)
  uint16_t a16;        // shared-memory offset of digit plane 0 of the A operand at this k-step, in 16-byte units
  uint16_t a_pstride16;  // distance between A digit planes, in 16-byte units
  uint16_t dcol;       // TMEM column of the level-0 accumulator of this (block, group)
  uint8_t lstride;     // columns between level accumulators
  uint8_t flags;       // 1 first k-step of its (block, group): overwrite; 2 last unit of its block: signal the
                       // epilogue; 4 first unit of its block: wait for the epilogue's go; 8 two k-steps (the second 32 k
                       // follow in the A chunk and in the weight image)
};
static_assert(sizeof(I8Unit) == 16, "I8Unit layout");

struct I8Blk {         // one (layer, n-block) as the epilogue sees it
  uint16_t n0, nb;     // first output column, width
  uint16_t dcol;       // TMEM column of level 0 of K-group 0
  uint16_t lstride, gstride;  // columns between levels / between K-groups
  uint16_t hlo, hhi;   // TMEM columns of the low / high words of h (column 0 of the block)
  uint16_t kbase;      // pass 2: k index of column 0 in the destination operand
  uint32_t dst_off;    // pass 2: shared-memory offset of digit plane 0 of the destination region
  uint16_t dst_pstride;  // bytes between destination planes / 16
  uint8_t layer, gin;
  uint8_t quant;       // 1: the row scale is fixed after this block and blocks [qfirst, this] are quantised
  uint8_t qfirst;
  uint8_t sslot;       // which of the two scales of the next layer's input the quantised block(s) define (its K-group)
  uint8_t lend;        // 1: last block of its layer (the next layer's scales become current)
  uint8_t last;        // 1: output layer
  uint8_t safe32;      // 1: acc_2 * 256 + acc_3 cannot overflow 32 bits for these weights (checked on the host)
};

struct I8Params {
  const uint8_t* wimg;
  const double2* sb;        // per layer, padded, concatenated: (2^24 * weight-row scale, bias)
  // the same in the kernel parameters (constant bank, 16 KB of the 32 KB a launch may carry): the epilogue reads the
  // 16 pairs of a batch through the uniform datapath instead of 16 LDG.128 into 64 vector registers whose latency sat
  // between the accumulator loads and their first use (dominant stall of the kernel: long scoreboard)
  double2 sbc[I8_MAX_COLS];
  const double *emu_lo, *emu_rng_inv;
  int n_units, n_blocks, n_in, last_out;
  int col_off[C1D_MAX_LAYERS];  // offset of the layer in sw / bias
  double slope;
  int exp;  // C1D_I8_EXP (measurements only, results are wrong): 1 = no weight copies, 2 = epilogues without arithmetic
  long long* dbg;  // C1D_I8_DEBUG=1: CTA 0 records the clock of its second tile per (layer, n-block), 8 slots each
  I8Blk blk[I8_MAX_BLOCKS];
  I8Unit units[I8_MAX_UNITS];
};

struct I8State {
  I8Params p;
  void *d_wimg, *d_sb, *d_rng_inv, *d_dbg;
  bool ready;
};

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(I8_EPI_THREADS) : "memory"); }

// int32 -> float64 without the conversion unit (I2F.F64 runs at 16 per clock per SM; the epilogue converts two or
// three integers per element): the integer is placed in the mantissa of 2^52 + 2^31 + x and the offset subtracted,
// one logic operation and one DADD.  Exact.
__device__ __forceinline__ double i2d_fast(int x) { return __hiloint2double(0x43300000, x ^ (int)0x80000000) - 4503601774854144.0; }

// recombined integer sum of 16 columns: T = ((acc0 * 256 + acc1) * 256 + acc2) * 256 + acc3, exact in float64.
// SAFE32: acc2 * 256 + acc3 fits 32 bits (two conversions and one DFMA per value instead of three and two).
template <bool SAFE32>
__device__ __forceinline__ void load_T16(uint32_t taddr, uint32_t ls, double* T) {
  uint32_t r0[16], r1[16], r2[16];
  tmem_ld16(taddr, r0);
  tmem_ld16(taddr + ls, r1);
  tmem_ld16(taddr + 2 * ls, r2);
  tc_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) r0[i] = (uint32_t)((int)r0[i] * 256 + (int)r1[i]);  // |.| < 2^31: 256 * 128^2 * 256 + 2 * 256 * 128^2
  tmem_ld16(taddr + 3 * ls, r1);
  tc_wait_ld();
  if (SAFE32) {
    // both halves are 32-bit integers: join them in 64-bit integer arithmetic (|T| < 2^48) and convert once through the
    // mantissa of 1.5 2^52 (one IMAD.WIDE, one add on the high word, one DADD instead of two conversions and a DFMA)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const long long t = (long long)(int)r0[i] * 65536 + (long long)((int)r2[i] * 256 + (int)r1[i]);
      T[i] = __longlong_as_double(t + 0x4338000000000000LL) - 6755399441055744.0;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) T[i] = fma(i2d_fast((int)r0[i]), 65536.0, fma(i2d_fast((int)r2[i]), 256.0, i2d_fast((int)r1[i])));
  }
}

// 16 quantised values (balanced digit bytes, most significant in byte 3) -> 16 bytes of each digit plane
__device__ __forceinline__ void store_digits16(uint32_t plane0, uint32_t pstride, uint32_t off, const uint32_t* v) {
  uint32_t pw[4][4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t a = __byte_perm(v[4 * g], v[4 * g + 1], 0x5140), b = __byte_perm(v[4 * g], v[4 * g + 1], 0x7362);
    const uint32_t c = __byte_perm(v[4 * g + 2], v[4 * g + 3], 0x5140), d = __byte_perm(v[4 * g + 2], v[4 * g + 3], 0x7362);
    pw[3][g] = __byte_perm(a, c, 0x5410);  // bytes 0: least significant digit = plane 3
    pw[2][g] = __byte_perm(a, c, 0x7632);
    pw[1][g] = __byte_perm(b, d, 0x5410);
    pw[0][g] = __byte_perm(b, d, 0x7632);
  }
#pragma unroll
  for (int p = 0; p < 4; ++p)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(plane0 + (uint32_t)p * pstride + off), "r"(pw[p][0]),
                 "r"(pw[p][1]), "r"(pw[p][2]), "r"(pw[p][3])
                 : "memory");
}

// 14 warps = 4 on two of the SM sub-partitions: 16 384 registers / (4 x 32) = 128 per thread
__global__ void __launch_bounds__(I8_THREADS, 1)
k_mlp_i8(I8Params p, const double* __restrict__ emu_in, int64_t nrows, double* __restrict__ out, int out_stride) {
  // the swizzled operands need a 1024-byte aligned base; the budget has no room for slack, so the kernel relies on
  // (and checks) the dynamic shared memory window starting on such a boundary when there is no static allocation
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t ring = base + I8_OFF_RING;
  const uint32_t bar0 = base + I8_OFF_BAR;
  const uint32_t bar_bfull = bar0, bar_bempty = bar0 + 8 * I8_NSLOT;
  const uint32_t bar_dfull = bar0 + 16 * I8_NSLOT, bar_go = bar_dfull + 8, tmem_slot = bar_go + 8;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(smem_raw + I8_OFF_BAR + 16 * I8_NSLOT + 16);
  uint32_t* rmax_s = reinterpret_cast<uint32_t*>(smem_raw + I8_OFF_RMAX);

  // the warp index through a shuffle: the compiler then knows that the role branches are warp-uniform and keeps
  // the MMA issuer's descriptors in uniform registers (a divergent `tid`-derived branch costs a seven-instruction
  // elect / broadcast loop per tcgen05.mma: 85 cycles per instruction, measured)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (tid == 0) {
    if (base & 1023u) {
      printf("cup1d_b200: k_mlp_i8 shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    for (int s = 0; s < I8_NSLOT; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    mbar_init(bar_dfull, 1);
    mbar_init(bar_go, I8_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == I8_EPI_WARPS) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_p, 0);

  const int64_t ntiles = (nrows + I8_ROWS - 1) / I8_ROWS;
  const int64_t tiles_per_cta = (ntiles + gridDim.x - 1) / gridDim.x;

  if (warp < I8_EPI_WARPS) {
    // ======================= row owners: quantisation and epilogues =======================
    const uint32_t quad = (uint32_t)warp & 3u, sub = (uint32_t)warp >> 2;
    const uint32_t row = quad * 32u + ((uint32_t)tid & 31u);
    const uint32_t lane_addr = tmem_base + ((quad * 32u) << 16);
    const int slope_hi = __double2hiint(p.slope), slope_lo = __double2loint(p.slope);
    uint32_t d_phase = 0;
    for (int64_t it = 0; it < tiles_per_cta; ++it) {
      const int64_t tile = it * gridDim.x + blockIdx.x;
      const int64_t grow = tile * I8_ROWS + row;
      bool bad = false;
      const bool dbg = p.dbg && blockIdx.x == 0 && tid == 0 && it == 1;
      double sa0 = 0.0, sa1 = 0.0;  // row scales of the current A operand (per K-group)
      double sn0 = 0.0, sn1 = 0.0;  // row scales of the operand being produced
      // ---- normalised inputs -> digit planes of the first A operand (k padded to 32 with zeros) ----
      {
        double xn[C1D_MAX_EMU];
        double m = 0.0;
#pragma unroll
        for (int j = 0; j < C1D_MAX_EMU; ++j) {
          xn[j] = 0.0;
          if (j < p.n_in) {
            const double xv = (grow < nrows) ? emu_in[grow * C1D_MAX_EMU + j] : 0.0;
            xn[j] = (xv - __ldg(p.emu_lo + j)) * __ldg(p.emu_rng_inv + j) - 0.5;
            m = fmax(m, fabs(xn[j]));
            if (!(fabs(xn[j]) < INFINITY)) bad = true;
          }
        }
        sa0 = m * (1.0 / I8_FULL);
        const double inv_s = (m > 0.0 && !bad) ? I8_FULL / m : 0.0;
        if (sub < 2) {
          uint32_t v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            uint32_t q = 0;
            if (sub == 0 && i < C1D_MAX_EMU) q = (uint32_t)__double2int_rn(xn[i] * inv_s);
            v[i] = (q + 0x80808080u) ^ 0x80808080u;
          }
          store_digits16(base + I8_OFF_X, I8_X_PSTRIDE, sw64_off(row, sub), v);
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(bar_go);

      uint32_t mhi = 0;  // running maximum of the high words of |h| over the blocks that share a scale
      for (int b = 0; b < p.n_blocks; ++b) {
        const I8Blk& B = p.blk[b];
        if (dbg) p.dbg[8 * b + 0] = clock64();
        mbar_wait(bar_dfull, d_phase);
        d_phase ^= 1;
        tc_fence_after();
        if (dbg) p.dbg[8 * b + 1] = clock64();
        const double2* sbp = p.sbc + p.col_off[B.layer] + B.n0;
        const uint32_t nb = B.nb, ls = B.lstride;
        if (B.last) {
          // ---- output layer: coefficients = T * scale + bias ----
          if (sub == 0) {
            double T[16];
            load_T16<false>(lane_addr + B.dcol, ls, T);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const double2 cb = sbp[i];
              T[i] = fma(T[i], sa0 * cb.x, cb.y);
            }
            if (B.gin == 2) {
              double U[16];
              load_T16<false>(lane_addr + B.dcol + B.gstride, ls, U);
#pragma unroll
              for (int i = 0; i < 16; ++i) T[i] = fma(U[i], sa1 * sbp[i].x, T[i]);
            }
            if (grow < nrows) {
#pragma unroll
              for (int o = 0; o < C1D_MAX_COEF; ++o)
                if (o < p.last_out) out[grow * out_stride + o] = bad ? NAN : T[o];
            }
          }
          tc_fence_before();
          if (dbg) p.dbg[8 * b + 2] = p.dbg[8 * b + 3] = clock64();
          continue;  // the next tile's input preparation gives the go
        }
        // ---- pass 1: accumulators -> h (float64, back into tensor memory), row maximum ----
        for (uint32_t c = 16u * sub; c < nb && !(p.exp & 2); c += 16u * I8_EPQ) {
          double T[16];
          if (B.safe32) load_T16<true>(lane_addr + B.dcol + c, ls, T);
          else load_T16<false>(lane_addr + B.dcol + c, ls, T);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const double2 cb = sbp[c + i];
            T[i] = fma(T[i], sa0 * cb.x, cb.y);
          }
          if (B.gin == 2) {
            double U[16];
            if (B.safe32) load_T16<true>(lane_addr + B.dcol + B.gstride + c, ls, U);
            else load_T16<false>(lane_addr + B.dcol + B.gstride + c, ls, U);
#pragma unroll
            for (int i = 0; i < 16; ++i) T[i] = fma(U[i], sa1 * sbp[c + i].x, T[i]);
          }
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            // LeakyReLU as a multiplication by 1 or the slope, chosen from the sign bit
            const bool neg = __double2hiint(T[i]) < 0;
            const double h = T[i] * __hiloint2double(neg ? slope_hi : 0x3ff00000, neg ? slope_lo : 0);
            lo[i] = (uint32_t)__double2loint(h);
            hi[i] = (uint32_t)__double2hiint(h);
            mhi = max(mhi, hi[i] & 0x7fffffffu);  // |h| orders like its high word; NaN / Inf carry the largest
          }
          tmem_st16(lane_addr + B.hlo + c, lo);
          tmem_st16(lane_addr + B.hhi + c, hi);
        }
        tc_wait_st();
        if (dbg) p.dbg[8 * b + 2] = clock64();
        if (B.quant) {
          // ---- row maximum over the I8_EPQ threads of the row ----
          epi_bar();  // earlier readers of rmax_s are done
          rmax_s[sub * I8_ROWS + row] = mhi;
          epi_bar();
          uint32_t m = 0;
#pragma unroll
          for (int s = 0; s < I8_EPQ; ++s) m = max(m, rmax_s[s * I8_ROWS + row]);
          const bool isbad = m >= 0x7ff00000u;
          bad = bad || isbad;
          // upper bound of max|h|: the next value of the high word (2^-20 above at most)
          const double mx = (m == 0u) ? 0.0 : __hiloint2double((int)(m + 1u), 0);
          const double inv_s = (m != 0u && !isbad) ? I8_FULL / mx : 0.0;
          if (B.sslot) sn1 = mx * (1.0 / I8_FULL);
          else sn0 = mx * (1.0 / I8_FULL);
          // ---- pass 2: h -> digits of the next A operand ----
          for (int bq = B.qfirst; bq <= b; ++bq) {
            const I8Blk& Q = p.blk[bq];
            const uint32_t plane0 = base + Q.dst_off, pstride = (uint32_t)Q.dst_pstride * 16u;
            for (uint32_t c = 16u * sub; c < Q.nb && !(p.exp & 2); c += 16u * I8_EPQ) {
              uint32_t lo[16], hi[16], v[16];
              tmem_ld16(lane_addr + Q.hlo + c, lo);
              tmem_ld16(lane_addr + Q.hhi + c, hi);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const double h = __hiloint2double((int)hi[i], (int)lo[i]);
                // round to nearest through the low word of h / s + 1.5 2^52 (|q| < 2^31): no conversion instruction
                const uint32_t q = (uint32_t)__double2loint(fma(h, inv_s, 6755399441055744.0));
                v[i] = (q + 0x80808080u) ^ 0x80808080u;
              }
              const uint32_t k = (uint32_t)Q.kbase + c;
              store_digits16(plane0, pstride, (k >> 6) * I8_CHUNK + sw64_off(row, (k & 63u) >> 4), v);
            }
          }
          mhi = 0;
          fence_async_smem();
        }
        if (B.lend) { sa0 = sn0; sa1 = sn1; }
        tc_fence_before();
        mbar_arrive(bar_go);
        if (dbg) p.dbg[8 * b + 3] = clock64();
      }
    }
  } else if (warp == I8_EPI_WARPS) {
    // ================================ MMA issuer ================================
    const bool leader = elect_one();
    const uint64_t desc_sw64_0 = make_desc_sw64(0), desc_sw32_0 = make_desc_sw32(0);
    const uint32_t base16 = base >> 4, ring16 = ring >> 4;
    uint32_t go_phase = 0, g = 0;
    // The fields of a unit (constant-bank loads indexed by the unit, descriptor arithmetic, the moves into uniform
    // registers) are prepared one unit AHEAD, before the MMAs of the current unit are issued: the issue of an MMA waits
    // for the tensor pipe, so whatever sits between the last MMA of one unit and the first of the next is idle tensor
    // time (measured: ~240 cycles per unit, 40 units per tile, before this change)
    uint32_t n_flags, n_idesc, n_d0, n_ls, n_slot;
    uint64_t n_ad0, n_astep, n_bd0, n_bstep;
    uint32_t r_flags, r_idesc, r_a16, r_apst, r_dcol, r_ls;  // raw fields of the unit after the current one
#define I8_LOAD(U_)                       \
  do {                                    \
    r_flags = p.units[U_].flags;          \
    r_idesc = p.units[U_].idesc;          \
    r_a16 = p.units[U_].a16;              \
    r_apst = p.units[U_].a_pstride16;     \
    r_dcol = p.units[U_].dcol;            \
    r_ls = p.units[U_].lstride;           \
  } while (0)
#define I8_DERIVE(G_)                                                                \
  do {                                                                               \
    n_flags = r_flags;                                                               \
    n_idesc = r_idesc;                                                               \
    n_slot = (G_) % I8_NSLOT;                                                        \
    n_ad0 = desc_sw64_0 | (uint64_t)(base16 + r_a16);                                \
    n_astep = r_apst;                                                                \
    n_bd0 = desc_sw32_0 | (uint64_t)(ring16 + n_slot * (I8_SLOT >> 4));              \
    n_bstep = (r_idesc >> 17) << 4; /* nb rows x 32 B per plane, / 16 */             \
    n_d0 = tmem_base + r_dcol;                                                       \
    n_ls = r_ls;                                                                     \
  } while (0)
    I8_LOAD(0);
    I8_DERIVE(g);
    for (int64_t it = 0; it < tiles_per_cta; ++it) {
      const bool dbg = p.dbg && blockIdx.x == 0 && leader && it == 1;
      const bool dbgw = p.dbg && blockIdx.x == 0 && it == 1;
      long long st0 = 0, st1 = 0, st2 = 0;  // per-unit clocks, one unit per lane (cheap: no memory traffic inside the loop)
      int bidx = -1;
      for (int u = 0; u < p.n_units; ++u, ++g) {
        const uint32_t uflags = n_flags, idesc = n_idesc, d0 = n_d0, ls = n_ls, slot = n_slot;
        // descriptors: constant fields | address >> 4 (addresses stay below 2^18, no masking needed)
        const uint64_t ad0 = n_ad0, astep = n_astep, bd0 = n_bd0, bstep = n_bstep;
        // the constant-bank loads of the next unit are issued here and consumed after this unit's MMAs are issued
        // (the asm pins them in program order: the MMAs below are volatile statements)
        const int un = (u + 1 < p.n_units) ? u + 1 : 0;
        I8_LOAD(un);
        asm volatile("" : "+r"(r_flags), "+r"(r_idesc), "+r"(r_a16), "+r"(r_apst), "+r"(r_dcol), "+r"(r_ls)::"memory");
        const bool dbgu = dbg && p.dbg && (p.exp & 4) && u < 64;
        long long q0 = 0, q1 = 0, q2 = 0, q3 = 0;
        if (dbgu) q0 = clock64();
        if (uflags & 4u) {
          ++bidx;
          if (dbg) p.dbg[8 * bidx + 4] = clock64();
          mbar_wait(bar_go, go_phase);
          go_phase ^= 1;
          tc_fence_after();
          if (dbg) p.dbg[8 * bidx + 5] = clock64();
        }
        // (no tcgen05 fence here: the barrier's complete_tx orders the bulk copy before the MMAs that read it; the fence
        // after the `go` wait above orders the epilogue warps' tensor-memory accesses)
        if (!(p.exp & 1)) mbar_wait_fast(bar_bfull + 8 * slot, (g / I8_NSLOT) & 1u);
        const uint32_t keep = (uflags & 1u) ? 0u : 1u;
        if (dbgu) q1 = clock64();
        if (leader) {
          // Digit planes j, j + 1 of the weights are consecutive row groups of one operand and the level accumulators
          // consecutive column groups (lstride = nb), so one instruction of N = 2 nb serves two pairs: six instructions
          // instead of ten, 64 KB instead of 80 KB of operand reads per unit (the SS form is shared-memory bound at
          // N <= 128: tools/i8_probe.cu)
          const uint32_t nfield = idesc & (63u << 17);
          const uint32_t idesc2 = idesc + nfield;  // N -> 2 N
          if (nfield <= (8u << 17)) {
            // blocks of at most 64 columns: plane i of the activations against ALL the weight planes it pairs with
            // (N = (4 - i) nb <= 256), four instructions per unit and each A plane read once
            mma_i8_ss(d0, ad0, bd0, idesc2 + 2 * nfield, keep);                   // (0,0..3) -> levels 0..3
            mma_i8_ss(d0 + ls, ad0 + astep, bd0, idesc2 + nfield, 1u);            // (1,0..2) -> levels 1..3
            mma_i8_ss(d0 + 2 * ls, ad0 + 2 * astep, bd0, idesc2, 1u);             // (2,0..1) -> levels 2, 3
            mma_i8_ss(d0 + 3 * ls, ad0 + 3 * astep, bd0, idesc, 1u);              // (3,0)    -> level 3
            if (uflags & 8u) {  // the unit's second k-step: 32 bytes on in the A chunk, four planes on in the slot
              const uint64_t ad1 = ad0 + 2, bd1 = bd0 + 4 * bstep;
              mma_i8_ss(d0, ad1, bd1, idesc2 + 2 * nfield, 1u);
              mma_i8_ss(d0 + ls, ad1 + astep, bd1, idesc2 + nfield, 1u);
              mma_i8_ss(d0 + 2 * ls, ad1 + 2 * astep, bd1, idesc2, 1u);
              mma_i8_ss(d0 + 3 * ls, ad1 + 3 * astep, bd1, idesc, 1u);
            }
          } else {
          mma_i8_ss(d0, ad0, bd0, idesc2, keep);                                  // (0,0) (0,1) -> levels 0, 1
          mma_i8_ss(d0 + 2 * ls, ad0, bd0 + 2 * bstep, idesc2, keep);             // (0,2) (0,3) -> levels 2, 3
          mma_i8_ss(d0 + ls, ad0 + astep, bd0, idesc2, 1u);                       // (1,0) (1,1) -> levels 1, 2
          mma_i8_ss(d0 + 3 * ls, ad0 + astep, bd0 + 2 * bstep, idesc, 1u);        // (1,2)       -> level 3
          mma_i8_ss(d0 + 2 * ls, ad0 + 2 * astep, bd0, idesc2, 1u);               // (2,0) (2,1) -> levels 2, 3
          mma_i8_ss(d0 + 3 * ls, ad0 + 3 * astep, bd0, idesc, 1u);                // (3,0)       -> level 3
          }
          if (dbgu) q2 = clock64();
          tc_commit(bar_bempty + 8 * slot);
          if (uflags & 2u) tc_commit(bar_dfull);
          if (dbg && (uflags & 2u)) p.dbg[8 * bidx + 6] = clock64();
          if (dbgu) {
            q3 = clock64();
            long long* o = p.dbg + 8 * I8_MAX_BLOCKS + 96 + 4 * u;
            o[0] = q0; o[1] = q1; o[2] = q2; o[3] = q3;
          }
        }
        I8_DERIVE(g + 1);
        asm volatile("" : "+r"(n_flags), "+r"(n_idesc), "+r"(n_d0), "+r"(n_ls), "+r"(n_slot), "+l"(n_ad0), "+l"(n_astep), "+l"(n_bd0), "+l"(n_bstep)::"memory");
        __syncwarp();
        if (dbgw) {
          const long long now = clock64();
          const bool mine = (u & 31) == (tid & 31);
          st0 = (mine && u < 32) ? now : st0;
          st1 = (mine && u >= 32 && u < 64) ? now : st1;
          st2 = (mine && u >= 64) ? now : st2;
        }
      }
      if (dbgw) {
        p.dbg[8 * I8_MAX_BLOCKS + (tid & 31)] = st0;
        p.dbg[8 * I8_MAX_BLOCKS + 32 + (tid & 31)] = st1;
        p.dbg[8 * I8_MAX_BLOCKS + 64 + (tid & 31)] = st2;
      }
    }
  } else {
    // ============================== weight streamer ==============================
    const bool leader = (tid & 31) == 0;
    uint32_t g = 0;
    for (int64_t it = 0; it < tiles_per_cta; ++it) {
      for (int u = 0; u < p.n_units; ++u, ++g) {
        const uint32_t slot = g % I8_NSLOT;
        if (g >= (uint32_t)I8_NSLOT) mbar_wait(bar_bempty + 8 * slot, ((g / I8_NSLOT) - 1u) & 1u);
        if (leader && !(p.exp & 1)) {
          const uint32_t bytes = (((p.units[u].idesc >> 17) & 63u) << 10) << ((p.units[u].flags >> 3) & 1u);  // nb x 128 B per k-step
          mbar_expect_tx(bar_bfull + 8 * slot, bytes);
          bulk_g2s(ring + slot * I8_SLOT, p.wimg + p.units[u].w_off, bytes, bar_bfull + 8 * slot);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == I8_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
int c1d_i8_prepare(c1d_ctx* ctx) {
  const c1d_config& c = ctx->cfg;
  if (ctx->i8_state) c1d_i8_destroy(ctx);
  if (c.emu_kind != 0 || c.n_layers < 1) return C1D_OK;
  I8State* st = new I8State();
  memset(st, 0, sizeof(*st));
  ctx->i8_state = st;
  const int L = c.n_layers;
  // ---- shape plan; unsupported shapes leave the state not ready (the float64 kernel serves them) ----
  std::vector<int> npad(L), kpad(L), gin(L), nblk(L), mode(L);
  {
    int g = 1;
    for (int l = 0; l < L; ++l) {
      const int N = c.layer_dims[l + 1];
      npad[l] = (l == L - 1) ? ((N + 15) & ~15) : ((N + 31) & ~31);
      kpad[l] = (l == 0) ? 32 : npad[l - 1];
      gin[l] = g;
      if (l == 0 && c.layer_dims[0] > C1D_MAX_EMU) return C1D_OK;
      if (g == 2 && npad[l] > 64) return C1D_OK;
      if (l == L - 1 && npad[l] > 16) return C1D_OK;
      if (npad[l] <= 128) { nblk[l] = 1; mode[l] = 0; g = 1; }
      else if (npad[l] <= 192) { nblk[l] = 2; mode[l] = 0; g = 1; }
      else {
        if (l == L - 1) return C1D_OK;
        nblk[l] = 2; mode[l] = 1; g = 2;
      }
    }
  }
  size_t nw = 0, nbias = 0;
  std::vector<size_t> woff(L), boff(L);
  for (int l = 0; l < L; ++l) {
    woff[l] = nw;
    boff[l] = nbias;
    nw += (size_t)c.layer_dims[l] * c.layer_dims[l + 1];
    nbias += c.layer_dims[l + 1];
  }
  std::vector<double> W(nw), Bv(nbias);
  C1D_CUDA(ctx, cudaMemcpy(W.data(), ctx->dev[C1D_F_NN_W], nw * sizeof(double), cudaMemcpyDeviceToHost));
  C1D_CUDA(ctx, cudaMemcpy(Bv.data(), ctx->dev[C1D_F_NN_B], nbias * sizeof(double), cudaMemcpyDeviceToHost));
  auto wat = [&](int l, int n, int k) { return W[woff[l] + (size_t)k * c.layer_dims[l + 1] + n]; };  // in-major storage

  I8Params& p = st->p;
  p.n_in = c.layer_dims[0];
  p.last_out = c.layer_dims[L];
  p.slope = c.nn_slope;
  std::vector<I8Unit> units;
  std::vector<I8Blk> blks;
  std::vector<uint8_t> img;
  std::vector<double2> sb;
  for (int l = 0; l < L; ++l) {
    const int K = c.layer_dims[l], N = c.layer_dims[l + 1];
    p.col_off[l] = (int)sb.size();
    const size_t sb0 = sb.size();
    // digits of the weight rows
    std::vector<uint32_t> qd((size_t)npad[l] * kpad[l], 0u);  // balanced digit bytes of q[n][k] (0 = value 0)
    for (int n = 0; n < npad[l]; ++n) {
      double m = 0.0;
      if (n < N)
        for (int k = 0; k < K; ++k) m = fmax(m, fabs(wat(l, n, k)));
      const double s = m / I8_FULL;
      sb.push_back(make_double2(s * 16777216.0, n < N ? Bv[boff[l] + n] : 0.0));
      if (n < N && m > 0.0)
        for (int k = 0; k < K; ++k) {
          const long long q = llrint(wat(l, n, k) / s);
          qd[(size_t)n * kpad[l] + k] = ((uint32_t)(int32_t)q + 0x80808080u) ^ 0x80808080u;
        }
    }
    // blocks.  On the N side a layer is padded to 16 columns only (the MMA shapes need no more): the k columns of the
    // next operand between that and its 32-k padding are never written, and meet zero weight digits.
    const int n16 = (N + 15) & ~15;
    int nbs[2] = {n16, 0};
    if (nblk[l] == 2) { nbs[0] = 128; nbs[1] = n16 - 128; }
    const bool out_groups = (mode[l] == 1);  // this layer's blocks become the K-groups of the next layer
    int n0 = 0;
    for (int b = 0; b < nblk[l]; ++b) {
      const int nb = nbs[b];
      I8Blk B;
      memset(&B, 0, sizeof(B));
      B.n0 = (uint16_t)n0;
      B.nb = (uint16_t)nb;
      B.layer = (uint8_t)l;
      B.gin = (uint8_t)gin[l];
      B.last = (l == L - 1) ? 1 : 0;
      B.lend = (b == nblk[l] - 1) ? 1 : 0;
      {
        // |acc_l| <= 128 * sum over the pairs of level l of sum_k |weight digit|: is acc_2 * 256 + acc_3 a 32-bit value?
        double worst = 0.0;
        for (int g = 0; g < gin[l]; ++g) {
          const int k_lo = (g == 0) ? 0 : 128, k_hi = (gin[l] == 2 && g == 0) ? 128 : kpad[l];
          for (int n = n0; n < n0 + nb; ++n) {
            double S[4] = {0, 0, 0, 0};
            for (int k = k_lo; k < k_hi; ++k) {
              const uint32_t v = qd[(size_t)n * kpad[l] + k];
              for (int pl = 0; pl < 4; ++pl) S[pl] += fabs((double)(int8_t)(uint8_t)(v >> (8 * (3 - pl))));
            }
            const double s012 = S[0] + S[1] + S[2];
            worst = fmax(worst, 128.0 * (256.0 * s012 + s012 + S[3]));
          }
        }
        B.safe32 = (worst < 2147483648.0) ? 1 : 0;
      }
      if (gin[l] == 2) { B.dcol = 0; B.lstride = 64; B.gstride = 256; }
      else if (b == 1 && mode[l] == 0) { B.dcol = 256; B.lstride = 64; B.gstride = 0; }
      else { B.dcol = 0; B.lstride = 128; B.gstride = 0; }
      B.lstride = (uint16_t)nb;  // consecutive level accumulators: one MMA covers two of them
      B.hlo = B.dcol;
      B.hhi = (uint16_t)(B.dcol + B.lstride);
      if (out_groups) {
        B.quant = 1; B.qfirst = (uint8_t)blks.size(); B.sslot = (uint8_t)b; B.kbase = 0;
        B.dst_off = (b == 0) ? I8_OFF_Y : I8_OFF_X;
        B.dst_pstride = (uint16_t)(((b == 0) ? I8_Y_PSTRIDE : I8_X_PSTRIDE) / 16);
      } else {
        B.quant = (b == nblk[l] - 1) ? 1 : 0;
        B.qfirst = (uint8_t)(blks.size() - b);
        B.sslot = 0;
        B.kbase = (uint16_t)n0;
        B.dst_off = I8_OFF_X;
        B.dst_pstride = (uint16_t)(I8_X_PSTRIDE / 16);
      }
      blks.push_back(B);
      // weight units of the block: K-groups in turn, 32 k each
      for (int g = 0; g < gin[l]; ++g) {
        const int k_lo = (g == 0) ? 0 : 128, k_hi = (gin[l] == 2 && g == 0) ? 128 : kpad[l];
        const uint32_t a_region = (gin[l] == 2 && g == 0) ? I8_OFF_Y : I8_OFF_X;
        const uint32_t a_pstride = (gin[l] == 2 && g == 0) ? I8_Y_PSTRIDE : I8_X_PSTRIDE;
        const int nks = (k_hi - k_lo) / 32;
        for (int ks = 0; ks < nks; ++ks) {
          // blocks of at most 64 columns take two k-steps per unit where a pair exists (their MMAs are short: the fixed
          // cost of a unit - barrier check, commit, loop - is what paces them)
          const int nk = (nb <= 64 && !(ks & 1) && ks + 1 < nks) ? 2 : 1;
          I8Unit u;
          memset(&u, 0, sizeof(u));
          u.w_off = (uint32_t)img.size();
          u.a16 = (uint16_t)((a_region + (uint32_t)(ks >> 1) * I8_CHUNK + (uint32_t)(ks & 1) * 32u) >> 4);
          u.a_pstride16 = (uint16_t)(a_pstride / 16);
          u.idesc = (2u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)nb >> 3) << 17) | ((128u >> 4) << 24);  // make_idesc_s8(nb)
          u.dcol = (uint16_t)(B.dcol + g * B.gstride);
          u.lstride = (uint8_t)B.lstride;
          u.flags = (uint8_t)((ks == 0 ? 1 : 0) | ((g == gin[l] - 1 && ks + nk == nks) ? 2 : 0) | ((g == 0 && ks == 0) ? 4 : 0) | (nk == 2 ? 8 : 0));
          units.push_back(u);
          for (int sk = 0; sk < nk; ++sk) {
            const size_t o = img.size();
            img.resize(o + (size_t)nb * 128, 0);
            for (int pl = 0; pl < 4; ++pl)
              for (int n = 0; n < nb; ++n)
                for (int kk = 0; kk < 32; ++kk) {
                  const uint32_t v = qd[(size_t)(n0 + n) * kpad[l] + (k_lo + 32 * (ks + sk) + kk)];
                  img[o + (size_t)pl * nb * 32 + sw32_off((uint32_t)n, (uint32_t)kk >> 4) + (kk & 15)] = (uint8_t)(v >> (8 * (3 - pl)));
                }
          }
          ks += nk - 1;
        }
      }
      n0 += nb;
    }
  }
  if ((int)units.size() > I8_MAX_UNITS || (int)blks.size() > I8_MAX_BLOCKS) return C1D_OK;  // not ready: float64 kernel
  p.n_units = (int)units.size();
  p.n_blocks = (int)blks.size();
  for (size_t q = 0; q < units.size(); ++q) p.units[q] = units[q];
  for (size_t q = 0; q < blks.size(); ++q) p.blk[q] = blks[q];
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
  UP(st->d_sb, sb.data(), sb.size() * sizeof(double2));
  UP(st->d_rng_inv, rinv.data(), rinv.size() * sizeof(double));
#undef UP
  p.wimg = (const uint8_t*)st->d_wimg;
  p.sb = (const double2*)st->d_sb;
  if (sb.size() > (size_t)I8_MAX_COLS) return C1D_OK;  // unsupported shape: the caller keeps the float64 kernel
  memcpy(p.sbc, sb.data(), sb.size() * sizeof(double2));
  p.dbg = nullptr;
  p.exp = getenv("C1D_I8_EXP") ? atoi(getenv("C1D_I8_EXP")) : 0;
  if (getenv("C1D_I8_DEBUG")) {
    C1D_CUDA(ctx, cudaMalloc(&st->d_dbg, (8 * I8_MAX_BLOCKS + 4 * I8_MAX_UNITS) * sizeof(long long)));
    C1D_CUDA(ctx, cudaMemset(st->d_dbg, 0, (8 * I8_MAX_BLOCKS + 4 * I8_MAX_UNITS) * sizeof(long long)));
    p.dbg = (long long*)st->d_dbg;
  }
  p.emu_lo = ctx->d.emu_lo;
  p.emu_rng_inv = (const double*)st->d_rng_inv;
  C1D_CUDA(ctx, cudaFuncSetAttribute(k_mlp_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8_SMEM));
  st->ready = true;
  return C1D_OK;
}

int c1d_i8_available(const c1d_ctx* ctx) {
  const I8State* st = (const I8State*)ctx->i8_state;
  return st && st->ready;
}

int c1d_i8_launch(c1d_ctx* ctx, const double* emu_in, int64_t nrows, double* out, int out_stride, cudaStream_t s) {
  I8State* st = (I8State*)ctx->i8_state;
  if (!st || !st->ready) {
    snprintf(ctx->err, sizeof(ctx->err),
             "int8 digit-slice emulator path unavailable (needs an NN emulator whose layers wider than 192 are followed by one of at most 64)");
    return C1D_ERR_STATE;
  }
  const int64_t ntiles = (nrows + I8_ROWS - 1) / I8_ROWS;
  const int grid = (int)(ntiles < ctx->num_sms ? ntiles : ctx->num_sms);
  k_mlp_i8<<<grid, I8_THREADS, I8_SMEM, s>>>(st->p, emu_in, nrows, out, out_stride);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  if (st->p.dbg && ntiles > grid) {
    // diagnostic: cycles of CTA 0's second tile per (layer, n-block); synchronises
    long long h[8 * I8_MAX_BLOCKS + 4 * I8_MAX_UNITS];
    C1D_CUDA(ctx, cudaStreamSynchronize(s));
    C1D_CUDA(ctx, cudaMemcpy(h, st->d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
    long long t_mma = 0, t_p1 = 0, t_q = 0;
    for (int b = 0; b < st->p.n_blocks; ++b) {
      const long long* q = h + 8 * b;
      fprintf(stderr, "i8 blk %2d layer %d nb %3d gin %d safe32 %d: go->issued %6lld  go->dfull %6lld  pass1 %6lld  quant %6lld  | MMA warp waited for go %6lld\n", b,
              (int)st->p.blk[b].layer, (int)st->p.blk[b].nb, (int)st->p.blk[b].gin, (int)st->p.blk[b].safe32, q[6] - q[5], q[1] - q[5], q[2] - q[1], q[3] - q[2], q[5] - q[4]);
      t_mma += q[1] - q[5]; t_p1 += q[2] - q[1]; t_q += q[3] - q[2];
    }
    if (getenv("C1D_I8_DEBUG")[0] == '2') {
      const long long* q = h + 8 * I8_MAX_BLOCKS;
      for (int u = 0; u < st->p.n_units; ++u)
        fprintf(stderr, "  unit %2d nb %3d flags %d: issued %6lld cycles after the previous unit\n", u, (int)(((st->p.units[u].idesc >> 17) & 63u) << 3),
                (int)st->p.units[u].flags, u ? q[u] - q[u - 1] : 0LL);
    }
    if (st->p.exp & 4) {
      const long long* q = h + 8 * I8_MAX_BLOCKS + 96;
      for (int u = 0; u < st->p.n_units && u < 64; ++u)
        fprintf(stderr, "  unit %2d nb %3d: top->waited %5lld  MMAs issued %5lld  commits %5lld | top since previous top %6lld\n", u,
                (int)(((st->p.units[u].idesc >> 17) & 63u) << 3), q[4 * u + 1] - q[4 * u], q[4 * u + 2] - q[4 * u + 1], q[4 * u + 3] - q[4 * u + 2],
                u ? q[4 * u] - q[4 * u - 4] : 0LL);
    }
    fprintf(stderr, "i8 tile: MMA (go->dfull) %lld  pass1 %lld  quant %lld  whole tile %lld cycles\n", t_mma, t_p1, t_q,
            h[8 * (st->p.n_blocks - 1) + 3] - h[4]);
  }
  return C1D_OK;
}

void c1d_i8_destroy(c1d_ctx* ctx) {
  I8State* st = (I8State*)ctx->i8_state;
  if (!st) return;
  cudaFree(st->d_wimg); cudaFree(st->d_sb); cudaFree(st->d_rng_inv); cudaFree(st->d_dbg);
  delete st;
  ctx->i8_state = nullptr;
}
