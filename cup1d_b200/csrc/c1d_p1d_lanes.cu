// Stage 3, walker-per-lane form: emulator polynomial -> contaminant templates -> rebinning ->
// residual (reference: lya_theory.py:690-784, si_mult.py:157-343, si_vid_final.py:157-221,
// si_add.py:130-221, hcd_model_rogers_class.py:94-140, resolution_class.py:88-114,
// likelihood.py:147-161, 939, 957).
//
// A warp owns a work item (z, 32 walkers): lane = walker.  All lanes walk the fine k grid of the
// redshift bin together, so every table value is ONE broadcast shared-memory read for the warp
// (the warp-per-walker kernel in c1d_kernels.cu reads a different point per lane: 4x the
// wavefronts), the walker's amplitudes are genuinely per-lane registers, the exp recurrences are
// set up once per (walker, z) per LANE instead of 32 times, and the rebinning needs no staging
// buffer: a fine point feeds at most two adjacent data bins (checked on the host), so two running
// sums per lane are flushed as the bins complete — in the same summation order as the banded form.
//
// When the per-z chi2 is requested (block-diagonal contexts, lnL_z outputs) the residuals are also
// written to a z-aligned buffer and a small tensor-core kernel (k_chi2z_dmma) forms d_z^T C_z^-1 d_z.
// Small batches, and rebinning operators that are not of the two-bin form, fall back to the
// warp-per-walker kernel in c1d_launch_stage3.
//
// Medium batches (fewer than ~4 waves of (z, 32 walkers) items) cut every redshift bin into up to 8
// segments along k: a segment starts on a recurrence anchor (multiple of PL_ANCHOR fine points), owns the
// data bins whose first fine point lies in it and runs on to the last fine point those bins need, so the
// few points around a cut are evaluated by both neighbours.  Every model value and every bin sum is formed
// by the same operations in the same order as in the uncut kernel: the results are bit-identical for any
// number of segments (tests/test_gpu_parity.py), only the work items get finer.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "c1d_internal.h"
#include "c1d_devmath.cuh"
#include "c1d_tc_common.cuh"

namespace {

constexpr int PL_ANCHOR = 64;  // fine points between exact re-evaluations of the exp recurrences
// U fine points per loop trip as independent instruction streams, with 64K / threads registers per
// thread.  Measured on B200 (C4, 65 536 walkers): U=1 x 512 threads 2.52 ms, U=2 x 384 3.00 ms,
// U=4 x 384 2.78 ms, U=4 x 256 3.57 ms, and U=1 with 640 / 768 / 1024 threads (spilling) 2.96 / 3.23 /
// 3.88 ms: warps beat unrolling, and 16 warps at the 128-register cap is the optimum.  (Later, with the table-based
// 10^x: U=1 x 512 2.11 ms, U=2 x 512 - no spills at 128 registers, but the two streams are serialised - 2.36 ms.)
constexpr int PL_U = 1;
// Cut the redshift bins until the batch holds PL_WAVES waves of items, and hand batches below PL_MIN_WAVES2 / 2
// waves of the finest items to the warp-per-walker kernel.  Measured on B200 (C4, stage 3 in ms, warp-per-walker
// kernel | 1, 2, 4, 8 segments): 2048 walkers 0.126 | 0.391 0.224 0.176 0.142; 4096: 0.223 | 0.405 0.274 0.233 0.207;
// 8192: 0.413 | 0.466 0.385 0.349 0.349; 16 384: 0.778 | 0.744 0.628 0.610 0.622; 32 768: 1.519 | 1.177 1.129 1.120
// 1.176; 65 536: 2.979 | 2.169 2.113 2.155 2.281 - the optimum sits at 10-20 waves of items.
constexpr int PL_WAVES = 12;
constexpr int PL_MIN_WAVES2 = 7;
__host__ __device__ constexpr int pl_threads(int U) { return U == 1 ? 512 : (U == 2 ? 384 : 256); }

constexpr int PL_NPART = 4;    // segment tables for 1, 2, 4 and 8 segments per redshift bin
constexpr int PL_SEGW = 8;     // ints per segment record: z, i_lo, i_hi (fine points), j_lo, j_hi (data bins)

struct P1dwState {
  int* d_ja;       // [nf] local index of the first data bin a fine point feeds
  double* d_wa;    // [nf] weight into bin ja
  double* d_wb;    // [nf] weight into bin ja + 1
  int* d_seg[PL_NPART];  // [nseg][PL_SEGW]
  int nseg[PL_NPART];
  double* d_wsw;   // [nf][2] (+-wa, wb) for the warp-specialised kernel: the sign of wa says "this point opens the next bin"
  int ws_ok;       // weights non-negative and the bin index advances by at most one per fine point
};

template <int NC, int KIND, int METAL, bool UNI, int U>
__global__ void __launch_bounds__(pl_threads(U), 1)
k_fused_p1d_lanes(DevCtx c, const int* __restrict__ seg, int nseg, const int* __restrict__ rja_g,
                  const double* __restrict__ rwa_g, const double* __restrict__ rwb_g, int64_t W, const int* __restrict__ status,
                  const double* __restrict__ amp, double* __restrict__ dvec, double* __restrict__ p_out,
                  double* __restrict__ dvz, int tzw, int nfz_max, int ndz_max) {
  constexpr int PL_WARPS = pl_threads(U) / 32;
  extern __shared__ __align__(16) double sm[];
  double* tabw = sm;                                     // [nfz_max][C1D_NCOL] table row of a fine point (nfz_max: multiple of U)
  double* rwa = tabw + (size_t)nfz_max * C1D_NCOL;       // [nfz_max]
  double* rwb = rwa + nfz_max;                           // [nfz_max]
  double* dat = rwb + nfz_max;                           // [ndz_max]
  double* etab = dat + ndz_max;                          // [32] 2^(j/32) for exp10_tab
  int* rja = reinterpret_cast<int*>(etab + 32);          // [nfz_max]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 32) etab[threadIdx.x] = c_exp2_tab[threadIdx.x];  // visible after the first __syncthreads below
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  // cost-balanced contiguous range of this CTA in units of fine points, items = (segment, 32 walkers)
  const int64_t nitems = (W + 31) / 32;
  int64_t total = 0;
  for (int sg = 0; sg < nseg; ++sg) total += nitems * (int64_t)(seg[sg * PL_SEGW + 2] - seg[sg * PL_SEGW + 1]);
  const int64_t t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1) / gridDim.x;

  int64_t cum = 0;
  for (int sg = 0; sg < nseg; ++sg) {
    const int z = seg[sg * PL_SEGW], i_lo = seg[sg * PL_SEGW + 1], np = seg[sg * PL_SEGW + 2] - i_lo;
    const int j_lo = seg[sg * PL_SEGW + 3], j_hi = seg[sg * PL_SEGW + 4];
    const int f0 = c.zoff_f[z], nfz = c.zoff_f[z + 1] - f0;
    const int d0 = c.zoff_d[z], ndz = c.zoff_d[z + 1] - d0;
    const int64_t cost = np;
    const int64_t zend = cum + nitems * cost;
    const int64_t c_lo = (t0 > cum) ? (t0 - cum + cost - 1) / cost : 0;
    const int64_t c_hi = (t1 < zend) ? ((t1 > cum) ? (t1 - cum + cost - 1) / cost : 0) : nitems;
    cum = zend;
    if (c_lo >= c_hi) continue;

    __syncthreads();  // previous segment fully processed before its tables are overwritten
    // the segment is padded to a multiple of U with copies of its last point that carry no
    // rebinning weight, so the loop body needs no bounds tests
    const int nfp = (np + U - 1) / U * U;
    const int g0 = f0 + i_lo;  // first fine point of the segment in the packed tables
    for (int t = threadIdx.x; t < C1D_NCOL * nfp; t += blockDim.x) {
      const int col = t / nfp, i = t - col * nfp;
      tabw[i * C1D_NCOL + col] = c.tab[(size_t)col * c.nf + g0 + (i < np ? i : np - 1)];
    }
    for (int t = threadIdx.x; t < nfp; t += blockDim.x) {
      const bool in = t < np;
      rwa[t] = in ? rwa_g[g0 + t] : 0.0;
      rwb[t] = in ? rwb_g[g0 + t] : 0.0;
      rja[t] = rja_g[g0 + (in ? t : np - 1)];
    }
    for (int t = threadIdx.x; t < ndz; t += blockDim.x) dat[t] = c.data[d0 + t];
    __syncthreads();
    // spacing of the whole bin's grid (the same value in every segment)
    const double* kcol = c.tab + (size_t)C1D_COL_K * c.nf + f0;
    const double dk = UNI ? (kcol[nfz - 1] - kcol[0]) / (double)(nfz - 1) : 0.0;
    const int jc0 = (i_lo == 0) ? 0 : rja[0];  // bins before j_lo are accumulated but never stored

    for (int64_t item = c_lo + warp; item < c_hi; item += PL_WARPS) {
      const int64_t w = item * 32 + lane;
      const bool live = w < W;
      const int64_t wv = live ? w : W - 1;
      const bool ok = live && status[wv] == ST_OK;
      double* dv = dvec ? dvec + wv * c.ldd + d0 : nullptr;
      double* po = p_out ? p_out + wv * c.nd + d0 : nullptr;
      double* dzp = dvz ? dvz + wv * ((int64_t)c.nz * tzw) + (int64_t)z * tzw : nullptr;  // z-aligned copy for the per-z chi2
      if (!__any_sync(0xffffffffu, ok)) {
        // rejected walkers carry no model (the reference returns None)
        if (live)
          for (int j = j_lo; j < j_hi; ++j) {
            if (dv) dv[j] = 0.0;
            if (dzp) dzp[j] = 0.0;
            if (po) po[j] = qnan;
          }
        continue;
      }
      // amplitude record of this lane's (z, walker); rejected lanes compute on whatever stage 1 left
      // there and their outputs are overridden
      const double2* a2p = reinterpret_cast<const double2*>(amp + ((int64_t)z * W + wv) * C1D_NAMP);
      double av[C1D_NAMP];
#pragma unroll
      for (int j = 0; j < C1D_NAMP / 2; ++j) {
        const double2 v = __ldg(a2p + j);
        av[2 * j] = v.x;
        av[2 * j + 1] = v.y;
      }
      double cf[C1D_MAX_COEF];
#pragma unroll
      for (int j = 0; j < C1D_MAX_COEF; ++j) cf[j] = av[A_COEF0 + j];
      const double s3 = av[A_S3], s2 = av[A_S2];
      const double m0 = av[A_M0], c1 = av[A_C1];
      const double c2a = av[A_C2A], c2b = av[A_C2B];
      const double c3a = av[A_C3A], c3b = av[A_C3B];
      const double h0 = av[A_H0];
      const double hA = av[A_HCD1 + 0], hB = av[A_HCD1 + 1];
      const double hC = av[A_HCD1 + 2], hD = av[A_HCD1 + 3];
      const double aa = av[A_AA], sa2 = av[A_SA2];
      const double res = av[A_RES];
      const double fagn = av[A_AGN];

      // exp(-s k) and exp(-s^2 k^2) on an evenly spaced grid advance by constant ratios; the products
      // are re-anchored with an exact evaluation every PL_ANCHOR points
      const double sg3 = (METAL == 0) ? -s3 : s3;  // SiVid grows with k (si_vid_final.py:208)
      double E3 = 0.0, R3 = 0.0, E2 = 0.0, R2 = 0.0, gg = 0.0, grho = 0.0, gq = 0.0;
      if (UNI) {
        R3 = exp_poly(sg3 * dk);
        R2 = exp_poly(-s2 * dk);
        gq = exp_poly(-2.0 * sa2 * (dk * dk));
      }
      // GP emulator with a k-table norm (Fig3a.py:83-85): this lane's row of the table is the linear interpolation of
      // two table rows in its mean flux; along k the fine points walk the norm's own grid, two row values at a time
      int gj = 0, gcp = -2;
      double gw = 0.0, gN0 = 0.0, gN1 = 0.0;
      const bool gnorm = (KIND == 1) && c.gp_nm > 0;
      if (gnorm) {
        const double mF = c.gp_emu_in[((int64_t)z * W + wv) * C1D_MAX_EMU + c.gp_imf];
        while (gj < c.gp_nm - 2 && mF >= __ldg(c.gp_nmf + gj + 1)) ++gj;
        const double m0n = __ldg(c.gp_nmf + gj), m1n = __ldg(c.gp_nmf + gj + 1);
        gw = fmin(fmax((mF - m0n) / (m1n - m0n), 0.0), 1.0);
      }
      auto grow = [&](int cc) {
        const double a = __ldg(c.gp_ntab + (size_t)gj * c.gp_nkn + cc), b = __ldg(c.gp_ntab + (size_t)(gj + 1) * c.gp_nkn + cc);
        return fma(gw, b - a, a);
      };
      int jc = jc0;
      double acc0 = 0.0, acc1 = 0.0;
      auto flush = [&](int j, double m) {
        const double dj = dat[j] - m;
        if (live && j >= j_lo) {
          if (dv) dv[j] = ok ? dj : 0.0;
          if (dzp) dzp[j] = ok ? dj : 0.0;
          if (po) po[j] = ok ? m : qnan;
        }
      };
      const double2* tp0 = reinterpret_cast<const double2*>(tabw);
      for (int i0 = 0; i0 < nfp; i0 += U, tp0 += U * (C1D_NCOL / 2)) {
        if (UNI && (i0 & (PL_ANCHOR - 1)) == 0) {
          const double k = tp0[C1D_COL_K / 2].x;
          E3 = exp_poly(sg3 * k);
          E2 = exp_poly(-s2 * k);
          gg = exp_poly(-1.0 * sa2 * (k * k));
          grho = exp_poly(-1.0 * sa2 * (2.0 * k * dk + dk * dk));
        }
        double P[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double2* tp = tp0 + u * (C1D_NCOL / 2);
          const double2 ks = tp[C1D_COL_K / 2], tt1 = tp[C1D_COL_T / 2];  // (k, S), (t, T1)
          const double k = ks.x;
          double e3, e2, g;
          if (UNI) {
            e3 = E3;
            e2 = E2;
            g = gg;
            E3 *= R3;
            E2 *= R2;
            gg *= grho;
            grho *= gq;
          } else {
            e3 = exp(sg3 * k);
            e2 = exp(-s2 * k);
            g = exp(-1.0 * sa2 * (k * k));
          }
          // emulator polynomial (Horner) -> P_emu in km/s (lya_theory.py:690-701)
          double pl = 0.0;
          if (NC > 0) {
#pragma unroll
            for (int j = NC - 1; j >= 0; --j) pl = fma(pl, tt1.x, cf[j]);
          } else {
#pragma unroll
            for (int j = C1D_MAX_COEF - 1; j >= 0; --j)
              if (j < c.nc) pl = fma(pl, tt1.x, cf[j]);
          }
          double p_emu = ks.y * (KIND == 0 ? exp10_tab(pl, etab) : exp(pl));
          if (gnorm) {
            const double v = tp[C1D_COL_AGN / 2].y;  // C1D_COL_SPARE: interval + weight in the norm's k grid
            const int cc = (int)v;
            if (cc != gcp) {  // the same interval for every lane: a uniform branch
              gN0 = (cc == gcp + 1) ? gN1 : grow(cc);
              gN1 = grow(cc + 1 < c.gp_nkn ? cc + 1 : cc);
              gcp = cc;
            }
            p_emu *= fma(v - (double)cc, gN1 - gN0, gN0);
          }
          double mult;
          if (METAL == 0) {
            // si_mult.py:212-218, 267-341:  G = 2 - 2/(1 + E)
            const double2 v2 = tp[C1D_COL_T2A / 2], v3 = tp[C1D_COL_T3A / 2];  // (T2A, T2BC), (T3A, T3BC)
            const double G3 = fma(-2.0, rcp_fast1(1.0 + e3), 2.0);
            const double G2 = fma(-2.0, rcp_fast1(1.0 + e2), 2.0);
            mult = m0 + (c1 * G3 * tt1.y + G2 * (c2a * v2.x + c2b * v2.y)) + (c3a * v3.x + c3b * v3.y);
          } else {
            // si_vid_final.py:205-219
            mult = 1.0 + c1 * e3 + c2a * tt1.y * e2;
          }
          if (c.has_agn) mult *= fma(fagn, tp[C1D_COL_AGN / 2].x, 1.0);  // AGN_model.py:116-120: 1 + f_AGN delta
          // hcd_model_rogers_class.py:115-135 (or hcd_boss.py:5-7 through D1)
          const double2 d12 = tp[C1D_COL_D1 / 2], d34 = tp[C1D_COL_D3 / 2], kq = tp[C1D_COL_KT / 2];  // (D1, D2), (D3, D4), (KT, Q)
          const double hcd = h0 + hA * d12.x + hB * d12.y + hC * d34.x + hD * d34.y;
          // si_add.py:168-219
          const double add = aa * kq.x * g;
          // resolution_class.py:99-108
          const double syst = 1.0 + res * kq.y;
          // lya_theory.py:778-784 (IC correction is folded into the S column)
          P[u] = (hcd * mult * p_emu + add) * syst;
        }
        // rebinning (likelihood.py:147-161): point i feeds bins ja and ja + 1
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int ja = rja[i0 + u];
          while (jc < ja) {
            flush(jc, acc0);
            acc0 = acc1;
            acc1 = 0.0;
            ++jc;
          }
          acc0 = fma(rwa[i0 + u], P[u], acc0);
          acc1 = fma(rwb[i0 + u], P[u], acc1);
        }
      }
      while (jc < j_hi) {
        flush(jc, acc0);
        acc0 = acc1;
        acc1 = 0.0;
        ++jc;
      }
    }
  }
}


// ------------------------------------------------------------------------------------
// Warp-specialised walker-per-lane form (k_p1d_split).  The single-role kernel above is bound by
// instruction issue and by the latency of its dependent fp64 chain at 4 warps per sub-partition
// (128 registers per thread): ncu issue slots 60 %, fp64 pipe 56 %, dominant stall "wait".
// Here a (z, 32 walkers) item is evaluated by a PAIR of warps.  The PRODUCER warp forms, per fine point,
//     A = HCD(k) P_emu(k),   P_emu = S 10^poly(t)   (or e^poly for the GP emulators)
// from the NC polynomial coefficients and the five HCD amplitudes of its lane's walker; the CONSUMER
// warp forms the metal terms (exp recurrences, G = 2 - 2/(1+E)), the SiAdd term,
// P = (A mult + add) (1 + R Q), the rebinning and the residual.  Each role keeps only its own
// amplitudes in registers (`setmaxnreg` moves what the producers do not need to the consumers), so 24
// warps (6 per sub-partition) are resident instead of 16, each with a shorter dependent chain and ~35 %
// fewer instructions per point (constant-bank polynomial coefficients, no conversion instructions in
// 10^x, one flag bit instead of a bin-index table for the rebinning).  A producer hands its consumer
// chunks of WS_CH points x 32 lanes through a double buffer in shared memory and one named-barrier
// rendezvous per chunk.  Every value is formed by the same operations whatever the k-segment cut
// (anchors sit on multiples of PL_ANCHOR points of the redshift bin), so the results stay bit-identical
// for any number of segments.
// ------------------------------------------------------------------------------------
#ifndef WS_CH_
#define WS_CH_ 16
#define WS_UNROLL_ 4
#define WS_PREG_ 64
#define WS_CREG_ 96
#endif
// 12 pairs: `setmaxnreg` acts on warpgroups (four consecutive warps), so the producer warps [0, WS_PAIRS) must fill
// whole warpgroups - 13 or 14 pairs put both roles into one warpgroup and hang - and 16 pairs would need 18 of the 16
// named barriers (and leave 64 registers per thread at launch)
constexpr int WS_PAIRS = 12, WS_THREADS = 64 * WS_PAIRS;
static_assert(WS_PAIRS % 4 == 0 && WS_PAIRS + 2 <= 16, "producer warps in whole warpgroups; one named barrier per pair");
constexpr int WS_CH = WS_CH_, WS_NBUF = 2;  // points per chunk, chunks in the ring (double buffer)
constexpr int WS_UNROLL = WS_UNROLL_;
constexpr int WS_WAVES = 3;
constexpr int WS_PROW = 6, WS_CROW = 12;     // doubles per producer / consumer table row
constexpr int WS_PREG = WS_PREG_, WS_CREG = WS_CREG_;   // registers per thread after setmaxnreg (launch: 80)
// producer row: (t, S) (D1, D2) (D3, D4);  consumer row: (k, T1) (T2A, T2BC) (T3A, T3BC) (KT, Q) (+-wa, wb) (AGN, -)
__device__ __constant__ int c_ws_psrc[6] = {C1D_COL_T, C1D_COL_S, C1D_COL_D1, C1D_COL_D2, C1D_COL_D3, C1D_COL_D4};
__device__ __constant__ int c_ws_csrc[8] = {C1D_COL_K, C1D_COL_T1, C1D_COL_T2A, C1D_COL_T2BC, C1D_COL_T3A, C1D_COL_T3BC, C1D_COL_KT, C1D_COL_Q};
// (ln 2 / 32)^k / k!, k = 0 .. 6
__device__ __constant__ double c_ws_e2[7] = {1.0, 0.02166084939249829, 0.0002345961982022468, 1.693850972437182e-06,
                                             9.172562701824643e-09, 3.973709984549416e-11, 1.4345655584131934e-13};

__device__ __noinline__ double ws_exp_slow(double x, int kind) { return kind == 0 ? exp10(x) : exp(x); }
// the exact re-evaluations at the recurrence anchors (one point in PL_ANCHOR): out of line, the loop body stays small
__device__ __noinline__ double ws_exp_anchor(double x) { return exp_poly(x); }

// 10^x (KIND 0) or e^x (KIND 1) = 2^e 2^(j/32) 2^(r/32): n = rint(32 x log2 b) = 32 e + j from the low word of
// x L + 1.5 2^52 (no conversion instructions), degree-6 Horner polynomial with constant-bank coefficients,
// the power of two added into the exponent field of the table value.  |x| < 280 (700): the caller checks.
template <int KIND>
__device__ __forceinline__ double ws_exp_fast(double x, const double* __restrict__ tab) {
  const double L_HI = KIND == 0 ? 106.30169903639559 : 46.16624130844683;
  const double L_LO = KIND == 0 ? 5.3171760543154944e-15 : 6.513687597097931e-16;
  const double MAGIC = 6755399441055744.0;
  const double tq = fma(x, L_HI, MAGIC);
  const int ni = __double2loint(tq);
  const double n = tq - MAGIC;
  double r = fma(x, L_HI, -n);
  r = fma(x, L_LO, r);
  const double tv = tab[ni & 31];
  const double sc = __hiloint2double(__double2hiint(tv) + ((ni >> 5) << 20), __double2loint(tv));
  // Estrin: (1 + c1 r) + r^2 (c2 + c3 r) + r^4 ((c4 + c5 r) + c6 r^2)
  const double r2 = r * r;
  const double a0 = fma(c_ws_e2[1], r, 1.0), a1 = fma(c_ws_e2[3], r, c_ws_e2[2]), a2 = fma(c_ws_e2[5], r, c_ws_e2[4]);
  const double r4 = r2 * r2;
  const double b0 = fma(a1, r2, a0), b1 = fma(c_ws_e2[6], r2, a2);
  return fma(b1, r4, b0) * sc;
}

// rendezvous of a producer / consumer pair once per chunk on the pair's own named barrier (ids 2 .. 13): the
// producer arrives after filling chunk q, the consumer before reading it, so the consumer reads chunk q while the
// producer fills chunk q + 1 in the other slot; no polling, the waiting warp sleeps in the barrier unit
// emulator polynomial sum_j c_j t^j in Estrin form: dependent depth 3 instead of NC - 1 (the producer's chain is
// what bounds it: ptxas does not interleave the points of a chunk)
template <int NC>
__device__ __forceinline__ double ws_poly(const double* cf, double t) {
  const double t2 = t * t;
  const double a0 = fma(cf[1], t, cf[0]), a1 = fma(cf[3], t, cf[2]);
  const double a2 = NC >= 6 ? fma(cf[5], t, cf[4]) : cf[4];
  if (NC <= 6) {
    const double t4 = t2 * t2;
    return fma(a2, t4, fma(a1, t2, a0));
  } else {
    const double a3 = cf[6], t4 = t2 * t2;
    return fma(fma(a3, t2, a2), t4, fma(a1, t2, a0));
  }
}

// rendezvous of a producer / consumer pair once per chunk on the pair's own named barrier (ids 2 .. 13): the
// producer arrives after filling chunk q, the consumer before reading it, so the consumer reads chunk q while the
// producer fills chunk q + 1 in the other slot; no polling, the waiting warp sleeps in the barrier unit
// the same for two arguments at once, statement by statement: two independent chains for the scheduler
template <int KIND>
__device__ __forceinline__ void ws_exp_fast2(double xa, double xb, const double* __restrict__ tab, double& ya, double& yb) {
  const double L_HI = KIND == 0 ? 106.30169903639559 : 46.16624130844683;
  const double L_LO = KIND == 0 ? 5.3171760543154944e-15 : 6.513687597097931e-16;
  const double MAGIC = 6755399441055744.0;
  const double ta = fma(xa, L_HI, MAGIC), tb = fma(xb, L_HI, MAGIC);
  const int na = __double2loint(ta), nb = __double2loint(tb);
  const double tva = tab[na & 31], tvb = tab[nb & 31];
  const double fa = ta - MAGIC, fb = tb - MAGIC;
  double ra = fma(xa, L_HI, -fa), rb = fma(xb, L_HI, -fb);
  ra = fma(xa, L_LO, ra);
  rb = fma(xb, L_LO, rb);
  double pa = fma(c_ws_e2[6], ra, c_ws_e2[5]), pb = fma(c_ws_e2[6], rb, c_ws_e2[5]);
#pragma unroll
  for (int k = 4; k >= 1; --k) {
    pa = fma(pa, ra, c_ws_e2[k]);
    pb = fma(pb, rb, c_ws_e2[k]);
  }
  pa = fma(pa, ra, 1.0);
  pb = fma(pb, rb, 1.0);
  ya = pa * __hiloint2double(__double2hiint(tva) + ((na >> 5) << 20), __double2loint(tva));
  yb = pb * __hiloint2double(__double2hiint(tvb) + ((nb >> 5) << 20), __double2loint(tvb));
}

__device__ __forceinline__ void ws_bar_pair(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 2) : "memory"); }
__device__ __forceinline__ void ws_bar_all() { asm volatile("bar.sync 1, %0;" ::"n"(WS_THREADS) : "memory"); }

struct WsRange {  // the CTA's share of a segment
  int z, i_lo, np, j_lo, j_hi;
  int c_lo, c_hi;
};

template <int NC, int KIND, int METAL, bool AGN>
__global__ void __launch_bounds__(WS_THREADS, 1)
k_p1d_split(DevCtx c, const int* __restrict__ seg, int nseg, const int* __restrict__ rja_g, const double2* __restrict__ wsw_g, int64_t W,
            const int* __restrict__ status, const double* __restrict__ amp, double* __restrict__ dvec, double* __restrict__ p_out,
            double* __restrict__ dvz, int tzw, int nfp_max, int ndz_max) {
  using namespace c1d_tc;
  extern __shared__ __align__(16) double sm[];
  double* crow = sm;                                       // [nfp_max][WS_CROW] consumer table rows
  double* prow = crow + (size_t)nfp_max * WS_CROW;         // [nfp_max][WS_PROW] producer table rows
  double* ring = prow + (size_t)nfp_max * WS_PROW;         // [WS_PAIRS][WS_NBUF][WS_CH][32] A
  double* dat = ring + (size_t)WS_PAIRS * WS_NBUF * WS_CH * 32;  // [ndz_max]
  double* etab = dat + ndz_max;                            // [32]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 32) etab[threadIdx.x] = c_exp2_tab[threadIdx.x];
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  const int nitems = (int)((W + 31) / 32);
  int64_t total = 0;
  for (int sg = 0; sg < nseg; ++sg) total += nitems * (int64_t)(seg[sg * PL_SEGW + 2] - seg[sg * PL_SEGW + 1]);
  const int64_t t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1) / gridDim.x;

  // both roles walk the segments the same way: range of items, table staging by all threads, two CTA barriers
  auto seg_range = [&](int sg, int64_t& cum, WsRange& r) {
    r.z = seg[sg * PL_SEGW];
    r.i_lo = seg[sg * PL_SEGW + 1];
    r.np = seg[sg * PL_SEGW + 2] - r.i_lo;
    r.j_lo = seg[sg * PL_SEGW + 3];
    r.j_hi = seg[sg * PL_SEGW + 4];
    const int64_t cost = r.np, zend = cum + nitems * cost;
    // boundaries on multiples of WS_PAIRS items (the same formula on both sides of a boundary): every pair of a CTA
    // gets the same number of items of a segment, the CTAs differ by at most one round
    auto cut = [&](int64_t t) {
      const int64_t it = ((t - cum + cost - 1) / cost + WS_PAIRS - 1) / WS_PAIRS * WS_PAIRS;
      return (int)(it < nitems ? it : nitems);
    };
    r.c_lo = (t0 > cum) ? cut(t0) : 0;
    r.c_hi = (t1 < zend) ? ((t1 > cum) ? cut(t1) : 0) : nitems;
    cum = zend;
    return r.c_lo < r.c_hi;
  };
  auto stage = [&](const WsRange& r, int nfp) {
    ws_bar_all();  // previous segment fully processed before its tables are overwritten
    const int f0 = c.zoff_f[r.z], d0 = c.zoff_d[r.z], ndz = c.zoff_d[r.z + 1] - d0;
    const int g0 = f0 + r.i_lo;
    for (int t = threadIdx.x; t < 14 * nfp; t += WS_THREADS) {
      const int col = t / nfp, i = t - col * nfp;
      const int is = g0 + (i < r.np ? i : r.np - 1);  // the padding repeats the last point (it carries no weight)
      if (col < 8)
        crow[i * WS_CROW + col] = c.tab[(size_t)c_ws_csrc[col] * c.nf + is];
      else
        prow[i * WS_PROW + col - 8] = c.tab[(size_t)c_ws_psrc[col - 8] * c.nf + is];
    }
    for (int t = threadIdx.x; t < nfp; t += WS_THREADS) {
      const int is = g0 + (t < r.np ? t : r.np - 1);
      const double2 w2 = t < r.np ? wsw_g[is] : make_double2(0.0, 0.0);  // padding points carry no weight and open no bin
      crow[t * WS_CROW + 8] = w2.x;
      crow[t * WS_CROW + 9] = w2.y;
      crow[t * WS_CROW + 10] = AGN ? c.tab[(size_t)C1D_COL_AGN * c.nf + is] : 0.0;
      crow[t * WS_CROW + 11] = 0.0;
    }
    for (int t = threadIdx.x; t < ndz; t += WS_THREADS) dat[t] = c.data[d0 + t];
    ws_bar_all();
  };

  if (warp < WS_PAIRS) {
    // ---------------- producers: A = HCD P_emu (1 + R Q) ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_PREG));
    double* const myring = ring + (size_t)warp * WS_NBUF * WS_CH * 32 + lane;
    uint32_t q = 0;  // chunks handed over so far
    int64_t cum = 0;
    for (int sg = 0; sg < nseg; ++sg) {
      WsRange r;
      if (!seg_range(sg, cum, r)) continue;
      const int nfp = (r.np + WS_CH - 1) / WS_CH * WS_CH;
      stage(r, nfp);
      for (int item = r.c_lo + warp; item < r.c_hi; item += WS_PAIRS) {
        const int64_t w = (int64_t)item * 32 + lane;
        const int64_t wv = w < W ? w : W - 1;
        if (!__any_sync(0xffffffffu, w < W && status[wv] == ST_OK)) continue;  // the consumer skips the item too
        const double2* a2p = reinterpret_cast<const double2*>(amp + ((int64_t)r.z * W + wv) * C1D_NAMP);
        double cf[NC + 1];
#pragma unroll
        for (int j = 0; j < (NC + 1) / 2; ++j) {
          const double2 v = __ldg(a2p + A_COEF0 / 2 + j);
          cf[2 * j] = v.x;
          cf[2 * j + 1] = v.y;
        }
        const double2 h01 = __ldg(a2p + A_H0 / 2), h23 = __ldg(a2p + A_H0 / 2 + 1), h4x = __ldg(a2p + A_H0 / 2 + 2);  // (h0, hA) (hB, hC) (hD, aa)
        const double2* tp = reinterpret_cast<const double2*>(prow);
        for (int i0 = 0; i0 < nfp; i0 += WS_CH, tp += WS_CH * (WS_PROW / 2)) {
          double* dst = myring + (q & (WS_NBUF - 1)) * (WS_CH * 32);
          bool bad = false;  // arguments outside the fast range (or NaN): redone below with the library function
#pragma unroll WS_UNROLL
          for (int u = 0; u < WS_CH; ++u) {
            const double2 ts = tp[u * (WS_PROW / 2)], d12 = tp[u * (WS_PROW / 2) + 1], d34 = tp[u * (WS_PROW / 2) + 2];
            const double pl = ws_poly<NC>(cf, ts.x);
            bad |= !(fabs(pl) < (KIND == 0 ? 280.0 : 700.0));
            // what does not depend on the exponential first: B = S HCD
            const double hcd = (h01.x + h01.y * d12.x) + (h23.x * d12.y + (h23.y * d34.x + h4x.x * d34.y));  // hcd_model_rogers_class.py:115-135
            dst[u * 32] = ws_exp_fast<KIND>(pl, etab) * (ts.y * hcd);                                       // lya_theory.py:690-701
          }
          if (__any_sync(0xffffffffu, bad)) {
#pragma unroll 1
            for (int u = 0; u < WS_CH; ++u) {
              const double2 ts = tp[u * (WS_PROW / 2)], d12 = tp[u * (WS_PROW / 2) + 1], d34 = tp[u * (WS_PROW / 2) + 2];
              const double pl = ws_poly<NC>(cf, ts.x);
              if (!(fabs(pl) < (KIND == 0 ? 280.0 : 700.0))) {
                const double hcd = (h01.x + h01.y * d12.x) + (h23.x * d12.y + (h23.y * d34.x + h4x.x * d34.y));
                dst[u * 32] = ws_exp_slow(pl, KIND) * (ts.y * hcd);
              }
            }
          }
          ws_bar_pair(warp);  // chunk q is complete; the consumer has finished chunk q - 1, whose slot is filled next
          ++q;
        }
      }
    }
  } else {
    // ---------------- consumers: metals, SiAdd, rebinning, residual ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_CREG));
    const int cw = warp - WS_PAIRS;
    const double* const myring = ring + (size_t)cw * WS_NBUF * WS_CH * 32 + lane;
    uint32_t q = 0;
    int64_t cum = 0;
    for (int sg = 0; sg < nseg; ++sg) {
      WsRange r;
      if (!seg_range(sg, cum, r)) continue;
      const int nfp = (r.np + WS_CH - 1) / WS_CH * WS_CH;
      stage(r, nfp);
      const int z = r.z, j_lo = r.j_lo, j_hi = r.j_hi;
      const int f0 = c.zoff_f[z], nfz = c.zoff_f[z + 1] - f0, d0 = c.zoff_d[z];
      const double* kcol = c.tab + (size_t)C1D_COL_K * c.nf + f0;
      const double dk = (kcol[nfz - 1] - kcol[0]) / (double)(nfz - 1);
      // bin the first point of the segment feeds; a set flag on that point belongs to the previous segment
      const int jc0 = (r.i_lo == 0) ? 0 : rja_g[f0 + r.i_lo] - (__double2hiint(crow[8]) < 0 ? 1 : 0);

      for (int item = r.c_lo + cw; item < r.c_hi; item += WS_PAIRS) {
        const int64_t w = (int64_t)item * 32 + lane;
        const bool live = w < W;
        const int64_t wv = live ? w : W - 1;
        const bool ok = live && status[wv] == ST_OK;
        if (!__any_sync(0xffffffffu, ok)) {
          // rejected walkers carry no model (the reference returns None)
          if (live)
            for (int j = j_lo; j < j_hi; ++j) {
              if (dvec) dvec[wv * c.ldd + d0 + j] = 0.0;
              if (dvz) dvz[wv * ((int64_t)c.nz * tzw) + (int64_t)z * tzw + j] = 0.0;
              if (p_out) p_out[wv * c.nd + d0 + j] = qnan;
            }
          continue;
        }
        const double2* a2p = reinterpret_cast<const double2*>(amp + ((int64_t)z * W + wv) * C1D_NAMP);
        const double2 s32 = __ldg(a2p + A_S3 / 2), mc1 = __ldg(a2p + A_M0 / 2), c2ab = __ldg(a2p + A_C2A / 2), c3ab = __ldg(a2p + A_C3A / 2);
        const double2 hdaa = __ldg(a2p + A_AA / 2), sar = __ldg(a2p + A_SA2 / 2), agn2 = __ldg(a2p + A_AGN / 2);  // (hD, aa) (sa2, res) (fagn, -)
        const double s3 = s32.x, s2 = s32.y, m0 = mc1.x, c1 = mc1.y, c2a = c2ab.x, c2b = c2ab.y, c3a = c3ab.x, c3b = c3ab.y;
        const double aa = hdaa.y, sa2 = sar.x, res = sar.y, fagn = agn2.x;
        const double sg3 = (METAL == 0) ? -s3 : s3;  // SiVid grows with k (si_vid_final.py:208)
        // exp(-s k) and exp(-s^2 k^2) on an evenly spaced grid advance by constant ratios; the products
        // are re-anchored with an exact evaluation every PL_ANCHOR points
        const double R3 = ws_exp_anchor(sg3 * dk), R2 = ws_exp_anchor(-s2 * dk), gq = ws_exp_anchor(-2.0 * sa2 * (dk * dk));
        double E3 = 0.0, E2 = 0.0, gg = 0.0, grho = 0.0;
        int jc = jc0;
        double acc0 = 0.0, acc1 = 0.0;
        auto flush = [&](int j, double m) {
          const double dj = dat[j] - m;
          if (live && j >= j_lo) {
            if (dvec) dvec[wv * c.ldd + d0 + j] = ok ? dj : 0.0;
            if (dvz) dvz[wv * ((int64_t)c.nz * tzw) + (int64_t)z * tzw + j] = ok ? dj : 0.0;
            if (p_out) p_out[wv * c.nd + d0 + j] = ok ? m : qnan;
          }
        };
        const double2* rp = reinterpret_cast<const double2*>(crow);
        for (int i0 = 0; i0 < nfp; i0 += WS_CH, rp += WS_CH * (WS_CROW / 2)) {
          if ((i0 & (PL_ANCHOR - 1)) == 0) {
            const double k = rp[0].x;
            E3 = ws_exp_anchor(sg3 * k);
            E2 = ws_exp_anchor(-s2 * k);
            gg = aa * ws_exp_anchor(-1.0 * sa2 * (k * k));  // the SiAdd amplitude rides on the recurrence
            grho = ws_exp_anchor(-1.0 * sa2 * (2.0 * k * dk + dk * dk));
          }
          ws_bar_pair(cw);  // chunk q has been produced
          const double* src = myring + (q & (WS_NBUF - 1)) * (WS_CH * 32);
#pragma unroll WS_UNROLL
          for (int u = 0; u < WS_CH; ++u) {
            const double2* tp = rp + u * (WS_CROW / 2);
            const double A = src[u * 32];
            const double e3 = E3, e2 = E2, g = gg;
            E3 *= R3;
            E2 *= R2;
            gg *= grho;
            grho *= gq;
            const double T1 = tp[0].y;
            double mult;
            if (METAL == 0) {
              // si_mult.py:212-218, 267-341:  G = 2 - 2/(1 + E)
              const double2 v2 = tp[1], v3 = tp[2];
              const double G3 = fma(-2.0, rcp_fast1(1.0 + e3), 2.0);
              const double G2 = fma(-2.0, rcp_fast1(1.0 + e2), 2.0);
              mult = m0 + (c1 * G3 * T1 + G2 * (c2a * v2.x + c2b * v2.y)) + (c3a * v3.x + c3b * v3.y);
            } else {
              mult = 1.0 + c1 * e3 + c2a * T1 * e2;  // si_vid_final.py:205-219
            }
            if (AGN) mult *= fma(fagn, tp[5].x, 1.0);  // AGN_model.py:116-120
            const double2 kq = tp[3], rw = tp[4];
            // si_add.py:168-219 (g = aa exp(-sa2 k^2)), resolution_class.py:99-108, lya_theory.py:778-784
            const double P = fma(A, mult, kq.x * g) * fma(res, kq.y, 1.0);
            // rebinning (likelihood.py:147-161): a negative sign on wa opens the next bin
            if (__double2hiint(rw.x) < 0) {
              flush(jc, acc0);
              acc0 = acc1;
              acc1 = 0.0;
              ++jc;
            }
            acc0 = fma(fabs(rw.x), P, acc0);
            acc1 = fma(rw.y, P, acc1);
          }
          ++q;
        }
        while (jc < j_hi) {
          flush(jc, acc0);
          acc0 = acc1;
          acc1 = 0.0;
          ++jc;
        }
      }
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
void c1d_p1dw_destroy(c1d_ctx* ctx) {
  P1dwState* st = (P1dwState*)ctx->p1dw_state;
  if (!st) return;
  cudaFree(st->d_ja);
  cudaFree(st->d_wa);
  cudaFree(st->d_wb);
  cudaFree(st->d_wsw);
  for (int q = 0; q < PL_NPART; ++q) cudaFree(st->d_seg[q]);
  delete st;
  ctx->p1dw_state = nullptr;
}

// Per fine point: the (at most two, adjacent) data bins it feeds.  Leaves p1dw_state null when the
// rebinning operator is not of that form (the caller then uses the warp-per-walker kernel).
int c1d_p1dw_prepare(c1d_ctx* ctx) {
  const c1d_config& c = ctx->cfg;
  c1d_p1dw_destroy(ctx);
  const int RW = c.rebin_width;
  std::vector<int> zf(c.nz + 1), zd(c.nz + 1), st(c.nd);
  std::vector<double> wg((size_t)c.nd * RW);
  C1D_CUDA(ctx, cudaMemcpy(zf.data(), ctx->dev[C1D_F_ZOFF_F], (c.nz + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  C1D_CUDA(ctx, cudaMemcpy(zd.data(), ctx->dev[C1D_F_ZOFF_D], (c.nz + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  C1D_CUDA(ctx, cudaMemcpy(st.data(), ctx->dev[C1D_F_RB_START], c.nd * sizeof(int), cudaMemcpyDeviceToHost));
  C1D_CUDA(ctx, cudaMemcpy(wg.data(), ctx->dev[C1D_F_RB_WGT], wg.size() * sizeof(double), cudaMemcpyDeviceToHost));
  std::vector<int> ja(c.nf, 0);
  std::vector<double> wa(c.nf, 0.0), wb(c.nf, 0.0);
  std::vector<int> segs[PL_NPART];
  for (int z = 0; z < c.nz; ++z) {
    const int f0 = zf[z], nfz = zf[z + 1] - f0, d0 = zd[z], ndz = zd[z + 1] - d0;
    std::vector<int> first(nfz, -1), count(nfz, 0);
    for (int j = 0; j < ndz; ++j)
      for (int t = 0; t < RW; ++t) {
        const double wt = wg[(size_t)(d0 + j) * RW + t];
        const int i = st[d0 + j] + t;
        if (wt == 0.0 || i >= nfz) continue;
        if (first[i] < 0) {
          first[i] = j;
          wa[f0 + i] = wt;
        } else if (j == first[i] + 1 && count[i] == 1) {
          wb[f0 + i] = wt;
        } else {
          return C1D_OK;  // a fine point feeds more than two (or non-adjacent) bins
        }
        count[i]++;
      }
    int last = 0;
    for (int i = 0; i < nfz; ++i) {
      if (first[i] < 0) first[i] = last;  // feeds no bin: stay on the current one with zero weights
      if (first[i] < last) return C1D_OK;  // bins must come in grid order
      last = first[i];
      ja[f0 + i] = first[i];
    }
    // segments of this z for 1, 2, 4, 8 cuts: first and last fine point that feed each data bin
    std::vector<int> bfirst(ndz, nfz), blast(ndz, -1);
    for (int i = 0; i < nfz; ++i) {
      const int j = ja[f0 + i];
      if (wa[f0 + i] != 0.0 && j < ndz) {
        if (i < bfirst[j]) bfirst[j] = i;
        blast[j] = i;
      }
      if (wb[f0 + i] != 0.0 && j + 1 < ndz) {
        if (i < bfirst[j + 1]) bfirst[j + 1] = i;
        blast[j + 1] = i;
      }
    }
    // a bin no fine point feeds goes with the next bin that is fed
    for (int j = ndz - 2; j >= 0; --j)
      if (blast[j] < 0) bfirst[j] = bfirst[j + 1];
    const int nblk = (nfz + PL_ANCHOR - 1) / PL_ANCHOR;
    for (int q = 0; q < PL_NPART; ++q) {
      const int want = 1 << q;
      const int P = want < nblk ? want : nblk;
      std::vector<int> jl(P + 1, ndz), il(P + 1, nfz);
      for (int p = 0; p < P; ++p) {
        il[p] = PL_ANCHOR * (int)(((long long)p * nblk) / P);
        int j = 0;
        while (j < ndz && bfirst[j] < il[p]) ++j;
        jl[p] = (p == 0) ? 0 : j;
      }
      for (int p = 0; p < P; ++p) {
        if (jl[p + 1] <= jl[p]) continue;  // owns no bin: its points are covered by the neighbours
        int ih = il[p] + 1;
        for (int j = jl[p]; j < jl[p + 1]; ++j)
          if (blast[j] + 1 > ih) ih = blast[j] + 1;
        if (p == P - 1 || jl[p + 1] >= ndz) ih = nfz;
        const int rec[PL_SEGW] = {z, il[p], ih, jl[p], jl[p + 1], 0, 0, 0};
        segs[q].insert(segs[q].end(), rec, rec + PL_SEGW);
      }
    }
  }
  P1dwState* s = new P1dwState();
  memset(s, 0, sizeof(*s));
  ctx->p1dw_state = s;
  C1D_CUDA(ctx, cudaMalloc(&s->d_ja, c.nf * sizeof(int)));
  C1D_CUDA(ctx, cudaMalloc(&s->d_wa, c.nf * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&s->d_wb, c.nf * sizeof(double)));
  C1D_CUDA(ctx, cudaMemcpy(s->d_ja, ja.data(), c.nf * sizeof(int), cudaMemcpyHostToDevice));
  C1D_CUDA(ctx, cudaMemcpy(s->d_wa, wa.data(), c.nf * sizeof(double), cudaMemcpyHostToDevice));
  C1D_CUDA(ctx, cudaMemcpy(s->d_wb, wb.data(), c.nf * sizeof(double), cudaMemcpyHostToDevice));
  // warp-specialised kernel: (+-wa, wb) pairs, the sign of wa set on the points that open the next bin
  {
    std::vector<double> wsw((size_t)c.nf * 2);
    int ok = 1;
    for (int z = 0; z < c.nz && ok; ++z) {
      const int f0 = zf[z], nfz = zf[z + 1] - f0;
      for (int i = 0; i < nfz; ++i) {
        const int step = ja[f0 + i] - (i ? ja[f0 + i - 1] : 0);
        if (step < 0 || step > 1 || wa[f0 + i] < 0.0 || wb[f0 + i] < 0.0) {
          ok = 0;
          break;
        }
        wsw[2 * (size_t)(f0 + i)] = step ? -wa[f0 + i] : wa[f0 + i];  // -0.0 keeps the flag on a zero weight
        wsw[2 * (size_t)(f0 + i) + 1] = wb[f0 + i];
      }
    }
    s->ws_ok = ok;
    C1D_CUDA(ctx, cudaMalloc(&s->d_wsw, wsw.size() * sizeof(double)));
    C1D_CUDA(ctx, cudaMemcpy(s->d_wsw, wsw.data(), wsw.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  for (int q = 0; q < PL_NPART; ++q) {
    s->nseg[q] = (int)(segs[q].size() / PL_SEGW);
    C1D_CUDA(ctx, cudaMalloc(&s->d_seg[q], segs[q].size() * sizeof(int)));
    C1D_CUDA(ctx, cudaMemcpy(s->d_seg[q], segs[q].data(), segs[q].size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  return C1D_OK;
}

// returns 1 when the launch was done here, 0 when the caller must use the warp-per-walker kernel,
// a negative c1d_status on error
int c1d_p1dw_launch(c1d_ctx* ctx, int64_t W, double* p_out, int want_dvec, int want_chi2z, cudaStream_t s) {
  P1dwState* st = (P1dwState*)ctx->p1dw_state;
  // C1D_P1D_LAYOUT=warp | lanes overrides the choice by batch size (tests run both forms on the same inputs)
  const char* layout = getenv("C1D_P1D_LAYOUT");
  // ("lanes" = the single-role walker-per-lane kernel, "split" = the warp-specialised one, both whatever the batch size)
  const bool force_warp = layout && !strcmp(layout, "warp"), force_split = layout && !strcmp(layout, "split");
  const bool force_lanes = (layout && !strcmp(layout, "lanes")) || force_split;
  const bool no_split = layout && !strcmp(layout, "lanes");
  if (!st || force_warp) return 0;
  const c1d_config& cfg = ctx->cfg;
  const int nfz_max = ctx->nfz_max, ndz_max = ctx->ndz_max;
  constexpr int U = PL_U;
  const int nfp_max = (nfz_max + 3) & ~3;
  const size_t smem = ((size_t)nfp_max * (C1D_NCOL + 2) + ndz_max + 32) * sizeof(double) + (size_t)nfp_max * sizeof(int) + 16;
  if (smem > 227 * 1024) return 0;
  typedef void (*kern_t)(DevCtx, const int*, int, const int*, const double*, const double*, int64_t, const int*, const double*, double*,
                         double*, double*, int, int, int);
  kern_t kern = nullptr;
  const int nc = (cfg.nc >= 5 && cfg.nc <= 7) ? cfg.nc : 0;
  const int sel = ((nc ? nc - 4 : 0) << 3) | (cfg.emu_kind << 2) | (cfg.metal_model << 1) | (cfg.uniform_k ? 1 : 0);
#define PL_PICK(...) k_fused_p1d_lanes<__VA_ARGS__, PL_U>
#define PL_CASE(NCI, NCV)                                              \
  case ((NCI) << 3) | 0: kern = PL_PICK(NCV, 0, 0, false); break;      \
  case ((NCI) << 3) | 1: kern = PL_PICK(NCV, 0, 0, true); break;       \
  case ((NCI) << 3) | 2: kern = PL_PICK(NCV, 0, 1, false); break;      \
  case ((NCI) << 3) | 3: kern = PL_PICK(NCV, 0, 1, true); break;       \
  case ((NCI) << 3) | 4: kern = PL_PICK(NCV, 1, 0, false); break;      \
  case ((NCI) << 3) | 5: kern = PL_PICK(NCV, 1, 0, true); break;       \
  case ((NCI) << 3) | 6: kern = PL_PICK(NCV, 1, 1, false); break;      \
  case ((NCI) << 3) | 7: kern = PL_PICK(NCV, 1, 1, true); break;
  switch (sel) {
    PL_CASE(0, 0)
    PL_CASE(1, 5)
    PL_CASE(2, 6)
    PL_CASE(3, 7)
  }
#undef PL_CASE
#undef PL_PICK
  if (!kern) return 0;
  // the warp-specialised kernel where it applies: evenly spaced k, a compile-time polynomial degree, no GP norm table,
  // a rebinning operator whose bin index moves by at most one per fine point
  typedef void (*kern_ws_t)(DevCtx, const int*, int, const int*, const double2*, int64_t, const int*, const double*, double*, double*, double*, int,
                            int, int);
  kern_ws_t kws = nullptr;
  const int nfp_ws = (nfz_max + WS_CH - 1) / WS_CH * WS_CH;
  const size_t smem_ws = ((size_t)nfp_ws * (WS_CROW + WS_PROW) + (size_t)WS_PAIRS * WS_NBUF * WS_CH * 32 + ndz_max + 32) * sizeof(double) + 16;
  if (st->ws_ok && !no_split && cfg.uniform_k && nc && !(cfg.emu_kind == 1 && cfg.gp_nm > 0) && smem_ws <= 227 * 1024) {
    const int wsel = ((nc - 5) << 3) | (cfg.emu_kind << 2) | (cfg.metal_model << 1) | (cfg.has_agn ? 1 : 0);
#define WS_CASE(NCV)                                                         \
  case (((NCV) - 5) << 3) | 0: kws = k_p1d_split<NCV, 0, 0, false>; break;   \
  case (((NCV) - 5) << 3) | 1: kws = k_p1d_split<NCV, 0, 0, true>; break;    \
  case (((NCV) - 5) << 3) | 2: kws = k_p1d_split<NCV, 0, 1, false>; break;   \
  case (((NCV) - 5) << 3) | 3: kws = k_p1d_split<NCV, 0, 1, true>; break;    \
  case (((NCV) - 5) << 3) | 4: kws = k_p1d_split<NCV, 1, 0, false>; break;   \
  case (((NCV) - 5) << 3) | 5: kws = k_p1d_split<NCV, 1, 0, true>; break;    \
  case (((NCV) - 5) << 3) | 6: kws = k_p1d_split<NCV, 1, 1, false>; break;   \
  case (((NCV) - 5) << 3) | 7: kws = k_p1d_split<NCV, 1, 1, true>; break;
    switch (wsel) {
      WS_CASE(5)
      WS_CASE(6)
      WS_CASE(7)
    }
#undef WS_CASE
  }
  if (kws)
    C1D_CUDA(ctx, cudaFuncSetAttribute(kws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
  else
    C1D_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = kws ? WS_THREADS : pl_threads(U), warps = kws ? WS_PAIRS : pl_threads(U) / 32;
  const int64_t zitems = ((W + 31) / 32) * cfg.nz;
  // a warp-item is 32 walkers long; batches below PL_WAVES waves of (z, 32 walkers) items cut the redshift
  // bins into 2, 4 or 8 segments (bit-identical results, finer items), and below PL_MIN_WAVES2 / 2 waves of
  // the finest items the warp-per-walker kernel takes over (C1D_P1D_PARTS=1|2|4|8 forces the cut)
  const int64_t wave = (int64_t)ctx->num_sms * warps;
  int q = 0;
  // (the pair kernel pays more per segment - staging behind two CTA barriers, per-item set-up in both roles: measured
  // optimum at ~3 waves of pair items against 10-20 for the single-role kernel)
  while (q + 1 < PL_NPART && (zitems << q) < (kws ? WS_WAVES : PL_WAVES) * wave) ++q;
  if (const char* e = getenv("C1D_P1D_PARTS")) {
    const int v = atoi(e);
    q = v >= 8 ? 3 : (v >= 4 ? 2 : (v >= 2 ? 1 : 0));
  }
  const int64_t items = ((W + 31) / 32) * st->nseg[q];
  // (pair kernel: only when even the finest cut leaves fewer than 3.5 waves; measured C4: 2048 walkers 0.125 ms warp-per-walker
  // against 0.139, 4096 walkers 0.223 against 0.19)
  if (2 * items < PL_MIN_WAVES2 * wave && (!kws || q == PL_NPART - 1) && !force_lanes) return 0;
  int64_t grid = (int64_t)ctx->num_sms;
  if (grid > (items + warps - 1) / warps) grid = (items + warps - 1) / warps;
  if (grid < 1) grid = 1;
  double* dvec = want_dvec ? ctx->ws.dvec : nullptr;
  double* dvz = nullptr;
  if (want_chi2z) {
    if (!ctx->ws.dvz) {
      // allocated on first use: [cap][nz * tzw], pad columns stay zero
      const size_t n = (size_t)ctx->ws.cap * cfg.nz * ctx->tzw;
      C1D_CUDA(ctx, cudaMalloc(&ctx->ws.dvz, n * sizeof(double)));
      C1D_CUDA(ctx, cudaMemsetAsync(ctx->ws.dvz, 0, n * sizeof(double), s));
    }
    dvz = ctx->ws.dvz;
  }
  ctx->d.gp_emu_in = ctx->ws.emu_in;
  if (kws)
    kws<<<(unsigned)grid, threads, smem_ws, s>>>(ctx->d, st->d_seg[q], st->nseg[q], st->d_ja, reinterpret_cast<const double2*>(st->d_wsw), W,
                                                    ctx->ws.status, ctx->ws.amp, dvec, p_out, dvz, ctx->tzw, nfp_ws, ndz_max);
  else
    kern<<<(unsigned)grid, threads, smem, s>>>(ctx->d, st->d_seg[q], st->nseg[q], st->d_ja, st->d_wa, st->d_wb, W, ctx->ws.status, ctx->ws.amp, dvec, p_out,
                                                  dvz, ctx->tzw, nfp_max, ndz_max);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return 1;
}
