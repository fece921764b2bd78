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
  const bool force_warp = layout && !strcmp(layout, "warp"), force_lanes = layout && !strcmp(layout, "lanes");
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
  C1D_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = pl_threads(U), warps = threads / 32;
  const int64_t zitems = ((W + 31) / 32) * cfg.nz;
  // a warp-item is 32 walkers long; batches below PL_WAVES waves of (z, 32 walkers) items cut the redshift
  // bins into 2, 4 or 8 segments (bit-identical results, finer items), and below PL_MIN_WAVES2 / 2 waves of
  // the finest items the warp-per-walker kernel takes over (C1D_P1D_PARTS=1|2|4|8 forces the cut)
  const int64_t wave = (int64_t)ctx->num_sms * warps;
  int q = 0;
  while (q + 1 < PL_NPART && (zitems << q) < PL_WAVES * wave) ++q;
  if (const char* e = getenv("C1D_P1D_PARTS")) {
    const int v = atoi(e);
    q = v >= 8 ? 3 : (v >= 4 ? 2 : (v >= 2 ? 1 : 0));
  }
  const int64_t items = ((W + 31) / 32) * st->nseg[q];
  if (2 * items < PL_MIN_WAVES2 * wave && !force_lanes) return 0;
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
  kern<<<(unsigned)grid, threads, smem, s>>>(ctx->d, st->d_seg[q], st->nseg[q], st->d_ja, st->d_wa, st->d_wb, W, ctx->ws.status, ctx->ws.amp, dvec, p_out,
                                                dvz, ctx->tzw, nfp_max, ndz_max);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return 1;
}
