// Hand-written sm_100a kernels of the batched P1D log-likelihood.
//
//   stage 0/1/5  k_stage1      cube -> parameters -> per-z emulator inputs, contaminant
//                              amplitudes, blobs, Gaussian prior, star box, hull; thread per (walker, z)
//   stage 2      k_mlp_f64     LaCE-style MLP on the fp64 tensor cores (DMMA, exact-parity mode; c1d_mlp_f64.cu)
//                k_mlp_tc      the same MLP on tcgen05 in 3xTF32 (c1d_mlp_tc.cu)
//                k_gp_f64      GP posterior mean = batched kernel-vector product
//   stage 3      k_fused_p1d   emulator polynomial -> contaminant templates -> rebin -> residual
//                              (-> per-z chi2), warp per walker, tables staged in shared memory per z:
//                              small batches and rebinning operators outside the two-bin form
//                k_fused_p1d_lanes  the same with lane = walker (c1d_p1d_lanes.cu): large batches
//   stage 4      k_chi2_dmma   chi2 = d^T C^-1 d: tiled fp64 tensor-core GEMM (DMMA), symmetric,
//                              row-dot fused into the epilogue
//                k_chi2z_dmma  per-z chi2 of z-aligned residuals (behind k_fused_p1d_lanes)
//   stage 5      k_finalize    regulate + prior, blobs conventions
//
// Reference lines are cited next to each formula.
#include <cstdlib>
#include <cstring>

#include "c1d_internal.h"
#include "c1d_devmath.cuh"

// ---------------------------------------------------------------------------------------
// stage 1
// ---------------------------------------------------------------------------------------
// the walker's cube coordinates: row w of x[W][ndim], or column w of the transposed copy xt[ndim][W] (large batches:
// the lanes of a warp are consecutive walkers, so a coordinate of 32 walkers is one 256-byte segment of xt against 32
// lines of x - the row-major gather kept the kernel at the L1 tag rate, ~100 loads x 32 lines per warp)
template <bool XT>
struct XRow {
  const double* p;
  int64_t stride;
  __device__ __forceinline__ double operator[](int i) const { return XT ? p[(int64_t)i * stride] : p[i]; }
};
template <bool XT>
__device__ __forceinline__ double param_value(const DevCtx& c, const XRow<XT>& xw, int i) {
  // likelihood_parameter.py:58-61
  return c.pmin[i] + xw[i] * c.pscale[i];
}

// xt[i][w] = x[w][i] through 32 x 33 shared-memory tiles
__global__ void __launch_bounds__(256)
k_transpose_x(const double* __restrict__ x, int64_t W, int ndim, double* __restrict__ xt) {
  __shared__ double tile[32][33];
  const int64_t w0 = (int64_t)blockIdx.x * 32;
  const int i0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t w = w0 + r;
    const int i = i0 + tx;
    tile[r][tx] = (w < W && i < ndim) ? x[w * ndim + i] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r;
    const int64_t w = w0 + tx;
    if (i < ndim && w < W) xt[(int64_t)i * W + w] = tile[tx][r];
  }
}

// One thread per (walker, z): blockIdx.y is the redshift bin.  The z = 0 thread of a walker also does
// the walker-level work (cube test, Gaussian prior, blobs, star-prior box).  The status word is zeroed
// before the launch and combined with atomicMax (out of cube > rejected > ok).
#define S1_THREADS 128
template <bool XT, int EW>  // EW: compile-time ELL width (0 = read it from the context)
__global__ void __launch_bounds__(S1_THREADS)
k_stage1(DevCtx c, const double* __restrict__ x, int64_t W, int* __restrict__ status,
         double* __restrict__ lnprior, double* __restrict__ blobs, double* __restrict__ emu_in,
         double* __restrict__ amp, uint32_t flags) {
  // the z-function rows of this block's redshift (ELL weights with the parameter offset and scale they
  // multiply) are staged in shared memory: the inner loop then has one dependent global load (the
  // walker's coordinate) instead of four
  extern __shared__ __align__(16) double s1[];
  const int z = blockIdx.y;
  const int nt = C1D_NQ * c.ell_width;
  double* s_wgt = s1;            // [NQ * ell]
  double* s_pm = s_wgt + nt;     // pmin[idx]
  double* s_ps = s_pm + nt;      // pscale[idx]
  int* s_idx = reinterpret_cast<int*>(s_ps + nt);
  for (int t = threadIdx.x; t < nt; t += blockDim.x) {
    const int iq = t / c.ell_width, e = t - iq * c.ell_width;
    const int g = (iq * c.nz + z) * c.ell_width + e;
    const int idx = c.zidx[g];
    s_idx[t] = idx;
    s_wgt[t] = c.zwgt[g];
    s_pm[t] = idx >= 0 ? c.pmin[idx] : 0.0;
    s_ps[t] = idx >= 0 ? c.pscale[idx] : 0.0;
  }
  __syncthreads();
  // the per-(walker, z) records of a block are contiguous in memory ([z][walker][slot]): they are
  // assembled in shared memory and written as whole sectors (one thread's 208-byte record scattered
  // as 8-byte stores costs read-modify-write traffic in the memory system: 445 MB against 224 MB)
  __shared__ double s_rec[S1_THREADS * (C1D_NAMP + 1)];
  __shared__ double s_emu[S1_THREADS * (C1D_MAX_EMU + 1)];
  const int64_t wb0 = (int64_t)blockIdx.x * S1_THREADS;
  const int64_t w = wb0 + threadIdx.x;
  const bool live = w < W;
  if (live) {
  const XRow<XT> xw = {XT ? x + w : x + w * c.ndim, W};

  // primordial-power rescaling (lya_theory.py:281-307, 542-580)
  double ratio_As = 1.0, d_ns = 0.0, d_nrun = 0.0;
  if (c.idx_As >= 0) ratio_As = param_value(c, xw, c.idx_As) / c.As_fid;
  if (c.idx_ns >= 0) d_ns = param_value(c, xw, c.idx_ns) - c.ns_fid;
  if (c.idx_nrun >= 0) d_nrun = param_value(c, xw, c.idx_nrun) - c.nrun_fid;
  const double ln_rA = log(ratio_As);
  const double e_emu = exp(ln_rA + (d_ns + 0.5 * d_nrun * c.L_emu) * c.L_emu);
  const double dn_emu = d_ns + d_nrun * c.L_emu;

  if (z == 0) {
    double* bl = blobs + w * 6;
    // compute_log_prob: out of the unit cube (likelihood.py:989-994)
    bool out = false;
    double lp = 0.0;
    for (int i = 0; i < c.ndim; ++i) {
      double xi = xw[i];
      out |= (xi > 1.0) | (xi < 0.0);
      if (c.has_gauss) {
        double d = c.prior_x0[i] - xi;  // get_log_prior, likelihood.py:1060-1063
        lp += d * d * c.prior_w[i];
      }
    }
    if (out && !(flags & C1D_FLAG_NO_CUBE)) {
      atomicMax(status + w, (int)ST_OUT_OF_CUBE);
      lnprior[w] = 0.0;
      double qnan = __longlong_as_double(0x7ff8000000000000LL);
      for (int i = 0; i < 6; ++i) bl[i] = qnan;  // Theory.get_blob(), lya_theory.py:514-523
    } else {
      lnprior[w] = -lp;
      double b0 = c.blob_fid[0] * exp(ln_rA + (d_ns + 0.5 * d_nrun * c.L_star) * c.L_star);
      double b1 = c.blob_fid[1] + d_ns + d_nrun * c.L_star;
      double b2 = c.blob_fid[2] + d_nrun;
      bl[0] = b0; bl[1] = b1; bl[2] = b2;
      bl[3] = c.blob_fid[3]; bl[4] = c.blob_fid[4]; bl[5] = c.blob_fid[5];
      // star-prior box (lya_theory.py:647-654)
      bool rej0 = (b0 > c.star_hi[0]) | (b0 < c.star_lo[0]) | (b1 > c.star_hi[1]) | (b1 < c.star_lo[1]) |
                  (b2 > c.star_hi[2]) | (b2 < c.star_lo[2]);
      if (rej0) atomicMax(status + w, (int)ST_REJECTED);
    }
  }

  const bool use_hull = (c.n_hull_facets > 0) && !(flags & C1D_FLAG_NO_HULL);
  const bool use_zmask = (flags & C1D_FLAG_ZMASK) != 0;
  bool rej = false;
  // rows of walkers that end up rejected are written too (finite or not): later stages skip them by status
  {
    double q[C1D_NQ];
    if (EW > 0) {
      // the coordinates of four quantities at a time: their loads are independent and in flight together (the kernel
      // is bound by the latency of these gathers: one dependent round trip per quantity otherwise)
      static_assert(C1D_NQ % 4 == 0, "quantities in groups of four");
#pragma unroll 1
      for (int iq0 = 0; iq0 < C1D_NQ; iq0 += 4) {
        double xv[4][EW > 0 ? EW : 1];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int e = 0; e < EW; ++e) {
            const int idx = s_idx[(iq0 + a) * EW + e];
            xv[a][e] = xw[idx >= 0 ? idx : 0];
          }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int iq = iq0 + a;
          double lin = c.zbase[iq * c.nz + z];
#pragma unroll
          for (int e = 0; e < EW; ++e) {
            const int t = iq * EW + e;
            if (s_idx[t] >= 0) lin += s_wgt[t] * (s_pm[t] + xv[a][e] * s_ps[t]);  // same operations in the same order as below
          }
          double v = c.zotype[iq] ? exp(lin) : lin;
          q[iq] = (lin <= c.znull[iq]) ? 0.0 : v;
        }
      }
    } else {
#pragma unroll 1
    for (int iq = 0; iq < C1D_NQ; ++iq) {
      int row = iq * c.nz + z;
      double lin = c.zbase[row];
      for (int e = 0; e < c.ell_width; ++e) {
        const int t = iq * c.ell_width + e;
        const int idx = s_idx[t];
        if (idx >= 0) lin += s_wgt[t] * (s_pm[t] + xw[idx] * s_ps[t]);  // wgt * value_from_cube (likelihood_parameter.py:58-61)
      }
      // zero the '<= null' values (si_mult.py:170-176), then the output type
      double v = c.zotype[iq] ? exp(lin) : lin;
      q[iq] = (lin <= c.znull[iq]) ? 0.0 : v;
    }
    }
    double M = c.mz[z];
    // IGM histories (mean_flux_class.py:51-61, thermal_class.py:52-72, pressure_class.py:51-56)
    double tau = q[C1D_Q_TAU] * c.igmfid[0 * c.nz + z];
    double mF = exp(-tau);
    double sigT_kms = q[C1D_Q_SIGT] * c.igmfid[1 * c.nz + z];
    double gamma = q[C1D_Q_GAMMA] * c.igmfid[2 * c.nz + z];
    double kF_kms = q[C1D_Q_KF] * c.igmfid[3 * c.nz + z];
    double ev[C1D_MAX_EMU];
    for (int j = 0; j < C1D_MAX_EMU; ++j) {
      double v = 0.0;
      if (j < c.n_emu) {
        switch (c.emu_code[j]) {  // lya_theory.py:447-487
          case 0: v = c.linp[z * 3 + 0] * e_emu; break;
          case 1: v = c.linp[z * 3 + 1] + dn_emu; break;
          case 2: v = c.linp[z * 3 + 2] + d_nrun; break;
          case 3: v = mF; break;
          case 4: v = sigT_kms / M; break;
          case 5: v = gamma; break;
          case 6: v = kF_kms * M; break;
          default: v = 1000.0 / (kF_kms * M); break;
        }
      }
      ev[j] = v;
    }
    if (use_hull && (!use_zmask || c.zmask[z])) {
      // every facet of every 2-D hull: A p + b <= tol (hull.py:9-10, 157-165)
      for (int f = 0; f < c.n_hull_facets; ++f) {
        double p0 = ev[c.hull_dim[2 * f]], p1 = ev[c.hull_dim[2 * f + 1]];
        double val = c.hull_abc[3 * f] * p0 + c.hull_abc[3 * f + 1] * p1 + c.hull_abc[3 * f + 2];
        if (!(val <= c.hull_tol)) { rej = true; break; }
      }
    }
    double* eo = s_emu + threadIdx.x * (C1D_MAX_EMU + 1);
    for (int j = 0; j < C1D_MAX_EMU; ++j) eo[j] = ev[j];

    double* a = s_rec + threadIdx.x * (C1D_NAMP + 1);
    for (int j = 0; j < A_S3; ++j) a[j] = 0.0;  // coefficient slots: stage 2 fills them
    a[C1D_NAMP - 1] = 0.0;
    double om = 1.0 - mF;
    a[A_S3] = q[C1D_Q_S_SIIII];
    a[A_S2] = q[C1D_Q_S_SIII];
    if (c.metal_model == 0) {
      // SiMult amplitudes (si_mult.py:242-247, 267-314)
      double a3 = q[C1D_Q_F_SIIII] / om;
      double a2 = c.rb3 * q[C1D_Q_F_SIII] / om;
      double r = c.ra3_over_rb3 * q[C1D_Q_F_SIIIA_SIIII];
      double gmm = q[C1D_Q_F_SIIIB_SIIII];
      a[A_M0] = 1.0 + (a3 * a3 * c.c0_o1 + a2 * a2 * (r * r * c.c0_o2 + c.c0_o3c));
      a[A_C1] = 2.0 * a3;
      a[A_C2A] = 2.0 * a2 * r;
      a[A_C2B] = 2.0 * a2;
      a[A_C3A] = 2.0 * a3 * a2 * gmm * r;
      a[A_C3B] = 2.0 * a3 * a2 * gmm;
    } else {
      // SiVid (si_vid_final.py:205-217)
      a[A_M0] = 1.0;
      a[A_C1] = q[C1D_Q_F_SIIII] / om;
      a[A_C2A] = q[C1D_Q_F_SIII] / om;
      a[A_C2B] = 0.0; a[A_C3A] = 0.0; a[A_C3B] = 0.0;
    }
    a[A_H0] = 1.0 + q[C1D_Q_HCD_CONST];  // hcd_model_rogers_class.py:116
    // HCD_Model_McDonald2005 (hcd_model_McDonald2005.py:71-92): A_damp(z) unless the constant coefficient is nulled
    a[A_HCD1 + 0] = (!c.hcd_gate || q[C1D_Q_HCD2] > 0.0) ? q[C1D_Q_HCD1] : 0.0;
    a[A_HCD1 + 1] = c.hcd_gate ? 0.0 : q[C1D_Q_HCD2];
    a[A_HCD1 + 2] = q[C1D_Q_HCD3];
    a[A_HCD1 + 3] = q[C1D_Q_HCD4];
    a[A_AA] = q[C1D_Q_F_SIADD] * q[C1D_Q_F_SIADD];  // si_add.py:219
    a[A_SA2] = q[C1D_Q_S_SIADD] * q[C1D_Q_S_SIADD]; // si_add.py:168-170
    a[A_RES] = c.apply_resolution ? q[C1D_Q_RES] : 0.0;  // lya_theory.py:712-722
    a[A_AGN] = (q[C1D_Q_AGN_CONST] > 0.0) ? q[C1D_Q_AGN] : 0.0;  // AGN_model.py:81-91
  }
  if (rej) atomicMax(status + w, (int)ST_REJECTED);
  }  // live
  __syncthreads();
  const int nrec = (int)((W - wb0 < S1_THREADS) ? (W - wb0) : S1_THREADS);
  double* ga = amp + ((int64_t)z * W + wb0) * C1D_NAMP;
  for (int t = threadIdx.x; t < nrec * C1D_NAMP; t += S1_THREADS) {
    const int r = t / C1D_NAMP;
    ga[t] = s_rec[r * (C1D_NAMP + 1) + (t - r * C1D_NAMP)];
  }
  double* ge = emu_in + ((int64_t)z * W + wb0) * C1D_MAX_EMU;
  for (int t = threadIdx.x; t < nrec * C1D_MAX_EMU; t += S1_THREADS) {
    const int r = t / C1D_MAX_EMU;
    ge[t] = s_emu[r * (C1D_MAX_EMU + 1) + (t - r * C1D_MAX_EMU)];
  }
}

int c1d_launch_stage1(c1d_ctx* ctx, const double* x, int64_t W, double* blobs_out, uint32_t flags,
                      cudaStream_t s) {
  int threads = S1_THREADS;
  int64_t blocks = (W + threads - 1) / threads;
  C1D_CUDA(ctx, cudaMemsetAsync(ctx->ws.status, 0, (size_t)W * sizeof(int), s));
  const size_t smem1 = (size_t)C1D_NQ * ctx->cfg.ell_width * (3 * sizeof(double) + sizeof(int));
  // from a few thousand walkers on the kernel reads a transposed copy of the coordinates (one more small launch)
  static const int64_t xt_min = getenv("C1D_S1_XT_MIN") ? atoll(getenv("C1D_S1_XT_MIN")) : 4096;
  if (W >= xt_min) {
    k_transpose_x<<<dim3((unsigned)((W + 31) / 32), (unsigned)((ctx->cfg.ndim + 31) / 32)), 256, 0, s>>>(x, W, ctx->cfg.ndim, ctx->ws.xt);
    ctx->launches++;
    auto kern = ctx->cfg.ell_width == 2 ? k_stage1<true, 2> : ctx->cfg.ell_width == 4 ? k_stage1<true, 4> : k_stage1<true, 0>;
    kern<<<dim3((unsigned)blocks, (unsigned)ctx->cfg.nz), threads, smem1, s>>>(ctx->d, ctx->ws.xt, W, ctx->ws.status, ctx->ws.lnprior, blobs_out,
                                                                               ctx->ws.emu_in, ctx->ws.amp, flags);
  } else {
    auto kern = ctx->cfg.ell_width == 2 ? k_stage1<false, 2> : ctx->cfg.ell_width == 4 ? k_stage1<false, 4> : k_stage1<false, 0>;
    kern<<<dim3((unsigned)blocks, (unsigned)ctx->cfg.nz), threads, smem1, s>>>(ctx->d, x, W, ctx->ws.status, ctx->ws.lnprior, blobs_out,
                                                                               ctx->ws.emu_in, ctx->ws.amp, flags);
  }
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return C1D_OK;
}

// ---------------------------------------------------------------------------------------
// stage 2b: GP posterior mean, one warp per row.  m_o = sum_i k(u, U_i) alpha[i, o] with
// an ARD squared-exponential kernel; fp64 because alpha = K^-1 y cancels heavily.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gp_f64(DevCtx c, const double* __restrict__ emu_in, int64_t nrows, double* __restrict__ out,
         int out_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < nrows; row += nwarps) {
    double u[C1D_MAX_EMU];
#pragma unroll
    for (int d = 0; d < C1D_MAX_EMU; ++d) {
      double xv = emu_in[row * C1D_MAX_EMU + d];
      u[d] = (d < c.n_emu) ? (xv - c.emu_lo[d]) / (c.emu_hi[d] - c.emu_lo[d]) : 0.0;
    }
    double acc[C1D_MAX_COEF];
#pragma unroll
    for (int o = 0; o < C1D_MAX_COEF; ++o) acc[o] = 0.0;
    for (int i = lane; i < c.gp_ntrain; i += 32) {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < C1D_MAX_EMU; ++d) {
        if (d < c.n_emu) {
          double t = (u[d] - __ldg(c.gp_x + (size_t)i * c.n_emu + d)) * c.gp_invell[d];
          s = fma(t, t, s);
        }
      }
      double kern = c.gp_sigma2 * exp(-0.5 * s);
#pragma unroll
      for (int o = 0; o < C1D_MAX_COEF; ++o)
        if (o < c.nc) acc[o] = fma(kern, __ldg(c.gp_alpha + (size_t)i * c.nc + o), acc[o]);
    }
#pragma unroll
    for (int o = 0; o < C1D_MAX_COEF; ++o)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
    if (lane == 0) {
      double mF = emu_in[row * C1D_MAX_EMU + c.gp_imf];
      double lnn = 0.0;  // ln norm(mF), Horner
      for (int j = c.gp_nnorm - 1; j >= 0; --j) lnn = lnn * mF + c.gp_norm[j];
      for (int o = 0; o < c.nc; ++o) {
        double y = c.gp_ymean[o] + c.gp_ystd[o] * acc[o];
        if (o == 0) y += lnn;
        out[row * out_stride + o] = y;
      }
    }
  }
}

// The same product with the training set resident in shared memory (structure of arrays, inputs
// pre-divided by the length scales, alpha pre-multiplied by sigma^2), two rows per warp so that every
// training point read feeds two rows, and e^x from a 32-entry table + degree-6 polynomial: 34 fp64
// operations per (row, training point) instead of ~73 with eight uncoalesced global loads.  One
// persistent 16-warp CTA per SM.  Used when the training set fits in shared memory.
#ifndef GP_THREADS
#define GP_THREADS 512
#endif
__global__ void __launch_bounds__(GP_THREADS, 1)
k_gp_f64_smem(DevCtx c, const double* __restrict__ emu_in, int64_t nrows, double* __restrict__ out,
              int out_stride, int ntp) {
  extern __shared__ __align__(16) double gsm[];
  double* Xs = gsm;                                   // [n_emu][ntp]
  double* As = Xs + (size_t)c.n_emu * ntp;            // [nc][ntp]
  double* etab = As + (size_t)c.nc * ntp;             // [32]
  for (int t = threadIdx.x; t < c.n_emu * ntp; t += blockDim.x) {
    const int d = t / ntp, i = t - d * ntp;
    // pad points sit far away (kernel value 0) and carry alpha = 0
    Xs[t] = (i < c.gp_ntrain) ? c.gp_x[(size_t)i * c.n_emu + d] * c.gp_invell[d] : 1.0e3;
  }
  for (int t = threadIdx.x; t < c.nc * ntp; t += blockDim.x) {
    const int o = t / ntp, i = t - o * ntp;
    As[t] = (i < c.gp_ntrain) ? c.gp_sigma2 * c.gp_alpha[(size_t)i * c.nc + o] : 0.0;
  }
  if (threadIdx.x < 32) etab[threadIdx.x] = c_exp2_tab[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t npairs = (nrows + 1) / 2;
  for (int64_t pr = (int64_t)blockIdx.x * (GP_THREADS / 32) + warp; pr < npairs; pr += (int64_t)gridDim.x * (GP_THREADS / 32)) {
    const int64_t row0 = 2 * pr, row1 = (2 * pr + 1 < nrows) ? 2 * pr + 1 : 2 * pr;
    double u0[C1D_MAX_EMU], u1[C1D_MAX_EMU];
#pragma unroll
    for (int d = 0; d < C1D_MAX_EMU; ++d) {
      u0[d] = u1[d] = 0.0;
      if (d < c.n_emu) {
        const double lo = c.emu_lo[d], sc = c.gp_invell[d] / (c.emu_hi[d] - lo);
        u0[d] = (emu_in[row0 * C1D_MAX_EMU + d] - lo) * sc;
        u1[d] = (emu_in[row1 * C1D_MAX_EMU + d] - lo) * sc;
      }
    }
    double acc0[C1D_MAX_COEF], acc1[C1D_MAX_COEF];
#pragma unroll
    for (int o = 0; o < C1D_MAX_COEF; ++o) acc0[o] = acc1[o] = 0.0;
    for (int i = lane; i < ntp; i += 32) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int d = 0; d < C1D_MAX_EMU; ++d) {
        if (d < c.n_emu) {
          const double xv = Xs[d * ntp + i];
          const double t0 = u0[d] - xv, t1 = u1[d] - xv;
          s0 = fma(t0, t0, s0);
          s1 = fma(t1, t1, s1);
        }
      }
      const double k0 = exp_tab_neg(-0.5 * s0, etab), k1 = exp_tab_neg(-0.5 * s1, etab);
#pragma unroll
      for (int o = 0; o < C1D_MAX_COEF; ++o) {
        if (o < c.nc) {
          const double a = As[o * ntp + i];
          acc0[o] = fma(k0, a, acc0[o]);
          acc1[o] = fma(k1, a, acc1[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < C1D_MAX_COEF; ++o)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        acc0[o] += __shfl_xor_sync(0xffffffffu, acc0[o], off);
        acc1[o] += __shfl_xor_sync(0xffffffffu, acc1[o], off);
      }
    if (lane < 2 && (lane == 0 || row1 != row0)) {
      const int64_t row = lane == 0 ? row0 : row1;
      const double mF = emu_in[row * C1D_MAX_EMU + c.gp_imf];
      double lnn = 0.0;  // ln norm(mF), Horner
      for (int j = c.gp_nnorm - 1; j >= 0; --j) lnn = lnn * mF + c.gp_norm[j];
#pragma unroll
      for (int o = 0; o < C1D_MAX_COEF; ++o) {
        if (o < c.nc) {
          double y = c.gp_ymean[o] + c.gp_ystd[o] * (lane == 0 ? acc0[o] : acc1[o]);
          if (o == 0) y += lnn;
          out[row * out_stride + o] = y;
        }
      }
    }
  }
}

// The same product with lane = row: a group of GPL_SPLIT warps owns 32 rows, each warp of the group walks
// its quarter of the training set and every training-point value is ONE broadcast shared-memory read for
// 32 rows (the two-rows-per-warp kernel above reads 13 doubles per lane and training point for 2 rows:
// its shared-memory wavefronts, not the fp64 pipe, pace it).  The partial sums of the group are added in
// a fixed order (part 0 + 1 + 2 + 3), so a row's result does not depend on the batch it travels in.
// Measured on B200 (C2: 1320 training points, 6 inputs, 6 outputs; two-rows-per-warp kernel | this one):
// 65 536 walkers 4.12 | 2.47 ms, 16 384: 1.06 | 0.67, 4096: 0.29 | 0.19, 1024: 0.100 | 0.105 (the older form
// keeps the batches below one item per group slot).  512 threads: 2.50 ms; run-time input / output counts
// (predicated loads and FMAs) and a branch in the exponential: 3.17 ms; unrolled by 2 instead of 4: 2.47 ms against 2.43.
#define GPL_THREADS 640
#define GPL_SPLIT 4
#define GPL_GROUPS (GPL_THREADS / 32 / GPL_SPLIT)
// NE / NO: emulator inputs / outputs at compile time (0 = run-time counts)
template <int NE, int NO>
__global__ void __launch_bounds__(GPL_THREADS, 1)
k_gp_f64_lanes(DevCtx c, const double* __restrict__ emu_in, int64_t nrows, double* __restrict__ out,
               int out_stride, int ntp) {
  const int n_emu = NE ? NE : c.n_emu, nc = NO ? NO : c.nc;
  extern __shared__ __align__(16) double gsm[];
  double* Xs = gsm;                                   // [ntp][8]: inputs of a training point, pre-divided by the length scales
  double* As = Xs + (size_t)C1D_MAX_EMU * ntp;        // [ntp][8]: sigma^2 alpha of a training point
  double* etab = As + (size_t)C1D_MAX_COEF * ntp;     // [32]
  double* part = etab + 32;                           // [GPL_GROUPS][GPL_SPLIT - 1][C1D_MAX_COEF][32]
  for (int t = threadIdx.x; t < C1D_MAX_EMU * ntp; t += blockDim.x) {
    const int i = t / C1D_MAX_EMU, d = t - i * C1D_MAX_EMU;
    // pad points sit far away (kernel value 0) and carry alpha = 0; unused input slots are 0 on both sides
    Xs[t] = (d < n_emu) ? ((i < c.gp_ntrain) ? c.gp_x[(size_t)i * n_emu + d] * c.gp_invell[d] : 1.0e3) : 0.0;
  }
  for (int t = threadIdx.x; t < C1D_MAX_COEF * ntp; t += blockDim.x) {
    const int i = t / C1D_MAX_COEF, o = t - i * C1D_MAX_COEF;
    As[t] = (o < nc && i < c.gp_ntrain) ? c.gp_sigma2 * c.gp_alpha[(size_t)i * nc + o] : 0.0;
  }
  if (threadIdx.x < 32) etab[threadIdx.x] = c_exp2_tab[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp / GPL_SPLIT, prt = warp % GPL_SPLIT;
  const int chunk = ntp / GPL_SPLIT;  // ntp is a multiple of 32
  const int i0 = prt * chunk, i1 = i0 + chunk;
  const int64_t nitems = (nrows + 31) / 32;
  for (int64_t item = (int64_t)blockIdx.x * GPL_GROUPS + grp; item < nitems; item += (int64_t)gridDim.x * GPL_GROUPS) {
    const int64_t row = item * 32 + lane;
    const int64_t rv = row < nrows ? row : nrows - 1;
    double u[C1D_MAX_EMU];
#pragma unroll
    for (int d = 0; d < C1D_MAX_EMU; ++d) {
      u[d] = 0.0;
      if (d < n_emu) {
        const double lo = c.emu_lo[d], sc = c.gp_invell[d] / (c.emu_hi[d] - lo);
        u[d] = (emu_in[rv * C1D_MAX_EMU + d] - lo) * sc;
      }
    }
    double acc[C1D_MAX_COEF];
#pragma unroll
    for (int o = 0; o < C1D_MAX_COEF; ++o) acc[o] = 0.0;
    const double2* xp = reinterpret_cast<const double2*>(Xs) + (size_t)i0 * (C1D_MAX_EMU / 2);
    const double2* ap = reinterpret_cast<const double2*>(As) + (size_t)i0 * (C1D_MAX_COEF / 2);
#pragma unroll 4
    for (int i = i0; i < i1; ++i, xp += C1D_MAX_EMU / 2, ap += C1D_MAX_COEF / 2) {
      double s2 = 0.0;
#pragma unroll
      for (int h = 0; h < C1D_MAX_EMU / 2; ++h) {
        if (2 * h < n_emu) {
          const double2 xv = xp[h];  // broadcast
          const double t0 = u[2 * h] - xv.x, t1 = u[2 * h + 1] - xv.y;
          s2 = fma(t0, t0, s2);
          s2 = fma(t1, t1, s2);
        }
      }
      const double kv = exp_tab_neg_clamped(-0.5 * s2, etab);
#pragma unroll
      for (int h = 0; h < C1D_MAX_COEF / 2; ++h) {
        if (2 * h < nc) {
          const double2 av = ap[h];  // broadcast
          acc[2 * h] = fma(kv, av.x, acc[2 * h]);
          acc[2 * h + 1] = fma(kv, av.y, acc[2 * h + 1]);
        }
      }
    }
    // fixed-order reduction over the parts of the group
    double* pg = part + (size_t)grp * (GPL_SPLIT - 1) * C1D_MAX_COEF * 32;
    if (prt > 0) {
#pragma unroll
      for (int o = 0; o < C1D_MAX_COEF; ++o) pg[((prt - 1) * C1D_MAX_COEF + o) * 32 + lane] = acc[o];
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(32 * GPL_SPLIT) : "memory");
    if (prt == 0) {
#pragma unroll
      for (int q = 0; q < GPL_SPLIT - 1; ++q)
#pragma unroll
        for (int o = 0; o < C1D_MAX_COEF; ++o) acc[o] += pg[(q * C1D_MAX_COEF + o) * 32 + lane];
      if (row < nrows) {
        const double mF = emu_in[row * C1D_MAX_EMU + c.gp_imf];
        double lnn = 0.0;  // ln norm(mF), Horner
        for (int j = c.gp_nnorm - 1; j >= 0; --j) lnn = lnn * mF + c.gp_norm[j];
#pragma unroll
        for (int o = 0; o < C1D_MAX_COEF; ++o) {
          if (o < nc) {
            double y = c.gp_ymean[o] + c.gp_ystd[o] * acc[o];
            if (o == 0) y += lnn;
            out[row * out_stride + o] = y;
          }
        }
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(32 * GPL_SPLIT) : "memory");  // partials consumed before the next item
  }
}

int c1d_launch_emulator(c1d_ctx* ctx, const double* emu_in, int64_t nrows, double* out,
                        int out_stride, uint32_t flags, cudaStream_t s) {
  if (nrows <= 0) return C1D_OK;
  if (ctx->cfg.emu_kind == 0) {
    if (flags & C1D_FLAG_EMU_I8) return c1d_i8_launch(ctx, emu_in, nrows, out, out_stride, s);
    if (flags & C1D_FLAG_EMU_TC) return c1d_tc_launch(ctx, emu_in, nrows, out, out_stride, s);
    return c1d_mlp64_launch(ctx, emu_in, nrows, out, out_stride, s);
  } else {
    const c1d_config& cfg = ctx->cfg;
    const int ntp = (cfg.gp_ntrain + 31) & ~31;
    const size_t smem = ((size_t)(cfg.n_emu + cfg.nc) * ntp + 32) * sizeof(double);
    const int64_t pairs = (nrows + 1) / 2;
    // lane = row kernel: tables [ntp][8] + [ntp][8], the partial sums of the groups
    const size_t smem_l = ((size_t)(C1D_MAX_EMU + C1D_MAX_COEF) * ntp + 32 + (size_t)GPL_GROUPS * (GPL_SPLIT - 1) * C1D_MAX_COEF * 32) * sizeof(double);
    const int64_t items = (nrows + 31) / 32;
    const char* gk = getenv("C1D_GP_KERNEL");  // lanes | pairs | global: forces one form (tests, measurements)
    const bool want_lanes = gk ? !strcmp(gk, "lanes") : (items >= (int64_t)ctx->num_sms * GPL_GROUPS);
    if (smem_l <= 227 * 1024 && want_lanes) {
      void (*kern)(DevCtx, const double*, int64_t, double*, int, int) = k_gp_f64_lanes<0, 0>;
      const int ne = cfg.n_emu, no = cfg.nc;
      if (ne == 6 && no == 5) kern = k_gp_f64_lanes<6, 5>;
      else if (ne == 6 && no == 6) kern = k_gp_f64_lanes<6, 6>;
      else if (ne == 6 && no == 7) kern = k_gp_f64_lanes<6, 7>;
      else if (ne == 7 && no == 5) kern = k_gp_f64_lanes<7, 5>;
      else if (ne == 7 && no == 6) kern = k_gp_f64_lanes<7, 6>;
      else if (ne == 7 && no == 7) kern = k_gp_f64_lanes<7, 7>;
      C1D_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
      int64_t blocks = (items + GPL_GROUPS - 1) / GPL_GROUPS;
      int grid = (int)(blocks < ctx->num_sms ? blocks : ctx->num_sms);
      kern<<<grid, GPL_THREADS, smem_l, s>>>(ctx->d, emu_in, nrows, out, out_stride, ntp);
    } else if (smem <= 227 * 1024 && (gk ? !strcmp(gk, "pairs") : pairs >= (int64_t)ctx->num_sms * (GP_THREADS / 32) / 2)) {
      // training set resident in shared memory, two rows per warp
      C1D_CUDA(ctx, cudaFuncSetAttribute(k_gp_f64_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int64_t blocks = (pairs + GP_THREADS / 32 - 1) / (GP_THREADS / 32);
      int grid = (int)(blocks < ctx->num_sms ? blocks : ctx->num_sms);
      k_gp_f64_smem<<<grid, GP_THREADS, smem, s>>>(ctx->d, emu_in, nrows, out, out_stride, ntp);
    } else {
      // small batches (the staging of the training set would dominate) and training sets beyond shared memory
      int64_t warps_needed = nrows;
      int64_t blocks = (warps_needed + 7) / 8;
      int64_t cap = (int64_t)ctx->num_sms * 8;
      int grid = (int)(blocks < cap ? blocks : cap);
      k_gp_f64<<<grid, 256, 0, s>>>(ctx->d, emu_in, nrows, out, out_stride);
    }
  }
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return C1D_OK;
}

// ---------------------------------------------------------------------------------------
// stage 3: fused emulator polynomial + contaminants + rebinning + residual.
// A CTA is bound to one redshift bin: it stages the walker-independent k tables of that z
// in shared memory once and streams a tile of walkers through them, one walker per warp.
// ---------------------------------------------------------------------------------------
#define P1D_THREADS 512
#define P1D_WARPS (P1D_THREADS / 32)
#define P1D_CHUNK 32   // walkers per work item
#define P1D_NPAIR (C1D_NCOL / 2)
// position (in double2 units) of column pair pr of fine point i in the staged tables
#define P1D_TIDX(pr, i) ((((i) >> 5) * P1D_NPAIR + (pr)) * 32 + ((i) & 31))

// A persistent CTA streams work items (z, chunk of 32 walkers) that are contiguous in a
// z-major order and balanced by cost (fine points per z), so the grid is exactly one wave.
// Compile-time specialisation: NC = emulator coefficients (0 = run-time count), KIND = emulator
// kind (0: 10^poly(log10 k), 1: exp(poly(k/kmax))), METAL = 0 SiMult / 1 SiVid, UNI = evenly
// spaced k grid (exp recurrences).
template <int NC, int KIND, int METAL, bool UNI>
__global__ void __launch_bounds__(P1D_THREADS, 1)
k_fused_p1d(DevCtx c, int64_t W, const int* __restrict__ status, const double* __restrict__ amp,
            double* __restrict__ dvec, double* __restrict__ p_out, double* __restrict__ chi2z,
            int nfz_max, int ndz_max) {
  extern __shared__ __align__(16) double sm[];
  const int RW = c.rebin_width, RWP = RW | 1;        // odd row stride: conflict-free
  const int pb_len = nfz_max + (nfz_max >> 3) + 2;   // P(i) lives at i + i/8
  double2* tab2 = reinterpret_cast<double2*>(sm);    // [blocks of 32 points][P1D_NPAIR][32] column pairs
  double* pbuf = sm + (size_t)2 * P1D_NPAIR * ((nfz_max + 31) & ~31);   // [P1D_WARPS][pb_len]
  double* rbw = pbuf + (size_t)P1D_WARPS * pb_len;   // [ndz_max][RWP]
  double* dat = rbw + (size_t)ndz_max * RWP;         // [ndz_max]
  double* dzb = dat + ndz_max;                       // [P1D_WARPS][ndz_max]
  int* rbs = reinterpret_cast<int*>(dzb + (size_t)P1D_WARPS * ndz_max);  // [ndz_max] first fine point of a bin
  int* rbn = rbs + ndz_max;                                              // [ndz_max] taps up to the last non-zero weight

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* pb = pbuf + (size_t)warp * pb_len;
  double* dz = dzb + (size_t)warp * ndz_max;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  // cost-balanced contiguous range of this CTA: [t0, t1) in units of (fine points + 2 data bins)
  const int64_t nchunks = (W + P1D_CHUNK - 1) / P1D_CHUNK;
  int64_t total = 0;
  for (int z = 0; z < c.nz; ++z)
    total += nchunks * (int64_t)((c.zoff_f[z + 1] - c.zoff_f[z]) + 2 * (c.zoff_d[z + 1] - c.zoff_d[z]));
  const int64_t t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1) / gridDim.x;

  int64_t cum = 0;
  size_t icov_base = 0;
  for (int z = 0; z < c.nz; ++z) {
    const int f0 = c.zoff_f[z], nfz = c.zoff_f[z + 1] - f0;
    const int d0 = c.zoff_d[z], ndz = c.zoff_d[z + 1] - d0;
    const int64_t cost = nfz + 2 * ndz;
    const int64_t zend = cum + nchunks * cost;
    // chunks of this z whose start cost falls in [t0, t1)
    int64_t c_lo = (t0 > cum) ? (t0 - cum + cost - 1) / cost : 0;
    int64_t c_hi = (t1 < zend) ? ((t1 > cum) ? (t1 - cum + cost - 1) / cost : 0) : nchunks;
    const double* icz = c.icov_blocks + icov_base;
    icov_base += (size_t)ndz * ndz;
    cum = zend;
    if (c_lo >= c_hi) continue;

    __syncthreads();  // previous z fully processed before its tables are overwritten
    // table layout [block of 32 points][8 column pairs][32 lanes]: a lane reads all pairs of its
    // point at compile-time offsets from one pointer that advances by one block per iteration
    for (int t = threadIdx.x; t < P1D_NPAIR * nfz; t += blockDim.x) {
      int pr = t / nfz, i = t % nfz;
      tab2[P1D_TIDX(pr, i)] = make_double2(c.tab[(size_t)(2 * pr) * c.nf + f0 + i],
                                           c.tab[(size_t)(2 * pr + 1) * c.nf + f0 + i]);
    }
    for (int t = threadIdx.x; t < ndz * RW; t += blockDim.x)
      rbw[(t / RW) * RWP + (t % RW)] = c.rb_wgt[(size_t)d0 * RW + t];
    for (int t = threadIdx.x; t < ndz; t += blockDim.x) {
      dat[t] = c.data[d0 + t];
      const int st = c.rb_start[d0 + t];
      rbs[t] = st;
      int cnt = RW;
      while (cnt > 1 && c.rb_wgt[(size_t)(d0 + t) * RW + cnt - 1] == 0.0) --cnt;
      if (st + cnt > nfz) cnt = nfz - st;  // taps beyond the grid carry no weight
      rbn[t] = cnt < 0 ? 0 : cnt;
    }
    __syncthreads();
    const int64_t w_begin = c_lo * P1D_CHUNK;
    int64_t w_end = c_hi * P1D_CHUNK;
    if (w_end > W) w_end = W;
    for (int64_t w = w_begin + warp; w < w_end; w += P1D_WARPS) {
      const bool ok = status[w] == ST_OK;
      if (!ok) {
        // rejected walkers carry no model (the reference returns None)
        for (int j = lane; j < ndz; j += 32) {
          if (dvec) dvec[w * c.ldd + d0 + j] = 0.0;
          if (p_out) p_out[w * c.nd + d0 + j] = qnan;
        }
        if (chi2z && lane == 0) chi2z[w * c.nz + z] = 0.0;
        continue;
      }
      // amplitude record of (z, walker): every lane loads the same 16-byte words (one broadcast
      // transaction each)
      const double2* a2p = reinterpret_cast<const double2*>(amp + ((int64_t)z * W + w) * C1D_NAMP);
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

      // exp(-s k) and exp(-s^2 k^2) on an evenly spaced grid advance by constant ratios:
      // one exp per term and lane instead of one per point
      constexpr bool uni = UNI;
      const double sg3 = (METAL == 0) ? -s3 : s3;  // SiVid grows with k (si_vid_final.py:208)
      double E3 = 0.0, R3 = 0.0, E2 = 0.0, R2 = 0.0, gg = 0.0, grho = 0.0, gq = 0.0;
      if (uni) {
        const double dk32 = 32.0 * ((tab2[P1D_TIDX(C1D_COL_K / 2, nfz - 1)].x - tab2[P1D_TIDX(C1D_COL_K / 2, 0)].x) / (double)(nfz - 1));
        const double kl = tab2[P1D_TIDX(C1D_COL_K / 2, lane < nfz ? lane : 0)].x;
        E3 = exp_poly(sg3 * kl);
        E2 = exp_poly(-s2 * kl);
        gg = exp_poly(-1.0 * sa2 * (kl * kl));
        grho = exp_poly(-1.0 * sa2 * (2.0 * kl * dk32 + dk32 * dk32));
        // the three lane-independent ratios come from ONE evaluation: lanes 0..2 take one argument each
        const double arg = (lane == 0) ? sg3 * dk32 : (lane == 1) ? -s2 * dk32 : -2.0 * sa2 * (dk32 * dk32);
        const double rr = exp_poly(arg);
        R3 = __shfl_sync(0xffffffffu, rr, 0);
        R2 = __shfl_sync(0xffffffffu, rr, 1);
        gq = __shfl_sync(0xffffffffu, rr, 2);
      }
      // GP emulator with a k-table norm (Fig3a.py:83-85): the walker's two table rows and its weight in mean flux
      int gj = 0;
      double gw = 0.0;
      const bool gnorm = (KIND == 1) && c.gp_nm > 0;
      if (gnorm) {
        const double mF = c.gp_emu_in[((int64_t)z * W + w) * C1D_MAX_EMU + c.gp_imf];
        while (gj < c.gp_nm - 2 && mF >= __ldg(c.gp_nmf + gj + 1)) ++gj;
        const double m0n = __ldg(c.gp_nmf + gj), m1n = __ldg(c.gp_nmf + gj + 1);
        gw = fmin(fmax((mF - m0n) / (m1n - m0n), 0.0), 1.0);
      }
      const double2* tp = tab2 + lane;  // this lane's point of the current block
      for (int i = lane; i < nfz; i += 32, tp += P1D_NPAIR * 32) {
        const double2 ks = tp[(C1D_COL_K / 2) * 32], tt1 = tp[(C1D_COL_T / 2) * 32];  // (k, S), (t, T1)
        const double k = ks.x;
        if (!uni) {
          E3 = exp(sg3 * k);
          E2 = exp(-s2 * k);
          gg = exp(-1.0 * sa2 * (k * k));
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
        double p_emu = ks.y * (KIND == 0 ? exp10_poly(pl) : exp(pl));
        if (gnorm) {
          const double v = tp[(C1D_COL_AGN / 2) * 32].y;  // C1D_COL_SPARE: interval + weight in the norm's k grid
          const int cc = (int)v, cn = cc + 1 < c.gp_nkn ? cc + 1 : cc;
          const double* r0 = c.gp_ntab + (size_t)gj * c.gp_nkn;
          const double* r1 = r0 + c.gp_nkn;
          const double a0 = __ldg(r0 + cc), a1 = __ldg(r0 + cn);
          const double n0 = fma(gw, __ldg(r1 + cc) - a0, a0), n1 = fma(gw, __ldg(r1 + cn) - a1, a1);
          p_emu *= fma(v - (double)cc, n1 - n0, n0);
        }
        double mult;
        if (METAL == 0) {
          // si_mult.py:212-218, 267-341:  G = 2 - 2/(1 + E)
          const double2 v2 = tp[(C1D_COL_T2A / 2) * 32], v3 = tp[(C1D_COL_T3A / 2) * 32];  // (T2A, T2BC), (T3A, T3BC)
          const double G3 = fma(-2.0, rcp_fast(1.0 + E3), 2.0);
          const double G2 = fma(-2.0, rcp_fast(1.0 + E2), 2.0);
          mult = m0 + (c1 * G3 * tt1.y + G2 * (c2a * v2.x + c2b * v2.y)) + (c3a * v3.x + c3b * v3.y);
        } else {
          // si_vid_final.py:205-219
          mult = 1.0 + c1 * E3 + c2a * tt1.y * E2;
        }
        if (c.has_agn) mult *= fma(fagn, tp[(C1D_COL_AGN / 2) * 32].x, 1.0);  // AGN_model.py:116-120: 1 + f_AGN delta
        // hcd_model_rogers_class.py:115-135 (or hcd_boss.py:5-7 through D1)
        const double2 d12 = tp[(C1D_COL_D1 / 2) * 32], d34 = tp[(C1D_COL_D3 / 2) * 32], kq = tp[(C1D_COL_KT / 2) * 32];  // (D1, D2), (D3, D4), (KT, Q)
        const double hcd = h0 + hA * d12.x + hB * d12.y + hC * d34.x + hD * d34.y;
        // si_add.py:168-219
        const double add = aa * kq.x * gg;
        // resolution_class.py:99-108
        const double syst = 1.0 + res * kq.y;
        // lya_theory.py:778-784 (IC correction is folded into the S column)
        pb[i + (i >> 3)] = (hcd * mult * p_emu + add) * syst;
        if (uni) {
          E3 *= R3;
          E2 *= R2;
          gg *= grho;
          grho *= gq;
        }
      }
      __syncwarp();
      // rebinning (likelihood.py:147-161) and residual (likelihood.py:939, 957)
      for (int j = lane; j < ndz; j += 32) {
        const int st = rbs[j], cnt = rbn[j];
        const double* wr = rbw + j * RWP;
        double m = 0.0;
        for (int t = 0; t < cnt; ++t) {
          const int ii = st + t;
          m = fma(wr[t], pb[ii + (ii >> 3)], m);
        }
        const double dj = dat[j] - m;
        if (dvec) dvec[w * c.ldd + d0 + j] = dj;
        if (p_out) p_out[w * c.nd + d0 + j] = m;
        dz[j] = dj;
      }
      __syncwarp();
      if (chi2z) {
        // chi2_z = d^T C_z^-1 d (likelihood.py:939-940): lane owns columns j, rows broadcast
        double part = 0.0;
        for (int j = lane; j < ndz; j += 32) {
          double acc = 0.0;
#pragma unroll 4
          for (int i = 0; i < ndz; ++i) acc = fma(__ldg(icz + (size_t)i * ndz + j), dz[i], acc);
          part = fma(acc, dz[j], part);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (lane == 0) chi2z[w * c.nz + z] = part;
      }
      __syncwarp();
    }
  }
}

int c1d_launch_stage3(c1d_ctx* ctx, int64_t W, double* p_out, int want_dvec, int want_chi2z,
                      cudaStream_t s) {
  const c1d_config& cfg = ctx->cfg;
  {
    // the walker-per-lane kernel (c1d_p1d_lanes.cu) when its rebinning form applies and the batch fills the
    // machine; a requested per-z chi2 then comes from a small tensor-core kernel over its z-aligned residuals
    const int rc = c1d_p1dw_launch(ctx, W, p_out, want_dvec, want_chi2z, s);
    if (rc < 0) return rc;
    if (rc > 0) return want_chi2z ? c1d_launch_chi2z(ctx, W, s) : C1D_OK;
  }
  int nfz_max = ctx->nfz_max, ndz_max = ctx->ndz_max;
  const int RWP = cfg.rebin_width | 1;
  const int pb_len = nfz_max + (nfz_max >> 3) + 2;
  size_t smem = ((size_t)C1D_NCOL * ((nfz_max + 31) & ~31) + (size_t)P1D_WARPS * pb_len + (size_t)ndz_max * RWP + ndz_max +
                 (size_t)P1D_WARPS * ndz_max) * sizeof(double) + 2 * (size_t)ndz_max * sizeof(int) + 16;
  if (smem > 227 * 1024) {
    snprintf(ctx->err, sizeof(ctx->err), "stage 3 needs %zu B of shared memory (z-bin too large)", smem);
    return C1D_ERR_ARG;
  }
  typedef void (*kern_t)(DevCtx, int64_t, const int*, const double*, double*, double*, double*, int, int);
  kern_t kern = nullptr;
  const int nc = (cfg.nc >= 5 && cfg.nc <= 7) ? cfg.nc : 0;
  const int sel = ((nc ? nc - 4 : 0) << 3) | (cfg.emu_kind << 2) | (cfg.metal_model << 1) | (cfg.uniform_k ? 1 : 0);
#define P1D_CASE(NCI, NCV)                                                                              \
  case ((NCI) << 3) | 0: kern = k_fused_p1d<NCV, 0, 0, false>; break;                                   \
  case ((NCI) << 3) | 1: kern = k_fused_p1d<NCV, 0, 0, true>; break;                                    \
  case ((NCI) << 3) | 2: kern = k_fused_p1d<NCV, 0, 1, false>; break;                                   \
  case ((NCI) << 3) | 3: kern = k_fused_p1d<NCV, 0, 1, true>; break;                                    \
  case ((NCI) << 3) | 4: kern = k_fused_p1d<NCV, 1, 0, false>; break;                                   \
  case ((NCI) << 3) | 5: kern = k_fused_p1d<NCV, 1, 0, true>; break;                                    \
  case ((NCI) << 3) | 6: kern = k_fused_p1d<NCV, 1, 1, false>; break;                                   \
  case ((NCI) << 3) | 7: kern = k_fused_p1d<NCV, 1, 1, true>; break;
  switch (sel) {
    P1D_CASE(0, 0)
    P1D_CASE(1, 5)
    P1D_CASE(2, 6)
    P1D_CASE(3, 7)
  }
#undef P1D_CASE
  if (!kern) {
    snprintf(ctx->err, sizeof(ctx->err), "no stage-3 kernel for this configuration");
    return C1D_ERR_ARG;
  }
  C1D_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // exactly one wave: one 16-warp CTA per SM (the tables are shared by all its warps)
  int64_t items = ((W + P1D_CHUNK - 1) / P1D_CHUNK) * cfg.nz;
  int64_t grid = (int64_t)ctx->num_sms;
  if (grid > items) grid = items;
  double* dvec = want_dvec ? ctx->ws.dvec : nullptr;
  double* chi2z = want_chi2z ? ctx->ws.chi2z : nullptr;
  ctx->d.gp_emu_in = ctx->ws.emu_in;
  kern<<<(unsigned)grid, P1D_THREADS, smem, s>>>(ctx->d, W, ctx->ws.status, ctx->ws.amp, dvec, p_out, chi2z, nfz_max,
                                                  ndz_max);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return C1D_OK;
}

// ---------------------------------------------------------------------------------------
// stage 4: chi2 = d^T C^-1 d with the dense inverse covariance (likelihood.py:956-960) as a
// tiled fp64 GEMM on the tensor cores (mma.sync m8n8k4 f64) with the row-dot fused into the
// epilogue.  CTA tile = 128 walkers x 64 columns j; the K loop runs over rows i of C^-1 in
// chunks of 16 staged with cp.async (double buffered).  C^-1 is symmetric: for column tile
// jt only i < (jt+1)*64 is visited, the strictly-lower part counted twice; the grid runs the
// costliest column tiles first and every (walker, column tile) leaves one partial sum that
// k_finalize adds in a fixed order.  Both operands are zero-padded to ndp = multiple of 64, so
// there is no edge handling.
// ---------------------------------------------------------------------------------------
#define Q_TW 128
#define Q_TJ 64
#define Q_KC 16
#define Q_ASTR 20   // doubles per A row in shared memory (16 + 4: conflict-free fragment loads)
#define Q_BSTR 68   // doubles per B row (64 + 4)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int NST>  // cp.async stages: 2 for batches that fill the machine (3 CTAs per SM), 4 for small ones (latency of the K loop)
__global__ void __launch_bounds__(256)
k_chi2_dmma(int64_t W, int ndp, int grp, const double* __restrict__ dvec, const double* __restrict__ Cp,
            double* __restrict__ partial) {
  extern __shared__ __align__(16) double qsm[];
  double (*As)[Q_TW][Q_ASTR] = reinterpret_cast<double (*)[Q_TW][Q_ASTR]>(qsm);
  double (*Bs)[Q_KC][Q_BSTR] = reinterpret_cast<double (*)[Q_KC][Q_BSTR]>(qsm + NST * Q_TW * Q_ASTR);
  double (*red)[Q_TW] = reinterpret_cast<double (*)[Q_TW]>(qsm + NST * Q_TW * Q_ASTR + NST * Q_KC * Q_BSTR);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  // Grid order: groups of `grp` walker tiles; inside a group the column tiles from the costliest down, the walker
  // tiles of the group fastest.  The CTAs in flight stream the same column tile of C^-1, and the residual rows of a
  // group (grp x 128 walkers x ndp doubles, 46 MB at grp = 64) stay in L2 while its column tiles re-read them, so they
  // come from HBM about once instead of once per column tile (2.16 GB -> see DESIGN.md; with one group = the whole
  // batch, the former order, every column-tile pass streams the residuals again).  The other extreme, column tiles
  // fastest, also reads them once but measured 16 % slower (1.28 ms against 1.10 ms).
  const int ntj = ndp / Q_TJ;
  const int nrb = (int)((W + Q_TW - 1) / Q_TW);
  const int per = grp * ntj;
  const int g = (int)blockIdx.x / per, rem = (int)blockIdx.x - g * per;
  const int gl = min(grp, nrb - g * grp);  // walker tiles in this group (the last one may be short)
  const int by = rem / gl;
  const int64_t w0 = (int64_t)(g * grp + rem - by * gl) * Q_TW;
  // one column tile per CTA, the costliest (last) tiles first.  (Measured against pairing tiles
  // (y, ntj-1-y) for equal work per CTA: the smaller units balance medium batches over the SMs, 0.135 ->
  // 0.088 ms at 4096 walkers and 0.199 -> 0.164 ms at 8192 on C4, and tie at 65 536.)
  {
    const int jt = ntj - 1 - by;
    const int j0 = jt * Q_TJ;
    const int nk = (jt + 1) * (Q_TJ / Q_KC);
    const int kdiag = jt * (Q_TJ / Q_KC);
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    auto load_chunk = [&](int kc, int buf) {
      const int i0 = kc * Q_KC;
#pragma unroll
      for (int t = 0; t < 4; ++t) {  // A: 128 rows x 8 pieces of 16 B
        int piece = tid + t * 256;
        int r = piece >> 3, cc = piece & 7;
        int64_t wr = w0 + r;
        if (wr >= W) wr = W - 1;
        cp_async16(&As[buf][r][cc * 2], dvec + wr * ndp + i0 + cc * 2);
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {  // B: 16 rows x 32 pieces
        int piece = tid + t * 256;
        int r = piece >> 5, cc = piece & 31;
        cp_async16(&Bs[buf][r][cc * 2], Cp + (size_t)(i0 + r) * ndp + j0 + cc * 2);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // NST - 1 chunks in flight; one (possibly empty) commit group per trip keeps the group count uniform
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) {
      if (p < nk)
        load_chunk(p, p);
      else
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kc = 0; kc < nk; ++kc) {
      const int buf = kc % NST;
      if (kc + NST - 1 < nk)
        load_chunk(kc + NST - 1, (kc + NST - 1) % NST);
      else
        asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(NST - 1) : "memory");
      __syncthreads();
      if (kc == kdiag) {
        // everything accumulated so far is strictly below the diagonal block: count it twice
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) { acc[a][b][0] *= 2.0; acc[a][b][1] *= 2.0; }
      }
#pragma unroll
      for (int ks = 0; ks < Q_KC / 4; ++ks) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = As[buf][32 * wm + 8 * a + (lane >> 2)][4 * ks + (lane & 3)];
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = Bs[buf][4 * ks + (lane & 3)][32 * wn + 8 * b + (lane >> 2)];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
      }
      __syncthreads();
    }
    // epilogue: sum_j Q[w][j] d[w][j] over this column tile
    double ps[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int64_t wr = w0 + 32 * wm + 8 * a + (lane >> 2);
      if (wr >= W) wr = W - 1;
      double sum = 0.0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double2 dv = *reinterpret_cast<const double2*>(dvec + wr * ndp + j0 + 32 * wn + 8 * b + 2 * (lane & 3));
        sum = fma(acc[a][b][0], dv.x, sum);
        sum = fma(acc[a][b][1], dv.y, sum);
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      ps[a] = sum;
    }
    if ((lane & 3) == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a) red[wn][32 * wm + 8 * a + (lane >> 2)] = ps[a];
    }
    __syncthreads();
    if (tid < Q_TW && w0 + tid < W) partial[(w0 + tid) * ntj + jt] = red[0][tid] + red[1][tid];  // thread r < 128 owns walker w0 + r
    __syncthreads();
  }
}

// Per-z chi2_z = d_z^T C_z^-1 d_z (likelihood.py:939-940) for the walker-per-lane stage 3: the residuals
// live in a z-aligned buffer dvz[w][z * tzw + j] (tzw = widest z block rounded up to 64, pads zero) and the
// per-z blocks in a zero-padded [nz][tzw][tzw] array, so every (walker tile, z) is a small dense
// [128 x tzw] . [tzw x tzw] product on the fp64 tensor cores with the row-dot fused, exactly as above but
// without the symmetric pairing (the blocks are only tzw / 64 column tiles wide).
__global__ void __launch_bounds__(256)
k_chi2z_dmma(int64_t W, int nz, int tzw, const double* __restrict__ dvz, const double* __restrict__ Cz,
             double* __restrict__ chi2z) {
  extern __shared__ __align__(16) double qsm[];
  double (*As)[Q_TW][Q_ASTR] = reinterpret_cast<double (*)[Q_TW][Q_ASTR]>(qsm);
  double (*Bs)[Q_KC][Q_BSTR] = reinterpret_cast<double (*)[Q_KC][Q_BSTR]>(qsm + 2 * Q_TW * Q_ASTR);
  double (*red)[Q_TW] = reinterpret_cast<double (*)[Q_TW]>(qsm + 2 * Q_TW * Q_ASTR + 2 * Q_KC * Q_BSTR);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int64_t w0 = (int64_t)blockIdx.x * Q_TW;
  const int z = blockIdx.y;
  const int ldz = nz * tzw;
  const double* dz = dvz + (size_t)z * tzw;           // column 0 of this z in a residual row
  const double* cz = Cz + (size_t)z * tzw * tzw;
  const int nk = tzw / Q_KC;
  double tot = 0.0;  // thread r < 128 owns walker w0 + r

  for (int jt = 0; jt < tzw / Q_TJ; ++jt) {
    const int j0 = jt * Q_TJ;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    auto load_chunk = [&](int kc, int buf) {
      const int i0 = kc * Q_KC;
#pragma unroll
      for (int t = 0; t < 4; ++t) {  // A: 128 rows x 8 pieces of 16 B
        int piece = tid + t * 256;
        int r = piece >> 3, cc = piece & 7;
        int64_t wr = w0 + r;
        if (wr >= W) wr = W - 1;
        cp_async16(&As[buf][r][cc * 2], dz + wr * ldz + i0 + cc * 2);
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {  // B: 16 rows x 32 pieces
        int piece = tid + t * 256;
        int r = piece >> 5, cc = piece & 31;
        cp_async16(&Bs[buf][r][cc * 2], cz + (size_t)(i0 + r) * tzw + j0 + cc * 2);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    load_chunk(0, 0);
    for (int kc = 0; kc < nk; ++kc) {
      const int buf = kc & 1;
      if (kc + 1 < nk) {
        load_chunk(kc + 1, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < Q_KC / 4; ++ks) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = As[buf][32 * wm + 8 * a + (lane >> 2)][4 * ks + (lane & 3)];
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = Bs[buf][4 * ks + (lane & 3)][32 * wn + 8 * b + (lane >> 2)];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
      }
      __syncthreads();
    }
    // epilogue: sum_j Q[w][j] d[w][j] over this column tile
    double ps[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int64_t wr = w0 + 32 * wm + 8 * a + (lane >> 2);
      if (wr >= W) wr = W - 1;
      double sum = 0.0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double2 dv = *reinterpret_cast<const double2*>(dz + wr * ldz + j0 + 32 * wn + 8 * b + 2 * (lane & 3));
        sum = fma(acc[a][b][0], dv.x, sum);
        sum = fma(acc[a][b][1], dv.y, sum);
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      ps[a] = sum;
    }
    if ((lane & 3) == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a) red[wn][32 * wm + 8 * a + (lane >> 2)] = ps[a];
    }
    __syncthreads();
    if (tid < Q_TW) tot += red[0][tid] + red[1][tid];
    __syncthreads();
  }
  if (tid < Q_TW && w0 + tid < W) chi2z[(w0 + tid) * nz + z] = tot;
}

int c1d_launch_chi2z(c1d_ctx* ctx, int64_t W, cudaStream_t s) {
  dim3 grid((unsigned)((W + Q_TW - 1) / Q_TW), (unsigned)ctx->cfg.nz);
  const size_t smem = (size_t)(2 * Q_TW * Q_ASTR + 2 * Q_KC * Q_BSTR + 2 * Q_TW) * sizeof(double);
  C1D_CUDA(ctx, cudaFuncSetAttribute(k_chi2z_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_chi2z_dmma<<<grid, 256, smem, s>>>(W, ctx->cfg.nz, ctx->tzw, ctx->ws.dvz, ctx->icovz_pad, ctx->ws.chi2z);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return C1D_OK;
}

int c1d_launch_chi2(c1d_ctx* ctx, int64_t W, uint32_t flags, cudaStream_t s) {
  const int ndp = ctx->d.ldd;
  const int ntj = ndp / Q_TJ;
  if (flags & C1D_FLAG_CHI2_I8) return c1d_chi2i8_launch(ctx, W, ntj, s);  // tcgen05 digit slices (c1d_chi2_i8.cu)
  const int nrb = (int)((W + Q_TW - 1) / Q_TW);
  int grp = 64;  // 8192 walkers: 46 MB of residuals at ndp = 704, well inside the 126 MB L2
  if (const char* e = getenv("C1D_CHI2_GROUP")) grp = atoi(e) > 0 ? atoi(e) : grp;
  if (grp > nrb) grp = nrb;
  // batches of at most one CTA per SM are bound by the latency of the K loop (up to 44 dependent cp.async round trips
  // with two stages): four stages there (C4, 1024 walkers: see DESIGN.md), the same arithmetic in the same order
  int nst = (nrb * ntj <= ctx->num_sms) ? 4 : 2;
  if (const char* e = getenv("C1D_CHI2_STAGES")) nst = atoi(e) == 4 ? 4 : 2;
  const size_t smem = (size_t)(nst * Q_TW * Q_ASTR + nst * Q_KC * Q_BSTR + 2 * Q_TW) * sizeof(double);
  // per launch, not once per process: the attribute is per device and a process may own several contexts
  if (nst == 4) {
    C1D_CUDA(ctx, cudaFuncSetAttribute(k_chi2_dmma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_chi2_dmma<4><<<(unsigned)(nrb * ntj), 256, smem, s>>>(W, ndp, grp, ctx->ws.dvec, ctx->icov_pad, ctx->ws.chi2p);
  } else {
    C1D_CUDA(ctx, cudaFuncSetAttribute(k_chi2_dmma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_chi2_dmma<2><<<(unsigned)(nrb * ntj), 256, smem, s>>>(W, ndp, grp, ctx->ws.dvec, ctx->icov_pad, ctx->ws.chi2p);
  }
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return C1D_OK;
}

// ---------------------------------------------------------------------------------------
// stage 5: regulate, add the prior, blob conventions (likelihood.py:966-1024)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_finalize(DevCtx c, int64_t W, const int* __restrict__ status, const double* __restrict__ lnprior,
           const double* __restrict__ chi2z, const double* __restrict__ chi2p, int ntj, int use_full,
           double* __restrict__ lnp, double* __restrict__ lnl_z, double* __restrict__ blobs,
           double* __restrict__ chi2_out, uint32_t flags) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const int st = status[w];
  const double ninf = __longlong_as_double(0xfff0000000000000LL);
  double loglike = ninf;
  if (st == ST_OK) {
    if (use_full) {
      double chi2 = 0.0;  // fixed summation order: column tiles (p, ntj-1-p) first, then over p
      for (int p = 0; 2 * p < ntj; ++p) {
        double v = chi2p[w * ntj + p];
        if (ntj - 1 - p != p) v += chi2p[w * ntj + ntj - 1 - p];
        chi2 += v;
      }
      loglike = -0.5 * chi2;
    } else {
      double sum = 0.0;
      for (int z = 0; z < c.nz; ++z) {
        bool on = !(flags & C1D_FLAG_ZMASK) || c.zmask[z];
        if (on) sum += -0.5 * chi2z[w * c.nz + z];
      }
      loglike = sum;
    }
    if (loglike != loglike) loglike = ninf;  // NaN -> null_out (likelihood.py:966-967)
  }
  const bool null_out = !(loglike > ninf);
  if (lnl_z) {
    for (int z = 0; z < c.nz; ++z) {
      bool on = !(flags & C1D_FLAG_ZMASK) || c.zmask[z];
      double v = 0.0;
      if (null_out) v = ninf;
      else if (on) v = -0.5 * chi2z[w * c.nz + z];
      lnl_z[w * c.nz + z] = v;
    }
  }
  if (chi2_out) chi2_out[w] = -2.0 * loglike;
  if (st == ST_OUT_OF_CUBE) {
    lnp[w] = c.min_log_like;  // no prior added (likelihood.py:989-994); blobs are NaN already
    return;
  }
  // regulate_log_like (likelihood.py:974-980), then the prior (:1019-1024)
  double reg = (loglike > c.min_log_like) ? loglike : c.min_log_like;
  lnp[w] = reg + lnprior[w];
  if (null_out && blobs) {
    for (int i = 0; i < 6; ++i) blobs[w * 6 + i] = 0.0;  // null blob (likelihood.py:833-836)
  }
}

int c1d_launch_finalize(c1d_ctx* ctx, int64_t W, double* lnp, double* lnl_z, double* blobs,
                        double* chi2, uint32_t flags, cudaStream_t s) {
  int use_full = ctx->cfg.has_full && !(flags & C1D_FLAG_ZMASK);
  int threads = 256;
  unsigned grid = (unsigned)((W + threads - 1) / threads);
  const int ntj = ctx->d.ldd / Q_TJ;
  k_finalize<<<grid, threads, 0, s>>>(ctx->d, W, ctx->ws.status, ctx->ws.lnprior, ctx->ws.chi2z,
                                      ctx->ws.chi2p, ntj, use_full, lnp, lnl_z, blobs, chi2, flags);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return C1D_OK;
}

__global__ void __launch_bounds__(256)
k_pack7(int64_t W, const double* __restrict__ lnp, const double* __restrict__ blobs, double* __restrict__ out7) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * 7) return;
  int64_t w = t / 7;
  int j = (int)(t % 7);
  out7[t] = (j == 0) ? lnp[w] : blobs[w * 6 + j - 1];
}

int c1d_launch_pack7(c1d_ctx* ctx, int64_t w0, int64_t W, cudaStream_t s) {
  unsigned grid = (unsigned)((W * 7 + 255) / 256);
  k_pack7<<<grid, 256, 0, s>>>(W, ctx->ws.lnp + w0, ctx->ws.blobs + w0 * 6, ctx->ws.out7 + w0 * 7);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return C1D_OK;
}
