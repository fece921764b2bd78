/*
 * cup1d_b200 — C ABI of the B200-native batched Lyman-alpha P1D log-likelihood.
 *
 * This is the drop-in boundary of the hot path.  The reference
 * is pure Python, so the binding a maintainer adds is a ctypes stub
 * (cup1d_b200/_abi.py; INTEGRATION.md shows it next to the reference call
 * sites).  Every entry point replaces a piece of
 *
 *   cup1d/likelihood/likelihood.py:1036-1047  Likelihood.log_prob_and_blobs
 *   cup1d/likelihood/likelihood.py:982-1024   Likelihood.compute_log_prob
 *   cup1d/likelihood/likelihood.py:822-972    Likelihood.get_log_like
 *   cup1d/likelihood/likelihood.py:722-785    Likelihood.get_p1d_kms
 *   cup1d/likelihood/lya_theory.py:610-814    Theory.get_p1d_kms
 *
 * evaluated for W walkers at once.  Plain pointers and sizes only; no torch or
 * C++ types cross this boundary.  All functions return 0 or a negative
 * c1d_status; numerical rejections (out of the unit cube, outside the hull or
 * the star-prior box, NaN log-like) are NOT errors: they produce exactly
 * min_log_like (-1e100) like the reference (likelihood.py:966-994).
 *
 * Threading: one context per device and per host thread; calls are
 * asynchronous with respect to the host and ordered on the CUDA stream given
 * (NULL = legacy default stream).  The *_host entry points synchronise.
 */
#ifndef CUP1D_B200_H
#define CUP1D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C1D_ABI_VERSION 8
#define C1D_MAX_LAYERS 12
#define C1D_MAX_EMU 8     /* emulator inputs */
#define C1D_MAX_COEF 8    /* polynomial coefficients the emulator returns */
#define C1D_NQ 20         /* z-dependent model quantities, see c1d_quantity */
#define C1D_NCOL 16       /* walker-independent k tables, see c1d_column */
#define C1D_NAMP 26       /* per-(walker,z) amplitudes handed from stage 1/2 to stage 3 */

typedef enum c1d_status {
  C1D_OK = 0,
  C1D_ERR_ARG = -1,      /* bad argument / inconsistent configuration */
  C1D_ERR_CUDA = -2,     /* CUDA runtime error, see c1d_last_error */
  C1D_ERR_STATE = -3,    /* context not finalised / field missing */
  C1D_ERR_ALLOC = -4
} c1d_status;

/* z-dependent quantities evaluated per walker from the coefficient families
 * (base_igm.py:245-319, base_contaminants.py:158-245) */
typedef enum c1d_quantity {
  C1D_Q_TAU = 0, C1D_Q_SIGT, C1D_Q_GAMMA, C1D_Q_KF,
  C1D_Q_F_SIIII, C1D_Q_S_SIIII, C1D_Q_F_SIII, C1D_Q_S_SIII,
  C1D_Q_F_SIIIA_SIIII, C1D_Q_F_SIIIB_SIIII,
  C1D_Q_F_SIADD, C1D_Q_S_SIADD,
  C1D_Q_HCD1, C1D_Q_HCD2, C1D_Q_HCD3, C1D_Q_HCD4, C1D_Q_HCD_CONST,
  C1D_Q_RES,
  C1D_Q_AGN,        /* f_AGN(z) = exp(poly(ln_AGN))           AGN_model.py:81-91 */
  C1D_Q_AGN_CONST   /* exp(constant coefficient): zero when it is <= the null value (:85-86) */
} c1d_quantity;

/* walker-independent tables over the packed model k grid (fine grid when
 * rebin_k != 1), one column each, evaluated on the host in float64 */
typedef enum c1d_column {
  C1D_COL_K = 0,     /* k [s/km] */
  C1D_COL_S,         /* emulator output scale: yscale * M(z) * IC(z,k) [* AGN shape] */
  C1D_COL_T,         /* polynomial abscissa: log10(k_Mpc) (NN) or k_Mpc/kmax (GP) */
  C1D_COL_T1,        /* off * cos(dv(SiIII-Lya) k)                     si_mult.py:273-279 */
  C1D_COL_T2A,       /* off * cos(dv(SiIIa-Lya) k)                     :286-288 */
  C1D_COL_T2BC,      /* off*cos(dv(SiIIb-Lya)k) + off*(rc3/rb3)cos(dv(SiIIc-Lya)k) */
  C1D_COL_T3A,       /* off * cos(dv(SiIII-SiIIa) k)                   :310-312 */
  C1D_COL_T3BC,      /* off*cos(dv(SiIII-SiIIb)k) + off*(rc3/rb3)cos(dv(SiIII-SiIIc)k) */
  C1D_COL_D1, C1D_COL_D2, C1D_COL_D3, C1D_COL_D4, /* HCD profiles  hcd_model_rogers_class.py:5-6 */
  C1D_COL_KT,        /* SiAdd ktot(k)                                  si_add.py:207-217 */
  C1D_COL_Q,         /* 2 R_z^2 k^2                                    resolution_class.py:99-108 */
  C1D_COL_AGN,       /* delta(z, k) of the AGN feedback template       AGN_model.py:99-116 */
  C1D_COL_SPARE      /* GP emulator with a k-table norm: c + t, interval and weight of k_Mpc in the norm's k grid
                        (np.interp(k_Mpc, input_norm["k_Mpc"], norm_imF(mF)), notebooks/figures_CM26/Fig3a.py:83-85) */
} c1d_column;

/* constant fields uploaded once with c1d_ctx_upload (float64 unless noted) */
typedef enum c1d_field {
  C1D_F_PMIN = 0,      /* [ndim]   LikelihoodParameter.min_value */
  C1D_F_PSCALE,        /* [ndim]   max_value - min_value */
  C1D_F_PRIOR_X0,      /* [ndim]   cube value of par.value            likelihood.py:1060 */
  C1D_F_PRIOR_W,       /* [ndim]   1/(2 sigma^2), 0 without priors    :1061-1063 */
  C1D_F_ZBASE,         /* [NQ*nz]  fixed part of each linear z-function */
  C1D_F_ZWGT,          /* [NQ*nz*ell_width] weights of the free coefficients */
  C1D_F_ZIDX,          /* int32 [NQ*nz*ell_width] parameter index, -1 = unused */
  C1D_F_ZNULL,         /* [NQ]     values with lin <= null are zeroed (-inf = never) */
  C1D_F_ZOTYPE,        /* int32 [NQ] 1 = exp, 0 = const */
  C1D_F_LINP,          /* [nz*3]   fiducial Delta2_p, n_p, alpha_p    lya_theory.py:309-320 */
  C1D_F_MZ,            /* [nz]     M(z) = H(z)/(1+z) */
  C1D_F_IGMFID,        /* [4*nz]   tau, sigT_kms, gamma, kF_kms fiducial histories */
  C1D_F_EMU_LO,        /* [n_emu]  emulator input normalisation */
  C1D_F_EMU_HI,        /* [n_emu] */
  C1D_F_HULL_ABC,      /* [n_hull_facets*3]  a_x, a_y, b              hull.py:9-10 */
  C1D_F_HULL_DIM,      /* int32 [n_hull_facets*2] emulator-input slots of the two axes */
  C1D_F_ZOFF_F,        /* int32 [nz+1] offsets into the packed model k grid */
  C1D_F_ZOFF_D,        /* int32 [nz+1] offsets into the packed data vector */
  C1D_F_TAB,           /* [NCOL*nf] column-major tables */
  C1D_F_RB_START,      /* int32 [nd] first fine cell of each data bin (relative to its z) */
  C1D_F_RB_WGT,        /* [nd*rebin_width] cover/sum_cover            likelihood.py:147-161 */
  C1D_F_DATA,          /* [nd]     measured P1D, concatenated over z  :957 */
  C1D_F_ICOV_BLOCKS,   /* [sum ndz^2] per-z inverse covariances       :933-951 */
  C1D_F_ICOV_FULL,     /* [nd*nd]  full inverse covariance (optional) :953-963 */
  C1D_F_NN_W,          /* NN weights, per layer in-major [K][N] */
  C1D_F_NN_B,          /* NN biases */
  C1D_F_GP_X,          /* [gp_ntrain*n_emu] training inputs (unit box) */
  C1D_F_GP_INVELL,     /* [n_emu] 1/lengthscale */
  C1D_F_GP_ALPHA,      /* [gp_ntrain*nc] K^-1 y */
  C1D_F_GP_YMEAN,      /* [nc] */
  C1D_F_GP_YSTD,       /* [nc] */
  C1D_F_GP_NORM,       /* [gp_nnorm] polynomial of ln norm(mF): k-independent norm (older form) */
  C1D_F_GP_NORM_MF,    /* [gp_nm] mean-flux nodes of the norm table, increasing */
  C1D_F_GP_NORM_TAB,   /* [gp_nm*gp_nkn] norm over (mF node, k node); last k column repeated */
  C1D_F_ZMASK,         /* int32 [nz] 1 = z-bin enters a zmask evaluation */
  C1D_F_COUNT
} c1d_field;

typedef struct c1d_config {
  int32_t abi_version;
  int32_t nz, ndim, n_emu, nd, nf;
  int32_t ell_width;          /* >= 1 */
  int32_t rebin_width;        /* 1 = no rebinning */
  int32_t emu_kind;           /* 0 = NN (MLP), 1 = GP */
  int32_t nc;                 /* coefficients returned by the emulator */
  int32_t n_layers;           /* NN */
  int32_t layer_dims[C1D_MAX_LAYERS + 1];
  int32_t gp_ntrain, gp_nnorm, gp_imf;
  int32_t metal_model;        /* 0 = SiMult, 1 = SiVid */
  int32_t apply_resolution;
  int32_t has_full;           /* full inverse covariance present */
  int32_t has_gauss;          /* Gaussian priors present */
  int32_t n_hull_facets;      /* 0 = hull prior off */
  int32_t idx_As, idx_ns, idx_nrun; /* parameter indices or -1 */
  int32_t emu_code[C1D_MAX_EMU]; /* 0 Delta2_p 1 n_p 2 alpha_p 3 mF 4 sigT_Mpc 5 gamma 6 kF_Mpc 7 lambda_P */
  int32_t device;             /* CUDA device ordinal */
  int32_t uniform_k;          /* 1 = the model k grid of every z is evenly spaced (linspace, likelihood.py:90-94) */
  int32_t has_agn;            /* AGN feedback term present (off in the reference: model_contaminants.py:138-154) */
  int32_t gp_nm, gp_nkn;      /* GP norm table: mean-flux nodes, k nodes + 1 (0, 0 = k-independent norm) */
  int32_t hcd_gate;           /* 1 = HCD_Model_McDonald2005 (hcd_model_McDonald2005.py:71-92): the amplitude quantity HCD_damp1 = A_damp(z)
                                 counts only while the gate quantity HCD_damp2 (the constant coefficient against the null value) is > 0;
                                 column D1 holds 0.018 + 1/(15000 k - 8.9).  0 = Rogers / BOSS (was reserved: zero-filled by older callers) */
  int32_t reserved_i[2];
  double As_fid, ns_fid, nrun_fid;
  double L_emu;               /* ln(kp_emu / ks)   lya_theory.py:299 */
  double L_star;              /* ln(kp_star / ks)  lya_theory.py:565 */
  double blob_fid[6];
  double star_lo[3], star_hi[3]; /* +-inf when open */
  double min_log_like;
  double hull_tol;
  double nn_slope;
  double gp_sigma2;
  /* SiMult constants (si_mult.py:181-183, 267-271) */
  double rb3;                 /* rat[SiIIb_SiIII] */
  double ra3_over_rb3;        /* rat[SiIIa_SiIII]/rat[SiIIb_SiIII] */
  double c0_o1, c0_o2, c0_o3c; /* switches entering C0: off(SiIII_Lya), off(SiIIa_Lya), off(SiIIb_Lya)+c^2 off(SiIIc_Lya) */
  double reserved_d[8];
} c1d_config;

typedef struct c1d_ctx c1d_ctx;

/* evaluation flags */
#define C1D_FLAG_ZMASK 1u     /* use the uploaded C1D_F_ZMASK: per-z chi2 of the selected bins only (likelihood.py:844-863, 953-954) */
#define C1D_FLAG_NO_HULL 2u   /* apply_hull=False (likelihood.py:730) */
#define C1D_FLAG_NO_CUBE 8u   /* no unit-cube test: Theory.get_p1d_kms evaluates any parameter values (lya_theory.py:610-814; the test lives in Likelihood, likelihood.py:989-994) */
#define C1D_FLAG_EMU_TC 4u    /* emulator layers on tcgen05 tensor cores in 3xTF32 (float32-class accuracy: misses the 1e-4 log-likelihood tolerance) */
#define C1D_FLAG_EMU_I8 16u   /* emulator layers on tcgen05 tensor cores as int8 digit slices with exact int32 accumulation and float64
                                 recombination (float64-class accuracy); without either flag the layers run on the fp64 tensor cores (DMMA) */
#define C1D_FLAG_CHI2_I8 32u  /* dense chi2 = d^T C^-1 d (likelihood.py:956-960) on tcgen05 tensor cores as five int8 digit planes per operand,
                                 exact int32 accumulation, float64 recombination (chi2 within 5e-11 relative of the float64 evaluation);
                                 without it the quadratic form runs on the fp64 tensor cores (DMMA).  Available up to 704 data points:
                                 see c1d_emulator_paths */

int c1d_abi_version(void);

/* replaces the construction half of Likelihood.__init__ / Theory.__init__ for
 * the device: sizes only, no data yet */
int c1d_ctx_create(const c1d_config* cfg, c1d_ctx** out);

/* copies one constant field to the device (host pointer, byte count checked) */
int c1d_ctx_upload(c1d_ctx* ctx, int field_id, const void* host_ptr, size_t bytes);

/* checks that every field the configuration needs has been uploaded */
int c1d_ctx_finalize(c1d_ctx* ctx);

/* Likelihood.log_prob_and_blobs for W walkers (likelihood.py:1036-1047).
 *   x_dev      [W, ndim] row-major unit-cube coordinates (device)
 *   lnp_dev    [W]       log-posterior
 *   lnl_z_dev  [W, nz]   -0.5*chi2_z (log_like_all, :948) or NULL
 *   blobs_dev  [W, 6]    Delta2_star n_star alpha_star f_star g_star H0 or NULL
 *   chi2_dev   [W]       -2*log_like incl. +inf for rejected walkers (:787-798) or NULL */
int c1d_loglike_batch(c1d_ctx* ctx, const double* x_dev, int64_t W, double* lnp_dev,
                      double* lnl_z_dev, double* blobs_dev, double* chi2_dev,
                      uint32_t flags, void* cuda_stream);

/* Likelihood.get_p1d_kms for W walkers (likelihood.py:722-785): the rebinned
 * model, concatenated over z like data.full_Pk_kms; rows of rejected walkers
 * are NaN (the reference returns None).
 *   p_dev      [W, nd]
 *   emu_dev    [W, nz, n_emu] emulator inputs (return_emu_params) or NULL */
int c1d_p1d_batch(c1d_ctx* ctx, const double* x_dev, int64_t W, double* p_dev,
                  double* emu_dev, uint32_t flags, void* cuda_stream);

/* stage 2 alone (emulator.emulate_p1d_Mpc, lya_theory.py:690) for profiling:
 *   emu_in_dev [nrows, C1D_MAX_EMU] raw emulator inputs, coef_dev [nrows, C1D_MAX_COEF] */
int c1d_emulator_only(c1d_ctx* ctx, const double* emu_in_dev, int64_t nrows,
                      double* coef_dev, uint32_t flags, void* cuda_stream);

/* host-buffer variant of c1d_loglike_batch: the call a Python caller with
 * NumPy arrays makes.  Copies x to the device, evaluates, copies
 * out_host[W,7] = (lnP, blob[6]) back and synchronises.  Batches of 32 768
 * walkers or more go up in two pieces on the context's own copy stream, so
 * that the upload of the second piece and the download of the first run under
 * the kernels (pinned host buffers make the copies asynchronous; pageable
 * ones work, without the overlap).  The kernels run on the legacy default stream. */
int c1d_loglike_batch_host(c1d_ctx* ctx, const double* x_host, int64_t W,
                           double* out_host, uint32_t flags);

/* Affine-invariant stretch move on the device (what emcee does around the reference's
 * per-walker calls, fitter.py:187-235).  The ensemble is split in two fixed
 * halves; each iteration updates both halves (proposal kernel, batched likelihood of the
 * proposed half, accept kernel), everything enqueued on cuda_stream without host round trips.
 *   x_dev [W, ndim], lnp_dev [W], blobs_dev [W, 6]: state, in/out (W even).  init != 0
 *     evaluates lnp/blobs of the initial positions first.
 *   step0, seed: Philox4x32-10 counters, so (seed, step) reproduces a chain.
 *   a: stretch scale (emcee default 2).  thin: every thin-th state is stored in
 *   chain_dev [nsteps/thin, W, ndim], lnp_chain_dev [nsteps/thin, W], blobs_chain_dev
 *   [nsteps/thin, W, 6] (each may be NULL).  naccept_dev [W] int64 counters or NULL. */
int c1d_sampler_run(c1d_ctx* ctx, double* x_dev, double* lnp_dev, double* blobs_dev, int64_t W,
                    int64_t nsteps, int64_t step0, uint64_t seed, double a, int init, int thin,
                    double* chain_dev, double* lnp_chain_dev, double* blobs_chain_dev,
                    long long* naccept_dev, uint32_t flags, void* cuda_stream);

/* the same for n_ens INDEPENDENT ensembles advanced together (the reference's production layout: one emcee ensemble of
 * max(nwalkers, 2 ndim + 1) walkers per MPI rank, 128 ranks, fitter.py:69-72, 216-272; launch_data_sampler.sh:69): rows
 * e*wn .. e*wn + wn - 1 of x_dev (wn = W / n_ens, even) are ensemble e; stretch-move partners come from the other half
 * of the same ensemble.  The random numbers of ensemble e are keyed on its global number ens0 + e, so that an ensemble
 * gives the same chain alone, in a batch, or on another rank (ensembles shard over GPUs with no communication;
 * chains concatenate on the walker axis like fitter.py:262-268). */
int c1d_sampler_run_ens(c1d_ctx* ctx, double* x_dev, double* lnp_dev, double* blobs_dev, int64_t W, int64_t n_ens,
                        int64_t ens0, int64_t nsteps, int64_t step0, uint64_t seed, double a, int init, int thin,
                        double* chain_dev, double* lnp_chain_dev, double* blobs_chain_dev,
                        long long* naccept_dev, uint32_t flags, void* cuda_stream);

/* one half-step of the same move restricted to proposals s0 .. s0 + ns - 1 of the active halves (propose, batched
 * likelihood, accept).  With it an ensemble that SPANS several GPUs advances with one all-gather of the updated rows per
 * half-step (cup1d_b200/dist.py): every rank holds the whole state, updates its share of the active half and the ranks
 * exchange their shares; the random numbers depend on the walker, not on the rank, so the chain is that of one GPU. */
int c1d_sampler_half_step(c1d_ctx* ctx, double* x_dev, double* lnp_dev, double* blobs_dev, int64_t W, int64_t n_ens,
                          int64_t ens0, int64_t step, int half, uint64_t seed, double a, int64_t s0, int64_t ns,
                          long long* naccept_dev, uint32_t flags, void* cuda_stream);

/* the sampler's random numbers on the host (tests replay a chain with them):
 * out2 = two uniforms in (0,1) of counter (step, half, walker, which); walker = (ens0 + e) * wn / 2 + index in the half */
int c1d_philox_u01(uint64_t seed, int64_t step, int half, int64_t walker, int which, double* out2);

/* arithmetic available to a finalised context: bit 0 = float64, C1D_FLAG_EMU_TC, C1D_FLAG_EMU_I8
 * (the int8 digit-slice kernel needs an MLP whose layers wider than 192 are followed by one of at most 64),
 * C1D_FLAG_CHI2_I8 (dense inverse covariance of at most 704 data points) */
int c1d_emulator_paths(const c1d_ctx* ctx);

/* number of kernels launched by this context since creation */
int64_t c1d_launch_count(const c1d_ctx* ctx);

/* number of c1d_loglike_batch calls served by replaying a captured CUDA graph: batches of at most 8192 walkers
 * (C1D_GRAPH=<walkers> in the environment, 0 = off) whose shape, flags and buffers repeat the previous call's are
 * captured once and replayed; their kernels count in c1d_launch_count as if launched one by one */
int64_t c1d_graph_replays(const c1d_ctx* ctx);

/* per-stage device time [ms], measured with CUDA events on the caller's stream when enabled
 * (0 = off, default).  The events are read back lazily, so the evaluation calls stay
 * asynchronous: c1d_get_timing waits for the evaluations recorded so far, returns the times
 * summed over them since the previous c1d_get_timing / c1d_set_timing and resets the sums. */
int c1d_set_timing(c1d_ctx* ctx, int enable);
int c1d_get_timing(c1d_ctx* ctx, double* ms_out /* [8] */);

/* diagnostic: D[128, N] = tf32(A[128, 32]) . tf32(B[N, 32])^T through one tcgen05.mma
 * chain (host pointers); mode 0 = A from shared memory, 1 = A from tensor memory */
int c1d_tc_selftest(int device, const float* A_host, const float* B_host, float* D_host, int N, int mode);

int c1d_ctx_destroy(c1d_ctx* ctx);
const char* c1d_last_error(const c1d_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CUP1D_B200_H */
