// Internal declarations shared by the translation units of libcup1d_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "cup1d_b200.h"

// slots of the per-(walker, z) amplitude record handed from stage 1/2 to stage 3
enum {
  A_COEF0 = 0,   // .. A_COEF0 + C1D_MAX_COEF - 1 : emulator polynomial coefficients
  A_S3 = 8,      // s_Lya_SiIII
  A_S2 = 9,      // s_Lya_SiII
  A_M0 = 10,     // 1 + C0                                  (si_mult.py:267-271)
  A_C1 = 11,     // 2 a3            | SiVid: f3/(1-mF)
  A_C2A = 12,    // 2 a2 r          | SiVid: f2/(1-mF)
  A_C2B = 13,    // 2 a2
  A_C3A = 14,    // 2 a3 a2 Gmm r
  A_C3B = 15,    // 2 a3 a2 Gmm
  A_H0 = 16,     // 1 + HCD_const
  A_HCD1 = 17,   // .. A_HCD1 + 3
  A_AA = 21,     // f_SiIIa_SiIIb^2
  A_SA2 = 22,    // s_SiIIa_SiIIb^2
  A_RES = 23,    // R_coeff(z) (0 when the resolution term is not applied)
  A_AGN = 24     // f_AGN(z) (0 when absent or nulled)
};

// walker status written by stage 1
// combined with atomicMax by the (walker, z) threads of stage 1: out-of-cube wins over a rejection
enum { ST_OK = 0, ST_REJECTED = 1, ST_OUT_OF_CUBE = 2 };

#define C1D_TRING 64

struct DevCtx {
  int nz, ndim, n_emu, nd, nf, ell_width, rebin_width, emu_kind, nc, n_layers;
  int ldd;  // row stride of the residual buffer = nd rounded up to a multiple of 64 (zero padded)
  int layer_dims[C1D_MAX_LAYERS + 1];
  int gp_ntrain, gp_nnorm, gp_imf, gp_nm, gp_nkn;
  int metal_model, apply_resolution, has_full, has_gauss, n_hull_facets, uniform_k, has_agn, hcd_gate;
  int idx_As, idx_ns, idx_nrun;
  int emu_code[C1D_MAX_EMU];
  double As_fid, ns_fid, nrun_fid, L_emu, L_star;
  double blob_fid[6], star_lo[3], star_hi[3];
  double min_log_like, hull_tol, nn_slope, gp_sigma2;
  double rb3, ra3_over_rb3, c0_o1, c0_o2, c0_o3c;
  const double *pmin, *pscale, *prior_x0, *prior_w;
  const double *zbase, *zwgt, *znull;
  const int *zidx, *zotype;
  const double *linp, *mz, *igmfid, *emu_lo, *emu_hi;
  const double* hull_abc;
  const int* hull_dim;
  const int *zoff_f, *zoff_d;
  const double* tab;
  const int* rb_start;
  const double *rb_wgt, *data, *icov_blocks, *icov_full;
  const double *nn_w, *nn_b;
  const double *gp_x, *gp_invell, *gp_alpha, *gp_ymean, *gp_ystd, *gp_norm;
  const double *gp_nmf, *gp_ntab;  // norm table of the GP emulator: mean-flux nodes, [gp_nm][gp_nkn] values
  const double* gp_emu_in;         // workspace emulator inputs [z][w][C1D_MAX_EMU] (stage 3 reads the mean flux), set per launch
  const int* zmask;
};

struct Workspace {
  int64_t cap;     // walkers the buffers hold
  int* status;     // [cap]
  double* lnprior; // [cap]
  double* blobs;   // [cap*6]
  double* emu_in;  // [nz*cap*C1D_MAX_EMU]
  double* amp;     // [nz*cap*C1D_NAMP]
  double* dvec;    // [cap*ldd], pad columns stay zero
  double* chi2z;   // [cap*nz]
  double* dvz;     // [cap*nz*tzw] residuals with every z block at a 64-aligned offset (lazily allocated), pads zero
  double* chi2p;   // [cap*(ldd/64)] partial dense chi2 per 64-column tile
  double* lnp;     // [cap]   (host-variant staging on the device)
  double* xdev;    // [cap*ndim]
  double* xt;      // [ndim*cap] transposed cube coordinates of a chunk (stage 1 of large batches)
  double* out7;    // [cap*7] packed (lnP, blob[6]) rows for the host variant
  double* factors; // [cap] (n-1) ln z of the stretch-move proposals
  double* blobs_q; // [cap*6] blobs of the proposals
};

// one captured evaluation (c1d_loglike_batch on a small batch): the key is everything the kernel parameters depend on
#define C1D_NGRAPH 8
struct c1d_graph_entry {
  int64_t W;
  uint32_t flags;
  const double* x;
  double *lnp, *lnl_z, *blobs, *chi2;
  uint64_t envh;     // hash of the diagnostic environment switches the launchers read
  void* exec;        // cudaGraphExec_t
  int64_t nlaunch;   // kernels in the graph
  uint64_t stamp;    // last use (least recently used entry is replaced)
};

struct c1d_ctx {
  c1d_config cfg;
  void* dev[C1D_F_COUNT];
  size_t bytes[C1D_F_COUNT];
  DevCtx d;
  Workspace ws;
  int finalized;
  int num_sms;
  int nfz_max, ndz_max;  // largest z-bin of the model grid / data vector
  int64_t launches;
  int timing;
  // stage timing: a ring of event sets (six events per evaluated chunk) read back lazily, so that enabling
  // it does not make the evaluation calls synchronous
  cudaEvent_t evr[C1D_TRING][6];
  int ev_pending;  // sets recorded and not yet read
  // host-buffer entry point: copy stream and events that order its uploads / downloads against the kernels
  cudaStream_t copy_stream;
  cudaEvent_t ev_up[2], ev_done[2];
  // CUDA graphs of small-batch evaluations (c1d_abi.cu)
  cudaStream_t graph_stream;
  c1d_graph_entry graphs[C1D_NGRAPH], graph_seen[C1D_NGRAPH];  // captured; keys met once (captured on their second call)
  uint64_t graph_clock;
  int64_t graph_replays;
  double last_ms[8];
  double* icov_pad;  // [ldd*ldd] zero-padded copy of the dense inverse covariance
  double* icovz_pad; // [nz][tzw*tzw] zero-padded per-z inverse covariance blocks, tzw = ndz_max rounded up to 64
  int tzw;
  // tensor-core emulator path (c1d_mlp_tc.cu)
  void* tc_state;
  // int8 digit-slice tensor-core emulator path (c1d_mlp_i8.cu)
  void* i8_state;
  // int8 digit-slice tensor-core dense chi2 (c1d_chi2_i8.cu)
  void* chi2i8_state;
  // fp64 tensor-core emulator path (c1d_mlp_f64.cu)
  void* mlp64_state;
  // walker-per-lane stage 3 (c1d_p1d_lanes.cu): per-fine-point rebinning maps, null when not applicable
  void* p1dw_state;
  char err[512];
};

#define C1D_CUDA(ctx, call)                                                              \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, \
               cudaGetErrorString(e_));                                                  \
      return C1D_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// launchers (each returns a c1d_status and bumps ctx->launches)
int c1d_launch_stage1(c1d_ctx* ctx, const double* x, int64_t W, double* blobs_out, uint32_t flags,
                      cudaStream_t s);
int c1d_launch_emulator(c1d_ctx* ctx, const double* emu_in, int64_t nrows, double* out,
                        int out_stride, uint32_t flags, cudaStream_t s);
int c1d_launch_stage3(c1d_ctx* ctx, int64_t W, double* p_out, int want_dvec, int want_chi2z,
                      cudaStream_t s);
int c1d_launch_chi2(c1d_ctx* ctx, int64_t W, uint32_t flags, cudaStream_t s);
int c1d_launch_finalize(c1d_ctx* ctx, int64_t W, double* lnp, double* lnl_z, double* blobs,
                        double* chi2, uint32_t flags, cudaStream_t s);

void c1d_graphs_clear(c1d_ctx* ctx);  // drop the captured small-batch evaluations (their parameters went stale)
int c1d_launch_pack7(c1d_ctx* ctx, int64_t w0, int64_t W, cudaStream_t s);  // rows [w0, w0 + W) of ws.lnp / ws.blobs -> ws.out7

// fp64 (DMMA) MLP: padded weight image + slab plan, built at finalize
int c1d_mlp64_prepare(c1d_ctx* ctx);
int c1d_mlp64_launch(c1d_ctx* ctx, const double* emu_in, int64_t nrows, double* out, int out_stride,
                     cudaStream_t s);
void c1d_mlp64_destroy(c1d_ctx* ctx);

// walker-per-lane stage 3: prepare builds the two-bin rebinning maps; launch returns 1 when it ran,
// 0 when the caller must use the warp-per-walker kernel
int c1d_p1dw_prepare(c1d_ctx* ctx);
int c1d_p1dw_launch(c1d_ctx* ctx, int64_t W, double* p_out, int want_dvec, int want_chi2z, cudaStream_t s);
int c1d_launch_chi2z(c1d_ctx* ctx, int64_t W, cudaStream_t s);  // per-z chi2 of the z-aligned residuals ws.dvz
void c1d_p1dw_destroy(c1d_ctx* ctx);

// tensor-core MLP (optional translation unit)
int c1d_tc_prepare(c1d_ctx* ctx);
int c1d_tc_launch(c1d_ctx* ctx, const double* emu_in, int64_t nrows, double* out, int out_stride,
                  cudaStream_t s);
void c1d_tc_destroy(c1d_ctx* ctx);

// int8 digit-slice tensor-core MLP (c1d_mlp_i8.cu): prepare leaves the path unavailable for unsupported shapes
int c1d_i8_prepare(c1d_ctx* ctx);
int c1d_i8_available(const c1d_ctx* ctx);
int c1d_i8_launch(c1d_ctx* ctx, const double* emu_in, int64_t nrows, double* out, int out_stride,
                  cudaStream_t s);
void c1d_i8_destroy(c1d_ctx* ctx);

// int8 digit-slice tensor-core dense chi2 (c1d_chi2_i8.cu): prepare digit-slices the folded inverse covariance (called
// whenever C1D_F_ICOV_FULL changes) and leaves the path unavailable beyond 704 data points; launch fills ws.chi2p
int c1d_chi2i8_prepare(c1d_ctx* ctx);
int c1d_chi2i8_available(const c1d_ctx* ctx);
int c1d_chi2i8_launch(c1d_ctx* ctx, int64_t W, int ntj, cudaStream_t s);
void c1d_chi2i8_destroy(c1d_ctx* ctx);
