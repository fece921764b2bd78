// C ABI of libcup1d_b200.so: context management and the batched drivers that chain the
// stage kernels on the caller's stream.  See include/cup1d_b200.h for the contract.
#include <new>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include "c1d_internal.h"

static const char* k_no_ctx = "null context";
static const int64_t CHUNK_MAX = 1 << 17;  // walkers evaluated per pass (workspace bound)

extern "C" int c1d_abi_version(void) { return C1D_ABI_VERSION; }

extern "C" const char* c1d_last_error(const c1d_ctx* ctx) { return ctx ? ctx->err : k_no_ctx; }

extern "C" int64_t c1d_launch_count(const c1d_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int64_t c1d_graph_replays(const c1d_ctx* ctx) { return ctx ? ctx->graph_replays : 0; }

static size_t field_elem_size(int f) {
  switch (f) {
    case C1D_F_ZIDX: case C1D_F_ZOTYPE: case C1D_F_HULL_DIM: case C1D_F_ZOFF_F: case C1D_F_ZOFF_D:
    case C1D_F_RB_START: case C1D_F_ZMASK:
      return sizeof(int32_t);
    default:
      return sizeof(double);
  }
}

// expected element count of a field, or -1 when the field is optional/unsized
static int64_t field_count(const c1d_config& c, int f, int64_t icov_blocks_n, int64_t nn_w_n, int64_t nn_b_n) {
  switch (f) {
    case C1D_F_PMIN: case C1D_F_PSCALE: case C1D_F_PRIOR_X0: case C1D_F_PRIOR_W: return c.ndim;
    case C1D_F_ZBASE: return (int64_t)C1D_NQ * c.nz;
    case C1D_F_ZWGT: case C1D_F_ZIDX: return (int64_t)C1D_NQ * c.nz * c.ell_width;
    case C1D_F_ZNULL: case C1D_F_ZOTYPE: return C1D_NQ;
    case C1D_F_LINP: return 3 * c.nz;
    case C1D_F_MZ: return c.nz;
    case C1D_F_IGMFID: return 4 * c.nz;
    case C1D_F_EMU_LO: case C1D_F_EMU_HI: return c.n_emu;
    case C1D_F_HULL_ABC: return 3 * (int64_t)c.n_hull_facets;
    case C1D_F_HULL_DIM: return 2 * (int64_t)c.n_hull_facets;
    case C1D_F_ZOFF_F: case C1D_F_ZOFF_D: return c.nz + 1;
    case C1D_F_TAB: return (int64_t)C1D_NCOL * c.nf;
    case C1D_F_RB_START: return c.nd;
    case C1D_F_RB_WGT: return (int64_t)c.nd * c.rebin_width;
    case C1D_F_DATA: return c.nd;
    case C1D_F_ICOV_BLOCKS: return icov_blocks_n;
    case C1D_F_ICOV_FULL: return c.has_full ? (int64_t)c.nd * c.nd : 0;
    case C1D_F_NN_W: return c.emu_kind == 0 ? nn_w_n : 0;
    case C1D_F_NN_B: return c.emu_kind == 0 ? nn_b_n : 0;
    case C1D_F_GP_X: return c.emu_kind == 1 ? (int64_t)c.gp_ntrain * c.n_emu : 0;
    case C1D_F_GP_INVELL: return c.emu_kind == 1 ? c.n_emu : 0;
    case C1D_F_GP_ALPHA: return c.emu_kind == 1 ? (int64_t)c.gp_ntrain * c.nc : 0;
    case C1D_F_GP_YMEAN: case C1D_F_GP_YSTD: return c.emu_kind == 1 ? c.nc : 0;
    case C1D_F_GP_NORM: return c.emu_kind == 1 ? c.gp_nnorm : 0;
    case C1D_F_GP_NORM_MF: return c.emu_kind == 1 ? c.gp_nm : 0;
    case C1D_F_GP_NORM_TAB: return c.emu_kind == 1 ? (int64_t)c.gp_nm * c.gp_nkn : 0;
    case C1D_F_ZMASK: return c.nz;
    default: return -1;
  }
}

extern "C" int c1d_ctx_create(const c1d_config* cfg, c1d_ctx** out) {
  if (!cfg || !out) return C1D_ERR_ARG;
  *out = nullptr;
  if (cfg->abi_version != C1D_ABI_VERSION) return C1D_ERR_ARG;
  c1d_ctx* ctx = new (std::nothrow) c1d_ctx();
  if (!ctx) return C1D_ERR_ALLOC;
  memset(ctx, 0, sizeof(*ctx));
  ctx->cfg = *cfg;
  *out = ctx;  // returned even on failure so that the caller can read the message
  const c1d_config& c = ctx->cfg;
  bool ok = c.nz > 0 && c.ndim > 0 && c.n_emu > 0 && c.n_emu <= C1D_MAX_EMU && c.nd > 0 && c.nf > 0 &&
            c.ell_width >= 1 && c.rebin_width >= 1 && c.nc >= 1 && c.nc <= C1D_MAX_COEF &&
            (c.emu_kind == 0 || c.emu_kind == 1);
  if (ok && c.emu_kind == 0) {
    ok = c.n_layers >= 1 && c.n_layers <= C1D_MAX_LAYERS && c.layer_dims[0] == c.n_emu &&
         c.layer_dims[c.n_layers] == c.nc;
    for (int l = 0; ok && l <= c.n_layers; ++l) ok = c.layer_dims[l] >= 1 && c.layer_dims[l] <= 256;
  }
  if (ok && c.emu_kind == 1) ok = c.gp_ntrain > 0 && c.gp_imf >= 0 && c.gp_imf < c.n_emu && c.gp_nnorm >= 0 &&
                                  ((c.gp_nm == 0 && c.gp_nkn == 0) || (c.gp_nm >= 2 && c.gp_nkn >= 3 && c.gp_nnorm == 0));
  if (!ok) {
    snprintf(ctx->err, sizeof(ctx->err), "inconsistent configuration (sizes / emulator layout; NN layers must be <= 256 wide)");
    return C1D_ERR_ARG;
  }
  C1D_CUDA(ctx, cudaSetDevice(c.device));
  cudaDeviceProp prop;
  C1D_CUDA(ctx, cudaGetDeviceProperties(&prop, c.device));
  ctx->num_sms = prop.multiProcessorCount;
  if (prop.major < 10) {
    snprintf(ctx->err, sizeof(ctx->err), "device %s is sm_%d%d; this library is built for sm_100a only",
             prop.name, prop.major, prop.minor);
    return C1D_ERR_CUDA;
  }
  for (int r = 0; r < C1D_TRING; ++r)
    for (int i = 0; i < 6; ++i) C1D_CUDA(ctx, cudaEventCreate(&ctx->evr[r][i]));
  C1D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  C1D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->graph_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    C1D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_up[i], cudaEventDisableTiming));
    C1D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
  }
  return C1D_OK;
}

// zero-padded copies of the inverse covariances for the tensor-core chi2 kernels; the allocations are kept across
// re-uploads of same-sized fields (set_icov on a live context)
static int refresh_icov_pad(c1d_ctx* ctx) {
  const c1d_config& c = ctx->cfg;
  if (!c.has_full) {
    if (ctx->icov_pad) { cudaFree(ctx->icov_pad); ctx->icov_pad = nullptr; }
    c1d_chi2i8_destroy(ctx);
    return C1D_OK;
  }
  const size_t ldd = ctx->d.ldd;
  if (!ctx->icov_pad) {
    C1D_CUDA(ctx, cudaMalloc(&ctx->icov_pad, ldd * ldd * sizeof(double)));
    C1D_CUDA(ctx, cudaMemset(ctx->icov_pad, 0, ldd * ldd * sizeof(double)));
  }
  C1D_CUDA(ctx, cudaMemcpy2D(ctx->icov_pad, ldd * sizeof(double), ctx->dev[C1D_F_ICOV_FULL], (size_t)c.nd * sizeof(double),
                             (size_t)c.nd * sizeof(double), c.nd, cudaMemcpyDeviceToDevice));
  return c1d_chi2i8_prepare(ctx);  // digit planes of the folded matrix for the tcgen05 chi2
}

static int refresh_icovz_pad(c1d_ctx* ctx, const int32_t* zd) {
  const c1d_config& c = ctx->cfg;
  const int tzw = (ctx->ndz_max + 63) / 64 * 64;
  if (ctx->icovz_pad && ctx->tzw != tzw) { cudaFree(ctx->icovz_pad); ctx->icovz_pad = nullptr; }
  ctx->tzw = tzw;
  if (!ctx->icovz_pad) {
    C1D_CUDA(ctx, cudaMalloc(&ctx->icovz_pad, (size_t)c.nz * tzw * tzw * sizeof(double)));
    C1D_CUDA(ctx, cudaMemset(ctx->icovz_pad, 0, (size_t)c.nz * tzw * tzw * sizeof(double)));
  }
  size_t off = 0;
  for (int z = 0; z < c.nz; ++z) {
    const int ndz = zd[z + 1] - zd[z];
    C1D_CUDA(ctx, cudaMemcpy2D(ctx->icovz_pad + (size_t)z * tzw * tzw, (size_t)tzw * sizeof(double),
                               (const double*)ctx->dev[C1D_F_ICOV_BLOCKS] + off, (size_t)ndz * sizeof(double),
                               (size_t)ndz * sizeof(double), ndz, cudaMemcpyDeviceToDevice));
    off += (size_t)ndz * ndz;
  }
  return C1D_OK;
}

extern "C" int c1d_ctx_upload(c1d_ctx* ctx, int field_id, const void* host_ptr, size_t bytes) {
  if (!ctx) return C1D_ERR_ARG;
  if (field_id < 0 || field_id >= C1D_F_COUNT || (!host_ptr && bytes)) {
    snprintf(ctx->err, sizeof(ctx->err), "bad field id %d", field_id);
    return C1D_ERR_ARG;
  }
  C1D_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  const bool same_size = ctx->dev[field_id] && ctx->bytes[field_id] == bytes;
  if (ctx->dev[field_id] && !same_size) {
    C1D_CUDA(ctx, cudaFree(ctx->dev[field_id]));
    ctx->dev[field_id] = nullptr;
  }
  if (bytes == 0) { ctx->bytes[field_id] = 0; return C1D_OK; }
  if (!ctx->dev[field_id]) C1D_CUDA(ctx, cudaMalloc(&ctx->dev[field_id], bytes));
  // re-uploads (new data vector, zmask) must not race with work in flight
  C1D_CUDA(ctx, cudaDeviceSynchronize());
  c1d_graphs_clear(ctx);  // the field may move, and derived tables are rebuilt
  C1D_CUDA(ctx, cudaMemcpy(ctx->dev[field_id], host_ptr, bytes, cudaMemcpyHostToDevice));
  ctx->bytes[field_id] = bytes;
  if (!ctx->finalized) return C1D_OK;
  // a live context: same-sized re-uploads of the data vector, the z mask and the inverse covariances (set_data,
  // set_icov, zmask evaluations) refresh only what depends on them; anything else re-runs the whole finalisation
  int rc = C1D_OK;
  if (same_size && (field_id == C1D_F_DATA || field_id == C1D_F_ZMASK)) {
    rc = C1D_OK;
  } else if (same_size && field_id == C1D_F_ICOV_FULL) {
    rc = refresh_icov_pad(ctx);
  } else if (same_size && field_id == C1D_F_ICOV_BLOCKS) {
    int32_t zd[1024];
    C1D_CUDA(ctx, cudaMemcpy(zd, ctx->dev[C1D_F_ZOFF_D], (ctx->cfg.nz + 1) * 4, cudaMemcpyDeviceToHost));
    rc = refresh_icovz_pad(ctx, zd);
  } else {
    return c1d_ctx_finalize(ctx);
  }
  if (rc) return rc;
  // the copies above ran on the legacy default stream: order them before evaluations on any (non-blocking) stream
  C1D_CUDA(ctx, cudaDeviceSynchronize());
  return C1D_OK;
}

extern "C" int c1d_ctx_finalize(c1d_ctx* ctx) {
  if (!ctx) return C1D_ERR_ARG;
  c1d_config& c = ctx->cfg;
  C1D_CUDA(ctx, cudaSetDevice(c.device));
  c1d_graphs_clear(ctx);
  // z offsets are needed on the host to size the per-z blocks
  int32_t zf[1024], zd[1024];
  if (c.nz + 1 > 1024) { snprintf(ctx->err, sizeof(ctx->err), "too many z bins"); return C1D_ERR_ARG; }
  if (!ctx->dev[C1D_F_ZOFF_F] || !ctx->dev[C1D_F_ZOFF_D] ||
      ctx->bytes[C1D_F_ZOFF_F] != (size_t)(c.nz + 1) * 4 || ctx->bytes[C1D_F_ZOFF_D] != (size_t)(c.nz + 1) * 4) {
    snprintf(ctx->err, sizeof(ctx->err), "z offsets not uploaded");
    return C1D_ERR_STATE;
  }
  C1D_CUDA(ctx, cudaMemcpy(zf, ctx->dev[C1D_F_ZOFF_F], (c.nz + 1) * 4, cudaMemcpyDeviceToHost));
  C1D_CUDA(ctx, cudaMemcpy(zd, ctx->dev[C1D_F_ZOFF_D], (c.nz + 1) * 4, cudaMemcpyDeviceToHost));
  int64_t blocks_n = 0;
  int nfz_max = 0, ndz_max = 0;
  for (int z = 0; z < c.nz; ++z) {
    int nfz = zf[z + 1] - zf[z], ndz = zd[z + 1] - zd[z];
    if (nfz <= 0 || ndz <= 0) { snprintf(ctx->err, sizeof(ctx->err), "empty z bin %d", z); return C1D_ERR_ARG; }
    blocks_n += (int64_t)ndz * ndz;
    if (nfz > nfz_max) nfz_max = nfz;
    if (ndz > ndz_max) ndz_max = ndz;
  }
  if (zf[0] != 0 || zd[0] != 0 || zf[c.nz] != c.nf || zd[c.nz] != c.nd) {
    snprintf(ctx->err, sizeof(ctx->err), "z offsets do not cover the grids");
    return C1D_ERR_ARG;
  }
  ctx->nfz_max = nfz_max;
  ctx->ndz_max = ndz_max;
  int64_t nn_w_n = 0, nn_b_n = 0;
  if (c.emu_kind == 0)
    for (int l = 0; l < c.n_layers; ++l) {
      nn_w_n += (int64_t)c.layer_dims[l] * c.layer_dims[l + 1];
      nn_b_n += c.layer_dims[l + 1];
    }
  for (int f = 0; f < C1D_F_COUNT; ++f) {
    int64_t n = field_count(c, f, blocks_n, nn_w_n, nn_b_n);
    if (f == C1D_F_ZMASK && !ctx->dev[f]) continue;  // optional until a zmask evaluation
    if (n < 0) continue;
    if ((size_t)n * field_elem_size(f) != ctx->bytes[f]) {
      snprintf(ctx->err, sizeof(ctx->err), "field %d: expected %lld elements, have %zu bytes", f,
               (long long)n, ctx->bytes[f]);
      return C1D_ERR_STATE;
    }
  }
  DevCtx& d = ctx->d;
  d.nz = c.nz; d.ndim = c.ndim; d.n_emu = c.n_emu; d.nd = c.nd; d.nf = c.nf;
  d.ldd = (c.nd + 63) / 64 * 64;
  d.ell_width = c.ell_width; d.rebin_width = c.rebin_width; d.emu_kind = c.emu_kind; d.nc = c.nc;
  d.n_layers = c.n_layers;
  for (int l = 0; l <= C1D_MAX_LAYERS; ++l) d.layer_dims[l] = c.layer_dims[l];
  d.gp_ntrain = c.gp_ntrain; d.gp_nnorm = c.gp_nnorm; d.gp_imf = c.gp_imf;
  d.gp_nm = c.emu_kind == 1 ? c.gp_nm : 0; d.gp_nkn = c.emu_kind == 1 ? c.gp_nkn : 0;
  d.metal_model = c.metal_model; d.apply_resolution = c.apply_resolution; d.has_full = c.has_full;
  d.has_gauss = c.has_gauss; d.n_hull_facets = c.n_hull_facets; d.uniform_k = c.uniform_k; d.has_agn = c.has_agn; d.hcd_gate = c.hcd_gate;
  d.idx_As = c.idx_As; d.idx_ns = c.idx_ns; d.idx_nrun = c.idx_nrun;
  for (int j = 0; j < C1D_MAX_EMU; ++j) d.emu_code[j] = c.emu_code[j];
  d.As_fid = c.As_fid; d.ns_fid = c.ns_fid; d.nrun_fid = c.nrun_fid; d.L_emu = c.L_emu; d.L_star = c.L_star;
  for (int i = 0; i < 6; ++i) d.blob_fid[i] = c.blob_fid[i];
  for (int i = 0; i < 3; ++i) { d.star_lo[i] = c.star_lo[i]; d.star_hi[i] = c.star_hi[i]; }
  d.min_log_like = c.min_log_like; d.hull_tol = c.hull_tol; d.nn_slope = c.nn_slope; d.gp_sigma2 = c.gp_sigma2;
  d.rb3 = c.rb3; d.ra3_over_rb3 = c.ra3_over_rb3; d.c0_o1 = c.c0_o1; d.c0_o2 = c.c0_o2; d.c0_o3c = c.c0_o3c;
#define DP(name, F) d.name = (const double*)ctx->dev[F]
#define IP(name, F) d.name = (const int*)ctx->dev[F]
  DP(pmin, C1D_F_PMIN); DP(pscale, C1D_F_PSCALE); DP(prior_x0, C1D_F_PRIOR_X0); DP(prior_w, C1D_F_PRIOR_W);
  DP(zbase, C1D_F_ZBASE); DP(zwgt, C1D_F_ZWGT); DP(znull, C1D_F_ZNULL); IP(zidx, C1D_F_ZIDX); IP(zotype, C1D_F_ZOTYPE);
  DP(linp, C1D_F_LINP); DP(mz, C1D_F_MZ); DP(igmfid, C1D_F_IGMFID); DP(emu_lo, C1D_F_EMU_LO); DP(emu_hi, C1D_F_EMU_HI);
  DP(hull_abc, C1D_F_HULL_ABC); IP(hull_dim, C1D_F_HULL_DIM); IP(zoff_f, C1D_F_ZOFF_F); IP(zoff_d, C1D_F_ZOFF_D);
  DP(tab, C1D_F_TAB); IP(rb_start, C1D_F_RB_START); DP(rb_wgt, C1D_F_RB_WGT); DP(data, C1D_F_DATA);
  DP(icov_blocks, C1D_F_ICOV_BLOCKS); DP(icov_full, C1D_F_ICOV_FULL); DP(nn_w, C1D_F_NN_W); DP(nn_b, C1D_F_NN_B);
  DP(gp_x, C1D_F_GP_X); DP(gp_invell, C1D_F_GP_INVELL); DP(gp_alpha, C1D_F_GP_ALPHA);
  DP(gp_ymean, C1D_F_GP_YMEAN); DP(gp_ystd, C1D_F_GP_YSTD); DP(gp_norm, C1D_F_GP_NORM); IP(zmask, C1D_F_ZMASK);
  DP(gp_nmf, C1D_F_GP_NORM_MF); DP(gp_ntab, C1D_F_GP_NORM_TAB);
  d.gp_emu_in = nullptr;
#undef DP
#undef IP
  if (ctx->icov_pad) { cudaFree(ctx->icov_pad); ctx->icov_pad = nullptr; }  // ldd may have changed
  { int rc = refresh_icov_pad(ctx); if (rc) return rc; }
  { int rc = refresh_icovz_pad(ctx, zd); if (rc) return rc; }
  if (ctx->ws.dvz) { cudaFree(ctx->ws.dvz); ctx->ws.dvz = nullptr; }  // its width may have changed
  ctx->finalized = 1;
  {
    int rc = c1d_p1dw_prepare(ctx);  // two-bin rebinning maps for the walker-per-lane stage 3
    if (rc != C1D_OK) return rc;
  }
  if (c.emu_kind == 0) {
    int rc = c1d_mlp64_prepare(ctx);  // padded fp64 weight image + slab plan
    if (rc != C1D_OK) return rc;
    rc = c1d_tc_prepare(ctx);  // builds the tensor-core weight images (no-op when unsupported)
    if (rc != C1D_OK) return rc;
    rc = c1d_i8_prepare(ctx);  // int8 digit planes of the weights (leaves the path unavailable for unsupported shapes)
    if (rc != C1D_OK) return rc;
  }
  // the memsets and copies of the finalisation ran on the legacy default stream; evaluations may be enqueued on
  // non-blocking streams, which do not synchronise with it
  C1D_CUDA(ctx, cudaDeviceSynchronize());
  return C1D_OK;
}

static void ws_free(c1d_ctx* ctx) {
  Workspace& w = ctx->ws;
  cudaFree(w.status); cudaFree(w.lnprior); cudaFree(w.blobs); cudaFree(w.emu_in); cudaFree(w.amp);
  cudaFree(w.dvec); cudaFree(w.dvz); cudaFree(w.chi2z); cudaFree(w.chi2p); cudaFree(w.lnp); cudaFree(w.xdev); cudaFree(w.xt); cudaFree(w.out7); cudaFree(w.factors); cudaFree(w.blobs_q);
  memset(&w, 0, sizeof(w));
}

static int ws_reserve(c1d_ctx* ctx, int64_t W) {
  Workspace& w = ctx->ws;
  if (W <= w.cap) return C1D_OK;
  const c1d_config& c = ctx->cfg;
  C1D_CUDA(ctx, cudaDeviceSynchronize());
  c1d_graphs_clear(ctx);  // captured launches point into the old workspace
  ws_free(ctx);
  int64_t cap = W;
  C1D_CUDA(ctx, cudaMalloc(&w.status, cap * sizeof(int)));
  C1D_CUDA(ctx, cudaMalloc(&w.lnprior, cap * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.blobs, cap * 6 * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.emu_in, (size_t)cap * c.nz * C1D_MAX_EMU * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.amp, (size_t)cap * c.nz * C1D_NAMP * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.dvec, (size_t)cap * ctx->d.ldd * sizeof(double)));
  C1D_CUDA(ctx, cudaMemset(w.dvec, 0, (size_t)cap * ctx->d.ldd * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.chi2z, (size_t)cap * c.nz * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.chi2p, (size_t)cap * (ctx->d.ldd / 64) * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.lnp, cap * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.xdev, (size_t)cap * c.ndim * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.xt, (size_t)cap * c.ndim * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.out7, (size_t)cap * 7 * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.factors, cap * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&w.blobs_q, (size_t)cap * 6 * sizeof(double)));
  // rows of rejected walkers are never written by stage 1: keep them finite
  C1D_CUDA(ctx, cudaMemset(w.emu_in, 0, (size_t)cap * c.nz * C1D_MAX_EMU * sizeof(double)));
  C1D_CUDA(ctx, cudaMemset(w.amp, 0, (size_t)cap * c.nz * C1D_NAMP * sizeof(double)));
  // the memsets ran on the legacy default stream: finish them before the caller's (possibly non-blocking) stream
  // writes into the buffers
  C1D_CUDA(ctx, cudaDeviceSynchronize());
  w.cap = cap;
  return C1D_OK;
}

// the sampler keeps proposals in the workspace across likelihood calls: size it once up front
int c1d_sampler_reserve(c1d_ctx* ctx, int64_t W) {
  if (!ctx->finalized) {
    snprintf(ctx->err, sizeof(ctx->err), "context not finalised");
    return C1D_ERR_STATE;
  }
  if (W > CHUNK_MAX) {
    snprintf(ctx->err, sizeof(ctx->err), "device sampler supports at most %lld walkers per half-ensemble", (long long)CHUNK_MAX);
    return C1D_ERR_ARG;
  }
  C1D_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  return ws_reserve(ctx, W);
}

static int check_ready(c1d_ctx* ctx, uint32_t flags) {
  if (!ctx) return C1D_ERR_ARG;
  if (!ctx->finalized) {
    snprintf(ctx->err, sizeof(ctx->err), "context not finalised");
    return C1D_ERR_STATE;
  }
  if ((flags & C1D_FLAG_ZMASK) && !ctx->dev[C1D_F_ZMASK]) {
    snprintf(ctx->err, sizeof(ctx->err), "C1D_FLAG_ZMASK without an uploaded C1D_F_ZMASK");
    return C1D_ERR_STATE;
  }
  return C1D_OK;
}

// reads the pending event sets (waits for the last one) and adds their stage times to last_ms
static int timing_drain(c1d_ctx* ctx) {
  for (int r = 0; r < ctx->ev_pending; ++r) {
    C1D_CUDA(ctx, cudaEventSynchronize(ctx->evr[r][5]));
    for (int i = 0; i < 5; ++i) {
      float ms = 0.f;
      C1D_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->evr[r][i], ctx->evr[r][i + 1]));
      ctx->last_ms[i] += ms;
    }
  }
  ctx->ev_pending = 0;
  return C1D_OK;
}

#define TICK(i)                                                                          \
  do {                                                                                   \
    if (ctx->timing) C1D_CUDA(ctx, cudaEventRecord(ctx->evr[ctx->ev_pending][i], s));    \
  } while (0)

// NVTX range around the launches of one stage (tracing): visible in nsys / ncu --nvtx timelines
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// the five stages of one chunk on stream s (timed per stage when timing is on)
static int run_chain(c1d_ctx* ctx, const double* x, int64_t n, double* lnp, double* lnl_z, double* blobs_out, double* chi2,
                     uint32_t flags, cudaStream_t s, bool timed) {
  const c1d_config& c = ctx->cfg;
  const int use_full = c.has_full && !(flags & C1D_FLAG_ZMASK);
  const int want_chi2z = !use_full || lnl_z != nullptr;
  double* blobs = blobs_out ? blobs_out : ctx->ws.blobs;
  int rc;
#define TICK_T(i) do { if (timed) TICK(i); } while (0)
  NvtxRange chunk_range("c1d_loglike_batch chunk");
  TICK_T(0);
  {
    NvtxRange r("c1d stage 1: parameters, priors, emulator inputs");
    if ((rc = c1d_launch_stage1(ctx, x, n, blobs, flags, s))) return rc;
  }
  TICK_T(1);
  {
    NvtxRange r("c1d stage 2: emulator");
    if ((rc = c1d_launch_emulator(ctx, ctx->ws.emu_in, n * c.nz, ctx->ws.amp, C1D_NAMP, flags, s))) return rc;
  }
  TICK_T(2);
  {
    NvtxRange r("c1d stage 3: contaminants, rebinning, residual");
    if ((rc = c1d_launch_stage3(ctx, n, nullptr, use_full, want_chi2z, s))) return rc;
  }
  TICK_T(3);
  {
    NvtxRange r("c1d stage 4: dense chi2");
    if (use_full && (rc = c1d_launch_chi2(ctx, n, flags, s))) return rc;
  }
  TICK_T(4);
  {
    NvtxRange r("c1d stage 5: finalize");
    if ((rc = c1d_launch_finalize(ctx, n, lnp, lnl_z, blobs, chi2, flags, s))) return rc;
  }
  TICK_T(5);
#undef TICK_T
  return C1D_OK;
}

// ---- CUDA graphs for small batches -------------------------------------------------------------------------------
// Below a few thousand walkers the five dependent launches and the gaps between them are a large part of a call
// (C1, 64 walkers: kernels ~40 us of a ~100 us call).  A call that repeats an earlier one's shape, flags
// and buffers - the sampler's half-steps, the minimiser's simplex evaluations, an emcee loop over torch's caching
// allocator - is captured once (on an internal stream: the caller's may be the legacy default stream, which cannot be
// captured) and replayed with one cudaGraphLaunch.  The first occurrence of a key runs normally, so every lazy
// allocation and function attribute of that configuration exists before the capture.  Anything that changes what the
// kernel parameters point to (workspace growth, uploads, re-finalisation) drops the cache.  C1D_GRAPH=0 disables it.
void c1d_graphs_clear(c1d_ctx* ctx) {
  bool any = false;
  for (int i = 0; i < C1D_NGRAPH; ++i) any = any || ctx->graphs[i].exec;
  if (any) cudaDeviceSynchronize();  // no replay in flight when its executable graph goes
  for (int i = 0; i < C1D_NGRAPH; ++i) {
    if (ctx->graphs[i].exec) cudaGraphExecDestroy((cudaGraphExec_t)ctx->graphs[i].exec);
    memset(&ctx->graphs[i], 0, sizeof(ctx->graphs[i]));
  }
  memset(ctx->graph_seen, 0, sizeof(ctx->graph_seen));
}

static bool graph_key_eq(const c1d_graph_entry& a, const c1d_graph_entry& b) {
  return a.W == b.W && a.flags == b.flags && a.x == b.x && a.lnp == b.lnp && a.lnl_z == b.lnl_z && a.blobs == b.blobs &&
         a.chi2 == b.chi2 && a.envh == b.envh;
}

// the launchers read diagnostic switches from the environment at every launch (kernel form, tile counts): a graph
// captured under one setting must not serve a call made under another (the tests flip them between calls)
static uint64_t graph_env_hash() {
  static const char* names[] = {"C1D_P1D_LAYOUT", "C1D_P1D_PARTS", "C1D_GP_KERNEL", "C1D_CHI2_GROUP", "C1D_CHI2_STAGES",
                                "C1D_MLP64_G", "C1D_TC_CLUSTER", "C1D_CHI2_SPLIT"};
  uint64_t h = 1469598103934665603ull;  // FNV-1a
  for (const char* n : names) {
    const char* v = getenv(n);
    for (const char* q = v ? v : ""; *q; ++q) h = (h ^ (uint8_t)*q) * 1099511628211ull;
    h = (h ^ 0xffu) * 1099511628211ull;
  }
  return h;
}

static int64_t graph_max_walkers() {
  static int64_t v = -1;
  if (v < 0) {
    const char* e = getenv("C1D_GRAPH");
    v = e ? atoll(e) : 1;
    if (v == 1) v = 8192;  // default: batches of at most 8192 walkers
  }
  return v;
}

extern "C" int c1d_loglike_batch(c1d_ctx* ctx, const double* x_dev, int64_t W, double* lnp_dev,
                                 double* lnl_z_dev, double* blobs_dev, double* chi2_dev,
                                 uint32_t flags, void* cuda_stream) {
  int rc = check_ready(ctx, flags);
  if (rc) return rc;
  if (W < 0 || (W > 0 && (!x_dev || !lnp_dev))) {
    snprintf(ctx->err, sizeof(ctx->err), "bad arguments to c1d_loglike_batch");
    return C1D_ERR_ARG;
  }
  if (W == 0) return C1D_OK;
  const c1d_config& c = ctx->cfg;
  C1D_CUDA(ctx, cudaSetDevice(c.device));
  cudaStream_t s = (cudaStream_t)cuda_stream;
  int64_t chunk = W < CHUNK_MAX ? W : CHUNK_MAX;
  rc = ws_reserve(ctx, chunk);
  if (rc) return rc;
  if (W <= graph_max_walkers() && !ctx->timing) {
    c1d_graph_entry key;
    memset(&key, 0, sizeof(key));
    key.W = W; key.flags = flags; key.x = x_dev; key.lnp = lnp_dev; key.lnl_z = lnl_z_dev; key.blobs = blobs_dev; key.chi2 = chi2_dev; key.envh = graph_env_hash();
    for (int i = 0; i < C1D_NGRAPH; ++i)
      if (ctx->graphs[i].exec && graph_key_eq(ctx->graphs[i], key)) {
        C1D_CUDA(ctx, cudaGraphLaunch((cudaGraphExec_t)ctx->graphs[i].exec, s));
        ctx->graphs[i].stamp = ++ctx->graph_clock;
        ctx->launches += ctx->graphs[i].nlaunch;
        ctx->graph_replays++;
        return C1D_OK;
      }
    int seen = -1;
    for (int i = 0; i < C1D_NGRAPH; ++i)
      if (ctx->graph_seen[i].stamp && graph_key_eq(ctx->graph_seen[i], key)) seen = i;
    if (seen >= 0) {
      // second call with this key (callers that alternate between two output buffers repeat every other call): capture it
      memset(&ctx->graph_seen[seen], 0, sizeof(ctx->graph_seen[seen]));
      int slot = 0;
      for (int i = 1; i < C1D_NGRAPH; ++i)
        if (ctx->graphs[i].stamp < ctx->graphs[slot].stamp) slot = i;
      if (ctx->graphs[slot].exec) cudaGraphExecDestroy((cudaGraphExec_t)ctx->graphs[slot].exec);
      memset(&ctx->graphs[slot], 0, sizeof(ctx->graphs[slot]));
      cudaGraph_t graph = nullptr;
      const int64_t l0 = ctx->launches;
      C1D_CUDA(ctx, cudaStreamBeginCapture(ctx->graph_stream, cudaStreamCaptureModeThreadLocal));
      rc = run_chain(ctx, x_dev, W, lnp_dev, lnl_z_dev, blobs_dev, chi2_dev, flags, ctx->graph_stream, false);
      cudaError_t ce = cudaStreamEndCapture(ctx->graph_stream, &graph);
      const int64_t nl = ctx->launches - l0;
      ctx->launches = l0;
      cudaGraphExec_t exec = nullptr;
      if (rc == C1D_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        cudaGraphDestroy(graph);
        key.exec = exec; key.nlaunch = nl; key.stamp = ++ctx->graph_clock;
        ctx->graphs[slot] = key;
        C1D_CUDA(ctx, cudaGraphLaunch(exec, s));
        ctx->launches += nl;
        ctx->graph_replays++;
        return C1D_OK;
      }
      // capture not possible (a launcher that allocates or synchronises): clear the error state and run normally
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (rc) return rc;
    } else {
      int slot = 0;  // remember the key: the oldest remembered one makes room
      for (int i = 1; i < C1D_NGRAPH; ++i)
        if (ctx->graph_seen[i].stamp < ctx->graph_seen[slot].stamp) slot = i;
      ctx->graph_seen[slot] = key;
      ctx->graph_seen[slot].stamp = ++ctx->graph_clock;
    }
  }
  for (int64_t w0 = 0; w0 < W; w0 += chunk) {
    int64_t n = (W - w0 < chunk) ? (W - w0) : chunk;
    if (ctx->timing && ctx->ev_pending == C1D_TRING && (rc = timing_drain(ctx))) return rc;
    if ((rc = run_chain(ctx, x_dev + w0 * c.ndim, n, lnp_dev + w0, lnl_z_dev ? lnl_z_dev + w0 * c.nz : nullptr,
                        blobs_dev ? blobs_dev + w0 * 6 : nullptr, chi2_dev ? chi2_dev + w0 : nullptr, flags, s, ctx->timing != 0)))
      return rc;
    if (ctx->timing) ctx->ev_pending++;
  }
  return C1D_OK;
}

extern "C" int c1d_p1d_batch(c1d_ctx* ctx, const double* x_dev, int64_t W, double* p_dev,
                             double* emu_dev, uint32_t flags, void* cuda_stream) {
  int rc = check_ready(ctx, flags);
  if (rc) return rc;
  if (W < 0 || (W > 0 && (!x_dev || !p_dev))) {
    snprintf(ctx->err, sizeof(ctx->err), "bad arguments to c1d_p1d_batch");
    return C1D_ERR_ARG;
  }
  if (W == 0) return C1D_OK;
  const c1d_config& c = ctx->cfg;
  C1D_CUDA(ctx, cudaSetDevice(c.device));
  cudaStream_t s = (cudaStream_t)cuda_stream;
  int64_t chunk = W < CHUNK_MAX ? W : CHUNK_MAX;
  rc = ws_reserve(ctx, chunk);
  if (rc) return rc;
  for (int64_t w0 = 0; w0 < W; w0 += chunk) {
    int64_t n = (W - w0 < chunk) ? (W - w0) : chunk;
    if ((rc = c1d_launch_stage1(ctx, x_dev + w0 * c.ndim, n, ctx->ws.blobs, flags, s))) return rc;
    if ((rc = c1d_launch_emulator(ctx, ctx->ws.emu_in, n * c.nz, ctx->ws.amp, C1D_NAMP, flags, s))) return rc;
    if ((rc = c1d_launch_stage3(ctx, n, p_dev + w0 * c.nd, 0, 0, s))) return rc;
    if (emu_dev) {
      // workspace layout is [z][w][8]; hand back [w][z][n_emu]
      for (int z = 0; z < c.nz; ++z)
        C1D_CUDA(ctx, cudaMemcpy2DAsync(emu_dev + (w0 * c.nz + z) * c.n_emu, (size_t)c.nz * c.n_emu * sizeof(double),
                                        ctx->ws.emu_in + (size_t)z * n * C1D_MAX_EMU, C1D_MAX_EMU * sizeof(double),
                                        c.n_emu * sizeof(double), n, cudaMemcpyDeviceToDevice, s));
    }
  }
  return C1D_OK;
}

extern "C" int c1d_emulator_only(c1d_ctx* ctx, const double* emu_in_dev, int64_t nrows, double* coef_dev,
                                 uint32_t flags, void* cuda_stream) {
  int rc = check_ready(ctx, 0);
  if (rc) return rc;
  if (nrows < 0 || (nrows > 0 && (!emu_in_dev || !coef_dev))) return C1D_ERR_ARG;
  C1D_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  return c1d_launch_emulator(ctx, emu_in_dev, nrows, coef_dev, C1D_MAX_COEF, flags, (cudaStream_t)cuda_stream);
}

extern "C" int c1d_loglike_batch_host(c1d_ctx* ctx, const double* x_host, int64_t W, double* out_host,
                                      uint32_t flags) {
  int rc = check_ready(ctx, flags);
  if (rc) return rc;
  if (W < 0 || (W > 0 && (!x_host || !out_host))) return C1D_ERR_ARG;
  if (W == 0) return C1D_OK;
  const c1d_config& c = ctx->cfg;
  C1D_CUDA(ctx, cudaSetDevice(c.device));
  cudaStream_t s = 0;
  int64_t chunk = W < CHUNK_MAX ? W : CHUNK_MAX;
  rc = ws_reserve(ctx, chunk);
  if (rc) return rc;
  cudaStream_t sc = ctx->copy_stream;
  for (int64_t w0 = 0; w0 < W; w0 += chunk) {
    int64_t n = (W - w0 < chunk) ? (W - w0) : chunk;
    // A large chunk goes up in two pieces (a quarter, then the rest) on the copy stream: the kernels of the
    // first piece hide the upload of the second, and the rows of the first piece come down under the kernels
    // of the second.  Small chunks keep one piece (their kernels are launch-bound, not copy-bound).
    const int np = (n >= 32768) ? 2 : 1;
    const int64_t n0 = (np == 2) ? ((n / 4 + 1023) / 1024) * 1024 : n;
    const int64_t off[3] = {0, n0, n};
    for (int p = 0; p < np; ++p) {
      C1D_CUDA(ctx, cudaMemcpyAsync(ctx->ws.xdev + off[p] * c.ndim, x_host + (w0 + off[p]) * c.ndim,
                                    (size_t)(off[p + 1] - off[p]) * c.ndim * sizeof(double), cudaMemcpyHostToDevice, sc));
      C1D_CUDA(ctx, cudaEventRecord(ctx->ev_up[p], sc));
    }
    for (int p = 0; p < np; ++p) {
      const int64_t m = off[p + 1] - off[p];
      C1D_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_up[p], 0));
      rc = c1d_loglike_batch(ctx, ctx->ws.xdev + off[p] * c.ndim, m, ctx->ws.lnp + off[p], nullptr, ctx->ws.blobs + off[p] * 6, nullptr,
                             flags, s);
      if (rc) return rc;
      // out_host rows are (lnP, blob[6]): pack on the device, one contiguous copy back
      if ((rc = c1d_launch_pack7(ctx, off[p], m, s))) return rc;
      C1D_CUDA(ctx, cudaEventRecord(ctx->ev_done[p], s));
      C1D_CUDA(ctx, cudaStreamWaitEvent(sc, ctx->ev_done[p], 0));
      C1D_CUDA(ctx, cudaMemcpyAsync(out_host + (w0 + off[p]) * 7, ctx->ws.out7 + off[p] * 7, (size_t)m * 7 * sizeof(double),
                                    cudaMemcpyDeviceToHost, sc));
    }
    C1D_CUDA(ctx, cudaStreamSynchronize(sc));
    C1D_CUDA(ctx, cudaStreamSynchronize(s));
  }
  return C1D_OK;
}

extern "C" int c1d_emulator_paths(const c1d_ctx* ctx) {
  if (!ctx || !ctx->finalized) return 0;
  int m = 0;
  if (ctx->cfg.emu_kind == 0) {
    m |= 1;                                // float64 (DMMA)
    if (ctx->tc_state) m |= (int)C1D_FLAG_EMU_TC;
    if (c1d_i8_available(ctx)) m |= (int)C1D_FLAG_EMU_I8;
  } else {
    m |= 1;
  }
  if (c1d_chi2i8_available(ctx)) m |= (int)C1D_FLAG_CHI2_I8;
  return m;
}

extern "C" int c1d_set_timing(c1d_ctx* ctx, int enable) {
  if (!ctx) return C1D_ERR_ARG;
  C1D_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  int rc = timing_drain(ctx);
  if (rc) return rc;
  for (int i = 0; i < 8; ++i) ctx->last_ms[i] = 0.0;
  ctx->timing = enable ? 1 : 0;
  return C1D_OK;
}

extern "C" int c1d_get_timing(c1d_ctx* ctx, double* ms_out) {
  if (!ctx || !ms_out) return C1D_ERR_ARG;
  C1D_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
  int rc = timing_drain(ctx);
  if (rc) return rc;
  for (int i = 0; i < 8; ++i) {
    ms_out[i] = ctx->last_ms[i];
    ctx->last_ms[i] = 0.0;
  }
  return C1D_OK;
}

extern "C" int c1d_ctx_destroy(c1d_ctx* ctx) {
  if (!ctx) return C1D_OK;
  cudaSetDevice(ctx->cfg.device);
  cudaDeviceSynchronize();
  c1d_tc_destroy(ctx);
  c1d_i8_destroy(ctx);
  c1d_chi2i8_destroy(ctx);
  c1d_mlp64_destroy(ctx);
  c1d_p1dw_destroy(ctx);
  c1d_graphs_clear(ctx);
  ws_free(ctx);
  if (ctx->icov_pad) cudaFree(ctx->icov_pad);
  if (ctx->icovz_pad) cudaFree(ctx->icovz_pad);
  for (int f = 0; f < C1D_F_COUNT; ++f)
    if (ctx->dev[f]) cudaFree(ctx->dev[f]);
  for (int r = 0; r < C1D_TRING; ++r)
    for (int i = 0; i < 6; ++i)
      if (ctx->evr[r][i]) cudaEventDestroy(ctx->evr[r][i]);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_up[i]) cudaEventDestroy(ctx->ev_up[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->graph_stream) cudaStreamDestroy(ctx->graph_stream);
  delete ctx;
  return C1D_OK;
}
