// Affine-invariant ensemble sampler on the device: the stretch move of
// Goodman & Weare (2010) in its two-half parallel form — what emcee's StretchMove does around
// the reference's per-walker likelihood calls (cup1d/likelihood/fitter.py:187-235) — with the
// proposal, the batched likelihood of the proposed half and the accept/reject all enqueued on
// one stream and no host round trip per step.  Random numbers: Philox4x32-10, counter
// (step, half, walker, 0), key = seed, so a chain is reproducible from (seed, step).
#include "c1d_internal.h"

namespace {

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in (0, 1) from 32 random bits + 21 bits of a second word (53-bit mantissa)
__host__ __device__ inline double u01(uint32_t a, uint32_t b) {
  const uint64_t v = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);
  return ((double)(v & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

// one warp per walker of the active half of its ensemble: q = c_j - (c_j - x_s) z with c_j drawn from the other
// half of the SAME ensemble.  The batch holds W / wn independent ensembles of wn walkers (rows e wn .. e wn + wn - 1,
// first half then second half); s runs over the active halves of all of them, the random numbers are keyed on the
// global half-index (ens0 + e) nh + sl, so an ensemble draws the same numbers alone, in a batch or on another rank.
__global__ void __launch_bounds__(256)
k_stretch_propose(const double* __restrict__ x, int64_t W, int64_t wn, int64_t ens0, int ndim, int half, int64_t step,
                  uint64_t seed, double a, int64_t s0, int64_t ns, double* __restrict__ q, double* __restrict__ factors) {
  const int lane = threadIdx.x & 31;
  const int64_t sl0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // proposal number within [s0, s0 + ns)
  const int64_t s = s0 + sl0;
  const int64_t nh = wn / 2;
  if (sl0 >= ns) return;
  const int64_t e = s / nh, sl = s - e * nh, sg = (ens0 + e) * nh + sl;
  uint32_t r[4];
  philox4x32_10((uint32_t)step, (uint32_t)(step >> 32) ^ ((uint32_t)half << 31), (uint32_t)sg, (uint32_t)(sg >> 32) << 1, (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  const double z = (a - 1.0) * u01(r[0], r[1]) + 1.0;
  const double zz = z * z / a;
  int64_t j = (int64_t)(u01(r[2], r[3]) * (double)nh);
  if (j >= nh) j = nh - 1;
  const double* xs = x + (e * wn + half * nh + sl) * ndim;
  const double* xc = x + (e * wn + (1 - half) * nh + j) * ndim;
  // proposals of this call are stored from row 0
  for (int d = lane; d < ndim; d += 32) q[sl0 * ndim + d] = xc[d] - (xc[d] - xs[d]) * zz;
  if (lane == 0) factors[sl0] = (ndim - 1.0) * log(zz);
}

// accept with probability min(1, z^(n-1) p(q)/p(x)); one warp per walker copies the row
__global__ void __launch_bounds__(256)
k_stretch_accept(double* __restrict__ x, double* __restrict__ lnp, double* __restrict__ blobs, int64_t W, int64_t wn, int64_t ens0,
                 int ndim, int half, int64_t step, uint64_t seed, int64_t s0, int64_t ns, const double* __restrict__ q,
                 const double* __restrict__ factors, const double* __restrict__ lnp_q,
                 const double* __restrict__ blobs_q, long long* __restrict__ naccept) {
  const int lane = threadIdx.x & 31;
  const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nh = wn / 2;
  if (s >= ns) return;
  const int64_t sq = s;  // row of this proposal in q / factors / lnp_q / blobs_q
  const int64_t sa = s0 + s;
  const int64_t e = sa / nh, sl = sa - e * nh, sg = (ens0 + e) * nh + sl;
  const int64_t w = e * wn + half * nh + sl;
  uint32_t r[4];
  philox4x32_10((uint32_t)step, (uint32_t)(step >> 32) ^ ((uint32_t)half << 31), (uint32_t)sg, ((uint32_t)(sg >> 32) << 1) | 1u, (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  const double lnu = log(u01(r[0], r[1]));
  // lane 0 decides and broadcasts: no lane reads lnp[w] after lane 0 has overwritten it
  int acc_i = 0;
  if (lane == 0) acc_i = lnu < factors[sq] + lnp_q[sq] - lnp[w];
  const bool acc = __shfl_sync(0xffffffffu, acc_i, 0) != 0;
  if (!acc) return;
  for (int d = lane; d < ndim; d += 32) x[w * ndim + d] = q[sq * ndim + d];
  if (lane < 6) blobs[w * 6 + lane] = blobs_q[sq * 6 + lane];
  if (lane == 0) {
    lnp[w] = lnp_q[sq];
    if (naccept) naccept[w] += 1;
  }
}

}  // namespace

// host-callable copy of the generator for tests (numpy replay of a chain)
extern "C" int c1d_philox_u01(uint64_t seed, int64_t step, int half, int64_t walker, int which, double* out2) {
  if (!out2) return C1D_ERR_ARG;
  uint32_t r[4];
  philox4x32_10((uint32_t)step, (uint32_t)(step >> 32) ^ ((uint32_t)half << 31), (uint32_t)walker,
                ((uint32_t)((uint64_t)walker >> 32) << 1) | (uint32_t)(which & 1), (uint32_t)seed, (uint32_t)(seed >> 32), r);
  out2[0] = u01(r[0], r[1]);
  out2[1] = u01(r[2], r[3]);
  return C1D_OK;
}

int c1d_sampler_reserve(c1d_ctx* ctx, int64_t W);  // c1d_abi.cu

// proposals s0 .. s0 + ns - 1 of the active halves: propose, evaluate, accept (everything on stream s)
static int sampler_half_step(c1d_ctx* ctx, double* x_dev, double* lnp_dev, double* blobs_dev, int64_t W, int64_t wn, int64_t ens0,
                             int64_t step, int half, uint64_t seed, double a, int64_t s0, int64_t ns, long long* naccept_dev,
                             uint32_t flags, cudaStream_t s) {
  if (ns <= 0) return C1D_OK;
  const int ndim = ctx->cfg.ndim;
  const unsigned grid = (unsigned)((ns * 32 + 255) / 256);
  k_stretch_propose<<<grid, 256, 0, s>>>(x_dev, W, wn, ens0, ndim, half, step, seed, a, s0, ns, ctx->ws.xdev, ctx->ws.factors);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  int rc = c1d_loglike_batch(ctx, ctx->ws.xdev, ns, ctx->ws.lnp, nullptr, ctx->ws.blobs_q, nullptr, flags, (void*)s);
  if (rc) return rc;
  k_stretch_accept<<<grid, 256, 0, s>>>(x_dev, lnp_dev, blobs_dev, W, wn, ens0, ndim, half, step, seed, s0, ns, ctx->ws.xdev,
                                        ctx->ws.factors, ctx->ws.lnp, ctx->ws.blobs_q, naccept_dev);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return C1D_OK;
}

extern "C" int c1d_sampler_half_step(c1d_ctx* ctx, double* x_dev, double* lnp_dev, double* blobs_dev, int64_t W, int64_t n_ens,
                                     int64_t ens0, int64_t step, int half, uint64_t seed, double a, int64_t s0, int64_t ns,
                                     long long* naccept_dev, uint32_t flags, void* cuda_stream) {
  if (!ctx) return C1D_ERR_ARG;
  if (!x_dev || !lnp_dev || !blobs_dev || n_ens < 1 || (W % n_ens) || ((W / n_ens) % 2) || ens0 < 0 || (half != 0 && half != 1) ||
      s0 < 0 || ns < 0 || s0 + ns > W / 2 || !(a > 1.0)) {
    snprintf(ctx->err, sizeof(ctx->err), "bad arguments to c1d_sampler_half_step");
    return C1D_ERR_ARG;
  }
  int rc = c1d_sampler_reserve(ctx, ns > 0 ? ns : 1);
  if (rc) return rc;
  return sampler_half_step(ctx, x_dev, lnp_dev, blobs_dev, W, W / n_ens, ens0, step, half, seed, a, s0, ns, naccept_dev, flags,
                           (cudaStream_t)cuda_stream);
}

extern "C" int c1d_sampler_run_ens(c1d_ctx* ctx, double* x_dev, double* lnp_dev, double* blobs_dev, int64_t W, int64_t n_ens,
                                   int64_t ens0, int64_t nsteps, int64_t step0, uint64_t seed, double a, int init, int thin,
                                   double* chain_dev, double* lnp_chain_dev, double* blobs_chain_dev,
                                   long long* naccept_dev, uint32_t flags, void* cuda_stream) {
  if (!ctx) return C1D_ERR_ARG;
  if (!x_dev || !lnp_dev || !blobs_dev || n_ens < 1 || W < 2 * n_ens || (W % n_ens) || ((W / n_ens) % 2) || ens0 < 0 || nsteps < 0 ||
      thin < 1 || !(a > 1.0)) {
    snprintf(ctx->err, sizeof(ctx->err), "bad arguments to c1d_sampler_run (an even number of walkers per ensemble, a > 1)");
    return C1D_ERR_ARG;
  }
  const int ndim = ctx->cfg.ndim;
  const int64_t wn = W / n_ens, nht = W / 2;  // proposals per half-step: the active halves of all the ensembles
  cudaStream_t s = (cudaStream_t)cuda_stream;
  int rc = c1d_sampler_reserve(ctx, nht);
  if (rc) return rc;
  if (init) {
    rc = c1d_loglike_batch(ctx, x_dev, W, lnp_dev, nullptr, blobs_dev, nullptr, flags, cuda_stream);
    if (rc) return rc;
  }
  int64_t stored = 0;
  for (int64_t it = 0; it < nsteps; ++it) {
    const int64_t step = step0 + it;
    for (int half = 0; half < 2; ++half) {
      rc = sampler_half_step(ctx, x_dev, lnp_dev, blobs_dev, W, wn, ens0, step, half, seed, a, 0, nht, naccept_dev, flags, s);
      if (rc) return rc;
    }
    if ((it + 1) % thin == 0) {
      if (chain_dev)
        C1D_CUDA(ctx, cudaMemcpyAsync(chain_dev + stored * W * ndim, x_dev, (size_t)W * ndim * sizeof(double),
                                      cudaMemcpyDeviceToDevice, s));
      if (lnp_chain_dev)
        C1D_CUDA(ctx, cudaMemcpyAsync(lnp_chain_dev + stored * W, lnp_dev, (size_t)W * sizeof(double),
                                      cudaMemcpyDeviceToDevice, s));
      if (blobs_chain_dev)
        C1D_CUDA(ctx, cudaMemcpyAsync(blobs_chain_dev + stored * W * 6, blobs_dev, (size_t)W * 6 * sizeof(double),
                                      cudaMemcpyDeviceToDevice, s));
      ++stored;
    }
  }
  return C1D_OK;
}

extern "C" int c1d_sampler_run(c1d_ctx* ctx, double* x_dev, double* lnp_dev, double* blobs_dev, int64_t W,
                               int64_t nsteps, int64_t step0, uint64_t seed, double a, int init, int thin,
                               double* chain_dev, double* lnp_chain_dev, double* blobs_chain_dev,
                               long long* naccept_dev, uint32_t flags, void* cuda_stream) {
  return c1d_sampler_run_ens(ctx, x_dev, lnp_dev, blobs_dev, W, 1, 0, nsteps, step0, seed, a, init, thin, chain_dev, lnp_chain_dev,
                             blobs_chain_dev, naccept_dev, flags, cuda_stream);
}
