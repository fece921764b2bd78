// Stage 2 (exact-parity mode): the LaCE-style MLP in fp64 on the fp64 tensor cores
// (mma.sync m8n8k4, DMMA).  Replaces, for all rows of a batch at once, the per-call
// `NNEmulator.emulate_p1d_Mpc` network forward pass that Theory.get_p1d_kms drives
// (reference call site cup1d/likelihood/lya_theory.py:794-812; LaCE itself is not in the
// reference tree, see oracle/emulators.py for the restated network).
//
// One CTA per SM works on three tiles of 24 rows ((walker, z) pairs) at a time, one per group of
// four warps (so every SM sub-partition holds one warp of each group), and chains all layers with the activations resident in shared memory (k-major, row
// stride 68 doubles).  The groups share the weight stream but synchronise separately, so one
// group's layer epilogue (pipeline drain, activation write-back, barrier) hides under the other
// group's DMMAs on the same SM sub-partitions.  The weights of all
// layers form one stream of k-slabs (<= 33 KB each) that a producer warp double buffers in
// shared memory with cp.async.bulk + mbarriers: slab s+1 (possibly of the next layer, or of the
// next tile) lands while slab s feeds the tensor cores, so no global-memory latency sits on the
// DMMA path and the compute warps only meet at the two barriers a layer needs.  The weight image
// is zero-padded on the host ([K to 4][N to 8, + 4 columns of row skew]) so the inner loop
// carries no bounds checks and a slab is one contiguous copy.
// Warp wn of a group owns the group's 32 rows and a run of 8-column blocks; the blocks of a layer
// are dealt to the four n-warps as evenly as they divide (the odd blocks go to opposite ends in
// the two groups, which balances the sub-partitions).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "c1d_internal.h"

// -DC1D_MLP64_TIMING compiles the per-layer cycle counters in (C1D_MLP64_DEBUG=1 prints them); the
// product build leaves them out: the counter array otherwise costs registers and local memory
#ifdef C1D_MLP64_TIMING
#define M64_DBG(x) x
#else
#define M64_DBG(x)
#endif

namespace {

constexpr int M64_GROUPS = 3;                         // row groups per CTA of the throughput variant (one warp of each per SM sub-partition)
constexpr int M64_RB = 3;                             // 8-row blocks per group
constexpr int M64_ROWS = 8 * M64_RB;                  // rows of one group's tile
// The kernel is instantiated three times: <3 groups, 2 slab buffers> for throughput (72 rows per CTA),
// <1 group, 4 slab buffers> for small batches, where a CTA holds one 24-row tile, the rows spread over
// up to 148 CTAs and four slabs in flight hide the L2 latency that otherwise paces a single tile, and
// <2 groups, 2 slab buffers> for the batch sizes in between whose last round of 72-row units would run
// nearly empty (the launcher picks by estimated time).
template <int G> struct M64Geom {
  static constexpr int CWARPS = 4 * G;             // compute warps
  static constexpr int THREADS = 32 * (CWARPS + 1);  // + the weight streamer
  static constexpr int STR = G * M64_ROWS + 4;     // doubles per k-row of the activation tile
  static_assert(STR % 16 == 4 || STR % 16 == 12, "k-row stride must spread fragments over all banks");
};
constexpr int M64_MAXW = 256;       // widest layer
constexpr int M64_NBUF = 2;         // weight slab buffers (ring) of the throughput variant
constexpr int M64_NBUF_SMALL = 4;   // ... of the small-batch variant
constexpr int M64_SLAB = 4224;      // doubles per weight slab buffer
constexpr int M64_MAX_SLABS = 120;
constexpr int M64_MAXBIAS = 1024;   // doubles of padded biases kept in shared memory
template <int G, int NBUF> constexpr size_t m64_smem() {
  return ((size_t)M64_MAXW * M64Geom<G>::STR + NBUF * (size_t)M64_SLAB + M64_MAXBIAS) * sizeof(double) + 16 * NBUF;
}
static_assert(m64_smem<M64_GROUPS, M64_NBUF>() <= 232448 && m64_smem<1, M64_NBUF_SMALL>() <= 232448, "shared memory budget");

struct M64Slab {
  uint32_t w_off;   // doubles from the start of the padded weight image
  uint16_t k0, kc;  // first k-row and number of k-rows (multiple of 4)
  uint8_t layer, first, last, pad;
};

struct M64Plan {
  int n_layers, n_slabs, n_in, n_out, debug, n_bias;
  double slope;
  const double* wimg;  // padded weights, layer after layer, [Kp][Np]
  const double* bimg;  // padded biases
  int np[C1D_MAX_LAYERS];       // N padded to 8
  int nbt[C1D_MAX_LAYERS];      // 8-column blocks of the layer
  int boff[C1D_MAX_LAYERS];
  M64Slab slabs[M64_MAX_SLABS];
};

struct M64State {
  M64Plan p;
  double *d_w, *d_b;
};

__device__ __forceinline__ void dmma8(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait (as in c1d_tc_common.cuh): every try_wait suspends for at most 10 ms; after ~20 s without the phase
// completing (a faulted bulk copy) the kernel traps - a CUDA error on the host's next
// call instead of a hang.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (int tries = 0; tries < 2048; ++tries) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();  // (no printf here: its call sequence costs the DMMA loop registers at the 128 cap)
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// barrier among the four warps of one group (the streamer and the other group never join)
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); }

// the two input values thread gtid of a group stages for its tile ([8 k-rows][M64_ROWS rows])
__device__ __forceinline__ void load_inputs(const M64Plan& p, const double* __restrict__ emu_in, const double* __restrict__ emu_lo,
                                            const double* __restrict__ emu_hi, int64_t row0, int64_t nrows, int gtid, double* xin) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int t = gtid + 128 * u;
    const int j = t / M64_ROWS;
    const int64_t row = row0 + (t % M64_ROWS);
    double v = 0.0;
    if (t < 8 * M64_ROWS && j < p.n_in && row < nrows) {
      const double lo = __ldg(emu_lo + j);
      v = (emu_in[row * C1D_MAX_EMU + j] - lo) / (__ldg(emu_hi + j) - lo) - 0.5;
    }
    xin[u] = v;
  }
}

// one layer on one tile; q is the running slab index (buffer parity), returns with every
// activation of the layer written (or, for the output layer, stored to `out`)
template <int NB, int G, int NBUF>
__device__ __forceinline__ void m64_layer(const M64Plan& p, int l, int& s, uint32_t& q, uint32_t bar_full, uint32_t bar_empty,
                                          double* act, const double* wbuf, const double* bsm, bool is_out, int64_t row0, int64_t nrows,
                                          double* __restrict__ out, int out_stride, long long* dbg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp >> 2, wn = warp & 3;
  const int kq = lane & 3, g = lane >> 2;
  const int r0 = M64_ROWS * grp;
  const int nbt = p.nbt[l];
  const int base = nbt >> 2, rem = nbt & 3;
  // blocks of this warp: NB or NB - 1; the rem odd blocks rotate over the n-warps from group to group so
  // that the sub-partitions (one warp of every group each) carry the same number of blocks
  const int wx = (wn + grp) & 3;
  const int cnt = base + (wx < rem ? 1 : 0);
  const int before = wx * base + (wx < rem ? wx : rem);
  const int n0 = 8 * before;  // first column
  const int ns = p.np[l] + 4;
  const bool last_blk = cnt == NB;  // the NB-th block exists for this warp
  // accumulators start from the bias (thread holds columns n0 + 8 nb + 2 kq + {0, 1} of rows r0 + 8 a + g)
  const double* bias = bsm + p.boff[l];
  double acc[M64_RB][NB][2];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const double b0 = (nb < cnt) ? bias[n0 + 8 * nb + 2 * kq] : 0.0;
    const double b1 = (nb < cnt) ? bias[n0 + 8 * nb + 2 * kq + 1] : 0.0;
#pragma unroll
    for (int a = 0; a < M64_RB; ++a) {
      acc[a][nb][0] = b0;
      acc[a][nb][1] = b1;
    }
  }

  for (;;) {
    const M64Slab sl = p.slabs[s];
    const uint32_t buf = q % NBUF;
    M64_DBG(const long long tw0 = dbg ? clock64() : 0;)
    mbar_wait(bar_full + 8 * buf, (q / NBUF) & 1u);  // slab q has landed
    M64_DBG(const long long tw1 = dbg ? clock64() : 0;)
    const double* wb = wbuf + buf * M64_SLAB;
    if (cnt > 0) {
      const double* ap = act + (size_t)(sl.k0 + kq) * M64Geom<G>::STR + r0 + g;
      const double* bp = wb + kq * ns + n0 + g;
#pragma unroll 2
      for (int kk = 0; kk < (int)sl.kc; kk += 4) {
        double af[M64_RB], bf[NB];
#pragma unroll
        for (int a = 0; a < M64_RB; ++a) af[a] = ap[kk * M64Geom<G>::STR + 8 * a];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) bf[nb] = (nb < NB - 1 || last_blk) ? bp[kk * ns + 8 * nb] : 0.0;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          if (nb < NB - 1 || last_blk) {
#pragma unroll
            for (int a = 0; a < M64_RB; ++a) dmma8(acc[a][nb][0], acc[a][nb][1], af[a], bf[nb]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * buf);  // this warp is done with the buffer
    M64_DBG(if (dbg) { dbg[0] += tw1 - tw0; dbg[1] += clock64() - tw1; })
    ++s;
    ++q;
    if (sl.last) break;
  }
  M64_DBG(const long long te0 = dbg ? clock64() : 0;)
  if (!is_out) {
    group_sync(grp);  // every read of the input activations is done
    const double slope = p.slope;
    // LeakyReLU on the sign bit, then a store order that touches every bank once: lanes with
    // kq < 2 store their even column first, the others their odd column (rows of the k-major tile
    // are 68 doubles apart, so columns 4 apart share banks)
    const bool swap = kq >= 2;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      if (nb < cnt) {
#pragma unroll
        for (int a = 0; a < M64_RB; ++a) {
          double v0 = acc[a][nb][0], v1 = acc[a][nb][1];
          if (__double2hiint(v0) < 0) v0 *= slope;  // padded columns stay exactly zero
          if (__double2hiint(v1) < 0) v1 *= slope;
          const double first = swap ? v1 : v0, second = swap ? v0 : v1;
          const int nf = n0 + 8 * nb + 2 * kq + (swap ? 1 : 0), nsx = n0 + 8 * nb + 2 * kq + (swap ? 0 : 1);
          act[(size_t)nf * M64Geom<G>::STR + r0 + 8 * a + g] = first;
          act[(size_t)nsx * M64Geom<G>::STR + r0 + 8 * a + g] = second;
        }
      }
    }
    group_sync(grp);  // activations of the next layer are visible
    M64_DBG(if (dbg) dbg[2] += clock64() - te0;)
  } else {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      if (nb < cnt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = n0 + 8 * nb + 2 * kq + e;
          if (n < p.n_out) {
#pragma unroll
            for (int a = 0; a < M64_RB; ++a) {
              const int64_t row = row0 + 8 * a + g;
              if (row < nrows) out[row * out_stride + n] = acc[a][nb][e];
            }
          }
        }
      }
    }
  }
}

// 13 warps are allocated as 16 (groups of four), so 128 registers per thread is the hardware cap for this
// block size (a 144- or 152-register build fails to launch); register double buffering of the fragments
// spills at 128 and was dropped.  Also measured: no streamer warp (the last warp out of a slab buffer
// starts its refill, 12 warps at 168 registers): 4.96 ms against 4.78 ms, 5.05 ms with double buffering.
// A non-blocking mbarrier.test_wait for the next slab issued two k-steps before the end of the current one
// (to skip the ~100-cycle blocking poll): 5.01 ms - the extra live state spills in the DMMA loop.
template <int G, int NBUF>
__global__ void __launch_bounds__(M64Geom<G>::THREADS, 1)
k_mlp_f64(M64Plan p, const double* __restrict__ emu_in, const double* __restrict__ emu_lo,
          const double* __restrict__ emu_hi, int64_t nrows, double* __restrict__ out, int out_stride) {
  extern __shared__ __align__(16) double smem_d[];
  double* act = smem_d;                                  // [M64_MAXW][M64Geom<G>::STR]
  double* wbuf = smem_d + (size_t)M64_MAXW * M64Geom<G>::STR;    // [NBUF][M64_SLAB]
  double* bsm = wbuf + NBUF * M64_SLAB;               // [M64_MAXBIAS] padded biases of all layers
  const uint32_t bar_full = smem_u32(bsm + M64_MAXBIAS), bar_empty = bar_full + 8 * NBUF;
  const int64_t ntiles = (nrows + M64_ROWS - 1) / M64_ROWS;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < p.n_bias; i += M64Geom<G>::THREADS) bsm[i] = __ldg(p.bimg + i);
  if (threadIdx.x == 0) {
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(bar_full + 8 * b, 1);
      mbar_init(bar_empty + 8 * b, M64Geom<G>::CWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // both groups walk the same number of tiles so that they consume the same slab stream
  // (surplus tiles hold no rows)
  const int64_t my_tiles = (ntiles + G * (int64_t)gridDim.x - 1) / (G * (int64_t)gridDim.x);
  if (warp == M64Geom<G>::CWARPS) {
    // ============================== weight streamer ==============================
    if ((threadIdx.x & 31) == 0) {
      uint32_t q = 0;
      for (int64_t it = 0; it < my_tiles; ++it)
        for (int s = 0; s < p.n_slabs; ++s, ++q) {
          const uint32_t b = q % NBUF;
          mbar_wait(bar_empty + 8 * b, ((q / NBUF) & 1u) ^ 1u);  // passes at once on the first lap
          const M64Slab sl = p.slabs[s];
          const uint32_t bytes = (uint32_t)sl.kc * (uint32_t)(p.np[sl.layer] + 4) * 8u;
          mbar_expect_tx(bar_full + 8 * b, bytes);
          bulk_g2s(smem_u32(wbuf + b * M64_SLAB), p.wimg + sl.w_off, bytes, bar_full + 8 * b);
        }
    }
    return;
  }
  // ================================ compute warps ================================
  const int grp = warp >> 2, gtid = threadIdx.x & 127;
  uint32_t q = 0;
  double xin[2];
  for (int64_t it = 0; it < my_tiles; ++it) {
    const int64_t tile = (it * gridDim.x + blockIdx.x) * G + grp;
    const int64_t row0 = tile * M64_ROWS;
    // normalised inputs (x - lo)/(hi - lo) - 0.5, k rows padded to 8 with zeros; they were fetched
    // while the previous tile was in its last hidden layers
    if (it == 0) load_inputs(p, emu_in, emu_lo, emu_hi, row0, nrows, gtid, xin);
    group_sync(grp);  // the previous tile's output layer has read its activations
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t = gtid + 128 * u;
      if (t < 8 * M64_ROWS) act[(t / M64_ROWS) * M64Geom<G>::STR + M64_ROWS * grp + (t % M64_ROWS)] = xin[u];
    }
    group_sync(grp);
    int s = 0;
#ifdef C1D_MLP64_TIMING
    long long dbg_store[3 * C1D_MAX_LAYERS];
    const bool dbg_on = p.debug && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && it == 1;
    const long long tt0 = dbg_on ? clock64() : 0;
#endif
    for (int l = 0; l < p.n_layers; ++l) {
      long long* dbg = nullptr;
      M64_DBG(dbg = dbg_on ? dbg_store + 3 * l : nullptr; if (dbg) dbg[0] = dbg[1] = dbg[2] = 0;)
      const bool is_out = l == p.n_layers - 1;
      if (l == p.n_layers - 2 && it + 1 < my_tiles)
        load_inputs(p, emu_in, emu_lo, emu_hi, (((it + 1) * gridDim.x + blockIdx.x) * G + grp) * M64_ROWS, nrows, gtid, xin);
      const int nb = (p.nbt[l] + 3) >> 2;
      switch (nb) {
        case 1: m64_layer<1, G, NBUF>(p, l, s, q, bar_full, bar_empty, act, wbuf, bsm, is_out, row0, nrows, out, out_stride, dbg); break;
        case 2: m64_layer<2, G, NBUF>(p, l, s, q, bar_full, bar_empty, act, wbuf, bsm, is_out, row0, nrows, out, out_stride, dbg); break;
        case 3: m64_layer<3, G, NBUF>(p, l, s, q, bar_full, bar_empty, act, wbuf, bsm, is_out, row0, nrows, out, out_stride, dbg); break;
        case 4: m64_layer<4, G, NBUF>(p, l, s, q, bar_full, bar_empty, act, wbuf, bsm, is_out, row0, nrows, out, out_stride, dbg); break;
        case 5: m64_layer<5, G, NBUF>(p, l, s, q, bar_full, bar_empty, act, wbuf, bsm, is_out, row0, nrows, out, out_stride, dbg); break;
        case 6: m64_layer<6, G, NBUF>(p, l, s, q, bar_full, bar_empty, act, wbuf, bsm, is_out, row0, nrows, out, out_stride, dbg); break;
        case 7: m64_layer<7, G, NBUF>(p, l, s, q, bar_full, bar_empty, act, wbuf, bsm, is_out, row0, nrows, out, out_stride, dbg); break;
        default: m64_layer<8, G, NBUF>(p, l, s, q, bar_full, bar_empty, act, wbuf, bsm, is_out, row0, nrows, out, out_stride, dbg); break;
      }
    }
#ifdef C1D_MLP64_TIMING
    if (dbg_on) {
      const long long tt1 = clock64();
      for (int w = 0; w < M64Geom<G>::CWARPS; ++w) {
        if (w == warp) {
          printf("mlp64 warp %d tile cycles %lld | per layer wait/compute/epilogue:", warp, tt1 - tt0);
          for (int l = 0; l < p.n_layers; ++l) printf(" L%d %lld/%lld/%lld", l, dbg_store[3 * l], dbg_store[3 * l + 1], dbg_store[3 * l + 2]);
          printf("\n");
        }
      }
    }
#endif
  }
}

}  // namespace

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
void c1d_mlp64_destroy(c1d_ctx* ctx) {
  M64State* st = (M64State*)ctx->mlp64_state;
  if (!st) return;
  cudaFree(st->d_w);
  cudaFree(st->d_b);
  delete st;
  ctx->mlp64_state = nullptr;
}

int c1d_mlp64_prepare(c1d_ctx* ctx) {
  const c1d_config& c = ctx->cfg;
  c1d_mlp64_destroy(ctx);
  if (c.emu_kind != 0 || c.n_layers < 1) return C1D_OK;
  M64State* st = new M64State();
  memset(st, 0, sizeof(*st));
  ctx->mlp64_state = st;
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

  M64Plan& p = st->p;
  p.n_layers = L;
  p.n_in = c.layer_dims[0];
  p.n_out = c.layer_dims[L];
  p.slope = c.nn_slope;
  p.debug = getenv("C1D_MLP64_DEBUG") ? 1 : 0;
  std::vector<double> img, bimg;
  int n_slabs = 0;
  for (int l = 0; l < L; ++l) {
    const int K = c.layer_dims[l], N = c.layer_dims[l + 1];
    // the input tile is zero-filled to 8 k-rows; hidden activations are exact zeros up to the
    // previous layer's padded width, which covers K rounded up to 4
    const int kp = (l == 0) ? ((K + 3) & ~3) : ((K + 3) & ~3);
    const int np = (N + 7) & ~7;
    p.np[l] = np;
    p.nbt[l] = np / 8;
    p.boff[l] = (int)bimg.size();
    for (int n = 0; n < np; ++n) bimg.push_back(n < N ? Bv[boff[l] + n] : 0.0);
    const size_t w0 = img.size();
    const int ns = np + 4;  // row skew: fragments of 4 k-rows x 8 columns touch every bank once
    img.resize(w0 + (size_t)kp * ns, 0.0);
    for (int k = 0; k < K; ++k)
      for (int n = 0; n < N; ++n) img[w0 + (size_t)k * ns + n] = W[woff[l] + (size_t)k * N + n];  // stored in-major: Wt[k*N + n]
    int ks = (M64_SLAB / (np + 4)) & ~3;  // k-rows per slab
    if (ks < 4) {
      snprintf(ctx->err, sizeof(ctx->err), "layer %d too wide for the fp64 MLP slab buffer", l);
      return C1D_ERR_ARG;
    }
    for (int k0 = 0; k0 < kp; k0 += ks) {
      if (n_slabs >= M64_MAX_SLABS) {
        snprintf(ctx->err, sizeof(ctx->err), "fp64 MLP slab table overflow");
        return C1D_ERR_ARG;
      }
      M64Slab& sl = p.slabs[n_slabs++];
      sl.w_off = (uint32_t)(w0 + (size_t)k0 * ns);
      sl.k0 = (uint16_t)k0;
      sl.kc = (uint16_t)((kp - k0 < ks) ? (kp - k0) : ks);
      sl.layer = (uint8_t)l;
      sl.first = k0 == 0;
      sl.last = k0 + ks >= kp;
      sl.pad = 0;
    }
  }
  p.n_slabs = n_slabs;
  p.n_bias = (int)bimg.size();
  if (p.n_bias > M64_MAXBIAS) {
    snprintf(ctx->err, sizeof(ctx->err), "fp64 MLP bias table overflow");
    return C1D_ERR_ARG;
  }
  C1D_CUDA(ctx, cudaMalloc(&st->d_w, img.size() * sizeof(double)));
  C1D_CUDA(ctx, cudaMalloc(&st->d_b, bimg.size() * sizeof(double)));
  C1D_CUDA(ctx, cudaMemcpy(st->d_w, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice));
  C1D_CUDA(ctx, cudaMemcpy(st->d_b, bimg.data(), bimg.size() * sizeof(double), cudaMemcpyHostToDevice));
  p.wimg = st->d_w;
  p.bimg = st->d_b;
  C1D_CUDA(ctx, cudaFuncSetAttribute(k_mlp_f64<M64_GROUPS, M64_NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)m64_smem<M64_GROUPS, M64_NBUF>()));
  C1D_CUDA(ctx, cudaFuncSetAttribute(k_mlp_f64<1, M64_NBUF_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)m64_smem<1, M64_NBUF_SMALL>()));
  C1D_CUDA(ctx, cudaFuncSetAttribute(k_mlp_f64<2, M64_NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)m64_smem<2, M64_NBUF>()));
  return C1D_OK;
}

int c1d_mlp64_launch(c1d_ctx* ctx, const double* emu_in, int64_t nrows, double* out, int out_stride, cudaStream_t s) {
  M64State* st = (M64State*)ctx->mlp64_state;
  if (!st) {
    snprintf(ctx->err, sizeof(ctx->err), "fp64 MLP not prepared");
    return C1D_ERR_STATE;
  }
  const int64_t ntiles = (nrows + M64_ROWS - 1) / M64_ROWS;  // 24-row tiles
  // One, two or three 24-row tiles per CTA and round: the variant with the shortest estimated time
  // = rounds x time per round.  Measured on C4 (us per round): one tile (four slabs in flight) 41, two 53,
  // three 72; e.g. 64 walkers 84 -> 47 us with one tile per CTA, 512 walkers 82 -> 64 us and 1024 walkers
  // 154 -> 115 us with two, three from ~3000 walkers on.  C1D_MLP64_G=1|2|3 forces a variant.
  const int64_t sms = ctx->num_sms;
  const double t_round[3] = {41.0, 53.0, 72.0};
  int g = 3;
  double best = 1e300;
  for (int cand = 3; cand >= 1; --cand) {
    const int64_t units = (ntiles + cand - 1) / cand;
    const double t = (double)((units + sms - 1) / sms) * t_round[cand - 1];
    if (t < 0.98 * best) {
      best = t;
      g = cand;
    }
  }
  if (const char* fg = getenv("C1D_MLP64_G")) {
    const int v = atoi(fg);
    if (v >= 1 && v <= 3) g = v;
  }
  const int64_t units = (ntiles + g - 1) / g;
  const int grid = (int)(units < sms ? units : sms);
  if (g == 1)
    k_mlp_f64<1, M64_NBUF_SMALL><<<grid, M64Geom<1>::THREADS, m64_smem<1, M64_NBUF_SMALL>(), s>>>(
        st->p, emu_in, ctx->d.emu_lo, ctx->d.emu_hi, nrows, out, out_stride);
  else if (g == 2)
    k_mlp_f64<2, M64_NBUF><<<grid, M64Geom<2>::THREADS, m64_smem<2, M64_NBUF>(), s>>>(
        st->p, emu_in, ctx->d.emu_lo, ctx->d.emu_hi, nrows, out, out_stride);
  else
    k_mlp_f64<M64_GROUPS, M64_NBUF><<<grid, M64Geom<M64_GROUPS>::THREADS, m64_smem<M64_GROUPS, M64_NBUF>(), s>>>(
        st->p, emu_in, ctx->d.emu_lo, ctx->d.emu_hi, nrows, out, out_stride);
  ctx->launches++;
  C1D_CUDA(ctx, cudaPeekAtLastError());
  return C1D_OK;
}
