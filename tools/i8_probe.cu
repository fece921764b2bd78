// Probes behind the int8 digit-slice (Ozaki-style) emulator kernel, run on a B200:
//   1. correctness of tcgen05.mma kind::i8 (s8 x s8 -> s32, M = 128, K = 32 per instruction) with A in
//      SWIZZLE_128B shared memory or in tensor memory and B in SWIZZLE_128B / 64B / 32B shared memory,
//   2. cycles per MMA against N, the operand form and the number of accumulators used in turn,
//   3. tcgen05.ld throughput with 1..4 warps per TMEM lane quadrant,
//   4. throughput of the fp64-pipe conversions the epilogue needs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Icup1d_b200/csrc -o tools/_bin/i8_probe tools/i8_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "c1d_tc_common.cuh"

using namespace c1d_tc;

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                      \
    }                                                                               \
  } while (0)

constexpr int A_BYTES = 4 * 16384;   // four digit planes, one 128-k chunk each
constexpr int B_BYTES = 4 * 32768;   // four planes of up to 256 rows x 128 B
constexpr int SMEM = 1024 + A_BYTES + B_BYTES + 64;

// mode: 0 correctness SS, 1 correctness TS; blay: B layout 0 = SW128 (row 128 B), 1 = SW64, 2 = SW32.
// K = 64 (two k-steps).  Aimg / Bimg are pre-swizzled images.
__global__ void __launch_bounds__(160, 1)
k_i8_check(const uint8_t* __restrict__ Aimg, const uint8_t* __restrict__ Bimg, const int8_t* __restrict__ Araw,
           int32_t* __restrict__ D, int N, int mode, int blay) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_smem = base, b_smem = base + A_BYTES, bar = b_smem + B_BYTES, tmem_slot = bar + 8;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  if (warp < 4) {
    for (int i = tid; i < 16384 / 16; i += 128) reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(Aimg)[i];
    for (int i = tid; i < 32768 / 16; i += 128)
      reinterpret_cast<uint4*>(gen + A_BYTES)[i] = reinterpret_cast<const uint4*>(Bimg)[i];
    if (mode == 1) {
      // A in tensor memory: lane = row, 32-bit column c holds k = 4c .. 4c + 3 (little endian)
      const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
      uint32_t w[16];
      for (int c = 0; c < 16; ++c) {
        uint32_t v = 0;
        for (int b = 0; b < 4; ++b) v |= (uint32_t)(uint8_t)Araw[tid * 64 + 4 * c + b] << (8 * b);
        w[c] = v;
      }
      tmem_st16(tmem_lane + 256, w);
      tc_wait_st();
    }
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 128) {
    const uint32_t idesc = make_idesc_s8((uint32_t)N);
    for (uint32_t ks = 0; ks < 2; ++ks) {
      uint64_t bd;
      if (blay == 0) bd = make_desc_sw128(b_smem + ks * 32u);
      else if (blay == 1) bd = make_desc_sw64(b_smem + ks * 32u);
      else bd = make_desc_sw32(b_smem + ks * (uint32_t)N * 32u);  // one 32-byte-row tile per k-step
      if (mode == 0) mma_i8_ss(tmem_base, make_desc_sw128(a_smem + ks * 32u), bd, idesc, ks ? 1u : 0u);
      else mma_i8_ts(tmem_base, tmem_base + 256 + ks * 8u, bd, idesc, ks ? 1u : 0u);
    }
    tc_commit(bar);
  }
  if (warp < 4) {
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int col0 = 0; col0 < N; col0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_lane + col0, r);
      tc_wait_ld();
      for (int i = 0; i < 16; ++i) D[tid * N + col0 + i] = (int32_t)r[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// issue-rate probe: REP MMAs of width N, accumulators used in turn (nacc of them, dstep columns apart),
// A planes / B planes cycled like the digit pairs of the real kernel.  form 0 = SS, 1 = TS.
// kind 0 = i8, 1 = tf32.  out[0] = cycles per MMA until completion, out[1] = cycles per MMA spent issuing.
__global__ void __launch_bounds__(160, 1) k_rate(float* out, int N, int nacc, int dstep, int form, int kind, int blay) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_smem = base, b_smem = base + A_BYTES, bar = b_smem + B_BYTES, tmem_slot = bar + 8;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512u);
  for (int i = tid; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x01010101u;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  if (tid == 128) {
    const uint32_t idesc = kind == 0 ? make_idesc_s8((uint32_t)N) : make_idesc_tf32((uint32_t)N);
    const int REP = 240;
    // everything an MMA needs sits in registers before the clock starts: the loop body is the MMA alone
    uint64_t ad[4], bd[4];
    uint32_t at[4], dd[4];
    for (uint32_t q = 0; q < 4; ++q) {
      ad[q] = make_desc_sw128(a_smem + q * 16384u + (q & 1u) * 32u);
      if (blay == 0) bd[q] = make_desc_sw128(b_smem + q * 32768u + (q & 1u) * 32u);
      else if (blay == 1) bd[q] = make_desc_sw64(b_smem + q * 16384u + (q & 1u) * 32u);
      else bd[q] = make_desc_sw32(b_smem + q * 8192u);
      at[q] = tmem_base + 384 + q * 16;
      dd[q] = tmem_base + (uint32_t)((q % nacc) * dstep);
    }
    const long long t0 = clock64();
    if (form == 0 && kind == 0) {
      for (int r = 0; r < REP; r += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) mma_i8_ss(dd[q & 3], ad[(q + (q >> 2)) & 3], bd[q & 3], idesc, (r | q) >= 4 ? 1u : 0u);
      }
    } else if (form == 1 && kind == 0) {
      for (int r = 0; r < REP; r += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) mma_i8_ts(dd[q & 3], at[(q + (q >> 2)) & 3], bd[q & 3], idesc, (r | q) >= 4 ? 1u : 0u);
      }
    } else if (form == 0) {
      for (int r = 0; r < REP; r += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) mma_ss(dd[q & 3], ad[(q + (q >> 2)) & 3], bd[q & 3], idesc, (r | q) >= 4 ? 1u : 0u);
      }
    } else {
      for (int r = 0; r < REP; r += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) mma_ts(dd[q & 3], at[(q + (q >> 2)) & 3], bd[q & 3], idesc, (r | q) >= 4 ? 1u : 0u);
      }
    }
    const long long t1 = clock64();
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = (float)(t2 - t0) / REP;
    out[1] = (float)(t1 - t0) / REP;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// the emulator kernel's own issue pattern: U units of ten MMAs (digit pairs, four level accumulators `ls` columns apart),
// A digit planes `apl` bytes apart in SWIZZLE_64B (alay 1) or SWIZZLE_128B (alay 0) chunks, k-steps walking the chunks,
// B planes of nb x 32 B in SWIZZLE_32B ring slots; commit per unit when `commits`.  order 0 = kernel order,
// 1 = A plane changes every MMA.
__global__ void __launch_bounds__(160, 1) k_rate2(float* out, int N, int ls, int alay, uint32_t apl, int commits, int order, int U, uint32_t aofs) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_smem = base + aofs, b_smem = base + 98304 + 65536, bar = b_smem + 65536, bar2 = bar + 8, tmem_slot = bar + 16;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512u);
  for (int i = tid; i < (98304 + 65536 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x01010101u;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_p, 0);
  if (warp == 4) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_s8((uint32_t)N);
    const long long t0 = clock64();
    for (int u = 0; u < U; ++u) {
      const uint32_t ks = (uint32_t)u;
      const uint32_t aoff = alay ? (ks >> 1) * 8192u + (ks & 1u) * 32u : (ks >> 2) * 16384u + (ks & 3u) * 32u;
      const uint64_t ad0 = alay ? make_desc_sw64(a_smem + aoff) : make_desc_sw128(a_smem + aoff), astep = apl >> 4;
      const uint64_t bd0 = make_desc_sw32(b_smem + (uint32_t)(u & 3) * 16384u), bstep = (uint32_t)N * 2u;
      const uint32_t d0 = tmem_base, keep = u ? 1u : 0u;
      if (leader) {
        if (order == 0) {
          mma_i8_ss(d0, ad0, bd0, idesc, keep);
          mma_i8_ss(d0 + ls, ad0, bd0 + bstep, idesc, keep);
          mma_i8_ss(d0 + 2 * ls, ad0, bd0 + 2 * bstep, idesc, keep);
          mma_i8_ss(d0 + 3 * ls, ad0, bd0 + 3 * bstep, idesc, keep);
          mma_i8_ss(d0 + ls, ad0 + astep, bd0, idesc, 1u);
          mma_i8_ss(d0 + 2 * ls, ad0 + astep, bd0 + bstep, idesc, 1u);
          mma_i8_ss(d0 + 3 * ls, ad0 + astep, bd0 + 2 * bstep, idesc, 1u);
          mma_i8_ss(d0 + 2 * ls, ad0 + 2 * astep, bd0, idesc, 1u);
          mma_i8_ss(d0 + 3 * ls, ad0 + 2 * astep, bd0 + bstep, idesc, 1u);
          mma_i8_ss(d0 + 3 * ls, ad0 + 3 * astep, bd0, idesc, 1u);
        } else {
          mma_i8_ss(d0, ad0, bd0, idesc, keep);
          mma_i8_ss(d0 + ls, ad0 + astep, bd0, idesc, keep);
          mma_i8_ss(d0 + 2 * ls, ad0 + 2 * astep, bd0, idesc, keep);
          mma_i8_ss(d0 + 3 * ls, ad0 + 3 * astep, bd0, idesc, keep);
          mma_i8_ss(d0 + ls, ad0, bd0 + bstep, idesc, 1u);
          mma_i8_ss(d0 + 2 * ls, ad0 + astep, bd0 + bstep, idesc, 1u);
          mma_i8_ss(d0 + 3 * ls, ad0 + 2 * astep, bd0 + bstep, idesc, 1u);
          mma_i8_ss(d0 + 2 * ls, ad0, bd0 + 2 * bstep, idesc, 1u);
          mma_i8_ss(d0 + 3 * ls, ad0 + astep, bd0 + 2 * bstep, idesc, 1u);
          mma_i8_ss(d0 + 3 * ls, ad0, bd0 + 3 * bstep, idesc, 1u);
        }
        if (commits) tc_commit(bar2);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (leader) tc_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (leader) {
      out[0] = (float)(t2 - t0) / (10 * U);
      out[1] = (float)(t1 - t0) / (10 * U);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// TMEM load throughput: wq warps per lane quadrant sweep the 512 columns REP times (x16 or x32 loads)
template <int X>
__device__ __forceinline__ uint32_t ld_sum(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_sum<16>(uint32_t taddr) {
  uint32_t r[16];
  tmem_ld16(taddr, r);
  tc_wait_ld();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  return s;
}
template <>
__device__ __forceinline__ uint32_t ld_sum<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  tc_wait_ld();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += r[i];
  return s;
}

template <int X>
__global__ void __launch_bounds__(544, 1) k_tmem_ld(float* out, uint32_t* sink, int wq, int pipelined) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 16) tmem_alloc(smem_u32(&slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = slot;
  const int REP = 16;
  long long t0 = 0, t1 = 0;
  uint32_t s = 0;
  __syncthreads();
  if (warp < 4 * wq) {
    const uint32_t quad = warp & 3, sub = warp >> 2;
    const uint32_t lane_addr = tmem_base + ((quad * 32u) << 16);
    t0 = clock64();
    for (int r = 0; r < REP; ++r) {
      if (pipelined && X == 16) {
        // two loads in flight per warp
        uint32_t ra[16], rb[16];
        for (uint32_t c = sub * 32; c < 512; c += wq * 32) {
          tmem_ld16(lane_addr + c, ra);
          tmem_ld16(lane_addr + c + 16, rb);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) s += ra[i] ^ rb[i];
        }
      } else {
        for (uint32_t c = sub * X; c < 512; c += wq * X) s += ld_sum<X>(lane_addr + c);
      }
    }
    t1 = clock64();
  }
  sink[blockIdx.x * blockDim.x + tid] = s;
  __syncthreads();
  if (tid == 0) out[0] = (float)(t1 - t0) / REP;
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// throughput of single instructions on the whole chip: 16 warps per SM, U independent chains per thread
template <int OP>
__global__ void __launch_bounds__(512, 1) k_ops(double* out, long long* cyc, int iters) {
  constexpr int U = 8;
  double d[U];
  long long q[U];
  int w[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    d[u] = 1.0 + threadIdx.x * 1e-3 + u;
    q[u] = (long long)threadIdx.x * 12345 + u;
    w[u] = threadIdx.x * 7 + u;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (OP == 0) d[u] = fma(d[u], 1.0000001, 1e-9);                                   // DFMA
      if (OP == 1) { asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d[u]) : "r"(w[u])); w[u] += i; }   // I2F.F64.S32
      if (OP == 2) { asm volatile("cvt.rn.f64.s64 %0, %1;" : "=d"(d[u]) : "l"(q[u])); q[u] += i; }   // I2F.F64.S64
      if (OP == 3) { asm volatile("cvt.rni.s32.f64 %0, %1;" : "=r"(w[u]) : "d"(d[u])); d[u] += 1.0; }  // F2I.S32.F64 (+ DADD)
      if (OP == 4) { asm volatile("cvt.rni.s64.f64 %0, %1;" : "=l"(q[u]) : "d"(d[u])); d[u] += 1.0; }  // F2I.S64.F64 (+ DADD)
      if (OP == 5) q[u] = (long long)w[u] * 256 + q[u];                                  // IMAD.WIDE
      if (OP == 6) { float f; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(w[u])); w[u] += __float_as_int(f) & 1; }  // I2F.F32
      if (OP == 7) d[u] = d[u] + 1.0;                                                    // DADD alone
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) s += d[u] + (double)q[u] + w[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

static uint32_t off_for(int blay, uint32_t row, uint32_t g) {
  if (blay == 0) return sw128_off(row, g);
  if (blay == 1) return sw64_off(row, g);
  return sw32_off(row, g);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  CK(cudaFuncSetAttribute(k_i8_check, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));

  // ---- 1. correctness ----
  srand(7);
  const int K = 64;
  for (int mode = 0; mode < 2; ++mode)
    for (int blay = 0; blay < 3; ++blay)
      for (int N : {16, 64, 80, 128, 256}) {
        std::vector<int8_t> A(128 * K), B(N * K);
        for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
        for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
        std::vector<uint8_t> Aimg(16384, 0), Bimg(32768, 0);
        for (int r = 0; r < 128; ++r)
          for (int k = 0; k < K; ++k) Aimg[sw128_off(r, k >> 4) + (k & 15)] = (uint8_t)A[r * K + k];
        for (int n = 0; n < N; ++n)
          for (int k = 0; k < K; ++k) {
            uint32_t o;
            if (blay == 2) o = (uint32_t)(k >> 5) * (uint32_t)N * 32u + sw32_off(n, (k & 31) >> 4) + (k & 15);
            else o = off_for(blay, n, k >> 4) + (k & 15);
            Bimg[o] = (uint8_t)B[n * K + k];
          }
        uint8_t *dA, *dB;
        int8_t* dAr;
        int32_t* dD;
        CK(cudaMalloc(&dA, 16384));
        CK(cudaMalloc(&dB, 32768));
        CK(cudaMalloc(&dAr, 128 * K));
        CK(cudaMalloc(&dD, 128 * N * 4));
        CK(cudaMemcpy(dA, Aimg.data(), 16384, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, Bimg.data(), 32768, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dAr, A.data(), 128 * K, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0xff, 128 * N * 4));
        k_i8_check<<<1, 160, SMEM>>>(dA, dB, dAr, dD, N, mode, blay);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("check mode %d blay %d N %d: CUDA error %s\n", mode, blay, N, cudaGetErrorString(e));
          return 1;
        }
        std::vector<int32_t> D(128 * N);
        CK(cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
        long bad = 0;
        int first = -1;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < N; ++n) {
            int32_t ref = 0;
            for (int k = 0; k < K; ++k) ref += (int32_t)A[r * K + k] * (int32_t)B[n * K + k];
            if (ref != D[r * N + n]) {
              if (first < 0) first = r * N + n;
              ++bad;
            }
          }
        printf("check %s B=%s N=%3d: %ld mismatches of %d", mode ? "TS" : "SS", blay == 0 ? "sw128" : blay == 1 ? "sw64" : "sw32", N, bad,
               128 * N);
        if (bad) printf(" (first at row %d col %d: got %d)", first / N, first % N, D[first]);
        printf("\n");
        cudaFree(dA); cudaFree(dB); cudaFree(dAr); cudaFree(dD);
      }

  // ---- 2. MMA rate ----
  float* dout;
  CK(cudaMalloc(&dout, 16));
  auto rate = [&](int N, int nacc, int dstep, int form, int kind, int blay) {
    k_rate<<<1, 160, SMEM>>>(dout, N, nacc, dstep, form, kind, blay);
    CK(cudaDeviceSynchronize());
    k_rate<<<1, 160, SMEM>>>(dout, N, nacc, dstep, form, kind, blay);
    CK(cudaDeviceSynchronize());
    float h[2];
    CK(cudaMemcpy(h, dout, 8, cudaMemcpyDeviceToHost));
    printf("rate kind=%s form=%s B=%s N=%3d nacc=%d dstep=%3d: %6.1f cycles/MMA (issue %5.1f)\n", kind ? "tf32" : "i8",
           form ? "TS" : "SS", blay == 0 ? "sw128" : blay == 1 ? "sw64" : "sw32", N, nacc, dstep, h[0], h[1]);
  };
  for (int kind = 0; kind < 2; ++kind)
    for (int form = 0; form < 2; ++form)
      for (int N : {16, 32, 64, 96, 128, 192, 256})
        for (int nacc : {1, 2, 4}) {
          int dstep = N <= 128 ? 128 : 256;
          if (form == 1 && N > 128) dstep = 0;  // A operand sits at columns 384..; keep D below it
          if (nacc * (dstep ? dstep : N) > (form == 1 ? 384 : 512)) continue;
          if (dstep == 0 && nacc > 1) continue;
          rate(N, nacc, dstep, form, kind, 0);
        }
  for (int blay = 1; blay < 3; ++blay)
    for (int N : {64, 128}) rate(N, 4, 128, 0, 0, blay);
  // accumulators packed tightly (levels N columns apart)
  for (int N : {32, 64, 96}) rate(N, 4, N, 0, 0, 0);

  {
    const int SMEM2 = 1024 + 98304 + 65536 + 65536 + 64;
    CK(cudaFuncSetAttribute(k_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2));
    auto rate2 = [&](int N, int ls, int alay, uint32_t apl, int commits, int order, int U, uint32_t aofs = 0) {
      k_rate2<<<1, 160, SMEM2>>>(dout, N, ls, alay, apl, commits, order, U, aofs);
      CK(cudaDeviceSynchronize());
      k_rate2<<<1, 160, SMEM2>>>(dout, N, ls, alay, apl, commits, order, U, aofs);
      CK(cudaDeviceSynchronize());
      float h[2];
      CK(cudaMemcpy(h, dout, 8, cudaMemcpyDeviceToHost));
      printf("unit pattern N=%3d ls=%3d A=%s at %6u planes %5u B apart commits=%d order=%d units=%2d: %6.1f cycles/MMA (issue %5.1f)\n", N, ls,
             alay ? "sw64 " : "sw128", aofs, apl, commits, order, U, h[0], h[1]);
    };
    for (int N : {32, 64, 128})
      for (int alay = 0; alay < 2; ++alay)
        for (int order = 0; order < 2; ++order) rate2(N, N <= 64 ? 64 : 128, alay, alay ? 24576u : 32768u, 1, order, 6);
    rate2(128, 128, 1, 24576u, 0, 0, 6);
    rate2(128, 128, 1, 16384u, 1, 0, 6);
    rate2(64, 64, 1, 16384u, 1, 0, 4);
    rate2(64, 64, 1, 24576u, 1, 0, 4);
    rate2(64, 64, 1, 24576u, 1, 0, 5);
    rate2(64, 64, 1, 24576u, 1, 0, 1);
    rate2(128, 128, 1, 24576u, 1, 0, 1);
    rate2(128, 128, 1, 24576u, 1, 0, 2);
    rate2(128, 128, 1, 24576u, 1, 0, 12);
    for (uint32_t aofs : {0u, 32768u, 65536u, 98304u, 131072u})
      for (int N : {64, 128}) rate2(N, N <= 64 ? 64 : 128, 1, 16384u, 1, 0, 4, aofs);
  }
  // ---- 3. TMEM load throughput ----
  uint32_t* sink;
  CK(cudaMalloc(&sink, 1024 * 4));
  for (int wq = 1; wq <= 4; ++wq) {
    k_tmem_ld<16><<<1, 544>>>(dout, sink, wq, 0);
    CK(cudaDeviceSynchronize());
    float h;
    CK(cudaMemcpy(&h, dout, 4, cudaMemcpyDeviceToHost));
    printf("tmem ld x16 %d warps/quadrant: %7.0f cycles per 128 x 512 sweep (%.1f B/cycle)\n", wq, h, 262144.0 / h);
    k_tmem_ld<16><<<1, 544>>>(dout, sink, wq, 1);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, dout, 4, cudaMemcpyDeviceToHost));
    printf("tmem ld x16 pipelined %d warps/quadrant: %7.0f cycles (%.1f B/cycle)\n", wq, h, 262144.0 / h);
    k_tmem_ld<32><<<1, 544>>>(dout, sink, wq, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, dout, 4, cudaMemcpyDeviceToHost));
    printf("tmem ld x32 %d warps/quadrant: %7.0f cycles (%.1f B/cycle)\n", wq, h, 262144.0 / h);
  }

  // ---- 4. instruction throughput ----
  double* dd;
  long long* cyc;
  CK(cudaMalloc(&dd, 148 * 512 * 8));
  CK(cudaMallocManaged(&cyc, 8));
  const char* names[] = {"DFMA", "I2F.F64.S32", "I2F.F64.S64", "F2I.S32.F64+DADD", "F2I.S64.F64+DADD", "IMAD.WIDE", "I2F.F32.S32", "DADD"};
  const int iters = 2000;
#define RUN_OP(OP)                                                                                               \
  k_ops<OP><<<148, 512>>>(dd, cyc, iters);                                                                       \
  CK(cudaDeviceSynchronize());                                                                                   \
  printf("op %-18s: %.2f thread-ops per cycle per SM\n", names[OP], 512.0 * 8 * iters / (double)*cyc);
  RUN_OP(0) RUN_OP(1) RUN_OP(2) RUN_OP(3) RUN_OP(4) RUN_OP(5) RUN_OP(6) RUN_OP(7)
  printf("done\n");
  return 0;
}
