// DMMA (mma.sync m8n8k4 f64) issue-rate probe: cycles per DMMA per SM sub-partition for
// W warps per sub-partition and C independent accumulator chains per warp.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void k(double* out, int iters, long long* cyc) {
  double acc[C][2];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c][0] = acc[c][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < C; ++c)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[c][0]), "+d"(acc[c][1])
                   : "d"(a), "d"(b));
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += acc[c][0] + acc[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int C>
void run(int warps_per_smsp) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaMallocManaged(&cyc, 8);
  int iters = 2000;
  int threads = 128 * warps_per_smsp;
  k<C><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k<C><<<148, threads>>>(out, iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  double per = (double)*cyc / ((double)iters * C * warps_per_smsp);
  printf("warps/SMSP %d chains %2d : %.2f cycles per DMMA per sub-partition (%s)\n", warps_per_smsp, C, per, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w = 1; w <= 4; ++w) {
    run<4>(w);
    run<8>(w);
    run<16>(w);
    run<32>(w);
  }
  return 0;
}
