// DFMA issue-rate / latency probe: cycles per warp-wide DFMA per SM sub-partition for W warps
// per sub-partition and C independent dependency chains per warp.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/dfma_probe tools/dfma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void k(double* out, int iters, long long* cyc, double a, double b) {
  double acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = threadIdx.x * 1e-3 + c;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fma(acc[c], a, b);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += acc[c];
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
  k<C><<<148, threads>>>(out, iters, cyc, 0.999999, 1e-7);
  cudaDeviceSynchronize();
  k<C><<<148, threads>>>(out, iters, cyc, 0.999999, 1e-7);
  cudaError_t e = cudaDeviceSynchronize();
  double per = (double)*cyc / ((double)iters * 8 * C * warps_per_smsp);
  printf("warps/SMSP %d chains %d : %.2f cycles per DFMA per sub-partition (%s)\n", warps_per_smsp, C, per, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w = 1; w <= 8; w *= 2) {
    run<1>(w);
    run<2>(w);
    run<4>(w);
    run<8>(w);
  }
  return 0;
}
