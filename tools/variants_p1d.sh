#!/bin/bash
# Build tuning variants of the stage-3 pair kernel into tools/_bin/ (git-ignored, travels with gpurun):
#   tools/variants_p1d.sh name "-DWS_CH_=16 -DWS_UNROLL_=8 -DWS_PREG_=64 -DWS_CREG_=96" [name flags ...]
# run one with C1D_LIB=tools/_bin/libv_<name>.so
cd "$(dirname "$0")/../cup1d_b200/csrc" || exit 1
mkdir -p ../../tools/_bin
while [ $# -ge 2 ]; do
  n=$1; f=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -gencode arch=compute_100a,code=sm_100a $f -c c1d_p1d_lanes.cu -o ../../tools/_bin/p1d_$n.o &&
    /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../tools/_bin/libv_$n.so c1d_abi.o c1d_kernels.o c1d_mlp_f64.o c1d_mlp_tc.o c1d_mlp_i8.o c1d_chi2_i8.o ../../tools/_bin/p1d_$n.o c1d_sampler.o -lcudart -lcuda -ldl && rm ../../tools/_bin/p1d_$n.o && echo built $n ) &
done
wait
