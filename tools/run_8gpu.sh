#!/bin/bash
# 8-GPU evidence in one gpurun --gpus 8 call: NCCL checks, weak and strong scaling bench lines (C4, C3), the sweep.
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
O=gpurun_out
mkdir -p $O
timeout 300 $T --master-port 29511 tests/dist_gpu_check.py > $O/r2_dist_check_${N}gpu.json 2> $O/r2_dist_${N}.err
timeout 400 $T --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_v10_bench_c4_${N}gpu.json 2> $O/r2_bench_${N}.err
timeout 300 $T --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-cpu-baseline --no-alt --no-sampler > $O/r2_v10_bench_c4_${N}gpu_strong.json 2>> $O/r2_bench_${N}.err
timeout 300 $T --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --workload c3 --scaling strong --walkers 65536 --no-cpu-baseline --no-alt --no-sampler > $O/r2_v10_bench_c3_${N}gpu_strong.json 2>> $O/r2_bench_${N}.err
tail -c 600 $O/r2_dist_check_${N}gpu.json; echo; for f in $O/r2_v10_bench_c4_${N}gpu.json $O/r2_v10_bench_c4_${N}gpu_strong.json $O/r2_v10_bench_c3_${N}gpu_strong.json; do python -c "
import json,sys
d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gather_check'), [ (s['ensembles'], s['evals_per_s']) for s in (d.get('sampler') or [])])"; done; tail -3 $O/r2_bench_${N}.err
