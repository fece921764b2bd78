#!/bin/bash
# ncu --set full of one stage-3 launch at C4 / 65 536 walkers (run under gpurun); layout from $1
export C1D_P1D_LAYOUT=${1:-split}
K=${2:-k_p1d_split}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/s3_$1 python tools/time_stage3.py > gpurun_out/s3_$1.log 2>&1
ncu -i gpurun_out/s3_$1.ncu-rep --page raw --csv > gpurun_out/s3_$1_raw.csv
ncu -i gpurun_out/s3_$1.ncu-rep --page source --csv > gpurun_out/s3_$1_src.csv 2>/dev/null
python tools/ncu_condense.py gpurun_out/s3_$1_raw.csv gpurun_out/s3_$1_cond.csv
cat gpurun_out/s3_$1_cond.csv
