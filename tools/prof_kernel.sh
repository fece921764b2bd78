#!/bin/bash
# ncu --set full of one launch of a kernel with the source page (run under gpurun):
#   bash tools/prof_kernel.sh <kernel regex> <tag> <skip> <python script and arguments ...>
K=$1; TAG=$2; SKIP=$3; shift 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/$TAG "$@" > gpurun_out/$TAG.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
python tools/ncu_condense.py gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_cond.csv
