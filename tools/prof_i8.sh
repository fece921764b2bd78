#!/bin/bash
# ncu --set full of one k_mlp_i8 launch (C4, 65 536 walkers) with the source page (run under gpurun)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mlp_i8 -s 4 -c 1 -f -o gpurun_out/i8_prof python tools/i8_time.py > gpurun_out/i8_prof.log 2>&1
ncu -i gpurun_out/i8_prof.ncu-rep --page raw --csv > gpurun_out/i8_prof_raw.csv
ncu -i gpurun_out/i8_prof.ncu-rep --page source --csv > gpurun_out/i8_prof_src.csv 2>/dev/null
python tools/ncu_condense.py gpurun_out/i8_prof_raw.csv gpurun_out/i8_prof_cond.csv
