#!/bin/bash
# DRAM bytes of one k_chi2_dmma launch at C4 / 65 536 walkers for several group sizes (run under gpurun)
for g in 32 64 128 100000; do
  C1D_CHI2_GROUP=$g timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_chi2_dmma -s 4 -c 1 --csv python tools/time_stage3.py 2>/dev/null | grep -i "k_chi2_dmma" | awk -F'","' -v g=$g '{print g, $(NF-2), $(NF-1), $NF}'
done
