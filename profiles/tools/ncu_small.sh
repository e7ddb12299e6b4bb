#!/bin/bash
# ncu --set full of one of the small kernels through post_sweep.py: ncu_small.sh <kernel regex> <post_sweep filter> <tag>
POST_SWEEP_ONLY="$2" python profiles/post_sweep.py 300000 > /dev/null 2>&1 && \
POST_SWEEP_ONLY="$2" ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -o /tmp/small_$3 -f python profiles/post_sweep.py 300000 > gpurun_out/ncu_small_$3.log 2>&1
ncu -i /tmp/small_$3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$3_raw.csv 2>/dev/null
ncu -i /tmp/small_$3.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02_ncu_$3_source.csv 2>/dev/null
rm -f /tmp/small_$3.ncu-rep
