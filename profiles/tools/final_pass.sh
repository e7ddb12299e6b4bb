#!/bin/bash
# the measurement pass of a kernel version: ncu of every workload, launch list of the bench command, the bench lines
bash profiles/tools/ncu_all.sh
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_lower_limb.json 2> gpurun_out/r02_bench_err.log && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_err.log
for w in pendulum_force knee full_body n_link_4; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02_bench_$w.json 2>> gpurun_out/r02_bench_err.log
done
python profiles/tools/show_sweep.py gpurun_out/r02_bench_*.json
tail -3 gpurun_out/r02_bench_err.log
