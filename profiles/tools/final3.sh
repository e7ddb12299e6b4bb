#!/bin/bash
# closing measurement pass of the round: GPU suite, bench lines of every workload (general-path ones with their
# model-specialised kernels), configuration sweep, launch list of the bench command, reference arm
python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_gpu.log 2>&1; tail -3 gpurun_out/r02_tests_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_lower_limb.json 2> gpurun_out/r02_bench_err.log
for w in pendulum_force knee n_link_4 full_body; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02_bench_$w.json 2>> gpurun_out/r02_bench_err.log
done
python profiles/config_sweep.py --out gpurun_out/r02_config_sweep.json > gpurun_out/r02_config_sweep.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_err.log
python profiles/tools/show_sweep.py gpurun_out/r02_bench_*.json
tail -3 gpurun_out/r02_bench_err.log
RK4_ARGS="--spec --batch=250000" bash profiles/tools/ncu_all.sh n_link_4
