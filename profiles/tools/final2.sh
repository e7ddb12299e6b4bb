#!/bin/bash
# closing pass of a kernel version without ncu --set full: tests, bench lines of every workload (executed-flop roofline
# from profiles/ncu_constants.json), the launch list of the bench command, reference arm, evaluator sweep, smoke
python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_gpu.log 2>&1; tail -2 gpurun_out/r02_tests_gpu.log
python bench.py > gpurun_out/r02_bench_lower_limb.json 2> gpurun_out/r02_bench_err.log && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_err.log
for w in pendulum_force knee full_body n_link_4; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02_bench_$w.json 2>> gpurun_out/r02_bench_err.log
done
python profiles/tools/show_sweep.py gpurun_out/r02_bench_*.json
python profiles/post_sweep.py > gpurun_out/r02_post_sweep.json 2>> gpurun_out/r02_bench_err.log; python profiles/tools/show_sweep.py gpurun_out/r02_post_sweep.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
tail -3 gpurun_out/r02_bench_err.log
