#!/bin/bash
# ncu --set full of the fused integrator at the bench configuration (1e6 systems x 100 steps), one capture per workload.
# The reports stay on the GPU box (gpurun brings back at most 64 MiB): only their raw-page CSV comes home, plus the
# per-source-line page of the headline workload.
# RK4_ARGS: extra arguments of rk4_launch.py (e.g. "--spec" for the model-specialised kernels, "--batch=250000")
for w in ${@:-lower_limb pendulum_force knee full_body n_link_4}; do
  python profiles/tools/rk4_launch.py $w $RK4_ARGS > gpurun_out/r02_launch_$w.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:rk4_kernel -s 3 -c 1 -o /tmp/r02_ncu_$w -f python profiles/tools/rk4_launch.py $w $RK4_ARGS > gpurun_out/r02_ncu_$w.log 2>&1
  tail -1 gpurun_out/r02_launch_$w.log | cut -c1-300
  ncu -i /tmp/r02_ncu_$w.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${w}_raw.csv 2>/dev/null
  if [ "$w" = "lower_limb" ]; then
    ncu -i /tmp/r02_ncu_$w.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02_ncu_${w}_source.csv 2>/dev/null
  fi
  rm -f /tmp/r02_ncu_$w.ncu-rep
done
