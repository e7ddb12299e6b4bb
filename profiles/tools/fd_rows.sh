#!/bin/bash
# forward_dynamics: bulk-copy rows kernel against the per-lane kernel, and its occupancy variants
python -m pytest tests/test_gpu_scale.py -q -x -k "bulk_copy or poison or written" 2>&1 | tail -3
for v in ""; do
  echo "== variant '$v'"
  POST_SWEEP_ONLY=forward_dynamics BIONC_CUDA_LIB=$PWD/bionc_b200/lib/libbionc_cuda$v.so python profiles/post_sweep.py 2>/dev/null > gpurun_out/fdr$v.json; python profiles/tools/show_sweep.py gpurun_out/fdr$v.json
done
