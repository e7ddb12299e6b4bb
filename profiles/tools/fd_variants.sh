#!/bin/bash
for v in "" _fd3 _fd2; do
  echo "== fd variant '$v'"
  POST_SWEEP_ONLY=forward_dynamics BIONC_CUDA_LIB=$PWD/bionc_b200/lib/libbionc_cuda$v.so python profiles/post_sweep.py 2>/dev/null > gpurun_out/fdv$v.json; python profiles/tools/show_sweep.py gpurun_out/fdv$v.json
done
python profiles/post_sweep.py > gpurun_out/r02_post_sweep_c.json 2>gpurun_out/r02_post_err.log; python profiles/tools/show_sweep.py gpurun_out/r02_post_sweep_c.json
python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "component or oracle or random" 2>&1 | tail -3
