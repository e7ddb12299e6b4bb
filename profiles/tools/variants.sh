#!/bin/bash
for v in "" _v1 _v2 _v3 _v4; do
  echo "== variant '$v'"
  BIONC_CUDA_LIB=$PWD/bionc_b200/lib/libbionc_cuda$v.so python profiles/tools/rk4_launch.py lower_limb 2>&1 | tail -1 | cut -c1-120
  BIONC_CUDA_LIB=$PWD/bionc_b200/lib/libbionc_cuda$v.so python profiles/tools/rk4_launch.py full_body 2>&1 | tail -1 | cut -c1-120
done
