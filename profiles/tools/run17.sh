#!/bin/bash
python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
python profiles/post_sweep.py > gpurun_out/r02_post_sweep_d.json 2>gpurun_out/r02_post_err.log; python profiles/tools/show_sweep.py gpurun_out/r02_post_sweep_d.json
python profiles/config_sweep.py --steps-per-launch 100 --launches 3 --out gpurun_out/r02_config_sweep_b.json > gpurun_out/r02_cs_b.log 2>&1; python profiles/tools/show_sweep.py gpurun_out/r02_cs_b.log
