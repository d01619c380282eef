#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (2 GPUs)"; timeout 600 python -m pytest tests -q -x -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_2gpu_round9.txt
echo "== api sweep nvlink"; timeout 300 python bench/api_sweep.py --pool 1 --iters 2 > gpurun_out/api_sweep_nvlink_round9.log 2>&1; tail -8 gpurun_out/api_sweep_nvlink_round9.log | cut -c1-100
