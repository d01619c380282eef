#!/bin/bash
# 2-GPU pass: NVLink paths with the bucketed index (peer claims / lookups), bench N=2
set -u
mkdir -p gpurun_out
echo "== pytest gpu (2 GPUs)"; timeout 900 python -m pytest tests -q -x -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_2gpu_round7.txt
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_round7.log 2>&1; tail -1 gpurun_out/bench_n2_round7.log | cut -c1-250
echo "== api sweep nvlink"; timeout 600 python bench/api_sweep.py --pool 1 --iters 2 > gpurun_out/api_sweep_nvlink_round7.log 2>&1; tail -8 gpurun_out/api_sweep_nvlink_round7.log | cut -c1-175
