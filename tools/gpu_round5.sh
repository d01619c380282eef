#!/bin/bash
# 1-GPU pass after the index redesign (bucketed index + eviction): build as the driver does,
# tests, epilogue trace, flagship bench and the API sweep (regression vs profiles/r1_*_final)
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build.log 2>&1; tail -2 gpurun_out/build.log
echo "== pytest gpu (1 GPU)"; timeout 900 python -m pytest tests -q -x -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_round5.txt
echo "== trace"; timeout 300 python bench/trace_tail.py > gpurun_out/trace_round5.log 2>&1; grep -v "^{" gpurun_out/trace_round5.log | tail -8
echo "== bench N=1"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_round5.log 2>&1; tail -1 gpurun_out/bench_n1_round5.log | cut -c1-300
echo "== api sweep local"; timeout 600 python bench/api_sweep.py --pool 0 --iters 2 > gpurun_out/api_sweep_round5.log 2>&1; tail -8 gpurun_out/api_sweep_round5.log | cut -c1-120
