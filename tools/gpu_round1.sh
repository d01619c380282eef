#!/bin/bash
# First GPU pass: tests, smoke, fabric probe, kernel sweep, flagship bench (1 GPU).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== probe"; timeout 120 python bench/probe_fabric.py > gpurun_out/probe.log 2>&1; tail -30 gpurun_out/probe.log
echo "== sweep"; timeout 600 python bench/sweep_copy.py ${SWEEP_ARGS:-} > gpurun_out/sweep.log 2>&1; tail -80 gpurun_out/sweep.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.log 2>&1; tail -5 gpurun_out/bench1.log
