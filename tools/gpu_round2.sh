#!/bin/bash
# 2-GPU pass: tests (incl. 2-GPU cases), host overhead, peer sweep, flagship bench at N=1,2.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== probe"; timeout 120 python bench/probe_fabric.py > gpurun_out/probe.log 2>&1; tail -12 gpurun_out/probe.log
echo "== host overhead"; timeout 300 python bench/host_overhead.py > gpurun_out/host_overhead.log 2>&1; tail -8 gpurun_out/host_overhead.log
echo "== sweep"; timeout 900 python bench/sweep_copy.py --peer ${SWEEP_ARGS:-} > gpurun_out/sweep.log 2>&1; grep -v "^$" gpurun_out/sweep.log | tail -150
echo "== bench N=1"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -3 gpurun_out/bench_n1.log
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; tail -3 gpurun_out/bench_n2.log
