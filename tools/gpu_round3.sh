#!/bin/bash
# tests + baselines + ncu capture of the top kernel + SASS listing
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== baselines"; timeout 300 python bench/baselines.py --iters 2 2>&1 | tail -4
timeout 300 python bench/baselines.py --iters 2 --block-kb 32 --size-mb 128 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench/baselines.py --nccl --iters 2 2>&1 | grep baseline
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log | cut -c1-200
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:kv_copy_ldst -s 40 -c 3 -o gpurun_out/prof_kv_copy python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kv_read_fused -s 10 -c 2 -o gpurun_out/prof_kv_read python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep 2>/dev/null
