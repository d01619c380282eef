#!/bin/bash
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -30 gpurun_out/build.log; }
echo "== bench N=1"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log
echo "== bench N=2 tma"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --variant tma --no-e2e > gpurun_out/bench_n2_tma.log 2>&1; tail -1 gpurun_out/bench_n2_tma.log
echo "== bench N=2 host lookup"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --host-lookup --no-e2e > gpurun_out/bench_n2_host.log 2>&1; tail -1 gpurun_out/bench_n2_host.log
