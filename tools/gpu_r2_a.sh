#!/bin/bash
# round 2, GPU call A (2 GPUs): tests of the new kernels, kernel lab, first bench numbers
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.txt 2>&1; tail -2 gpurun_out/r2a_smoke.txt
timeout 600 python bench/r2_lab.py --out gpurun_out/r2a_lab.json > gpurun_out/r2a_lab.txt 2>&1
echo "lab rc=$?"; tail -3 gpurun_out/r2a_lab.txt
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
echo "bench1 rc=$?"; tail -c 600 gpurun_out/r2a_bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2a_bench_n2.json 2> gpurun_out/r2a_bench_n2.err
echo "bench2 rc=$?"; tail -c 600 gpurun_out/r2a_bench_n2.json
