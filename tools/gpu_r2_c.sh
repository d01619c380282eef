#!/bin/bash
# round 2, GPU call C (2 GPUs): peer-path tests, kernel lab over NVLink, bench N=2
mkdir -p gpurun_out
step() {
    local name=$1 secs=$2; shift 2
    timeout -k 10 "$secs" "$@" > "gpurun_out/r2c_$name.txt" 2> "gpurun_out/r2c_$name.err"
    local rc=$?
    echo "== $name rc=$rc"; tail -4 "gpurun_out/r2c_$name.txt"
    if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "ABORT: $name timed out"; tail -5 "gpurun_out/r2c_$name.err"; exit 1; fi
}
step pytest 300 python -m pytest tests -m gpu -q
step smoke 90 python __graft_entry__.py smoke
step lab 300 python bench/r2_lab.py --out gpurun_out/r2c_lab.json
step bench_n2 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 1
