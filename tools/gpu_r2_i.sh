#!/bin/bash
# round 2, GPU call I (2 GPUs): full GPU suite with the sharded index, bench N=2, reference arm with e2e
mkdir -p gpurun_out
step() {
    local name=$1 secs=$2; shift 2
    timeout -k 10 "$secs" "$@" > "gpurun_out/r2i_$name.txt" 2> "gpurun_out/r2i_$name.err"
    local rc=$?
    echo "== $name rc=$rc"; tail -3 "gpurun_out/r2i_$name.txt" | cut -c1-700; [ $rc -ne 0 ] && tail -3 "gpurun_out/r2i_$name.err" | cut -c1-400
    if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "ABORT: $name timed out"; exit 1; fi
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
step pytest 400 python -m pytest tests -m gpu -q
step smoke 90 python __graft_entry__.py smoke
step n2 300 $TR --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 1 --no-extra --no-e2e
step ref_n1 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
step lat 200 $TR --master-port 29567 bench/configs.py latency
