#!/bin/bash
# round 2, 8-GPU verification: flagship + extras at N=8, reference arm at N=8
mkdir -p gpurun_out
step() {
    local name=$1 secs=$2; shift 2
    timeout -k 10 "$secs" "$@" > "gpurun_out/r2n8_$name.txt" 2> "gpurun_out/r2n8_$name.err"
    local rc=$?
    echo "== $name rc=$rc"; tail -2 "gpurun_out/r2n8_$name.txt" | cut -c1-600; [ $rc -ne 0 ] && tail -3 "gpurun_out/r2n8_$name.err" | cut -c1-400
    if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "ABORT: $name timed out"; exit 1; fi
}
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
step b200_n8 500 $TR --master-port 29517 bench.py --gpus 8 --steps 6 --warmup 2
step ref_n8 400 $TR --master-port 29557 bench.py --impl reference --gpus 8 --steps 3 --warmup 1
