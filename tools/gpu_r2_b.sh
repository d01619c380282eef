#!/bin/bash
# round 2, GPU call B (1 GPU): every step under a hard timeout; a step that hangs aborts the
# script at once (a wedged GPU must not burn the budget)
mkdir -p gpurun_out
step() {  # name seconds command...
    local name=$1 secs=$2; shift 2
    timeout -k 10 "$secs" "$@" > "gpurun_out/r2b_$name.txt" 2> "gpurun_out/r2b_$name.err"
    local rc=$?
    echo "== $name rc=$rc"; tail -4 "gpurun_out/r2b_$name.txt"
    if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "ABORT: $name timed out"; tail -5 "gpurun_out/r2b_$name.err"; exit 1; fi
}
step pytest 400 python -m pytest tests -m gpu -x -q
step smoke 120 python __graft_entry__.py smoke
step lab 300 python bench/r2_lab.py --out gpurun_out/r2b_lab.json
step bench_n1 400 python bench.py --gpus 1 --steps 4 --warmup 1
