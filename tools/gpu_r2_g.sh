#!/bin/bash
# round 2, GPU call G (1 GPU): the reference arm (unmodified reference, LOCAL_GPU path)
mkdir -p gpurun_out
step() {
    local name=$1 secs=$2; shift 2
    timeout -k 10 "$secs" "$@" > "gpurun_out/r2g_$name.txt" 2> "gpurun_out/r2g_$name.err"
    local rc=$?
    echo "== $name rc=$rc"; tail -3 "gpurun_out/r2g_$name.txt" | cut -c1-1200; tail -3 "gpurun_out/r2g_$name.err" | cut -c1-600
    if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "ABORT: $name timed out"; cat /tmp/ref_server_0.log | tail -20; exit 1; fi
}
step ref_n1 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
cat /tmp/ref_server_0.log 2>/dev/null | tail -15
step b200_n1 300 python bench.py --gpus 1 --steps 4 --warmup 1
