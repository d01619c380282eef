#!/bin/bash
# round 2, GPU call F (2 GPUs): NVLS flags test, bench after the reply-before-apply change,
# ncu profiles of every kernel on peer memory with NVLink counters
mkdir -p gpurun_out
step() {
    local name=$1 secs=$2; shift 2
    timeout -k 10 "$secs" "$@" > "gpurun_out/r2f_$name.txt" 2> "gpurun_out/r2f_$name.err"
    local rc=$?
    echo "== $name rc=$rc"; tail -2 "gpurun_out/r2f_$name.txt" | cut -c1-400
    if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "ABORT: $name timed out"; exit 1; fi
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="--steps 4 --warmup 1 --no-extra --no-e2e"
step pytest 300 python -m pytest tests/test_gpu_features.py tests/test_gpu_store.py -q
step driver 120 python bench/ncu_driver.py
step n2 300 $TR --master-port 29517 bench.py --gpus 2 $B
step n1 200 python bench.py --gpus 1 $B
step n2_full 400 $TR --master-port 29527 bench.py --gpus 2 --steps 4 --warmup 1
step ncu_full 900 ncu --set full --section Nvlink --clock-control none --import-source on -k regex:'kv_' -o gpurun_out/r2_prof_all python bench/ncu_driver.py
ls -la gpurun_out/*.ncu-rep
