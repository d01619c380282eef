#!/bin/bash
# round 2, GPU call H (2 GPUs): reference arm at N=1 and N=2, fp8 kernel lab, bench N=2 full
mkdir -p gpurun_out
step() {
    local name=$1 secs=$2; shift 2
    timeout -k 10 "$secs" "$@" > "gpurun_out/r2h_$name.txt" 2> "gpurun_out/r2h_$name.err"
    local rc=$?
    echo "== $name rc=$rc"; tail -3 "gpurun_out/r2h_$name.txt" | cut -c1-900; tail -2 "gpurun_out/r2h_$name.err" | cut -c1-300
    if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "ABORT: $name timed out"; tail -20 /tmp/ref_server_0.log; exit 1; fi
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
step ref_n1 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
tail -5 /tmp/ref_server_0.log
step ref_n2 600 $TR --master-port 29557 bench.py --impl reference --gpus 2 --steps 3 --warmup 1
step lab_fp8 200 python bench/r2_lab.py --only fp8 --out gpurun_out/r2h_lab_fp8.json
step pytest 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_features.py -q -k "fp8 or multi or swizzle"
step n2_full 400 $TR --master-port 29527 bench.py --gpus 2 --steps 4 --warmup 1
