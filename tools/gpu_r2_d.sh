#!/bin/bash
# round 2, GPU call D (2 GPUs): full GPU suite, smoke on a peer pool, bench N=2 and N=1 with
# two ring geometries
mkdir -p gpurun_out
step() {
    local name=$1 secs=$2; shift 2
    timeout -k 10 "$secs" "$@" > "gpurun_out/r2d_$name.txt" 2> "gpurun_out/r2d_$name.err"
    local rc=$?
    echo "== $name rc=$rc"; tail -3 "gpurun_out/r2d_$name.txt" | cut -c1-1500
    if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "ABORT: $name timed out"; tail -5 "gpurun_out/r2d_$name.err"; exit 1; fi
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
step pytest 300 python -m pytest tests -m gpu -q
step smoke 90 python __graft_entry__.py smoke
step bench_n2 400 $TR --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 1
step bench_n2_r64 300 $TR --master-port 29527 bench.py --gpus 2 --steps 4 --warmup 1 --stage-kb 16 --ring-kb 64 --no-extra --no-e2e
step bench_n1_r64 200 python bench.py --gpus 1 --steps 4 --warmup 1 --stage-kb 16 --ring-kb 64 --no-extra --no-e2e
step bench_n1_def 200 python bench.py --gpus 1 --steps 4 --warmup 1 --no-extra --no-e2e
