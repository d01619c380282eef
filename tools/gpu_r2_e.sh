#!/bin/bash
# round 2, GPU call E (2 GPUs): A/B of ring geometry with one CTA per SM per launch
mkdir -p gpurun_out
step() {
    local name=$1 secs=$2; shift 2
    timeout -k 10 "$secs" "$@" > "gpurun_out/r2e_$name.txt" 2> "gpurun_out/r2e_$name.err"
    local rc=$?
    echo "== $name rc=$rc"; tail -1 "gpurun_out/r2e_$name.txt" | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['breakdown']['write_phase_GBps'], d['breakdown']['read_phase_GBps'], d['roofline'])
except Exception as e: print('unparsed', e)"
    if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "ABORT: $name timed out"; exit 1; fi
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="--steps 4 --warmup 1 --no-extra --no-e2e"
step n2_r64 300 $TR --master-port 29517 bench.py --gpus 2 $B
step n2_r128 300 $TR --master-port 29527 bench.py --gpus 2 $B --stage-kb 32 --ring-kb 128
step n2_r64_s8 300 $TR --master-port 29537 bench.py --gpus 2 $B --streams 8
step n2_r48 300 $TR --master-port 29547 bench.py --gpus 2 $B --stage-kb 16 --ring-kb 48
step n1_r64 200 python bench.py --gpus 1 $B
step n1_r128 200 python bench.py --gpus 1 $B --stage-kb 32 --ring-kb 128
step n1_r64_s8 200 python bench.py --gpus 1 $B --streams 8
