#!/bin/bash
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e "${@:2}" 2>&1 | grep -o "\"value\": [0-9.]*\|write_phase_GBps\": [0-9.]*\|read_phase_GBps\": [0-9.]*" | tr '\n' ' '; echo " <- ${@:2}"; }
run 29601 --streams 4
run 29602 --streams 8
run 29603 --streams 8 --max-ctas 148
run 29604 --streams 4 --max-ctas 148
run 29605 --streams 2 --max-ctas 592
run 29606 --streams 1
run 29607 --streams 8 --max-ctas 74
run 29608 --streams 6 --max-ctas 222
