#!/bin/bash
# usage: gpu_configs.sh N   -- BASELINE configs 2..5 at N GPUs (+ scaling points below N)
set -u
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
for n in 1 2 4 8; do
  [ "$n" -gt "$N" ] && break
  echo "== bench N=$n"
  if [ "$n" -eq 1 ]; then timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1; fi
  tail -1 gpurun_out/bench_n$n.log | grep -o "\"value\": [0-9.]*\|\"verified\": [a-z]*\|write_phase_GBps\": [0-9.]*\|read_phase_GBps\": [0-9.]*" | tr '\n' ' '; echo
done
echo "== config3 fanin"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench/configs.py fanin --pages 32 2>&1 | grep '^{' | tee gpurun_out/config3_fanin_n$N.json
echo "== config4 bcast"; timeout 600 python bench/configs.py bcast 2>&1 | grep '^{' | tee gpurun_out/config4_bcast_n$N.json
echo "== config5 fp8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29722 bench/configs.py fp8 2>&1 | grep '^{' | tee gpurun_out/config5_fp8_n$N.json
