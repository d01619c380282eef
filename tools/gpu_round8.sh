#!/bin/bash
# last 1-GPU pass: verify the compact index search, A/B the lookup kernel and the read path
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -x -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_round8.txt
for side in new old new old; do
  if [ $side = old ]; then dir=build/ab_old; else dir=.; fi
  echo "== $side lookup"; (cd $dir && timeout 300 python bench/lookup_latency.py 2>&1 | tail -1)
done
echo "== new sweep"; timeout 600 python bench/api_sweep.py --pool 0 --iters 2 > gpurun_out/api_sweep_round8.log 2>&1; tail -8 gpurun_out/api_sweep_round8.log | cut -c1-175
echo "== bench N=1"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_round8.log 2>&1; tail -1 gpurun_out/bench_n1_round8.log | cut -c1-160
