#!/bin/bash
# A/B on ONE box: previous commit (build/ab_old, linear-probing index) vs this tree
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
for side in new old new old; do
  if [ $side = old ]; then dir=build/ab_old; else dir=.; fi
  echo "== $side sweep"; (cd $dir && timeout 600 python bench/api_sweep.py --pool 0 --iters 3 2>&1 | tail -8 | head -5 | cut -c1-95)
  echo "== $side bench"; (cd $dir && timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-120)
done
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -x -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_round5.txt
