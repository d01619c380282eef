#!/bin/bash
# final 1-GPU pass of the round: driver-style build+smoke, tests, churn benchmark, flagship
# bench, and fresh ncu captures of the two hot kernels (their index code changed)
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -x -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_round6.txt
echo "== churn"; timeout 600 python bench/evict_churn.py > gpurun_out/evict_churn.log 2>&1; tail -32 gpurun_out/evict_churn.log
echo "== bench N=1"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_round6.log 2>&1; tail -1 gpurun_out/bench_n1_round6.log | cut -c1-250
echo "== ncu"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:kv_copy_ldst -s 40 -c 1 -o gpurun_out/prof_kv_copy_r6 python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_r6a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kv_read_fused -s 10 -c 1 -o gpurun_out/prof_kv_read_r6 python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_r6b.log 2>&1
ls -la gpurun_out/*_r6.ncu-rep
