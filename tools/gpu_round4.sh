#!/bin/bash
# 1-GPU pass as the driver does it at round end + sanitizer + final ncu captures
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/build.log 2>&1; tail -2 gpurun_out/build.log
echo "== pytest gpu (1 GPU)"; timeout 900 python -m pytest tests -q -x -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -3
echo "== compute-sanitizer memcheck"; timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 600 -p no:cacheprovider > gpurun_out/sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck.log
echo "== compute-sanitizer racecheck"; timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 600 -p no:cacheprovider -k "publish or fp8_write or copy_small" > gpurun_out/sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/sanitizer_racecheck.log
echo "== compute-sanitizer synccheck"; timeout 900 compute-sanitizer --tool synccheck --print-limit 10 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 600 -p no:cacheprovider -k "publish or tma" > gpurun_out/sanitizer_synccheck.log 2>&1; tail -4 gpurun_out/sanitizer_synccheck.log
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 150 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_launch.log 2>&1; grep -c kv_ gpurun_out/launches.csv
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:kv_copy_ldst -s 40 -c 2 -o gpurun_out/prof_kv_copy python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kv_read_fused -s 10 -c 2 -o gpurun_out/prof_kv_read python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kv_copy_tma -s 10 -c 2 -o gpurun_out/prof_kv_tma python bench.py --steps 2 --warmup 1 --no-e2e --variant tma > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo "== bench N=1"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-400
