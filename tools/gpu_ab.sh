#!/bin/bash
# A/B on ONE box (box-to-box spread is +-10 %): a previous commit unpacked and built under
# build/ab_old (git archive <rev> | tar -x -C build/ab_old; python tools/build_native.py there)
# against this tree, quick bench at N=2 twice, interleaved, and at N=1.
mkdir -p gpurun_out
OUT=$PWD/gpurun_out
run() {  # name dir gpus
    local name=$1 dir=$2 n=$3
    if [ $n -gt 1 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400))"; else L=python; fi
    (cd $dir && timeout -k 10 200 $L bench.py --gpus $n --steps 4 --warmup 3 --no-extra --no-e2e > $OUT/ab_$name.txt 2> $OUT/ab_$name.err)
    rc=$?
    echo "== $name rc=$rc $(tail -1 $OUT/ab_$name.txt | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); b=d['breakdown']; print(d['value'], b['write_phase_GBps'], b['read_phase_GBps'], b['host_issue_write_ms_per_round'], b['host_issue_read_ms_per_round'])
except Exception as e: print('unparsed', e)")"
    if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo ABORT; exit 1; fi
}
for i in 1 2; do
  run old_n2_$i build/ab_old 2
  run new_n2_$i . 2
done
run old_n1 build/ab_old 1
run new_n1 . 1
