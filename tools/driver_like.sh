mkdir -p gpurun_out
( time timeout -k 10 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/drv_n1.txt 2> gpurun_out/drv_n1.err ) 2> gpurun_out/drv_n1.time; echo "b200 rc=$?"; tail -1 gpurun_out/drv_n1.txt | cut -c1-400; cat gpurun_out/drv_n1.time | grep real
( time timeout -k 10 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/drv_ref1.txt 2> gpurun_out/drv_ref1.err ) 2> gpurun_out/drv_ref1.time; echo "ref rc=$?"; tail -1 gpurun_out/drv_ref1.txt | cut -c1-300; cat gpurun_out/drv_ref1.time | grep real
