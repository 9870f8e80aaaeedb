#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "r2c or fft or c2r" > $O/t23.log 2>&1; echo "rc=$?" >> $O/t23.log; tail -n 2 $O/t23.log
for a in "1024 f8" "512 f8" "1024 f4"; do timeout 200 python tools/fft_probe.py $a 2>&1 | tail -n 2; done
NBK_FFT_TMA_L2=0 timeout 200 python tools/fft_probe.py 1024 f8 2>&1 | tail -n 2
NBK_FFT_TMA_L2=0 timeout 200 python tools/fft_probe.py 512 f8 2>&1 | tail -n 2
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv -k regex:^k_ -c 45"
timeout 900 $NCU --log-file $O/r02_launches_headline_final.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu23.log 2>&1
python tools/ncu_summary.py $O/r02_launches_headline_final.csv 60 | tail -n 17
echo "elapsed $(( $(date +%s) - T0 )) s"
