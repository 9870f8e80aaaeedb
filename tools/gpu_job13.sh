#!/bin/bash
# TMA-pipelined FFT line pass: correctness + per-pass timing vs the register-I/O kernels; paint poll knob; headline launch list
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "r2c or fft or c2r" > $O/t13.log 2>&1; echo "rc=$?" >> $O/t13.log; tail -n 5 $O/t13.log
for m in tma rg smem; do
  for a in "1024 f8" "512 f8" "1024 f4"; do NBK_FFT_LINES=$m timeout 200 python tools/fft_probe.py $a 2>&1 | tail -n 2; done
done
NBK_FFT_TMA_NS=20 timeout 200 python tools/fft_probe.py 1024 f8 2>&1 | tail -n 2
NBK_FFT_TMA_NS=10 timeout 200 python tools/fft_probe.py 512 f8 2>&1 | tail -n 2
NBK_FFT_TMA_NS=16 timeout 200 python tools/fft_probe.py 512 f8 2>&1 | tail -n 2
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 300 python tools/paint_bench.py 1e8 512 cic f8 --check 2>&1 | grep -v "sum =\|identical"
NBK_PAINT_POLL=acquire timeout 300 python tools/paint_bench.py 1e8 512 cic f8 2>&1 | grep -v "sum =\|identical"
timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 2>&1 | grep -v "sum =\|identical"
NBK_PAINT_POLL=acquire timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
echo "elapsed $(( $(date +%s) - T0 )) s"
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv -k regex:^k_ -c 40"
timeout 900 $NCU --log-file $O/r02_launches_headline.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu13a.log 2>&1
python tools/ncu_summary.py $O/r02_launches_headline.csv 60 | tail -n 14
echo "elapsed $(( $(date +%s) - T0 )) s"
