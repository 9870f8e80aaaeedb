#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/debug_dk0.py > $O/debug_dk0.log 2>&1; tail -n 8 $O/debug_dk0.log
for cfg in "NBK_FFT_PREFETCH=0" "NBK_FFT_PREFETCH=1"; do env $cfg timeout 200 python tools/fft_probe.py 1024 f8 2>&1 | tail -n 2; env $cfg timeout 200 python tools/fft_probe.py 512 f8 2>&1 | tail -n 2; env $cfg timeout 200 python tools/fft_probe.py 1024 f4 2>&1 | tail -n 2; done
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_bucket|k_tile -c 8 --csv"
timeout 600 $NCU --log-file $O/l9_default.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > $O/pb9.log 2>&1
python tools/ncu_summary.py $O/l9_default.csv 8
timeout 300 python tools/paint_bench.py 1e8 512 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fft_lines_rg|k_fft_z_r2c_rg" -s 6 -c 3 -o $O/fft_full_r02 -f python tools/fft_probe.py 1024 f8 > $O/ncu_fft9.log 2>&1; tail -n 2 $O/ncu_fft9.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "fft or r2c or c2r" > $O/t9_fft.log 2>&1; echo "fft rc=$?" >> $O/t9_fft.log; tail -n 3 $O/t9_fft.log
