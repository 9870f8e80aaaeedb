#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint" > $O/t6_default.log 2>&1; echo "default rc=$?" >> $O/t6_default.log; tail -n 3 $O/t6_default.log
NBK_PAINT_BUCKET=coherent timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled" > $O/t6_coh.log 2>&1; echo "coherent rc=$?" >> $O/t6_coh.log; tail -n 3 $O/t6_coh.log
NBK_PAINT_BUCKET=scattered timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled" > $O/t6_sca.log 2>&1; echo "scattered rc=$?" >> $O/t6_sca.log; tail -n 3 $O/t6_sca.log
timeout 900 python -m pytest tests/test_gpu_meshapi.py tests/test_gpu_fftpower.py -q -m gpu > $O/t6_misc.log 2>&1; echo "misc rc=$?" >> $O/t6_misc.log; tail -n 12 $O/t6_misc.log
P="timeout 300 python tools/paint_bench.py 1e8 512 cic f8"
$P --check > $O/pb6_512.log 2>&1
NBK_PAINT_SPREAD=1 $P --only-sorted >> $O/pb6_512.log 2>&1
grep -v "sum =\|identical" $O/pb6_512.log
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_bucket|k_tile -c 8 --csv"
timeout 600 $NCU --log-file $O/l6_default.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > /dev/null 2>&1
python tools/ncu_summary.py $O/l6_default.csv 8
timeout 600 python tools/paint_bench.py 1e9 1024 cic f8 > $O/pb6_1024.log 2>&1
grep -v "sum =\|identical" $O/pb6_1024.log
timeout 600 python tools/paint_bench.py 1e8 512 tsc f4 --only-sorted > $O/pb6_tsc.log 2>&1; grep -v "sum =\|identical" $O/pb6_tsc.log
for cfg in "" "NBK_FFT_LINE_B=4" "NBK_FFT_LINE_B=4 NBK_FFT_LINE_NT=512" "NBK_FFT_LINE_B=8 NBK_FFT_LINE_NT=256" "NBK_FFT_LINE_B=2"; do echo "fft [$cfg]"; env $cfg timeout 120 python tools/fftbench.py 1024 f8 2>&1 | tail -n 2; done
timeout 120 python tools/fftbench.py 1024 f4 2>&1 | tail -n 2
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled_slab_ghosts or tiled_is_order" > $O/racecheck6.log 2>&1; tail -n 3 $O/racecheck6.log
