#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint" > $O/t5_default.log 2>&1; echo "default rc=$?" >> $O/t5_default.log; tail -n 3 $O/t5_default.log
NBK_PAINT_BUCKET=coherent timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled" > $O/t5_coh.log 2>&1; echo "coherent rc=$?" >> $O/t5_coh.log; tail -n 3 $O/t5_coh.log
timeout 900 python -m pytest tests/test_gpu_convpower.py tests/test_gpu_meshapi.py tests/test_gpu_fftpower.py tests/test_gpu_lognormal.py tests/test_gpu_recon.py -q -m gpu > $O/t5_misc.log 2>&1; echo "misc rc=$?" >> $O/t5_misc.log; tail -n 25 $O/t5_misc.log
P="timeout 300 python tools/paint_bench.py 1e8 512 cic f8"
$P --check > $O/pb5_512.log 2>&1
grep -v "sum =\|identical" $O/pb5_512.log
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_bucket|k_tile -c 8 --csv"
timeout 600 $NCU --log-file $O/l5_default.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > /dev/null 2>&1
python tools/ncu_summary.py $O/l5_default.csv 8
timeout 600 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted > $O/pb5_1024.log 2>&1
grep -v "sum =\|identical" $O/pb5_1024.log
timeout 600 python tools/paint_bench.py 1e8 512 tsc f4 --only-sorted > $O/pb5_tsc.log 2>&1; grep -v "sum =\|identical" $O/pb5_tsc.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled_slab_ghosts or tiled_is_order" > $O/racecheck5.log 2>&1; tail -n 3 $O/racecheck5.log
