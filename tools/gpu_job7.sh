#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint" > $O/t7_default.log 2>&1; echo "default rc=$?" >> $O/t7_default.log; tail -n 3 $O/t7_default.log
NBK_PAINT_BUCKET=coherent timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled" > $O/t7_coh.log 2>&1; echo "coherent rc=$?" >> $O/t7_coh.log; tail -n 3 $O/t7_coh.log
timeout 900 python -m pytest tests/test_gpu_meshapi.py -q -m gpu > $O/t7_misc.log 2>&1; echo "misc rc=$?" >> $O/t7_misc.log; tail -n 8 $O/t7_misc.log
P="timeout 300 python tools/paint_bench.py 1e8 512 cic f8"
NBK_PAINT_SPREAD=0 $P --check > $O/pb7_512.log 2>&1
NBK_PAINT_SPREAD=1 $P --only-sorted >> $O/pb7_512.log 2>&1
grep -v "sum =\|identical" $O/pb7_512.log
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_bucket|k_tile -c 8 --csv"
NBK_PAINT_SPREAD=0 timeout 600 $NCU --log-file $O/l7_default.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > /dev/null 2>&1
python tools/ncu_summary.py $O/l7_default.csv 8
NBK_PAINT_SPREAD=0 timeout 600 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted > $O/pb7_1024.log 2>&1
grep -v "sum =\|identical" $O/pb7_1024.log
NBK_PAINT_SPREAD=0 timeout 600 python tools/paint_bench.py 1e8 512 tsc f4 --only-sorted > $O/pb7_tsc.log 2>&1; grep -v "sum =\|identical" $O/pb7_tsc.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled_slab_ghosts or tiled_is_order" > $O/racecheck7.log 2>&1; tail -n 3 $O/racecheck7.log
