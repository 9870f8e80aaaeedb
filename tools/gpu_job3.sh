#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint" > $O/t3_default.log 2>&1; echo "default rc=$?" >> $O/t3_default.log; tail -n 3 $O/t3_default.log
NBK_PAINT_BUCKET=coherent timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled" > $O/t3_coh.log 2>&1; echo "coherent rc=$?" >> $O/t3_coh.log; tail -n 3 $O/t3_coh.log
P="timeout 300 python tools/paint_bench.py 1e8 512 cic f8"
$P --check > $O/pb3_512.log 2>&1
NBK_PAINT_SPREAD=0 $P --only-sorted >> $O/pb3_512.log 2>&1
NBK_PAINT_THREADS=256 $P --only-sorted >> $O/pb3_512.log 2>&1
NBK_PAINT_THREADS=1024 $P --only-sorted >> $O/pb3_512.log 2>&1
grep -v "sum =\|identical" $O/pb3_512.log
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_bucket|k_tile -c 8 --csv"
timeout 600 $NCU --log-file $O/l3_default.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > /dev/null 2>&1
python tools/ncu_summary.py $O/l3_default.csv 8
timeout 600 python tools/paint_bench.py 1e9 1024 cic f8 > $O/pb3_1024.log 2>&1
grep -v "sum =\|identical" $O/pb3_1024.log
timeout 600 python tools/paint_bench.py 1e8 512 tsc f4 --only-sorted > $O/pb3_tsc.log 2>&1; grep -v "sum =\|identical" $O/pb3_tsc.log
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu > $O/t3_configs.log 2>&1; echo "configs rc=$?" >> $O/t3_configs.log; tail -n 15 $O/t3_configs.log
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 > $O/bench3_c2.json 2> $O/bench3_c2.err; tail -c 3000 $O/bench3_c2.json; tail -n 5 $O/bench3_c2.err
