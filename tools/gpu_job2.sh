#!/bin/bash
# paint v3 validation + bucketing sweep (scratch)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint" > $O/t2_default.log 2>&1; echo "default rc=$?" >> $O/t2_default.log; tail -n 3 $O/t2_default.log
NBK_PAINT_BUCKET=coherent NBK_PAINT_NST=4 NBK_PAINT_THREADS=256 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled" > $O/t2_coh.log 2>&1; echo "coherent rc=$?" >> $O/t2_coh.log; tail -n 3 $O/t2_coh.log
P="timeout 300 python tools/paint_bench.py 1e8 512 cic f8"
$P --check > $O/pb2_512.log 2>&1
NBK_PAINT_STAGED=0 $P --only-sorted >> $O/pb2_512.log 2>&1
NBK_PAINT_STAGED=0 NBK_PAINT_THREADS=1024 $P --only-sorted >> $O/pb2_512.log 2>&1
NBK_PAINT_NST=3 $P --only-sorted >> $O/pb2_512.log 2>&1
NBK_PAINT_NST=4 NBK_PAINT_THREADS=256 $P --only-sorted >> $O/pb2_512.log 2>&1
NBK_PAINT_NST=6 NBK_PAINT_THREADS=256 $P --only-sorted >> $O/pb2_512.log 2>&1
NBK_PAINT_NST=8 NBK_PAINT_THREADS=128 $P --only-sorted >> $O/pb2_512.log 2>&1
NBK_PAINT_SPREAD=0 $P --only-sorted >> $O/pb2_512.log 2>&1
grep -v "sum =\|identical" $O/pb2_512.log
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_bucket|k_tile -c 8 --csv"
timeout 600 $NCU --log-file $O/l2_default.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > /dev/null 2>&1
NBK_PAINT_STAGED=0 timeout 600 $NCU --log-file $O/l2_nostage.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > /dev/null 2>&1
NBK_PAINT_NST=4 NBK_PAINT_THREADS=256 timeout 600 $NCU --log-file $O/l2_nst4.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > /dev/null 2>&1
for f in default nostage nst4; do echo $f; python tools/ncu_summary.py $O/l2_$f.csv 8; done
timeout 600 python tools/paint_bench.py 1e9 1024 cic f8 > $O/pb2_1024.log 2>&1
NBK_PAINT_STAGED=0 timeout 600 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted >> $O/pb2_1024.log 2>&1
NBK_PAINT_NST=4 NBK_PAINT_THREADS=256 timeout 600 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted >> $O/pb2_1024.log 2>&1
grep -v "sum =\|identical" $O/pb2_1024.log
