#!/bin/bash
# A/B: corner rotation 7 (default lib) / 0 / 1 (alternate builds) + scatter diet (all builds); at both sizes
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
C=nbodykit_b200/csrc
cp $C/libnbk_b200.so /tmp/lib_rot7.so
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint or tiled or interlace" > $O/t19.log 2>&1; echo "rc=$?" >> $O/t19.log; tail -n 2 $O/t19.log
for v in 7 0 1; do
  if [ $v != 7 ]; then cp $C/libnbk_b200_rot$v.so $C/libnbk_b200.so; fi
  echo "== ROT=$v"
  timeout 300 python tools/paint_bench.py 1e8 512 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
  timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
done
cp /tmp/lib_rot7.so $C/libnbk_b200.so
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv -k regex:k_bucket|k_tile|k_apply -c 9"
timeout 600 $NCU --log-file $O/l19_paint.csv python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted > $O/ncu19.log 2>&1
python tools/ncu_summary.py $O/l19_paint.csv 9
echo "elapsed $(( $(date +%s) - T0 )) s"
