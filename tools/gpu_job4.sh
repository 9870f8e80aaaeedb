#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint" > $O/t4_default.log 2>&1; echo "default rc=$?" >> $O/t4_default.log; tail -n 3 $O/t4_default.log
NBK_PAINT_BUCKET=coherent timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled" > $O/t4_coh.log 2>&1; echo "coherent rc=$?" >> $O/t4_coh.log; tail -n 3 $O/t4_coh.log
timeout 900 python -m pytest tests/test_gpu_convpower.py tests/test_gpu_meshapi.py tests/test_gpu_fftpower.py -x -q -m gpu > $O/t4_conv.log 2>&1; echo "conv rc=$?" >> $O/t4_conv.log; tail -n 12 $O/t4_conv.log
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "c5 or c2" > $O/t4_c5.log 2>&1; echo "c5 rc=$?" >> $O/t4_c5.log; tail -n 12 $O/t4_c5.log
P="timeout 300 python tools/paint_bench.py 1e8 512 cic f8"
$P --check > $O/pb4_512.log 2>&1
NBK_PAINT_SPREAD=1 $P --only-sorted >> $O/pb4_512.log 2>&1
grep -v "sum =\|identical" $O/pb4_512.log
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_bucket|k_tile -c 8 --csv"
timeout 600 $NCU --log-file $O/l4_default.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > /dev/null 2>&1
python tools/ncu_summary.py $O/l4_default.csv 8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile_paint|k_bucket_scatter|k_bucket_count" -s 14 -c 5 -o $O/paint_full_r02 -f python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > $O/ncu_full4.log 2>&1; tail -n 2 $O/ncu_full4.log
timeout 600 python tools/paint_bench.py 1e9 1024 cic f8 > $O/pb4_1024.log 2>&1
grep -v "sum =\|identical" $O/pb4_1024.log
timeout 600 python tools/paint_bench.py 1e8 512 tsc f4 --only-sorted > $O/pb4_tsc.log 2>&1; grep -v "sum =\|identical" $O/pb4_tsc.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench4_headline.json 2> $O/bench4_headline.err; tail -c 4000 $O/bench4_headline.json; tail -n 5 $O/bench4_headline.err
