#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_meshapi.py tests/test_gpu_kernels.py -x -q -m gpu > $O/t11.log 2>&1; echo "rc=$?" >> $O/t11.log; tail -n 4 $O/t11.log
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "c2 or 1024" > $O/t11b.log 2>&1; echo "rc=$?" >> $O/t11b.log; tail -n 3 $O/t11b.log
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_bucket|k_tile -c 8 --csv"
timeout 600 $NCU --log-file $O/l11_default.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > $O/pb11.log 2>&1
python tools/ncu_summary.py $O/l11_default.csv 8
timeout 300 python tools/paint_bench.py 1e8 512 cic f8 --check 2>&1 | grep -v "sum =\|identical"
timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 2>&1 | grep -v "sum =\|identical"
timeout 300 python tools/paint_bench.py 1e8 512 tsc f4 --only-sorted 2>&1 | grep -v "sum =\|identical"
