#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bucket_count|k_bucket_scatter" -s 8 -c 4 -o $O/r02_full_bucket_c2 -f python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > $O/ncu20.log 2>&1; tail -n 1 $O/ncu20.log
ncu -i $O/r02_full_bucket_c2.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,launch__grid_size 2>/dev/null | cut -d, -f1,5,9,12,13 | head
