#!/bin/bash
# tile paint: byte-address rotation + mass-free loop; fresh ncu --set full of the paint kernels
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint or tiled or interlace" > $O/t18.log 2>&1; echo "rc=$?" >> $O/t18.log; tail -n 3 $O/t18.log
timeout 300 python tools/paint_bench.py 1e8 512 cic f8 --check 2>&1 | grep -v "sum =\|identical"
timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bucket|k_tile_paint" -s 14 -c 5 -o $O/r02_full_paint_c2 -f python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > $O/ncu18.log 2>&1; tail -n 1 $O/ncu18.log
echo "elapsed $(( $(date +%s) - T0 )) s"
