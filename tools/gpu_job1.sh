#!/bin/bash
# paint v2 validation + timing (scratch)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint" > $O/t_default.log 2>&1; echo "default rc=$?" >> $O/t_default.log
NBK_PAINT_BUCKET=coherent timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled" > $O/t_coh.log 2>&1; echo "coherent rc=$?" >> $O/t_coh.log
NBK_PAINT_BUCKET=scattered timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled" > $O/t_sca.log 2>&1; echo "scattered rc=$?" >> $O/t_sca.log
tail -3 $O/t_default.log $O/t_coh.log $O/t_sca.log
timeout 600 python tools/paint_bench.py 1e8 512 cic f8 --check > $O/pb_512.log 2>&1; cat $O/pb_512.log
timeout 600 python tools/paint_bench.py 1e9 1024 cic f8 > $O/pb_1024.log 2>&1; cat $O/pb_1024.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_bucket|k_tile" -c 24 --csv --log-file $O/launches_paint_512.csv python tools/paint_bench.py 1e8 512 cic f8 > $O/ncu_512.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tiled_slab_ghosts or tiled_is_order" > $O/racecheck.log 2>&1; tail -5 $O/racecheck.log
