#!/bin/bash
# final ncu --set full summary of one C2 step (all hot kernels of the final build)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bucket|k_tile|k_fft|k_power|k_apply" -s 60 -c 16 -o $O/r02_full_c2_final -f python bench.py --config c2 --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu25.log 2>&1; tail -n 2 $O/ncu25.log
python tools/ncu_full_summary.py $O/r02_full_c2_final.ncu-rep $O/r02_ncu_full_c2_final.csv
