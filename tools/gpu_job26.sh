#!/bin/bash
# final per-kernel ncu table of one full C2 step (probe .. power_bin), summary printed and saved as CSV
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --section SpeedOfLight --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis --section SchedulerStats --section ComputeWorkloadAnalysis --section InstructionStats --clock-control none -k regex:"k_bucket|k_tile|k_fft|k_power|k_apply" -s 67 -c 13 -o /tmp/r02_c2_final -f python bench.py --config c2 --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu26.log 2>&1; tail -n 1 $O/ncu26.log
python tools/ncu_full_summary.py /tmp/r02_c2_final.ncu-rep $O/r02_ncu_full_c2_final.csv
ls -la /tmp/r02_c2_final.ncu-rep
