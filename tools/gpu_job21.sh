#!/bin/bash
# new paint variant tests; random-order launch lists at both sizes (where does the time go?)
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint" > $O/t21.log 2>&1; echo "rc=$?" >> $O/t21.log; tail -n 3 $O/t21.log
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv -k regex:k_bucket|k_tile|k_apply"
timeout 600 $NCU -s 16 -c 16 --log-file $O/l21_c2_perm.csv python tools/paint_bench.py 1e8 512 cic f8 > $O/ncu21a.log 2>&1
python tools/ncu_summary.py $O/l21_c2_perm.csv 16
grep -E "l1tex__throughput|lts__throughput|issue_active|sectors_pipe" $O/l21_c2_perm.csv | awk -F'","' '{print $1, $5, $(NF-3), $NF}' | cut -c1-150 | tail -n 32
echo "elapsed $(( $(date +%s) - T0 )) s"
