#!/bin/bash
# paint: lane-rotated CIC corners (this build) vs job 15's numbers; launch list of the paint kernels; C2 bench
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "paint or tiled or interlace" > $O/t16.log 2>&1; echo "rc=$?" >> $O/t16.log; tail -n 3 $O/t16.log
timeout 300 python tools/paint_bench.py 1e8 512 cic f8 --check 2>&1 | grep -v "sum =\|identical"
NBK_PAINT_SPREAD=1 timeout 300 python tools/paint_bench.py 1e8 512 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 2>&1 | grep -v "sum =\|identical"
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum --clock-control none --csv -k regex:k_bucket|k_tile|k_apply -c 9"
timeout 600 $NCU --log-file $O/l16_paint.csv python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted > $O/ncu16.log 2>&1
python tools/ncu_summary.py $O/l16_paint.csv 9
grep "k_tile_paint\|k_bucket_scatter" $O/l16_paint.csv | grep "wavefronts\|sectors_pipe" | cut -d, -f5,13- | head -6
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu > $O/bench16_c2.json 2> $O/bench16_c2.err; echo "c2 rc=$?"
python - <<PY
import json
try:
    d = json.load(open('$O/bench16_c2.json'))
    print('c2', {k: d[k] for k in ('value', 'ms_per_step')}, d['parity']['ok'])
    print({k: round(v, 3) for k, v in d['stage_ms'].items()})
except Exception as e:
    print('no c2 line', e)
PY
echo "elapsed $(( $(date +%s) - T0 )) s"
