#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
P="timeout 300 python tools/paint_bench.py 1e8 512 cic f8"
$P --check > $O/pb8_512.log 2>&1
grep -v "sum =\|identical" $O/pb8_512.log
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_bucket|k_tile -c 8 --csv"
timeout 600 $NCU --log-file $O/l8_default.csv python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > /dev/null 2>&1
python tools/ncu_summary.py $O/l8_default.csv 8
timeout 600 python tools/paint_bench.py 1e9 1024 cic f8 > $O/pb8_1024.log 2>&1
grep -v "sum =\|identical" $O/pb8_1024.log
timeout 600 python tools/paint_bench.py 1e8 512 tsc f4 --only-sorted > $O/pb8_tsc.log 2>&1; grep -v "sum =\|identical" $O/pb8_tsc.log
timeout 1500 python -m pytest tests -x -q -m gpu > $O/t8_all.log 2>&1; echo "all rc=$?" >> $O/t8_all.log; tail -n 12 $O/t8_all.log
timeout 600 python bench.py --config c2 --steps 10 --warmup 3 > $O/bench8_c2.json 2> $O/bench8_c2.err; python -c "
import json; d=json.load(open('$O/bench8_c2.json')); print({k:d[k] for k in ('value','ms_per_step','stage_ms','parity')}); print(d['roofline']['frac'], d['e2e']['ms_per_step'], d.get('cpu_baseline',{}).get('seconds'))"
