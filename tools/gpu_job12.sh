#!/bin/bash
# session-2 baseline: full GPU suite, headline bench line, ncu launch lists (headline + c2), ncu --set full of the C2 step
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -x -q -m gpu > $O/t12.log 2>&1; echo "rc=$?" >> $O/t12.log; tail -n 6 $O/t12.log
echo "tests took $(( $(date +%s) - T0 )) s"
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench12_headline.json 2> $O/bench12_headline.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open('$O/bench12_headline.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'parity', 'clocks', 'gpu_launches')})
    print({k: round(v, 3) for k, v in d['stage_ms'].items()})
    print('frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], {k: v for k, v in d.get('cpu_baseline', {}).items() if k != 'sample'})
    print({k: round(v['frac'], 3) for k, v in d['roofline']['other_stages'].items()})
except Exception as e:
    print('no bench line', e)
PY
tail -n 3 $O/bench12_headline.err
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu > $O/bench12_c2.json 2> $O/bench12_c2.err; echo "c2 rc=$?"
python - <<PY
import json
try:
    d = json.load(open('$O/bench12_c2.json'))
    print('c2', {k: d[k] for k in ('value', 'ms_per_step', 'parity')})
    print({k: round(v, 3) for k, v in d['stage_ms'].items()})
    print('frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'])
except Exception as e:
    print('no c2 line', e)
PY
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv"
timeout 900 $NCU -c 60 --log-file $O/r02_launches_headline.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu12a.log 2>&1
python tools/ncu_summary.py $O/r02_launches_headline.csv 60 | tail -n 16
timeout 600 $NCU -c 80 --log-file $O/r02_launches_c2.csv python bench.py --config c2 --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu12b.log 2>&1
python tools/ncu_summary.py $O/r02_launches_c2.csv 80 | tail -n 16
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bucket|k_tile|k_fft|k_power" -s 48 -c 12 -o $O/r02_full_c2 -f python bench.py --config c2 --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu12c.log 2>&1; tail -n 2 $O/ncu12c.log
ls -la $O/*.ncu-rep
echo "elapsed $(( $(date +%s) - T0 )) s"
