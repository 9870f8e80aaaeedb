#!/bin/bash
# final validation of the round: full GPU suite, smoke, headline bench line, launch list of one step
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -x -q -m gpu > $O/t24.log 2>&1; echo "rc=$?" >> $O/t24.log; tail -n 3 $O/t24.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench24_headline.json 2> $O/bench24_headline.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open('$O/bench24_headline.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, d['parity'].get('ok'), d['parity'].get('vs_cpu_oracle'))
    print({k: round(v, 3) for k, v in d['stage_ms'].items()})
    print('frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d.get('cpu_baseline', {}).get('seconds'))
    print({k: round(v['frac'], 3) for k, v in d['roofline']['other_stages'].items()})
except Exception as e:
    print('no bench line', e)
PY
for k in 1 0; do NBK_BIN_LEAN=$k timeout 600 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 LEAN=$k', d['ms_per_step'], d['stage_ms']['power_bin'])"; done
for k in 1 0; do NBK_BIN_LEAN=$k timeout 600 python bench.py --config c4 --steps 3 --warmup 3 --no-cpu --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4 LEAN=$k', d['ms_per_step'], d['stage_ms']['power_bin'])"; done
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv -k regex:^k_ -c 48"
timeout 900 $NCU --log-file $O/r02_launches_headline_final.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu24.log 2>&1
python tools/ncu_summary.py $O/r02_launches_headline_final.csv 60 | tail -n 20
echo "elapsed $(( $(date +%s) - T0 )) s"
