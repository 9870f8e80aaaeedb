#!/bin/bash
# validation of the current build: full GPU suite, smoke(), headline bench line (with the CPU baseline leg), C2, random order
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -x -q -m gpu > $O/t22.log 2>&1; echo "rc=$?" >> $O/t22.log; tail -n 3 $O/t22.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
echo "elapsed $(( $(date +%s) - T0 )) s"
for a in "1024 f8" "512 f8"; do timeout 200 python tools/fft_probe.py $a 2>&1 | tail -n 2; done
NBK_FFT_TMA_L2=0 timeout 200 python tools/fft_probe.py 1024 f8 2>&1 | tail -n 2
timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench22_headline.json 2> $O/bench22_headline.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open('$O/bench22_headline.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, d['parity'].get('ok'), d['parity'].get('vs_cpu_oracle'))
    print({k: round(v, 3) for k, v in d['stage_ms'].items()})
    print('frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'e2e', d['e2e']['ms_per_step'], {k: v for k, v in d.get('cpu_baseline', {}).items() if k != 'sample'})
except Exception as e:
    print('no bench line', e)
PY
tail -n 3 $O/bench22_headline.err
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 --no-cpu > $O/bench22_c2.json 2> $O/bench22_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --order random --steps 5 --warmup 3 --no-cpu --no-parity > $O/bench22_random.json 2> $O/bench22_random.err; echo "random rc=$?"
python - <<PY
import json
for f in ('c2', 'random'):
    try:
        d = json.load(open('$O/bench22_%s.json' % f))
        print(f, {k: d[k] for k in ('value', 'ms_per_step')}, {k: round(v, 3) for k, v in d['stage_ms'].items() if k in ('paint', 'r2c', 'power_bin')})
    except Exception as e:
        print('no line', f, e)
PY
echo "elapsed $(( $(date +%s) - T0 )) s"
