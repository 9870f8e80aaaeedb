#!/bin/bash
# full GPU suite on the current build; binning shortcuts + pow2 count path; ncu --set full of the paint kernels (C2 size);
# C3 / C4 / C5 bench lines
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -x -q -m gpu > $O/t17.log 2>&1; echo "rc=$?" >> $O/t17.log; tail -n 3 $O/t17.log
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bucket|k_tile_paint" -s 14 -c 5 -o $O/r02_full_paint_c2 -f python tools/paint_bench.py 1e8 512 cic f8 --only-sorted > $O/ncu17.log 2>&1; tail -n 1 $O/ncu17.log
echo "elapsed $(( $(date +%s) - T0 )) s"
for cfg in headline c3 c4 c5; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu > $O/bench17_$cfg.json 2> $O/bench17_$cfg.err; echo "$cfg rc=$?"
  python - <<PY
import json
try:
    d = json.load(open('$O/bench17_$cfg.json'))
    print('$cfg', {k: d[k] for k in ('value', 'ms_per_step')}, d['parity'].get('ok'), 'e2e', round(d['e2e']['ms_per_step'], 2))
    print({k: round(v, 3) for k, v in d['stage_ms'].items()})
except Exception as e:
    print('no line', e)
PY
  tail -n 2 $O/bench17_$cfg.err
done
echo "elapsed $(( $(date +%s) - T0 )) s"
