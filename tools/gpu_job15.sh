#!/bin/bash
# paint: warp-staged record stores + deferred halo adds (A/B by knob); z pass twiddle tables; headline bench
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > $O/t15.log 2>&1; echo "rc=$?" >> $O/t15.log; tail -n 4 $O/t15.log
for a in "1024 f8" "512 f8" "1024 f4"; do timeout 200 python tools/fft_probe.py $a 2>&1 | tail -n 2; done
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 300 python tools/paint_bench.py 1e8 512 cic f8 --check 2>&1 | grep -v "sum =\|identical"
NBK_PAINT_WSTAGE=0 timeout 300 python tools/paint_bench.py 1e8 512 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
NBK_PAINT_DEFER=0 timeout 300 python tools/paint_bench.py 1e8 512 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
timeout 300 python tools/paint_bench.py 1e8 512 tsc f4 --only-sorted 2>&1 | grep -v "sum =\|identical"
timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
NBK_PAINT_WSTAGE=0 timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
NBK_PAINT_DEFER=0 timeout 300 python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted 2>&1 | grep -v "sum =\|identical"
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_fftpower.py -x -q -m gpu > $O/t15b.log 2>&1; echo "rc=$?" >> $O/t15b.log; tail -n 3 $O/t15b.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench15_headline.json 2> $O/bench15_headline.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open('$O/bench15_headline.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'parity')})
    print({k: round(v, 3) for k, v in d['stage_ms'].items()})
except Exception as e:
    print('no bench line', e)
PY
tail -n 3 $O/bench15_headline.err
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv -k regex:^k_ -c 14"
timeout 600 $NCU --log-file $O/l15_paint.csv python tools/paint_bench.py 1e9 1024 cic f8 --only-sorted > $O/ncu15.log 2>&1
python tools/ncu_summary.py $O/l15_paint.csv 14
echo "elapsed $(( $(date +%s) - T0 )) s"
