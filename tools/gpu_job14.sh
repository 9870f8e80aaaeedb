#!/bin/bash
# FFT: 64-byte-row tiles at N=1024 (2 CTAs/SM) + warp-per-row TMA z pass: correctness, per-pass timing, headline bench
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "r2c or fft or c2r" > $O/t14.log 2>&1; echo "rc=$?" >> $O/t14.log; tail -n 4 $O/t14.log
for a in "1024 f8" "512 f8" "1024 f4" "256 f8"; do timeout 200 python tools/fft_probe.py $a 2>&1 | tail -n 2; done
NBK_FFT_TMA_B=8 timeout 200 python tools/fft_probe.py 1024 f8 2>&1 | tail -n 2
NBK_FFT_TMA_B=4 timeout 200 python tools/fft_probe.py 512 f8 2>&1 | tail -n 2
NBK_FFT_TMA_B=16 timeout 200 python tools/fft_probe.py 1024 f4 2>&1 | tail -n 2
NBK_FFT_TMA_NS=20 timeout 200 python tools/fft_probe.py 1024 f8 2>&1 | tail -n 2
NBK_FFT_Z_NBUF=1 timeout 200 python tools/fft_probe.py 1024 f8 2>&1 | tail -n 2
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 900 python -m pytest tests/test_gpu_meshapi.py tests/test_gpu_fftpower.py -x -q -m gpu > $O/t14b.log 2>&1; echo "rc=$?" >> $O/t14b.log; tail -n 3 $O/t14b.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench14_headline.json 2> $O/bench14_headline.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open('$O/bench14_headline.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'parity')})
    print({k: round(v, 3) for k, v in d['stage_ms'].items()})
except Exception as e:
    print('no bench line', e)
PY
tail -n 3 $O/bench14_headline.err
echo "elapsed $(( $(date +%s) - T0 )) s"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fft" -s 6 -c 3 -o $O/r02_full_fft1024 -f python tools/fft_probe.py 1024 f8 > $O/ncu14.log 2>&1; tail -n 2 $O/ncu14.log
echo "elapsed $(( $(date +%s) - T0 )) s"
