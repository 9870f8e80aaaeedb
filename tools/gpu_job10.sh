#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_meshapi.py tests/test_gpu_kernels.py -x -q -m gpu > $O/t10.log 2>&1; echo "rc=$?" >> $O/t10.log; tail -n 4 $O/t10.log
timeout 200 python tools/fft_probe.py 1024 f8 2>&1 | tail -n 2
timeout 200 python tools/fft_probe.py 1024 f4 2>&1 | tail -n 2
timeout 200 python tools/fft_probe.py 512 f8 2>&1 | tail -n 2
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench10_headline.json 2> $O/bench10_headline.err; python -c "
import json; d=json.load(open('$O/bench10_headline.json')); print({k:d[k] for k in ('value','ms_per_step','stage_ms','parity')}); print(d['roofline']['frac'], d['e2e']['ms_per_step'], {k:v for k,v in d.get('cpu_baseline',{}).items() if k!='sample'})"; tail -n 3 $O/bench10_headline.err
