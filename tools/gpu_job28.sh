#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --config c2 --steps 3 --warmup 3 > gpurun_out/bench28_c2.json 2> gpurun_out/bench28_c2.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench28_c2.json')); print(d['ms_per_step'], d['parity'].get('ok'), d['parity'].get('vs_cpu_oracle')); print({k:v for k,v in d['cpu_baseline'].items() if k!='sample'})"
tail -n 2 gpurun_out/bench28_c2.err
