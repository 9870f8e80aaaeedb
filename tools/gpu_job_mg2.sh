#!/bin/bash
# usage: gpu_job_mg2.sh N   (multi-GPU parity check + headline bench on N GPUs)
N=$1
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/mgpu_check.py > $O/mgpu_check_$N.log 2>&1; echo "mgpu_check rc=$?" >> $O/mgpu_check_$N.log; grep -E "case|rc=|Error|error|FAIL|ok" $O/mgpu_check_$N.log | cut -c1-200 | tail -n 30
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_mg${N}.json 2> $O/bench_mg${N}.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open('$O/bench_mg${N}.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'parity')})
    print({k: round(v, 3) for k, v in d['stage_ms'].items()})
    print('e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['other_stages'].get('fft_y_scatter'))
except Exception as e:
    print('no bench line', e)
PY
tail -n 5 $O/bench_mg${N}.err
