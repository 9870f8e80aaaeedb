#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/mgpu_check_2c.log 2>&1; echo "rc=$?"; grep -E "case|Error" gpurun_out/mgpu_check_2c.log | cut -c1-160 | tail -n 14
