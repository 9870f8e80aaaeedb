#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t27.log 2>&1; echo "rc=$?" >> gpurun_out/t27.log; tail -n 3 gpurun_out/t27.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
