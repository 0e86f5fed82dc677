#!/bin/bash
mkdir -p gpurun_out
EXL_DS_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/tp_step_check.py --layers 4 > gpurun_out/d_tp_step.log 2>&1; echo "rc=$?" >> gpurun_out/d_tp_step.log
grep -v Warning gpurun_out/d_tp_step.log | grep "{\|token\|rc=" | cut -c1-600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/tp_step_check.py --layers 32 --reps 30 > gpurun_out/d_tp_step32.log 2>&1
grep -v Warning gpurun_out/d_tp_step32.log | grep "{\|rc=\|OK\|FAIL" | cut -c1-400
