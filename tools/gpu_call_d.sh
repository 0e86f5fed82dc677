#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_step.py -x -q --timeout 250 > gpurun_out/d_step_pytest.log 2>&1; tail -4 gpurun_out/d_step_pytest.log
EXL_DS_TP_REDUCE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/tp_step_check.py --layers 8 > gpurun_out/d_tp_reduce.log 2>&1
grep "{\|token\|OK\|FAIL" gpurun_out/d_tp_reduce.log | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/tp_step_check.py --layers 8 > gpurun_out/d_tp_direct.log 2>&1
grep "{\|OK\|FAIL" gpurun_out/d_tp_direct.log | cut -c1-300
timeout 300 python tools/step_bench.py --model 13b --ctx 1920 2>&1 | tail -1 | cut -c1-500
