#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c_trace.log
timeout 600 python -m pytest tests/test_gpu_decode_step.py -x -q --timeout 200 > gpurun_out/c_step_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c_step_pytest.log
tail -4 gpurun_out/c_step_pytest.log
for dbg in 0 1 7; do
  EXL_DS_DEBUG=$dbg timeout 200 python tools/step_trace.py --ctx 1920 >> gpurun_out/c_trace.log 2>&1
done
EXL_DS_DEBUG=0 EXL_DS_DEPTH=2 timeout 200 python tools/step_trace.py --ctx 1920 >> gpurun_out/c_trace.log 2>&1
EXL_DS_DEBUG=0 timeout 200 python tools/step_trace.py --ctx 4 >> gpurun_out/c_trace.log 2>&1
grep -v Warning gpurun_out/c_trace.log | cut -c1-400
timeout 300 python tools/step_bench.py --model 7b --ctx 1920 2>&1 | tail -1
