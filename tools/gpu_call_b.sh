#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_step.py -x -q --timeout 200 > gpurun_out/b_step_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b_step_pytest.log
tail -15 gpurun_out/b_step_pytest.log
timeout 300 python tools/step_bench.py --model 7b --ctx 1920 > gpurun_out/b_step_bench.log 2>&1; tail -3 gpurun_out/b_step_bench.log
timeout 300 python tools/step_bench.py --model 7b --ctx 4 --no-per-op >> gpurun_out/b_step_bench.log 2>&1; tail -1 gpurun_out/b_step_bench.log
timeout 600 python -m pytest tests/test_gpu_ref_pin.py -q --timeout 300 > gpurun_out/b_pin_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b_pin_pytest.log
tail -8 gpurun_out/b_pin_pytest.log
