#!/bin/bash
# GPU call A (round 2): full GPU parity suite incl. the reference-pin tests, then the drop-in run of the reference's
# unchanged test_benchmark_inference.py over both extensions on a synthetic 7B model directory.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 400 python tools/make_synth_model.py --model 7b --out /tmp/synth7b > gpurun_out/a_synth.log 2>&1
tail -2 gpurun_out/a_synth.log
timeout 900 python tools/run_dropin.py --model-dir /tmp/synth7b --out gpurun_out/dropin_7b.json --tag "7b g128 no-act" > gpurun_out/a_dropin.log 2>&1
tail -3 gpurun_out/a_dropin.log
