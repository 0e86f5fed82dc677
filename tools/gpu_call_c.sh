#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c_trace.log
EXL_DS_DEBUG=0 timeout 200 python tools/step_trace.py --ctx 1920 --layers 3 >> gpurun_out/c_trace.log 2>&1
grep -v Warning gpurun_out/c_trace.log | tail -2 | cut -c1-600
