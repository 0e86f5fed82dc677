#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/tp_debug.py --model 33b --groupsize 32 --act-order --layers 2 2>&1 | grep -v Warn | tail -16
timeout 200 python tools/tp_debug.py --model 13b --groupsize 128 --act-order --layers 2 2>&1 | grep -v Warn | tail -4
timeout 200 python tools/tp_debug.py --model 33b --groupsize 32 --layers 2 2>&1 | grep -v Warn | tail -4
timeout 600 python bench.py --model 13b --act-order --steps 32 --warmup 4 > gpurun_out/bench_r2_13b_act.json 2> gpurun_out/bench_r2_13b_act.err; echo rc=$?; tail -c 1500 gpurun_out/bench_r2_13b_act.json; tail -3 gpurun_out/bench_r2_13b_act.err
