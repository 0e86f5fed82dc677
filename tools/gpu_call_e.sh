#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "rc=$?"; tail -c 3000 gpurun_out/e_bench.json; tail -3 gpurun_out/e_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/e_bench_ref.json 2>> gpurun_out/e_bench.err; tail -c 600 gpurun_out/e_bench_ref.json
# ncu: launch list of the same command (short), then a full capture of the persistent kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 2 --warmup 3 --no-prefill --no-cpu-baseline --no-per-op > gpurun_out/e_ncu_bench.log 2>&1
tail -2 gpurun_out/e_ncu_bench.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 3 -c 1 -o gpurun_out/e_step_full python tools/step_bench.py --model 7b --ctx 1920 --no-per-op --reps 3 > gpurun_out/e_ncu_full.log 2>&1
tail -2 gpurun_out/e_ncu_full.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep
