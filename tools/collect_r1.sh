# round-1 evidence collection (run under gpurun, one GPU): bench line, launch list of the same command, kernel sweeps
set -x
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 600 gpurun_out/bench_r1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_bench_r1.csv \
    python bench.py --steps 2 --warmup 3 --no-prefill --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/decode_layer_r1 -f \
    python tools/profile_decode.py --layers 1 --steps 1 > gpurun_out/ncu_full.log 2>&1
timeout 300 python tools/kbench.py --shapes 7b 13b 33b 65b --reps 5 --ref --json gpurun_out/kbench_r1.json > /dev/null 2>&1
timeout 300 python tools/kbench.py --shapes 7b --m 2 4 7 --reps 5 --json gpurun_out/kbench_r1_m.json > /dev/null 2>&1
timeout 300 python tools/abench.py > gpurun_out/abench_r1.jsonl 2>&1
timeout 600 python tools/pbench.py --ref > gpurun_out/pbench_r1.jsonl 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2>&1; tail -c 400 gpurun_out/bench_r1_reference.json
