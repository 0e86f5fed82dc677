set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/kbench.py --shapes 7b 33b --reps 5 2>&1 | cut -c1-110
timeout 600 python bench.py --no-prefill --no-cpu-baseline > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; head -c 300 gpurun_out/bench_r1d.json; tail -3 gpurun_out/bench_r1d.err
