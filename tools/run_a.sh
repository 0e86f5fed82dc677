set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; head -c 300 gpurun_out/bench_r1c.json; tail -3 gpurun_out/bench_r1c.err
