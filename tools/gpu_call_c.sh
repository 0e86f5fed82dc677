#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --model 13b --act-order --steps 32 --warmup 4 --no-prefill --no-cpu-baseline > gpurun_out/r2_bench_13b_act_final.json 2> gpurun_out/r2_bench_13b_act_final.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_13b_act_final.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["step_parity"], d["decode_per_op_graph"]["value"])
PY
