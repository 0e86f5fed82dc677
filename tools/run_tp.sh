#!/bin/bash
# usage: tools/run_tp.sh N tag [bench args...]   -> gpurun_out/bench_r2_<tag>_tpN.json
N=$1; tag=$2; shift 2
mkdir -p gpurun_out
timeout ${RUN_TIMEOUT:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/bench_r2_${tag}_tp$N.json 2> gpurun_out/bench_r2_${tag}_tp$N.err
echo "rc=$? $tag tp$N"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_r2_${tag}_tp$N.json") if l.startswith("{")][-1])
    print({k:d.get(k) for k in ("metric","value","ms_per_step","n_gpus")}, d["e2e"]["value"], d["run"]["decode_mode"][:60], d.get("prefill",{}) and d["prefill"].get("value"), d.get("tp"), d.get("step_parity"))
except Exception as e:
    print("no json line:", e); print(open("gpurun_out/bench_r2_${tag}_tp$N.err").read()[-1500:])
PY
