#!/bin/bash
# Round-2 evidence run on one B200: smoke, GPU tests, bench (7B headline + 13B act-order), reference arm, ncu launch list of the
# bench command, drop-in run of the reference's unchanged benchmark over both extensions.  Outputs -> gpurun_out/ (scratch);
# the summaries worth keeping are copied into profiles/ by hand afterwards.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2_pytest.log 2>&1; tail -3 gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/r2_bench.err
timeout 600 python bench.py --model 13b --act-order --steps 32 --warmup 4 > gpurun_out/r2_bench_13b_act.json 2> gpurun_out/r2_bench_13b_act.err; echo "13b rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-prefill --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
# full capture of the persistent kernel (-> profiles/decode_step_ncu_r2.json, traffic_r2.json via tools/summarise_ncu.py)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 3 -c 1 -o gpurun_out/r2_step_full python tools/step_bench.py --model 7b --ctx 1920 --no-per-op --reps 3 > gpurun_out/r2_ncu_full.log 2>&1
timeout 400 python tools/make_synth_model.py --model 7b --out /tmp/synth7b > gpurun_out/r2_synth.log 2>&1
timeout 900 python tools/run_dropin.py --model-dir /tmp/synth7b --out gpurun_out/dropin_7b.json --tag "7b g128 no-act, zero-mean synthetic weights" > gpurun_out/r2_dropin.log 2>&1; tail -1 gpurun_out/r2_dropin.log | cut -c1-900
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench.json","gpurun_out/r2_bench_13b_act.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], (d.get("prefill") or {}).get("value"), d.get("cpu_baseline",{}) and d["cpu_baseline"].get("value"), d.get("decode_per_op_graph"))
    except Exception as e: print(f, "FAILED", e)
PY
# multi-GPU lines (separate gpurun --gpus N calls; charged N x):
#   bash tools/run_tp.sh 2 7b --steps 64 --warmup 8 --no-per-op --no-prefill
#   bash tools/run_tp.sh 4 33b_g32_act --model 33b --groupsize 32 --act-order --steps 16 --warmup 4 --no-prefill
#   bash tools/run_tp.sh 8 65b_seq4096 --model 65b --seq 4096 --steps 16 --warmup 4 --no-per-op --no-prefill
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/tp_step_check.py --layers 8
