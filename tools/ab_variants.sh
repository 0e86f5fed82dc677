#!/bin/bash
# A/B of persistent-kernel builds on ONE box (same GPU, same clocks, minutes apart): for every tag given, the prebuilt
# exllama_b200/libexl_b200_<tag>.so is copied over the main library, then tools/step_bench.py (7B, ctx 1920, 100 graph replays) and the
# decode-step parity tests run against it.  The fastest build whose tests pass becomes the main library for the full GPU suite and a
# bench.py run.  Outputs -> gpurun_out/ab_*.  (ptxas scheduling of this kernel moves by +-10 % with unrelated source changes -- DESIGN.md
# 3c -- so a change is only kept if THIS comparison says so.)
mkdir -p gpurun_out
cp exllama_b200/libexl_b200.so /tmp/exl_main.so
BUDGET=${AB_BUDGET_S:-150}            # stop starting new variants after this many seconds: the final suite + bench need the rest
for tag in "$@"; do
  if [ $SECONDS -gt $BUDGET ]; then echo "$tag: skipped (time)"; continue; fi
  cp exllama_b200/libexl_b200_$tag.so exllama_b200/libexl_b200.so || continue
  timeout 60 python tools/step_bench.py --model 7b --ctx 1920 --no-per-op --reps 100 > gpurun_out/ab_${tag}_bench.json 2> gpurun_out/ab_${tag}_bench.err
  timeout 100 python -m pytest tests/test_gpu_decode_step.py -x -q --timeout 90 > gpurun_out/ab_${tag}_pytest.log 2>&1
  echo "rc=$?" >> gpurun_out/ab_${tag}_pytest.log
  echo "$tag: $(cut -c1-400 gpurun_out/ab_${tag}_bench.json | grep -o '"fused_ms": [0-9.]*') $(tail -2 gpurun_out/ab_${tag}_pytest.log | tr '\n' ' ')"
done
best=$(python - "$@" <<'PY'
import json, sys
best, best_ms = None, 1e9
for tag in sys.argv[1:]:
    try:
        ok = open(f"gpurun_out/ab_{tag}_pytest.log").read().strip().endswith("rc=0")
        ms = json.loads([l for l in open(f"gpurun_out/ab_{tag}_bench.json") if l.startswith("{")][-1])["fused_ms"]
    except Exception:
        continue
    if ok and ms < best_ms:
        best, best_ms = tag, ms
print(best or "")
PY
)
echo "winner: ${best:-none}" | tee gpurun_out/ab_winner.txt
if [ -n "$best" ]; then cp exllama_b200/libexl_b200_$best.so exllama_b200/libexl_b200.so; else cp /tmp/exl_main.so exllama_b200/libexl_b200.so; fi
timeout 120 python -m pytest tests -m gpu -x -q --timeout 100 > gpurun_out/ab_full_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ab_full_pytest.log; tail -3 gpurun_out/ab_full_pytest.log
timeout 100 python bench.py --steps 64 --warmup 8 > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/ab_bench.json") if l.startswith("{")][-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["step_parity"], d.get("decode_best_ctx4"))
except Exception as e:
    print("bench line missing:", e)
PY
# phase timeline of the winner and of the baseline build (what moved)
timeout 40 python tools/step_trace.py --layers 4 > gpurun_out/ab_trace_winner.log 2>&1; tail -20 gpurun_out/ab_trace_winner.log | cut -c1-200
if [ -f exllama_b200/libexl_b200_base.so ]; then
  cp exllama_b200/libexl_b200_base.so exllama_b200/libexl_b200.so
  timeout 40 python tools/step_trace.py --layers 4 > gpurun_out/ab_trace_base.log 2>&1
fi
