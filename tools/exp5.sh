run() { python tools/kbench.py "$@" --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l); print('   ', r['K'], r['N'], r['gs'], r['us'], r['GBps'], r['frac'])
    except Exception: pass
"; }
timeout 200 python -m pytest tests/test_gpu_q4_matmul.py -m gpu -q -x 2>&1 | tail -2
for cap in 2 1; do echo "cap=$cap"; EXL_GV_CAP=$cap run --shapes 7b 33b; done
echo "cap=2 dbg4"; EXL_GV_DEBUG=4 run --shapes 7b 33b
