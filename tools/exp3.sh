run() { python tools/kbench.py "$@" --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l); print('   ', r['K'], r['N'], r['gs'], r['us'], r['GBps'])
    except Exception: pass
"; }
for dbg in 0 4 1; do echo "dbg=$dbg"; EXL_GV_DEBUG=$dbg run --shapes 7b 33b; done
for cs in 1 2 4 8; do echo "cs=$cs"; EXL_GV_CS=$cs run --shapes 7b; done
echo "pdl off"; EXL_GV_PDL=0 run --shapes 7b
