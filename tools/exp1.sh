for cps in 3 2 1; do for dbg in 0 1 2 3; do echo "cps=$cps dbg=$dbg"; EXL_GV_CPS=$cps EXL_GV_DEBUG=$dbg python tools/kbench.py --shapes 7b --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l); print('   ', r['K'], r['N'], r['us'], r['GBps'])
    except Exception: pass
"; done; done
