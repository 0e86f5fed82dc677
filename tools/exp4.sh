run() { python tools/kbench.py "$@" --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l); print('   ', r['K'], r['N'], r['gs'], r['us'], r['GBps'])
    except Exception: pass
"; }
for dbg in 0 8 16 24 4; do echo "dbg=$dbg"; EXL_GV_DEBUG=$dbg run --only 6656 17920 128; EXL_GV_DEBUG=$dbg run --only 11008 4096 128; done
