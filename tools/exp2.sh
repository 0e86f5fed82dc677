echo "== default lib, debug bits"
for dbg in 0 1 2 3; do echo "dbg=$dbg"; EXL_GV_DEBUG=$dbg python tools/kbench.py --shapes 7b --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l); print('   ', r['K'], r['N'], r['us'], r['GBps'])
    except Exception: pass
"; done
for v in u4_b2 u8_b2 u6_b2 u2_b3; do echo "== variant $v"; EXL_B200_LIB=exllama_b200/_obj/libexl_$v.so python tools/kbench.py --shapes 7b --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l); print('   ', r['K'], r['N'], r['us'], r['GBps'])
    except Exception: pass
"; done
