#!/bin/bash
for v in m1 m2 m3 m4 f5ef730; do
EXL_B200_LIB=$PWD/exllama_b200/libexl_b200_$v.so timeout 300 python tools/step_bench.py --model 7b --ctx 1920 --no-per-op 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['fused_ms'])"
done
