for v in "EXL_GV_CAP=1" "EXL_GV_CAP=1 EXL_GV_PDL=0"; do
  echo "== $v"
  env $v timeout 200 python bench.py --no-prefill --no-cpu-baseline --steps 32 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], [(k['kernel'][:8],k['us']) for k in d['kernels']])"
done
