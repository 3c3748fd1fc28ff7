# A/B of an environment switch: VAR=name VALS="0 1" bash scripts/gpu_ab_env.sh
mkdir -p gpurun_out
for v in $VALS; do
  env $VAR=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$v.log').read().strip().splitlines()[-1])
k=d['roofline']['kernel_ms_per_step']
print('$VAR=$v: fits/s=%.0f ms/step=%.3f e2e_ms=%.3f' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step']), {a: round(b,3) for a,b in list(k.items())[:4]})
PY
done
