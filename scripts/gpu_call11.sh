mkdir -p gpurun_out
for n in 3 4 6 8; do
  SMPLFIT_B200_HOST_SLOTS=$n timeout 300 python scripts/e2e_diag.py > gpurun_out/e2e_slots$n.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/e2e_slots$n.json'))
print('slots $n:', {k.replace('from_host_','').replace('_ms',''): round(v,2) for k,v in d.items() if k.startswith('from_host') and k.endswith('_ms') and 'issue' not in k})
PY
done
