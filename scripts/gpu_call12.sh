mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 -k "from_host" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
for t in 0 1; do for n in 4 6; do
  SMPLFIT_B200_HOST_TAPER=$t SMPLFIT_B200_HOST_SLOTS=$n timeout 300 python scripts/e2e_diag.py > gpurun_out/e2e_t${t}_s$n.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/e2e_t${t}_s$n.json'))
print('taper $t slots $n:', {k.replace('from_host_','').replace('_ms',''): round(v,2) for k,v in d.items() if k.startswith('from_host') and k.endswith('_ms') and 'issue' not in k})
PY
done; done
