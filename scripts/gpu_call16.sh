mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
for v in 1 0; do
  SMPLFIT_B200_GEMM_SPLIT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$v.log').read().strip().splitlines()[-1])
k=d['roofline']['kernel_ms_per_step']
print('split=$v: fits/s=%.0f ms/step=%.3f e2e_ms=%.3f vposed_tc=%.3f fwd=%.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], k['k_vposed_tc'], d['lbs_forward']['ms_per_call']), {a: round(b,3) for a,b in k.items()})
PY
done
