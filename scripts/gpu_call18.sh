mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab.log 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/ab.log').read().strip().splitlines()[-1])
k=d['roofline']['kernel_ms_per_step']
print('fits/s=%.0f ms/step=%.3f e2e_ms=%.3f fwd=%.3f ms clocks=%s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['lbs_forward']['ms_per_call'], d['clocks']), {a: round(b,3) for a,b in list(k.items())[:6]})
PY
