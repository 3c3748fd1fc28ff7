mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
for lib in libsmplfit_b200.so variant_k32.so; do
  SMPLFIT_B200_LIB=$PWD/smplfitter_b200/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$lib.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$lib.log').read().strip().splitlines()[-1])
k=d['roofline']['kernel_ms_per_step']
print('$lib: fits/s=%.0f ms/step=%.3f e2e_ms=%.3f vposed_tc=%.3f transpose=%.3f fwd=%.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], k['k_vposed_tc'], k.get('k_transpose_v',0), d['lbs_forward']['ms_per_call']))
PY
done
