#!/bin/bash
# A/B of kernel variants on the GPU box: bench with each SMPLFIT_B200_SHAPE_VARIANT
mkdir -p gpurun_out
for v in ${VARIANTS:-0 1 2 3}; do
  SMPLFIT_B200_SHAPE_VARIANT=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$v.log').read().strip().splitlines()[-1])
    k=d['roofline']['kernel_ms_per_step']
    print('variant $v: fits/s=%.0f ms/step=%.3f shape=%.3f stats=%.3f tc=%.3f' % (d['value'], d['ms_per_step'], k.get('k_shape_pass_rec',0), k.get('k_stats_rec',0), k.get('k_vposed_tc',0)))
except Exception as e:
    print('variant $v failed', e); print(open('gpurun_out/ab_$v.log').read()[-500:])
PY
done
if [ -n "$TEST_VARIANT" ]; then
  SMPLFIT_B200_SHAPE_VARIANT=$TEST_VARIANT timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_variant.log 2>&1
  echo "pytest variant $TEST_VARIANT rc=$?"; tail -4 gpurun_out/pytest_variant.log
fi
