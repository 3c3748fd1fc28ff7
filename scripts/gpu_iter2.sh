#!/bin/bash
# round-2 development iteration: GPU parity tests, then resident bench lines (default + A/B environment settings)
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/res_*.json
timeout ${PYTEST_LIMIT:-600} python -m pytest tests -m gpu -q -x --timeout 180 ${PYTEST_ARGS} > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 --resident-only > gpurun_out/res_default.json 2> gpurun_out/res.err
for v in ${AB_ENVS}; do
  timeout 200 env $v python bench.py --steps 10 --warmup 3 --resident-only > gpurun_out/res_${v//[^A-Za-z0-9]/_}.json 2>> gpurun_out/res.err
done
tail -25 gpurun_out/pytest.log; for f in gpurun_out/res_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['ms_per_step'], d.get('gpu_launches'))
    print(d.get('kernel_ms_per_step'))
except Exception as e:
    print('unreadable', e)
PY
done; tail -5 gpurun_out/res.err
