#!/bin/bash
# per-kernel device times of the other BASELINE configurations (resident steps): SMPL-X fit, 1024-vertex subset, converter
mkdir -p gpurun_out
for c in ${CONFIGS:-converter smplx subset}; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --resident-only > gpurun_out/res_cfg_$c.json 2>> gpurun_out/res_cfg.err
  python - gpurun_out/res_cfg_$c.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d['ms_per_step'], d.get('gpu_launches'))
    print(d.get('kernel_ms_per_step'))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
tail -5 gpurun_out/res_cfg.err
