#!/bin/bash
# N GPUs of one box (gpurun --gpus N): NCCL scatter -> fit -> gather test, then bench lines at N ranks
# (default config weak + strong scaling, and the BASELINE configs that are quoted per GPU on several GPUs)
N=${NGPUS:-2}
mkdir -p gpurun_out
if [ -z "${SKIP_DIST}" ]; then
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x --timeout 500 > gpurun_out/pytest_dist.log 2>&1; echo "pytest_dist rc=$?" | tee -a gpurun_out/pytest_dist.log
fi
run() { # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-reference-gpu $2 > gpurun_out/multi_${N}_$1.json 2>> gpurun_out/multi.err
  python - gpurun_out/multi_${N}_$1.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    e=d.get('e2e') or {}; sg=d.get('scatter_gather') or {}
    print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(e.get('value',0)), 'e2e_ms', e.get('ms_per_step'), 'sg', sg.get('fits_per_s'), sg.get('ms_per_step'))
except Exception as ex:
    print(sys.argv[1], 'unreadable', ex)
PY
}
run weak ""
for c in ${EXTRA:-strong}; do
  case $c in
    strong) run strong "--scaling strong";;
    *) run $c "--config $c";;
  esac
done
tail -3 gpurun_out/pytest_dist.log 2>/dev/null; grep -v Warning gpurun_out/multi.err | tail -5
