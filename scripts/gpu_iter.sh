#!/bin/bash
# one development iteration on the GPU box: parity tests, forward per-kernel times (+ A/B env), resident fit step (+ A/B env)
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout ${PYTEST_LIMIT:-420} python -m pytest tests -m gpu -q -x --timeout 120 ${PYTEST_ARGS} > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
timeout 90 python scripts/fwd_profile.py smpl 4096 > gpurun_out/fwd_smpl.json 2> gpurun_out/fwd.err
timeout 90 python scripts/fwd_profile.py smplx 4096 > gpurun_out/fwd_smplx.json 2>> gpurun_out/fwd.err
for v in ${FWD_ENVS}; do
  timeout 90 env $v python scripts/fwd_profile.py smpl 4096 > gpurun_out/fwd_smpl_${v//[^A-Za-z0-9]/_}.json 2>> gpurun_out/fwd.err
  timeout 90 env $v python scripts/fwd_profile.py smplx 4096 > gpurun_out/fwd_smplx_${v//[^A-Za-z0-9]/_}.json 2>> gpurun_out/fwd.err
done
if [ -z "${SKIP_FIT}" ]; then
timeout 200 python bench.py --steps 10 --warmup 3 --resident-only > gpurun_out/res_default.json 2> gpurun_out/res.err
for v in ${AB_ENVS}; do
  timeout 200 env $v python bench.py --steps 10 --warmup 3 --resident-only > gpurun_out/res_${v//[^A-Za-z0-9]/_}.json 2>> gpurun_out/res.err
done
fi
if [ -n "${NCU_FWD}" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fwd_fused -s 3 -c 1 -f -o gpurun_out/prof_fwd_r02 python scripts/fwd_profile.py smpl 4096 --short > gpurun_out/ncu_fwd.log 2>&1
fi
tail -15 gpurun_out/pytest.log; for f in gpurun_out/fwd_smpl*.json; do echo $f; cat $f; done; tail -3 gpurun_out/fwd.err; cat gpurun_out/res_*.json; tail -3 gpurun_out/res.err
