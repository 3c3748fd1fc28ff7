#!/bin/bash
# gradient tests alone (all failures reported), the cost of the backward, then the whole GPU suite and a resident bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_grad.py -q --timeout 240 > gpurun_out/pytest_grad.log 2>&1; echo "grad rc=$?" | tee -a gpurun_out/pytest_grad.log
tail -40 gpurun_out/pytest_grad.log
timeout 300 python scripts/grad_timing.py > gpurun_out/grad_timing.json 2> gpurun_out/grad_timing.err; cat gpurun_out/grad_timing.json; tail -3 gpurun_out/grad_timing.err
[ -n "$SKIP_SUITE" ] || bash scripts/gpu_iter2.sh
