#!/bin/bash
# gradient tests alone (all failures reported), then the whole GPU suite and a resident bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_grad.py -q --timeout 240 > gpurun_out/pytest_grad.log 2>&1; echo "grad rc=$?" | tee -a gpurun_out/pytest_grad.log
tail -40 gpurun_out/pytest_grad.log
bash scripts/gpu_iter2.sh
