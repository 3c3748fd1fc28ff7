#!/bin/bash
# Run on the GPU box (under gpurun): smoke, GPU parity tests, a short bench + reference arm; logs -> gpurun_out/
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ${PYTEST_ARGS} > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/bench.err
if [ -n "${REF_ARM}" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err; echo "ref rc=$?" | tee -a gpurun_out/bench.err
fi
tail -5 gpurun_out/smoke.log; tail -40 gpurun_out/pytest.log; cat gpurun_out/bench.log; cat gpurun_out/bench_ref.log 2>/dev/null; tail -5 gpurun_out/bench.err
