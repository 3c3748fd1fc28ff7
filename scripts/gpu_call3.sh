mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
timeout 300 python scripts/e2e_diag.py > gpurun_out/e2e_diag.json 2> gpurun_out/e2e_diag.err; cat gpurun_out/e2e_diag.json; tail -3 gpurun_out/e2e_diag.err
timeout 600 python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; cat gpurun_out/configs.json; tail -3 gpurun_out/configs.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
