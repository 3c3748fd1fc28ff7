mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.log
timeout 600 python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; cut -c1-200 gpurun_out/configs.json
