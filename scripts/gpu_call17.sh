mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; cut -c1-600 gpurun_out/bench_ref.log
bash scripts/gpu_launchlist.sh
NCU_KERNELS="k_shape_lite|k_stats_lite|k_vposed_tc|k_stats_tmpl|k_transpose_v" NCU_SKIP=5 NCU_COUNT=7 bash scripts/gpu_ncu_full.sh
timeout 600 python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; cat gpurun_out/configs.json
