mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python scripts/e2e_diag.py > gpurun_out/e2e_diag.json 2> gpurun_out/e2e_diag.err; cat gpurun_out/e2e_diag.json; tail -3 gpurun_out/e2e_diag.err
timeout 600 python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; cat gpurun_out/configs.json; tail -3 gpurun_out/configs.err
NCU_KERNELS='k_shape_lite|k_stats_lite|k_vposed_tc|k_fwd_skin' NCU_SKIP=0 NCU_COUNT=12 bash scripts/gpu_profile.sh > gpurun_out/profile.log 2>&1; tail -3 gpurun_out/profile.log
