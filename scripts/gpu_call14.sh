mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
bash scripts/gpu_launchlist.sh
NCU_KERNELS="k_shape_lite|k_stats_lite|k_vposed_tc|k_stats_tmpl|k_transpose|k_gram_entries" NCU_SKIP=6 NCU_COUNT=8 bash scripts/gpu_ncu_full.sh
ncu --set full --clock-control none --import-source on -k regex:"k_fwd_skin_tma|k_fwd_prep" -s 2 -c 2 -f -o gpurun_out/prof_fwd python scripts/bench_configs.py > gpurun_out/prof_fwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
