mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
VAR=SMPLFIT_B200_SLOT_MASK VALS="1 0" bash scripts/gpu_ab_env.sh
