#!/bin/bash
# end-of-round record on one B200: smoke, every GPU test, the bench line + reference arm, launch list and a full ncu
# capture of the dominant kernels (digest with: python scripts/profile_digest.py TAG)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.log
bash scripts/gpu_launchlist.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fit_fused|k_fwd_fused" -s 18 -c 2 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --resident-only > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out/prof.ncu-rep
