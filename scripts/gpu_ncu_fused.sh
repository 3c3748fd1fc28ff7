# ncu --set full of one shape-mode and one statistics-mode launch of k_fit_fused during resident bench steps
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_fit_fused}" -s ${NCU_SKIP:-18} -c ${NCU_COUNT:-2} \
    -f -o gpurun_out/prof_fused python bench.py --steps 1 --warmup 3 --no-cpu-baseline --resident-only > gpurun_out/prof_fused.log 2>&1
ls -la gpurun_out/prof_fused.ncu-rep; tail -2 gpurun_out/prof_fused.log | cut -c1-200
