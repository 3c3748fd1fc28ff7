# one ncu --set full capture of the kernels matching NCU_KERNELS (regex) during one bench step
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS}" -s ${NCU_SKIP:-8} -c ${NCU_COUNT:-2} \
    -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out/prof.ncu-rep
