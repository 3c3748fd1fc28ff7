#!/bin/bash
# ncu captures (under gpurun, 1 GPU): launch list of two bench steps + full-set capture of the top kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_shape_pass|k_stats|k_vposed_tc}" -s ${NCU_SKIP:-11} -c ${NCU_COUNT:-3} \
    -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out/
