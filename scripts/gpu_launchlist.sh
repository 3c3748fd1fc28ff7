# ncu launch list (per-launch gpu__time_duration) of a few bench steps -> gpurun_out/launches.csv
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_C:-700} --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --resident-only > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-300
