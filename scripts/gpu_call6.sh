mkdir -p gpurun_out
for B in 32 512; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_B$B.csv python scripts/small_batch.py $B > gpurun_out/small_$B.log 2>&1
done
ls gpurun_out
