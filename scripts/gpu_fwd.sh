#!/bin/bash
# forward LBS: per-kernel times, one full ncu capture of k_fwd_fused, launch list
mkdir -p gpurun_out
python scripts/fwd_profile.py smpl 4096 > gpurun_out/fwd_smpl.json 2> gpurun_out/fwd.err
python scripts/fwd_profile.py smplx 4096 > gpurun_out/fwd_smplx.json 2>> gpurun_out/fwd.err
python scripts/fwd_profile.py smpl 256 > gpurun_out/fwd_smpl_256.json 2>> gpurun_out/fwd.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd_fused -s 3 -c 1 -f -o gpurun_out/prof_fwd_r02 python scripts/fwd_profile.py smpl 4096 --short > gpurun_out/ncu_fwd.log 2>&1
cat gpurun_out/fwd_smpl.json gpurun_out/fwd_smplx.json gpurun_out/fwd_smpl_256.json; tail -3 gpurun_out/fwd.err; tail -3 gpurun_out/ncu_fwd.log
