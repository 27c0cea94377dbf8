#!/bin/bash
# Round-2 final captures (run under gpurun, one GPU), each preceded by the same command without ncu:
#  1. k_grid_persist<3, cube>: one whole 2000-point selection at 128^3 (ncu --set full)
#  2. launch list of one bench step (the 4000-point job is ONE k_grid_persist launch + set-up / export kernels)
set -u
mkdir -p gpurun_out
bash scripts/profile_r02c.sh
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-tf32 --no-sampling --no-extras"
$CMD > gpurun_out/r02_launches_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_grid.csv $CMD \
    > gpurun_out/r02_launches_ncu.log 2>&1
ls -la gpurun_out/r02_k_grid_persist_cube.ncu-rep gpurun_out/r02_launches_grid.csv
