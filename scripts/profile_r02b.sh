#!/bin/bash
# Round-2 ncu captures of the persistent selection kernel (run under gpurun, one GPU).  Each ncu run is
# preceded by the same command without ncu exiting 0.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
# 1. k_grid_persist, one whole selection of 1200 points on the 128^3 lattice (the second of the script's two)
CMD="python scripts/persist_time.py 128 1200"
$CMD > gpurun_out/r02_persist_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_grid_persist -s 1 -c 1 \
    -o gpurun_out/r02_k_grid_persist -f $CMD > gpurun_out/r02_persist_ncu.log 2>&1
# 2. launch list of one bench step (the whole 4000-point job is ONE k_grid_persist launch + set-up / export kernels)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-tf32 --no-sampling --no-extras"
$CMD > gpurun_out/r02_launches_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_grid.csv $CMD \
    > gpurun_out/r02_launches_ncu.log 2>&1
ls -la gpurun_out/r02_k_grid_persist.ncu-rep gpurun_out/r02_launches_grid.csv
