#!/bin/bash
# Round-2 ncu capture of the persistent selection kernel on 16^3 cubes (run under gpurun, one GPU): one whole selection
# of 2000 points on the 128^3 lattice (the second of the script's two), preceded by the same command without ncu.
set -u
mkdir -p gpurun_out
CMD="python scripts/persist_time.py 128 2000"
$CMD > gpurun_out/r02_persist_cube_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_grid_persist -s 1 -c 1 \
    -o gpurun_out/r02_k_grid_persist_cube -f $CMD > gpurun_out/r02_persist_cube_ncu.log 2>&1
ls -la gpurun_out/r02_k_grid_persist_cube.ncu-rep
