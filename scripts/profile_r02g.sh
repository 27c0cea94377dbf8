#!/bin/bash
# Round-2 ncu capture of the final k_grid_persist<3, cube> on the WHOLE headline job (4000 points at 128^3, the second of
# the script's two selections), preceded by the same command without ncu (run under gpurun, one GPU).
set -u
mkdir -p gpurun_out
CMD="python scripts/persist_time.py 128 4000"
$CMD > gpurun_out/r02_persist_full_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_grid_persist -s 1 -c 1 \
    -o gpurun_out/r02_k_grid_persist_full -f $CMD > gpurun_out/r02_persist_full_ncu.log 2>&1
ls -la gpurun_out/r02_k_grid_persist_full.ncu-rep
