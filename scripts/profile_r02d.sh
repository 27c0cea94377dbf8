#!/bin/bash
# Round-2 ncu capture of the object-batched selection kernel, tensor-pipe form with 16 warps (run under gpurun, one GPU):
# 148 objects = one wave, 128 picks each; preceded by the same command without ncu.
set -u
mkdir -p gpurun_out
CMD="python scripts/bench_batch.py --fused --objects 148 --active 128 --variants dmma"
$CMD > gpurun_out/r02_batch_dmma_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_batch_select -s 1 -c 1 \
    -o gpurun_out/r02_k_batch_select_dmma -f $CMD > gpurun_out/r02_batch_dmma_ncu.log 2>&1
ls -la gpurun_out/r02_k_batch_select_dmma.ncu-rep
