#!/bin/bash
# Round-2 ncu captures (run under gpurun, one GPU).  Each ncu run is preceded by the same command
# without ncu exiting 0, as the profiling recipe requires.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
B="python bench.py --grid 128 --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-tf32 --no-sampling"
# 1. TF32 update GEMM (tcgen05) of the per-step path at r ~ 2000
CMD="$B --active 2100 --precision tf32"
$CMD > gpurun_out/r02_tf32_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_grid_gemm_tf32 -s 4100 -c 1 \
    -o gpurun_out/r02_k_grid_gemm_tf32 -f $CMD > gpurun_out/r02_tf32_ncu.log 2>&1
# 2. factor append (k_grid_solve2) and epilogue of the per-step path at r ~ 2000
CMD="$B --active 2100 --loop steps"
$CMD > gpurun_out/r02_steps_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_grid_solve2|k_grid_epilogue" -s 8200 -c 2 \
    -o gpurun_out/r02_k_grid_solve2_epilogue -f $CMD > gpurun_out/r02_steps_ncu.log 2>&1
# 3. object-batched selection kernel: 148 objects = one wave, short selection
CMD="python scripts/bench_batch.py --fused --objects 148 --active 96 --variants tuned"
$CMD > gpurun_out/r02_batch_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_batch_select -s 1 -c 1 \
    -o gpurun_out/r02_k_batch_select -f $CMD > gpurun_out/r02_batch_ncu.log 2>&1
python scripts/bench_batch.py --fused --variants plain tuned dmma > gpurun_out/r02_batch_variants.json 2> gpurun_out/r02_batch_variants.err
ls -la gpurun_out/*.ncu-rep
