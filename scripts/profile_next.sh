#!/bin/bash
# ncu captures that round 1 did not get to (run under gpurun, one GPU; each ncu run is preceded by the
# same command without ncu, as the profiling recipe requires).  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
# 1. object-batched selection kernel (never profiled): 148 objects = one wave, short selection
CMD="python scripts/bench_batch.py --fused --objects 148 --active 96 --variants tuned"
$CMD > gpurun_out/batch_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_batch_select -s 1 -c 1 \
    -o gpurun_out/k_batch_select $CMD > gpurun_out/batch_ncu.log 2>&1
# 1b. the tensor-pipe form of the update (opt-in, never run on a GPU in round 1): parity, then timing
python -m pytest tests/test_zz_gpu_batch_fused.py -m gpu -q > gpurun_out/batch_dmma_tests.log 2>&1
python scripts/bench_batch.py --fused --variants plain tuned dmma > gpurun_out/batch_variants.json 2> gpurun_out/batch_variants.err
# 2. shape sampling after the two-level Cholesky (the committed launch list predates it)
CMD="python scripts/bench_sampling.py --grid 25 --reps 0"
$CMD > gpurun_out/sampling_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"k_chol|k_gemm_tn|k_prior|k_cross|k_symm|k_zero|k_draws|k_rows|k_logpdf|k_full_mean|k_prepare" \
    -c 3000 --csv --log-file gpurun_out/launches_sampling.csv $CMD > gpurun_out/sampling_ncu.log 2>&1
