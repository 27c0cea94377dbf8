#!/bin/bash
# Round-2 launch list of posterior shape sampling after the two-level blocked Cholesky (run under gpurun, one GPU);
# preceded by the same command without ncu.  Output lands in gpurun_out/.
set -u
mkdir -p gpurun_out
CMD="python scripts/bench_sampling.py --grid 25 --reps 0"
$CMD > gpurun_out/r02_sampling_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"k_chol|k_gemm_tn|k_prior|k_cross|k_symm|k_zero|k_draws|k_rows|k_logpdf|k_full_mean|k_prepare" \
    -c 3000 --csv --log-file gpurun_out/r02_launches_sampling.csv $CMD > gpurun_out/r02_sampling_ncu.log 2>&1
tail -2 gpurun_out/r02_sampling_plain.log
