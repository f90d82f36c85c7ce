#!/bin/bash
# Round-2 GPU call 24: softplus stack in the scaled domain (no FMUL per activation; weights / logit row repacked): full parity suite with
# the measured-error log (no asserts), then A/B against the previous library (psnerf_b200/lib_v_old).
mkdir -p gpurun_out
rm -f gpurun_out/r2sc_errlog.jsonl
PSNERF_B200_ERRLOG=gpurun_out/r2sc_errlog.jsonl PSNERF_B200_ERRLOG_NOASSERT=1 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2sc_tests.log 2>&1; tail -4 gpurun_out/r2sc_tests.log
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/r2sc_ab.log
