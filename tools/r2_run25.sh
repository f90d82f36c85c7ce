#!/bin/bash
# Round-2 GPU call 25: lo operand by one mixed-precision FHFMA per element (fma.rn.f32.f16) instead of HADD2.F32 + FADD: tensor-path tests
# (results must be bit-identical to before: x - hi is exact), then A/B against the previous library.
mkdir -p gpurun_out
rm -f gpurun_out/r2fh_errlog.jsonl
PSNERF_B200_ERRLOG=gpurun_out/r2fh_errlog.jsonl timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2fh_tests.log 2>&1; tail -3 gpurun_out/r2fh_tests.log
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/r2fh_ab.log
