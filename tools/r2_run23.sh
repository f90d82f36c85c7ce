#!/bin/bash
# Round-2 GPU call 23: point encoding by angle doubling (PSN_PE_DOUBLING=1) against the direct sincosf per octave: measured errors + A/B.
mkdir -p gpurun_out
rm -f gpurun_out/r2pe_errlog.jsonl
PSNERF_B200_LIB=$PWD/psnerf_b200/lib_v_pe2/libpsnerf_b200.so PSNERF_B200_ERRLOG=gpurun_out/r2pe_errlog.jsonl PSNERF_B200_ERRLOG_NOASSERT=1 timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stage1.py tests/test_gpu_parity_at_size.py tests/test_gpu_tc_two_level.py tests/test_gpu_pipeline.py -q > gpurun_out/r2pe_tests.log 2>&1; tail -3 gpurun_out/r2pe_tests.log
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/r2pe_ab.log
