#!/bin/bash
# Round-2 GPU call 20: clusters of 4 CTAs sharing one weight stream (PSN_CLUSTER=4) against the pairs.
mkdir -p gpurun_out
PSNERF_B200_LIB=$PWD/psnerf_b200/lib_v_cl4/libpsnerf_b200.so timeout 240 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_two_level.py tests/test_gpu_tc_mixed.py -x -q > gpurun_out/r2v_cl4_tests.log 2>&1; tail -2 gpurun_out/r2v_cl4_tests.log
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/r2v_ab.log
