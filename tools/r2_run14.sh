#!/bin/bash
# Round-2 GPU call 14: second-generation train GEMM (cp.async staging + converter warps + fused epilogues): tests, microbenchmark v2 vs v1, train steps.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q > gpurun_out/r2p_gemm_tests.log 2>&1; tail -5 gpurun_out/r2p_gemm_tests.log
timeout 200 python tools/time_gemm.py > gpurun_out/r2p_time_gemm_v2.json 2> gpurun_out/r2p_time_gemm_v2.err; tr -d '\n' < gpurun_out/r2p_time_gemm_v2.json | cut -c1-1400; echo; tail -2 gpurun_out/r2p_time_gemm_v2.err
PSNERF_B200_GEMM_V1=1 timeout 200 python tools/time_gemm.py > gpurun_out/r2p_time_gemm_v1.json 2>/dev/null
timeout 400 python -m pytest tests/test_gpu_train_stage1.py tests/test_gpu_train.py -x -q > gpurun_out/r2p_train_tests.log 2>&1; tail -5 gpurun_out/r2p_train_tests.log
PROFILE_TRAIN_REPS=2 timeout 300 python tools/profile_train.py > gpurun_out/r2p_train.log 2>&1; tail -1 gpurun_out/r2p_train.log | cut -c1-600
PSNERF_B200_GEMM_V1=1 PROFILE_TRAIN_REPS=2 timeout 300 python tools/profile_train.py > gpurun_out/r2p_train_v1.log 2>&1; tail -1 gpurun_out/r2p_train_v1.log | cut -c1-600
