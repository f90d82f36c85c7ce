#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_tc_gemm.py -q > gpurun_out/r2i_gemm_mn.log 2>&1; echo "gemm (MN-major tiles) rc=$?"; tail -4 gpurun_out/r2i_gemm_mn.log
PSNERF_B200_GEMM_MN=0 timeout -k 5 300 python -m pytest tests/test_gpu_tc_gemm.py -q > gpurun_out/r2i_gemm_k.log 2>&1; echo "gemm (K-major transposing) rc=$?"; tail -2 gpurun_out/r2i_gemm_k.log
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_stage1.py -q > gpurun_out/r2i_tests.log 2>&1; tail -3 gpurun_out/r2i_tests.log
PROFILE_TRAIN_REPS=2 timeout 300 python tools/profile_train.py > gpurun_out/r2i_train_mn.log 2>&1; tail -2 gpurun_out/r2i_train_mn.log
PSNERF_B200_GEMM_MN=0 PROFILE_TRAIN_REPS=2 timeout 300 python tools/profile_train.py > gpurun_out/r2i_train_k.log 2>&1; tail -1 gpurun_out/r2i_train_k.log
timeout 600 python tools/strong_breakdown.py 2 8 > gpurun_out/r2i_strong_breakdown.json 2> gpurun_out/r2i_strong_breakdown.err; python -c "
import json; d=json.load(open('gpurun_out/r2i_strong_breakdown.json'))
for k,v in d.items(): print(k, round(v['chain+rows']['ms'],1), v['chain+rows']['kernels'])"
