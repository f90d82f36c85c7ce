#!/bin/bash
mkdir -p gpurun_out
for m in 2 3; do
  PSNERF_B200_GEMM_MN=$m timeout -k 5 300 python -m pytest tests/test_gpu_tc_gemm.py -q > gpurun_out/r2j_gemm_mn$m.log 2>&1; echo "gemm MN layout $m rc=$?"; tail -2 gpurun_out/r2j_gemm_mn$m.log
  grep -m3 "AssertionError:" gpurun_out/r2j_gemm_mn$m.log
  PSNERF_B200_GEMM_MN=$m PROFILE_TRAIN_REPS=2 timeout 300 python tools/profile_train.py > gpurun_out/r2j_train_mn$m.log 2>&1; tail -1 gpurun_out/r2j_train_mn$m.log
done
PROFILE_TRAIN_REPS=2 timeout 300 python tools/profile_train.py > gpurun_out/r2j_train_k.log 2>&1; tail -1 gpurun_out/r2j_train_k.log
