#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/strong_breakdown.py 1 2 8 > gpurun_out/r2h_strong_breakdown.json 2> gpurun_out/r2h_strong_breakdown.err; cat gpurun_out/r2h_strong_breakdown.json | head -80
PROFILE_TRAIN_REPS=2 timeout 300 python tools/profile_train.py > gpurun_out/r2h_train_x2.log 2>&1; tail -2 gpurun_out/r2h_train_x2.log
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_train.py tests/test_gpu_train_stage1.py -q > gpurun_out/r2h_tests.log 2>&1; tail -3 gpurun_out/r2h_tests.log
