#!/bin/bash
# Round-2 GPU call 17: coalesced GEMM epilogue (transpose buffer): tests, microbenchmark of both generations, train-step tests and timings.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q > gpurun_out/r2s_gemm_tests.log 2>&1; tail -2 gpurun_out/r2s_gemm_tests.log
PSNERF_B200_GEMM_GEN=2 timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q > gpurun_out/r2s_gemm_tests_gen2.log 2>&1; tail -2 gpurun_out/r2s_gemm_tests_gen2.log
for g in 1 2; do
  PSNERF_B200_GEMM_GEN=$g timeout 200 python tools/time_gemm.py > gpurun_out/r2s_gemm_gen$g.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/r2s_gemm_gen$g.json"))
print("gen$g", " ".join("%s %.3f"%(k.split("_")[0]+k[-4:],v["ms"]) for k,v in d.items()))
PY
done
timeout 400 python -m pytest tests/test_gpu_train_stage1.py tests/test_gpu_train.py -x -q > gpurun_out/r2s_train_tests.log 2>&1; tail -3 gpurun_out/r2s_train_tests.log
for g in 1 2; do
PSNERF_B200_GEMM_GEN=$g PROFILE_TRAIN_REPS=3 timeout 300 python tools/profile_train.py > gpurun_out/r2s_train_gen$g.log 2>&1; tail -1 gpurun_out/r2s_train_gen$g.log | cut -c1-500
done
PSNERF_B200_TRAIN_FUSED=0 PROFILE_TRAIN_REPS=3 timeout 300 python tools/profile_train.py > gpurun_out/r2s_train_unfused.log 2>&1; tail -1 gpurun_out/r2s_train_unfused.log | cut -c1-500
