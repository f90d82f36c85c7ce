#!/bin/bash
# Round-2 GPU call 19: final train GEMM (single-buffer loaders, run-time epilogue choice): tests, microbenchmark, train steps fused vs unfused.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_train_stage1.py tests/test_gpu_train.py -x -q > gpurun_out/r2u_tests.log 2>&1; tail -2 gpurun_out/r2u_tests.log
timeout 200 python tools/time_gemm.py > gpurun_out/r2u_time_gemm.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2u_time_gemm.json"))
print("final", " ".join("%s %.3f"%(k.split("_")[0]+k[-4:],v["ms"]) for k,v in d.items()))
PY
for f in 1 0 1 0; do
PSNERF_B200_TRAIN_FUSED=$f PROFILE_TRAIN_REPS=3 timeout 300 python tools/profile_train.py > gpurun_out/r2u_train_fused$f.log 2>&1; echo fused=$f; tail -1 gpurun_out/r2u_train_fused$f.log | cut -c1-400
done
