#!/bin/bash
# Round-2 GPU call 16: mbarrier suspend-hint A/B (100 ns / 2000 ns / 20 us) on the train GEMM (both generations) and on the inference kernels;
# generation 1 now has double-buffered loader registers + the cheaper split.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q > gpurun_out/r2r_gemm_tests.log 2>&1; tail -2 gpurun_out/r2r_gemm_tests.log
for d in lib lib_v_hint100 lib_v_hint20000; do
  for g in 1 2; do
    PSNERF_B200_LIB=$PWD/psnerf_b200/$d/libpsnerf_b200.so PSNERF_B200_GEMM_GEN=$g timeout 200 python tools/time_gemm.py > gpurun_out/r2r_gemm_${d}_gen$g.json 2>/dev/null
    python - <<PY
import json
d=json.load(open("gpurun_out/r2r_gemm_${d}_gen$g.json"))
print("$d gen$g", " ".join("%s %.3f"%(k.split("_")[0]+k[-4:],v["ms"]) for k,v in d.items()))
PY
  done
done
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/r2r_ab.log
PROFILE_TRAIN_REPS=3 timeout 300 python tools/profile_train.py > gpurun_out/r2r_train.log 2>&1; tail -1 gpurun_out/r2r_train.log | cut -c1-600
