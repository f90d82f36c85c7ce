#!/bin/bash
# Round-2 GPU call 18: train-GEMM 2 x 2: epilogue {transposed/coalesced, row-per-thread} x loader registers {double, single}, both generations.
mkdir -p gpurun_out
for d in lib_g_r0d1 lib_g_r1d1 lib_g_r0d0 lib_g_r1d0; do
  for g in 1 2; do
    PSNERF_B200_LIB=$PWD/psnerf_b200/$d/libpsnerf_b200.so PSNERF_B200_GEMM_GEN=$g timeout 200 python tools/time_gemm.py > gpurun_out/r2t_gemm_${d}_gen$g.json 2>/dev/null
    python - <<PY
import json
d=json.load(open("gpurun_out/r2t_gemm_${d}_gen$g.json"))
print("$d gen$g", " ".join("%s %.3f"%(k.split("_")[0]+k[-4:],v["ms"]) for k,v in d.items()))
PY
  done
done
