#!/bin/bash
# A/B helper: bench.py (no extras, no CPU baseline) once per library build found under psnerf_b200/lib*/
mkdir -p gpurun_out
for d in psnerf_b200/lib psnerf_b200/lib_*; do
  [ -f $d/libpsnerf_b200.so ] || continue
  n=$(basename $d)
  PSNERF_B200_LIB=$PWD/$d/libpsnerf_b200.so timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 4 --warmup 3 > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$n.json"))
k=d["kernels"]; o=d.get("other_workloads",{})
s=d.get("secondary",{}).get("stage1_unisurf_512x512x128spp",{})
print("$n", "relit step %.1f ms"%d["ms_per_step"], "march %.1f"%k["occ_march"]["ms_per_step"], "shadow %.1f"%k["shadow"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"],
      "| stage-1 render %.1f ms"%s.get("ms_per_step",-1), "march %.1f"%s.get("kernels",{}).get("occ_march",{}).get("ms_per_step",-1))
PY
done
