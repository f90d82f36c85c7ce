#!/bin/bash
# A/B helper: bench.py (no extras, no CPU baseline) once per library build found under psnerf_b200/lib*/
mkdir -p gpurun_out
for d in psnerf_b200/lib psnerf_b200/lib_*; do
  [ -f $d/libpsnerf_b200.so ] || continue
  n=$(basename $d)
  PSNERF_B200_LIB=$PWD/$d/libpsnerf_b200.so timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$n.json"))
k=d["kernels"]; o=d.get("other_workloads",{})
print("$n", "step %.1f ms"%d["ms_per_step"], "march %.1f"%k["occ_march"]["ms_per_launch"], "rad %.1f"%k["radiance"]["ms_per_launch"], "clk", d["clocks"]["sm_mhz"],
      "| s2 %.1f ms"%o.get("stage2_shade_512x512x96L",{}).get("ms",-1), "shadow %.0f ms"%o.get("shadow_visibility_96L_x128",{}).get("ms",-1))
PY
done
