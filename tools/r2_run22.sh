#!/bin/bash
# Round-2 GPU call 22: radiance kernel back on the rolled pass loop (DRAM write-backs), fused-vs-separate train test, radiance-path tests, bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train_stage1.py tests/test_gpu_tc_mixed.py tests/test_gpu_stage1.py -x -q > gpurun_out/r2x_tests.log 2>&1; tail -2 gpurun_out/r2x_tests.log
timeout 300 ncu --csv --log-file gpurun_out/r2x_rad_dram.csv --clock-control none -k regex:k_tc_rad -c 2 \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    python tools/profile_step.py --steps 2 --precision tc_two_level > gpurun_out/r2x_rad_dram.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r2x_rad_dram.csv')))
hdr=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
h=rows[hdr]
for r in rows[hdr+1:]:
    if len(r)==len(h):
        d=dict(zip(h,r)); print(d['ID'], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
timeout 400 python bench.py --no-cpu-baseline --no-extras --steps 4 --warmup 3 > gpurun_out/r2x_bench.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/r2x_bench.json")); s=d["secondary"]["stage1_unisurf_512x512x128spp"]
print("relit %.1f ms"%d["ms_per_step"], "shadow %.1f"%d["kernels"]["shadow"]["ms_per_step"], "| render %.1f ms"%s["ms_per_step"], "radiance %.1f"%s["kernels"]["radiance"]["ms_per_step"], d["clocks"]["sm_mhz"])
PY
