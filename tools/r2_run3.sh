#!/bin/bash
# Round-2 GPU call 3: tf32x3 GEMM bring-up (under a timeout: a protocol slip would hang), full suite with the error log, bench.
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q > gpurun_out/r2c_gemm.log 2>&1; echo "gemm rc=$?"; tail -12 gpurun_out/r2c_gemm.log
rm -f gpurun_out/r2c_errlog.jsonl
(time PSNERF_B200_ERRLOG=gpurun_out/r2c_errlog.jsonl PSNERF_B200_ERRLOG_NOASSERT=1 timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_tc_gemm.py) > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2c_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 900 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
PSNERF_B200_TRAIN_GEMM=ffma timeout 300 python - > gpurun_out/r2c_train_ffma.log 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda:0')
cfg, net, rend, conf, ps = bench.build_models(dev, 'tc_two_level')
print(json.dumps(bench.train_steps(dev, rend, ps, bench.scene(0), 1, 0)))
PY
tail -2 gpurun_out/r2c_train_ffma.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tc_occ" --launch-skip 9 -c 1 -o gpurun_out/r2c_prof_shadow \
  python tools/profile_step.py --steps 1 --relit > gpurun_out/r2c_ncu_shadow.log 2>&1; tail -2 gpurun_out/r2c_ncu_shadow.log
