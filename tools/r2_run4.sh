#!/bin/bash
# Round-2 GPU call 4: fp32 fused secant + MODE_FEAT infer_occ + GEMM gates: full suite with the error log, bench, train-step launch list.
mkdir -p gpurun_out
rm -f gpurun_out/r2d_errlog.jsonl
(time PSNERF_B200_ERRLOG=gpurun_out/r2d_errlog.jsonl timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2d_pytest_gpu.log 2>&1; tail -12 gpurun_out/r2d_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 600 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2d_train_launches.csv \
  python tools/profile_train.py > gpurun_out/r2d_train_launches.log 2>&1; tail -2 gpurun_out/r2d_train_launches.log
python tools/launch_summary.py gpurun_out/r2d_train_launches.csv 30
