#!/bin/bash
# Round-2 GPU call 5: dead-ray cull + GEMM loader rework: suite with the final gates, smoke, bench, train-step timing x3, train launch list.
mkdir -p gpurun_out
rm -f gpurun_out/r2e_errlog.jsonl
(time PSNERF_B200_ERRLOG=gpurun_out/r2e_errlog.jsonl timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2e_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; tail -2 gpurun_out/r2e_smoke.log
timeout 600 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 700 gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e_bench.err
PROFILE_TRAIN_REPS=3 timeout 300 python tools/profile_train.py > gpurun_out/r2e_train_x3.log 2>&1; tail -3 gpurun_out/r2e_train_x3.log
PSNERF_B200_TRAIN_GEMM=ffma PROFILE_TRAIN_REPS=2 timeout 300 python tools/profile_train.py > gpurun_out/r2e_train_ffma.log 2>&1; tail -2 gpurun_out/r2e_train_ffma.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2e_train_launches.csv \
  python tools/profile_train.py > gpurun_out/r2e_train_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r2e_train_launches.csv 12
