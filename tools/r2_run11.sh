#!/bin/bash
# Round-2 GPU call 11: full suite (final gates) + smoke + default bench + launch list of the bench + full ncu of the top kernels.
mkdir -p gpurun_out
rm -f gpurun_out/r2k_errlog.jsonl
(time PSNERF_B200_ERRLOG=gpurun_out/r2k_errlog.jsonl timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2k_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2k_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -1 gpurun_out/r2k_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -c 500 gpurun_out/r2k_bench.json; tail -5 gpurun_out/r2k_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2k_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2k_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2k_launch_list.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2k_launch_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2k_launch_list.csv 16
