#!/bin/bash
# One gpurun call: tensor-path bring-up report (per-layer error vs the CPU oracle), the GPU tests, a tile timeline and the bench;
# PSN_NCU=1 adds the `ncu --set full` capture of the stage-1 tensor kernels.
mkdir -p gpurun_out
echo "== bringup std" > gpurun_out/bringup.log
timeout 240 python tests/tc_bringup.py >> gpurun_out/bringup.log 2>&1; echo "rc=$?" >> gpurun_out/bringup.log
cat gpurun_out/bringup.log
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python tools/tc_trace.py > gpurun_out/trace.log 2>&1; cat gpurun_out/trace.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ -n "$PSN_NCU" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_(occ|rad)" -c 11 -o gpurun_out/prof_s1 python tools/profile_step.py --steps 1 > gpurun_out/ncu_s1.log 2>&1; tail -3 gpurun_out/ncu_s1.log
fi
