#!/bin/bash
# Round-2 GPU call 15: A/B of the epilogue variants on top of the unrolled pass loop; full ncu capture of the second-generation GEMM (NT form).
mkdir -p gpurun_out
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/r2q_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm2 -s 9 -c 1 -o gpurun_out/r2q_gemm2_full python tools/time_gemm.py 131072 > gpurun_out/r2q_ncu.log 2>&1; tail -2 gpurun_out/r2q_ncu.log
