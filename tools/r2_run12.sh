#!/bin/bash
# Round-2 GPU call 12: epilogue-skew A/B (lib, lib_v_skew*), train-GEMM microbenchmark + one full ncu capture per operand form.
mkdir -p gpurun_out
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/r2n_ab.log
timeout 300 python tools/time_gemm.py > gpurun_out/r2n_time_gemm.json 2> gpurun_out/r2n_time_gemm.err; cat gpurun_out/r2n_time_gemm.json | tr -d '\n' | cut -c1-1500; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 9 -c 3 -o gpurun_out/r2n_gemm_full python tools/time_gemm.py 131072 > gpurun_out/r2n_ncu.log 2>&1; tail -3 gpurun_out/r2n_ncu.log
