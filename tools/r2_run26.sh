#!/bin/bash
# Round-2 GPU call 26 (last of the budget): next-tile point prefetch in k_tc_occ: full GPU suite (points are bit-identical by construction),
# then A/B against the previous library.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/r2gp_tests.log 2>&1; tail -2 gpurun_out/r2gp_tests.log
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/r2gp_ab.log
