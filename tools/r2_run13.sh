#!/bin/bash
# Round-2 GPU call 13: epilogue A/B (skew with test_wait detection, fully unrolled pass loop), train steps after colsum / sync-free losses.
mkdir -p gpurun_out
bash tools/ab_bench.sh 2>&1 | tee gpurun_out/r2o_ab.log
PROFILE_TRAIN_REPS=2 timeout 300 python tools/profile_train.py > gpurun_out/r2o_train.log 2>&1; tail -2 gpurun_out/r2o_train.log | cut -c1-600
