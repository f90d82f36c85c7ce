#!/bin/bash
# Round-2 multi-GPU call (gpurun --gpus N): NCCL equality check of the sharded paths, then the bench line with its multi_gpu block.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/run_sharded_check.py > gpurun_out/r2g_sharded_check_${N}gpu.log 2>&1; tail -4 gpurun_out/r2g_sharded_check_${N}gpu.log
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2g_bench_${N}gpu.json 2> gpurun_out/r2g_bench_${N}gpu.err; tail -c 2500 gpurun_out/r2g_bench_${N}gpu.json; tail -5 gpurun_out/r2g_bench_${N}gpu.err
