#!/bin/bash
# Round-2 GPU call 6: dead-ray cull (lead 16), GEMM integer rounding, flip-aware gradient gates; ncu of the train GEMM and the stash discard A/B.
mkdir -p gpurun_out
rm -f gpurun_out/r2f_errlog.jsonl
(time PSNERF_B200_ERRLOG=gpurun_out/r2f_errlog.jsonl timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2f_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 600 gpurun_out/r2f_bench.json; tail -5 gpurun_out/r2f_bench.err
PROFILE_TRAIN_REPS=2 timeout 300 python tools/profile_train.py > gpurun_out/r2f_train_x2.log 2>&1; tail -2 gpurun_out/r2f_train_x2.log
for hint in 0 1; do
  PSNERF_B200_STASH_HINT=$hint timeout 300 ncu --csv --log-file gpurun_out/r2f_rad_hint${hint}.csv --clock-control none -k regex:k_tc_rad -c 2 \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    python tools/profile_step.py --steps 1 --precision tc_two_level > gpurun_out/r2f_rad_hint${hint}.log 2>&1
  grep -E "dram__bytes|gpu__time" gpurun_out/r2f_rad_hint${hint}.csv | cut -d, -f5,13-15 | cut -c1-160
done
PSNERF_B200_STASH_HINT=1 timeout 300 python -m pytest tests/test_gpu_tc_mixed.py tests/test_gpu_parity_at_size.py -q -k "mixed or stage1_render" > gpurun_out/r2f_hint_tests.log 2>&1; tail -2 gpurun_out/r2f_hint_tests.log
PSNERF_B200_STASH_HINT=1 timeout 300 python bench.py --steps 3 --no-extras --no-cpu-baseline > gpurun_out/r2f_bench_hint.json 2>/dev/null
python -c "
import json
for f in ('r2f_bench','r2f_bench_hint'):
    d=json.load(open('gpurun_out/%s.json'%f)); s=d['secondary']['stage1_unisurf_512x512x128spp']
    print(f, 'relit %.1f ms'%d['ms_per_step'], 'stage1 render %.1f ms'%s['ms_per_step'], 'rad %.1f'%s['kernels']['radiance']['ms_per_launch'], d['clocks'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm --launch-skip 300 -c 40 -o gpurun_out/r2f_prof_gemm \
  python tools/profile_train.py > gpurun_out/r2f_ncu_gemm.log 2>&1; tail -2 gpurun_out/r2f_ncu_gemm.log
