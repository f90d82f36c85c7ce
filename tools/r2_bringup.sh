#!/bin/bash
# First GPU call of the next round (gpurun --timeout 1500 -- 'bash tools/r2_bringup.sh'): everything that was prepared without a GPU.
#   1. the full GPU suite (includes tc_mixed, the fused optimizers and the Trainer end-to-end test)
#   2. the two-level march (PSN_PREC_TC_TWOLEVEL) - gated tests under a timeout (a protocol error would hang, not fail)
#   3. bench lines for tc / tc_mixed / tc_two_level (no extras) and the default bench
#   3b. the H16 variant of the mixed radiance program (PSNERF_B200_RAD_H16=1): mixed tests + bench A/B
#   3c. L2 evict_last hints on the radiance kernel's scratch (PSNERF_B200_STASH_HINT=1): tests, bench A/B, DRAM bytes under ncu
#   4. the ncu launch list of the default bench command and a --set full capture of the two tensor kernels at tc_mixed
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu.log
PSNERF_B200_TEST_TWOLEVEL=1 timeout -k 5 180 python -m pytest tests/test_gpu_tc_two_level.py -x -q > gpurun_out/r2_two_level.log 2>&1
rc=$?; echo "two-level rc=$rc" | tee -a gpurun_out/r2_two_level.log; tail -4 gpurun_out/r2_two_level.log
for p in tc tc_mixed $([ $rc -eq 0 ] && echo tc_two_level); do
  timeout -k 5 120 python bench.py --precision $p --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_$p.json 2> gpurun_out/r2_bench_$p.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_$p.json")); k = d["kernels"]
    print("$p", "step %.1f ms" % d["ms_per_step"], "march %.1f" % k["occ_march"]["ms_per_launch"], "rad %.1f" % k["radiance"]["ms_per_launch"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$p failed:", e)
PY
done
# H16 variant of the mixed radiance program (sigma' rebuilt from the fp16 activations in the reverse layers): same tests, then A/B
PSNERF_B200_RAD_H16=1 timeout -k 5 180 python -m pytest tests/test_gpu_tc_mixed.py -x -q > gpurun_out/r2_h16.log 2>&1
rch=$?; echo "h16 rc=$rch" | tee -a gpurun_out/r2_h16.log; tail -3 gpurun_out/r2_h16.log
if [ $rch -eq 0 ]; then
  PSNERF_B200_RAD_H16=1 timeout -k 5 120 python bench.py --precision tc_mixed --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_tc_mixed_h16.json 2> gpurun_out/r2_bench_tc_mixed_h16.err
  python -c "import json; d=json.load(open('gpurun_out/r2_bench_tc_mixed_h16.json')); print('tc_mixed+h16 step %.1f ms rad %.1f' % (d['ms_per_step'], d['kernels']['radiance']['ms_per_launch']))"
fi
# L2 evict_last policy on the stash / parked scratch of the radiance kernel (PSNERF_B200_STASH_HINT=1): results must not change at all
# (run the mixed tests), then the step time and the DRAM bytes of the kernel (was 20.5 GB read + 162 GB written per launch)
PSNERF_B200_STASH_HINT=1 timeout -k 5 180 python -m pytest tests/test_gpu_tc_mixed.py -x -q > gpurun_out/r2_hint.log 2>&1
rcs=$?; echo "stash-hint rc=$rcs" | tee -a gpurun_out/r2_hint.log
if [ $rcs -eq 0 ]; then
  PSNERF_B200_STASH_HINT=1 timeout -k 5 120 python bench.py --precision tc_mixed --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_tc_mixed_hint.json 2> gpurun_out/r2_bench_tc_mixed_hint.err
  python -c "import json; d=json.load(open('gpurun_out/r2_bench_tc_mixed_hint.json')); print('tc_mixed+stash-hint step %.1f ms rad %.1f clk %s' % (d['ms_per_step'], d['kernels']['radiance']['ms_per_launch'], d['clocks']['sm_mhz']))"
  PSNERF_B200_STASH_HINT=1 timeout -k 5 200 ncu --csv --log-file gpurun_out/r2_rad_hint_metrics.csv --clock-control none -k regex:k_tc_rad -c 1 \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct python tools/profile_step.py --steps 1 --precision tc_mixed > gpurun_out/r2_rad_hint_ncu.log 2>&1
  grep -E "dram__bytes|gpu__time" gpurun_out/r2_rad_hint_metrics.csv | cut -d, -f13-15
fi
timeout 400 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 600 gpurun_out/r2_bench_default.json
timeout 300 python tools/tc_trace_rad.py --mixed > gpurun_out/r2_trace_rad_mixed.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launch_list.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_launch_list.log 2>&1
if [ -n "$PSN_NCU" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_(occ|rad)" -c 4 -o gpurun_out/r2_prof_s1 \
    python tools/profile_step.py --steps 1 --precision tc_mixed > gpurun_out/r2_ncu_s1.log 2>&1; tail -3 gpurun_out/r2_ncu_s1.log
fi
