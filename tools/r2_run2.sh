#!/bin/bash
# Round-2 GPU call 2: the GPU suite with the measured-error log, the new headline bench, launch list + --set full capture of the relit step.
mkdir -p gpurun_out
rm -f gpurun_out/r2_errlog.jsonl
(time PSNERF_B200_ERRLOG=gpurun_out/r2_errlog.jsonl PSNERF_B200_ERRLOG_NOASSERT=1 timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r2b_pytest_gpu.log 2>&1; tail -15 gpurun_out/r2b_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launch_list.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_launch_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tc_occ" -c 6 -o gpurun_out/r2b_prof_relit \
  python tools/profile_step.py --steps 1 --relit > gpurun_out/r2b_ncu_relit.log 2>&1; tail -3 gpurun_out/r2b_ncu_relit.log
timeout 300 ncu --clock-control none -k regex:"k_shadow_plan|k_shadow_composite_list|k_march_scan|k_s2_shade|k_march_refine_select|k_sphere_far|k_rays_from" -c 12 --csv --log-file gpurun_out/r2b_hbm_kernels.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
  python tools/profile_step.py --steps 1 --relit > gpurun_out/r2b_hbm_kernels.log 2>&1
