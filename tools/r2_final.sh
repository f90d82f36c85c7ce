#!/bin/bash
# Round-2 closing GPU call: full suite (measured-error log), smoke, default bench + reference arm, launch list of the bench, and the
# --set full captures of the top kernels (relit chain, stage-1 render, train GEMM).  Outputs -> gpurun_out/r2z_*; copied to profiles/ by hand.
mkdir -p gpurun_out
rm -f gpurun_out/r2z_errlog.jsonl
(time PSNERF_B200_ERRLOG=gpurun_out/r2z_errlog.jsonl timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2z_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 400 gpurun_out/r2z_bench.json; tail -3 gpurun_out/r2z_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r2z_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2z_launch_list.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-secondary > gpurun_out/r2z_launch_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2z_launch_list.csv 12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_occ -c 8 -o gpurun_out/r2z_relit_occ python tools/profile_step.py --relit --steps 1 > gpurun_out/r2z_ncu_relit.log 2>&1; tail -1 gpurun_out/r2z_ncu_relit.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tc_rad -c 1 -o gpurun_out/r2z_render_rad python tools/profile_step.py --steps 1 > gpurun_out/r2z_ncu_rad.log 2>&1; tail -1 gpurun_out/r2z_ncu_rad.log
for s in 9 22 35; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s $s -c 1 -o gpurun_out/r2z_gemm_s$s python tools/time_gemm.py 131072 > gpurun_out/r2z_ncu_gemm_$s.log 2>&1; tail -1 gpurun_out/r2z_ncu_gemm_$s.log
done
