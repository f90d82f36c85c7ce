#!/bin/bash
# Round-2 GPU call 21: launch list of the train steps (fused epilogues), metrics-only DRAM capture of the radiance kernel (apples to apples with
# r2_ncu_rad_stash_discard.json, which was a 5-metric capture: --set full replays the kernel ~40 times and changes what the L2 holds).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2w_train_launches.csv python tools/profile_train.py > gpurun_out/r2w_train_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r2w_train_launches.csv 14
timeout 300 ncu --csv --log-file gpurun_out/r2w_rad_dram.csv --clock-control none -k regex:k_tc_rad -c 2 \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    python tools/profile_step.py --steps 2 --precision tc_two_level > gpurun_out/r2w_rad_dram.log 2>&1
grep -E "dram__bytes|gpu__time|tensor" gpurun_out/r2w_rad_dram.csv | cut -d, -f5,13-15 | cut -c1-160
