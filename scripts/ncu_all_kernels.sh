#!/bin/bash
# GPU box: per-launch counters of EVERY kernel of one step (pyramid + encoder, 8 x 30k fragments), as CSV.
#   bash scripts/ncu_all_kernels.sh gpurun_out/<run>            -> <run>/all_kernels.csv (+ fused variant)
set -u
out=$1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $out/all_kernels.csv python scripts/one_step.py > $out/ncu_all.log 2>&1
[ "${FUSED:-0}" = 1 ] && D3F_FUSED_KPCONV=1 timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $out/all_kernels_fused.csv python scripts/one_step.py > $out/ncu_all_fused.log 2>&1
# the level-0 KPConv alone, three calls: DRAM bytes of the whole operator (bench.py's roofline.traffic)
ONLY=kpconv timeout 200 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $out/kpconv_op.csv python scripts/ncu_targets.py > $out/ncu_kpconv_op.log 2>&1
tail -2 $out/ncu_all.log
