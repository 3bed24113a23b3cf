#!/bin/bash
for c in 0 1 2 3 4; do
  echo "== category $c"
  timeout 300 ncu --metrics sm__icc_request_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:sort_track_kernel -s 1 -c 1 python scripts/icc_by_class.py 30 $c 2>&1 | grep -E "icc|time_duration|inst_executed|no_instruction|issue_active|rows"
done
