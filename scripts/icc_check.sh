#!/bin/bash
# quick look at the SORT kernel's instruction-cache behaviour (30-segment launch) + parity + full-size breakdown
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 ncu --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:sort_track_kernel -s 1 -c 1 python bench.py --segments 30 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "icc|gcc|time_duration|inst_executed|no_instruction|issue_active"
timeout 300 python scripts/breakdown.py 150 2>&1 | grep it2
