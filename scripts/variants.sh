#!/bin/bash
# occupancy experiment for the SORT kernel (see sort.cu W2T_SORT_VARIANT)
for v in ${2:-0 1 2 3}; do
  echo "variant $v"; W2T_SORT_VARIANT=$v timeout 300 python scripts/breakdown.py ${1:-150} 2>&1 | grep it2
done
