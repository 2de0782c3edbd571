#!/bin/bash
# A/B of the field kernel's op granularity: NSFF_TC_FINE=1 (one K-chunk per op, order a0 b0 a1 | a2 a3 | b1 b2 b3) vs 0
# (two K-chunks per op, a01 b01 a23 b23).  usage: tools/ab_fine.sh <tag>
tag=${1:-ab}
for mode in 1 0; do
  echo "=== NSFF_TC_FINE=$mode"
  NSFF_TC_FINE=$mode tools/quick_check.sh ${tag}_fine${mode} 2>&1 | tail -4
done
