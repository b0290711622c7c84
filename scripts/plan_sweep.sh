#!/bin/bash
# operator-application time for different stage plans of the structured solve
for cfg in "4 2 8" "4 1 8" "4 2 4" "4 2 16" "4 3 8" "3 2 8" "3 1 8" "4 1 4" "3 2 4" "4 2 2" "5 2 8"; do
  set -- $cfg
  LGPU_SLU_MU0=$1 LGPU_SLU_MU1=$2 LGPU_SLU_TOP=$3 timeout 120 python scripts/op_time.py 10001 2>&1 | tail -1
done
