#!/bin/bash
# operator-application time for different stage plans of the structured solve
for cfg in "4 3 32" "4 3 8" "4 3 4" "4 2 32" "4 2 8" "4 4 32" "4 5 32" "5 3 32" "5 2 32" "5 4 32" "5 5 32" "3 3 32" "3 4 32" "5 3 8" "6 3 32" "6 2 32"; do
  set -- $cfg
  LGPU_SLU_MU0=$1 LGPU_SLU_MU1=$2 LGPU_SLU_TOP=$3 timeout 120 python scripts/op_time.py 10001 2>&1 | tail -1
done
