#!/bin/bash
REPS=${1:-100}
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; env LGPU_DBG_DUAL=1 "$@" timeout 250 python scripts/inflight/modes.py persist $REPS > gpurun_out/dual_$name.txt 2>&1
  echo "solves with mismatch: $(grep -c 'lgpu dual' gpurun_out/dual_$name.txt)"
  grep "lgpu dual" gpurun_out/dual_$name.txt | sed 's/.*operator \([0-9]*\), step w \([0-9]*\), step h \([0-9]*\).*/op \1 w \2 h \3/' | sort | uniq -c | sort -rn | head -4
  tail -1 gpurun_out/dual_$name.txt; }
{
run l2only LGPU_LIB=$PWD/legolas_b200/liblegolas_b200_l2only.so
run nonc LGPU_LIB=$PWD/legolas_b200/liblegolas_b200_nonc.so
} 2>&1 | tee gpurun_out/inflight_bisect8.txt
