#!/bin/bash
REPS=${1:-120}
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; env LGPU_DBG_DUAL=1 "$@" timeout 250 python scripts/inflight/modes.py persist $REPS > gpurun_out/dual_$name.txt 2>&1
  echo "solves with mismatch: $(grep -c 'lgpu dual' gpurun_out/dual_$name.txt)"
  grep "lgpu dual" gpurun_out/dual_$name.txt | sed 's/.*operator \([0-9]*\), step w \([0-9]*\), step h \([0-9]*\).*/op \1 w \2 h \3/' | sort | uniq -c | sort -rn | head -6
  tail -1 gpurun_out/dual_$name.txt; }
{
run pdl0 LGPU_PDL=0
run upper0 LGPU_SLU_UPPER=0
run bxfuse0 LGPU_BX_FUSE=0
run alloff LGPU_PDL=0 LGPU_SLU_UPPER=0 LGPU_SLU_FUSE=0 LGPU_BX_FUSE=0 LGPU_CGS2_FUSED=0
} 2>&1 | tee gpurun_out/inflight_bisect6.txt
