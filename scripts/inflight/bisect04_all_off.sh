#!/bin/bash
REPS=${1:-200}
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name" ; env "$@" timeout 200 python scripts/inflight/modes.py persist $REPS 2>&1 | tail -5; }
{
run alloff LGPU_PDL=0 LGPU_SLU_UPPER=0 LGPU_SLU_FUSE=0 LGPU_BX_FUSE=0 LGPU_CGS2_FUSED=0
run alloff_but_cgs2 LGPU_PDL=0 LGPU_SLU_UPPER=0 LGPU_SLU_FUSE=0 LGPU_BX_FUSE=0
run alloff_but_upper LGPU_PDL=0 LGPU_BX_FUSE=0 LGPU_CGS2_FUSED=0
} > gpurun_out/inflight_bisect4.txt 2>&1
cat gpurun_out/inflight_bisect4.txt
