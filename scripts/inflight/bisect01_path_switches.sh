#!/bin/bash
# Scratch: three-in-flight vs sequential (scripts/inflight/check.py) under one path switch at a time.
REPS=${1:-250}
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name" ; env "$@" timeout 300 python scripts/inflight/check.py $REPS 2>&1 | grep -v identical | tail -8; }
{
run default LGPU_NOP=1
run pdl0 LGPU_PDL=0
run nonewcol LGPU_DBG_NONEWCOL=1
run upper0 LGPU_SLU_UPPER=0
run bxfuse0 LGPU_BX_FUSE=0
run cgsfused0 LGPU_CGS2_FUSED=0
} > gpurun_out/inflight_bisect.txt 2>&1
cat gpurun_out/inflight_bisect.txt
