#!/bin/bash
REPS=${1:-150}
mkdir -p gpurun_out
run() { name=$1; mode=$2; shift 2; echo "== $name" ; env "$@" timeout 300 python scripts/inflight/modes.py $mode $REPS 2>&1 | tail -8; }
{
run fresh fresh LGPU_NOP=1
run persist persist LGPU_NOP=1
run inflight_poison255 inflight LGPU_DBG_POISON=255
run fresh_poison255 fresh LGPU_DBG_POISON=255
run fresh_poison127 fresh LGPU_DBG_POISON=127
} > gpurun_out/inflight_bisect2.txt 2>&1
cat gpurun_out/inflight_bisect2.txt
