#!/bin/bash
REPS=${1:-150}
mkdir -p gpurun_out
run() { name=$1; mode=$2; shift 2; echo "== $name" ; env "$@" timeout 200 python scripts/inflight/modes.py $mode $REPS 2>&1 | tail -6; }
{
run serial_assembly persist LGPU_DBG_SERIAL=1
run serial_factor persist LGPU_DBG_SERIAL=2
run serial_arnoldi persist LGPU_DBG_SERIAL=4
echo "== racecheck"
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python scripts/inflight/race_small.py 2>&1 | grep -v "^=========     at\|^=========         in\|^=========     by\|Host Frame\|\.so\|^=========$" | head -120
} > gpurun_out/inflight_bisect3.txt 2>&1
tail -150 gpurun_out/inflight_bisect3.txt
