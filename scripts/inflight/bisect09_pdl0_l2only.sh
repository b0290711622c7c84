#!/bin/bash
REPS=${1:-300}
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; env "$@" timeout 250 python scripts/inflight/modes.py persist $REPS 2>&1 | grep -v Warning | grep -v "omega, _" | tail -4; }
{
run pdl0 LGPU_PDL=0
run pdl0_l2only LGPU_PDL=0 LGPU_LIB=$PWD/legolas_b200/liblegolas_b200_l2only.so
} 2>&1 | tee gpurun_out/inflight_bisect9.txt
