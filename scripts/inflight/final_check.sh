#!/bin/bash
# In-flight reproducibility of the product library: amplified check (every operator application and Gram-Schmidt step
# evaluated twice, LGPU_DBG_DUAL) and plain repetitions, three contexts in flight (scripts/inflight/modes.py).
mkdir -p gpurun_out
{
echo "== dual evaluation, three contexts in flight"
LGPU_DBG_DUAL=1 timeout 120 python scripts/inflight/modes.py persist ${1:-100} > gpurun_out/dual_final.txt 2>&1
echo "solves with mismatch: $(grep -c 'lgpu dual' gpurun_out/dual_final.txt)"; tail -1 gpurun_out/dual_final.txt
echo "== plain repetitions"
timeout 100 python scripts/inflight/modes.py persist ${2:-400} 2>&1 | grep -v Warning | grep -v "omega, _" | tail -3
} 2>&1 | tee gpurun_out/inflight_final.txt
