#!/bin/bash
mkdir -p gpurun_out
{
echo "== dual, programmatic launches, fence + hints (product library)"
LGPU_DBG_DUAL=1 LGPU_DBG_SHARED_PDL=1 timeout 120 python scripts/inflight/modes.py persist 100 > gpurun_out/dual_fence.txt 2>&1
echo "solves with mismatch: $(grep -c 'lgpu dual' gpurun_out/dual_fence.txt)"; tail -1 gpurun_out/dual_fence.txt
echo "== plain runs, programmatic launches allowed"
LGPU_DBG_SHARED_PDL=1 timeout 100 python scripts/inflight/modes.py persist 400 2>&1 | grep -v Warning | grep -v "omega, _" | tail -3
echo "== plain runs, library default"
timeout 100 python scripts/inflight/modes.py persist 400 2>&1 | grep -v Warning | grep -v "omega, _" | tail -3
} 2>&1 | tee gpurun_out/inflight_bisect11.txt
