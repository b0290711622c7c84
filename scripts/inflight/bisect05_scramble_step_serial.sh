#!/bin/bash
mkdir -p gpurun_out
{
echo "== scramble, one context"; LGPU_DBG_SCRAMBLE=1 timeout 200 python scripts/inflight/scramble_check.py 30 2>&1 | tail -5
echo "== no scramble, one context"; timeout 200 python scripts/inflight/scramble_check.py 30 2>&1 | tail -3
echo "== steps serialised (persist)"; LGPU_DBG_SERIAL=8 timeout 250 python scripts/inflight/modes.py persist 200 2>&1 | tail -5
} > gpurun_out/inflight_bisect5.txt 2>&1
cat gpurun_out/inflight_bisect5.txt
