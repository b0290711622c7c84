#!/bin/bash
# quick GPU check used while tuning: a few parity tests, then the bench's per-kernel table
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("step ms", round(d["ms_per_step"], 2), "e2e", round(1e3 * d["e2e"]["value"], 2), "phases", {k: round(v, 2) for k, v in d["phases_ms"].items()})
print("n_op", d["ranks"][0]["n_op"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "op us (batch)", d["op_roofline"]["us_per_op"], "frac", d["op_roofline"]["frac"], "sum of kernels", d["op_roofline"]["sum_of_kernel_events"]["us_per_op"], "step frac", d["step_roofline"]["frac"])
for k, v in d["kernels"].items():
    print(f"  {k:12s} {v['launches']:5d} x {v['us_avg']:8.2f} us = {v['ms_total']:7.2f} ms")
PY
