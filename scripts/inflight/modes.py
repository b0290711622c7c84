"""Scratch: where does the rare in-flight / sequential difference come from?
mode fresh   : ONE context at a time, a new context for every repetition (allocator history varies, no concurrency)
mode persist : three contexts in flight, the SAME three contexts for every repetition
mode inflight: three contexts in flight, new contexts every repetition (scripts/inflight/check.py)
"""
import sys
import numpy as np
sys.path.insert(0, ".")  # run from the repository root
from legolas_b200 import sweep, workloads as wl
mode = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
units = wl.sweep_units(12)
seq = wl.SweepSolver(sm_limit=148 // 3)
ref = np.stack([seq(u) for u in units])
seq.close()
print("reference has NaN rows:", int(np.isnan(ref.real).sum()), flush=True)
bad = 0
persist = [wl.SweepSolver(sm_limit=148 // 3) for _ in range(3)] if mode == "persist" else None
for rep in range(reps):
    if mode == "fresh":
        s = wl.SweepSolver(sm_limit=148 // 3)
        table = np.stack([s(u) for u in units])
        s.close()
    else:
        solvers = persist or [wl.SweepSolver(sm_limit=148 // 3) for _ in range(3)]
        table, _ = sweep.run_queue(units, None, wl.SWEEP_NEV, solvers=solvers)
        if persist is None:
            for sv in solvers:
                sv.close()
    same = np.array_equal(np.nan_to_num(ref), np.nan_to_num(table)) and np.array_equal(np.isnan(ref), np.isnan(table))
    if not same:
        bad += 1
        d = np.abs(np.nan_to_num(ref) - np.nan_to_num(table))
        print(f"rep {rep}: DIFFERENT rows {sorted(set(int(i[0]) for i in np.argwhere(d > 0)))} max abs diff {d.max():.3e} nan same {np.array_equal(np.isnan(ref), np.isnan(table))}", flush=True)
print(mode, "different:", bad, "of", reps)
