"""Scratch: sequential vs three-in-flight config-5 units, repeated; prints which entries differ and by how much."""
import sys
import numpy as np
sys.path.insert(0, ".")  # run from the repository root
from legolas_b200 import sweep, workloads as wl
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
units = wl.sweep_units(12)
seq = wl.SweepSolver(sm_limit=148 // 3)
ref = np.stack([seq(u) for u in units])
ref2 = np.stack([seq(u) for u in units])
print("sequential twice identical:", np.array_equal(np.nan_to_num(ref), np.nan_to_num(ref2)))
seq.close()
bad = 0
for rep in range(reps):
    solvers = [wl.SweepSolver(sm_limit=148 // 3) for _ in range(3)]
    table, mine = sweep.run_queue(units, None, wl.SWEEP_NEV, solvers=solvers)
    for sv in solvers:
        sv.close()
    same = np.array_equal(np.nan_to_num(ref), np.nan_to_num(table)) and np.array_equal(np.isnan(ref), np.isnan(table))
    if not same:
        bad += 1
        d = np.abs(np.nan_to_num(ref) - np.nan_to_num(table))
        idx = np.argwhere(d > 0)
        print(f"rep {rep}: DIFFERENT rows {sorted(set(int(i[0]) for i in idx))} max abs diff {d.max():.3e} rel {d.max() / np.abs(np.nan_to_num(ref)).max():.3e} nan pattern same {np.array_equal(np.isnan(ref), np.isnan(table))}")
    else:
        print(f"rep {rep}: identical")
print("different:", bad, "of", reps)
