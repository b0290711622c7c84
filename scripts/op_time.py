"""Scratch: time the operator application y = (A - sigma B)^-1 B x per kernel class (CUDA events)."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
s, grid, fields = heq.magnetothermal_instabilities(G)
s.solvers = lb.SolverSettings(solver="arnoldi", arpack_mode="shift-invert", number_of_eigenvalues=20, sigma=0.02 + 0.03j, maxiter=3)
ctx = lb.Context()
mats = lb.build_matrices(s, grid.base_grid, grid.gaussian_grid, fields, ctx=ctx)
cfg = lb.new_arpack_config(ctx.dim, 2, "I", s.solvers); cfg.maxiter = 3
ctx.shift_invert(cfg, 0.02 + 0.03j, want_vectors=False)
ctx.set_profiling(True)
ctx.shift_invert(cfg, 0.02 + 0.03j, want_vectors=False)
p = ctx.profile()
ops = ("matvec", "fwd_stage0", "fwd_stage", "top_stage", "bwd_stage", "bwd_stage0")
tot = sum(p[k][0] for k in ops) / max(p["fwd_stage0"][1], 1) * 1e3
print("env", {k: v for k, v in os.environ.items() if k.startswith("LGPU_")}, "us/op %.1f" % tot,
      " ".join(f"{k}={1e3*p[k][0]/max(p[k][1],1):.1f}" for k in ops + ("cgs2_step", "dots", "update", "scale")))
