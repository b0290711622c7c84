"""Scratch: BCR solve accuracy vs LAPACK at a given G (GPU)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
from oracle import assembly as asm, equilibria as oeq, solvers as osolvers
G = int(sys.argv[1]); sigma = 0.02 + 0.03j
s, grid, fields = heq.magnetothermal_instabilities(G)
so, go, xgo, fo = oeq.magnetothermal_eq(gridpts=G)
A, B = asm.build_matrices(so, go, xgo, fo)
ctx = lb.Context(); ctx.assemble(s, grid.base_grid, grid.gaussian_grid, fields)
print("lu_info", ctx.factorize(sigma))
rng = np.random.default_rng(0)
# smooth rhs (like an eigenvector) and random rhs
xs = np.tile(np.sin(np.linspace(0, 3, G))[:, None], (1, 16)).reshape(-1) * (1 + 0.5j)
for nm, x in (("smooth", xs), ("random", rng.standard_normal(A.n) + 1j * rng.standard_normal(A.n))):
    b = B.matvec(x)
    lu = osolvers.BandedLU(A.to_band() - sigma * B.to_band(), 31, 31)
    xl = lu.solve(b); xg = ctx.solve(b); xg1 = ctx.solve(b, refine_steps=1)
    Mm = asm.BlockTriMatrix(G, 16, "M"); Mm.blocks = A.blocks - sigma * B.blocks
    f = lambda v: np.linalg.norm(Mm.matvec(v) - b) / np.linalg.norm(b)
    print(nm, "res lapack %.2e gpu %.2e gpu+ir %.2e | gpu-lapack %.2e gpu+ir-lapack %.2e" % (f(xl), f(xg), f(xg1), np.linalg.norm(xg - xl) / np.linalg.norm(xl), np.linalg.norm(xg1 - xl) / np.linalg.norm(xl)))
