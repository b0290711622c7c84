"""Debug build only (LGPU_NVCC_EXTRA=-DLGPU_SLU_CLOCKS): a few solves at G=10001, cycle stamps printed by the kernels."""
import sys
import numpy as np
sys.path.insert(0, ".")
import legolas_b200 as lb
from legolas_b200 import equilibria as heq
G = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
s, grid, fields = heq.magnetothermal_instabilities(G)
ctx = lb.Context(); ctx.assemble(s, grid.base_grid, grid.gaussian_grid, fields)
ctx.factorize(0.02 + 0.03j)
b = np.ones(ctx.dim, dtype=np.complex128)
for _ in range(4):
    x = ctx.solve(b)
ctx.synchronize()
