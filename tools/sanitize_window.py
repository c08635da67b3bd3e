import sys, numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import amgcl_b200 as ab
from test_gpu_window import banded
ctx = ab.Context(0)
ctx.set_option("window_min_nnz", 0)
for per_row in (6, 30, 100):
    ptr, col, val = banded(1500, 1500, per_row, 1)
    A = ctx.csr(1500, 1500, ptr, col, val)
    print(per_row, A.plan(), A.window())
    x, f, d = ctx.vector(np.ones(1500)), ctx.vector(np.ones(1500)), ctx.vector(np.full(1500, 0.5))
    y, t = ctx.vector(1500), ctx.vector(1500)
    ctx.spmv(1.0, A, x, 0.0, y); ctx.spmv(1.0, A, x, 1.0, y); ctx.residual(f, A, x, y)
    ctx.relax(A, f, x, t, d, 0.7)
    ctx.clear(x); ctx.relax(A, f, x, t, d, 0.7); ctx.residual(f, A, x, y)
    print(float(np.abs(y.numpy()).sum()))
ptr, col, val, rhs = ab.poisson3d(14)
for prec in ("f64", "mixed"):
    S = ab.DropinSolver(ptr, col, val, "damped_jacobi", "cg", coarse_enough=200, ctx=ctx, precision=prec)
    print(prec, S.solve(rhs)[1:])
    S.close()
print("SAN_WIN_DONE")
