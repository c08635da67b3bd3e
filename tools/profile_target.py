#!/usr/bin/env python
"""Short profiling target for ncu: set up the 256^3 solver, run N solves, nothing else.

    ncu ... python tools/profile_target.py [n] [solves]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amgcl_b200 as ab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
solves = int(sys.argv[2]) if len(sys.argv) > 2 else 1
relax = sys.argv[3] if len(sys.argv) > 3 else "damped_jacobi"
krylov = sys.argv[4] if len(sys.argv) > 4 else "cg"
ctx = ab.Context(0)
ptr, col, val, rhs = ab.poisson3d(n)
S = ab.DropinSolver(ptr, col, val, relax, krylov, ctx=ctx)
S.upload_rhs(rhs)
for _ in range(solves):
    it, res = S.solve_resident()
print("iters", it, "resid", res, "launches", ctx.launches)
