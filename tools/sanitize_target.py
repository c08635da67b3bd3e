#!/usr/bin/env python
"""Small workload for compute-sanitizer: every kernel family once, tiny sizes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amgcl_b200 as ab  # noqa: E402

ctx = ab.Context(0)
rng = np.random.default_rng(0)
for variant in (1, 0):
    ctx.set_option("spmv_variant", variant)
    for lanes in (0, 2, 8, 32):
        ctx.set_option("lanes", lanes)
        nr, nc = 1500, 1300
        lens = rng.integers(0, 30, nr)
        lens[5] = 4000
        ptr = np.zeros(nr + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        col = rng.integers(0, nc, ptr[-1])
        val = rng.uniform(-1, 1, ptr[-1])
        A = ctx.csr(nr, nc, ptr, col, val)
        x, y, f = ctx.vector(rng.uniform(-1, 1, nc)), ctx.vector(nr), ctx.vector(rng.uniform(-1, 1, nr))
        ctx.spmv(1.0, A, x, 0.0, y)
        ctx.spmv(1.0, A, x, 0.5, y)
        ctx.residual(f, A, x, y)
ctx.set_option("lanes", 0)
ctx.set_option("spmv_variant", 1)
for relax, krylov, prec in (("damped_jacobi", "cg", "f64"), ("spai0", "bicgstab", "f64"),
                            ("damped_jacobi", "cg", "mixed")):
    ptr, col, val, rhs = ab.poisson3d(14)
    S = ab.DropinSolver(ptr, col, val, relax, krylov, coarse_enough=200, ctx=ctx, precision=prec)
    x, it, res = S.solve(rhs)
    print(relax, krylov, prec, it, res)
    S.close()
# recorded V-cycles (CUDA graph replays) and the ILU(0) sweeps built from the primitives
ptr, col, val, rhs = ab.poisson3d(14)
for relax, krylov, prec, graph in (("spai0", "bicgstab", "f64", True), ("damped_jacobi", "cg", "mixed", True),
                                   ("ilu0", "bicgstab", "f64", False)):
    S = ab.DropinSolver(ptr, col, val, relax, krylov, coarse_enough=200, ctx=ctx, precision=prec, graph=graph)
    for _ in range(2):
        x, it, res = S.solve(rhs)
    print(relax, krylov, prec, "graph" if graph else "", it, res, S.graph_stats())
    S.close()
# round 2: the coarse tail as one cooperative kernel, the reference's Krylov sequence on the same
# backend (fused_krylov = 0), and the stand-alone Krylov steps
for opt, val_ in (("coarse_tail", 1), ("fused_krylov", 0), ("fuse_first_sweep", 0)):
    ctx.set_option(opt, val_)
    S = ab.DropinSolver(ptr, col, val, "damped_jacobi", "cg", coarse_enough=200, ctx=ctx)
    x, it, res = S.solve(rhs)
    print(opt, val_, it, res, ctx.tail_stats())
    S.close()
    ctx.set_option(opt, 1 - val_)
n = ptr.size - 1
A = ctx.csr(n, n, ptr, col, val)
K = ab.Krylov(ctx, n)
vs = [ctx.vector(rng.uniform(-1, 1, n)) for _ in range(8)]
K.residual(vs[0], A, vs[1], vs[2])
K.cg_direction(vs[2], vs[3], vs[4])
K.cg_step(A, vs[4], vs[5], vs[1], vs[2])
K.bicg_start(vs[2], vs[6])
K.bicg_direction(vs[2], vs[5], vs[4])
K.bicg_step_s(A, vs[6], vs[3], vs[5], vs[2], vs[7], vs[1])
K.bicg_step_r(A, vs[6], vs[3], vs[0], vs[7], vs[2], vs[1])
print("krylov steps", K.scalars())
K.close()
print("SANITIZE_TARGET_DONE")
