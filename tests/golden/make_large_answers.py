#!/usr/bin/env python
"""Known answers of the REAL reference for the large BASELINE.json configurations.

Runs AMGCL's builtin backend (oracle/_ref/libamgcl_ref.so, the reference compiled in place by
oracle/Makefile) on 3-D Poisson 128^3 / 256^3 (smoothed_aggregation + damped_jacobi + CG and
+ spai0 + BiCGStab, all defaults, rhs = 1, x0 = 0) and -- with --big, on a host with >= 100 GB
of RAM -- 512^3 spai0 + BiCGStab (BASELINE.json config #3), and records iteration count,
final relative residual and samples of the solution in tests/golden/large_answers.json:

    python tests/golden/make_large_answers.py [--big] [--only N]

Existing entries for sizes not run are kept, so the 512^3 entry can be produced on the GPU
box's host (the prebuilt oracle/_ref travels there) and merged into the committed file.
"""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from amgcl_b200 import poisson3d  # noqa: E402

PATH = os.path.join(HERE, "large_answers.json")
NSAMPLES = 257


def sample_index(nrows):
    return np.linspace(0, nrows - 1, NSAMPLES).astype(np.int64)


def run(n, relax, krylov):
    ptr, col, val, rhs = poisson3d(n)
    S = oracle.RefSolver(ptr, col, val, relax, krylov)
    x, iters, resid = S.solve(rhs)
    rep = S.report()
    S.close()
    idx = sample_index(x.size)
    return {"n": n, "relax": relax, "krylov": krylov, "iters": int(iters), "resid": float(resid),
            "x_first": float(x[0]), "x_mid": float(x[x.size // 2]), "x_norm2": float(np.linalg.norm(x)),
            "x_max": float(np.abs(x).max()), "x_samples": [float(v) for v in x[idx]],
            "levels": [int(t.split()[1]) for t in rep.splitlines()
                       if t.strip() and t.split()[0].isdigit() and len(t.split()) >= 3]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="also run 512^3 spai0+bicgstab (>= 100 GB RAM)")
    ap.add_argument("--only", type=int, default=0, help="run only this grid size")
    ap.add_argument("--out", default=PATH)
    args = ap.parse_args()
    assert oracle.have_ref(), "needs oracle/_ref/libamgcl_ref.so (built where /root/reference exists)"
    plan = [(128, "damped_jacobi", "cg"), (128, "spai0", "bicgstab"),
            (256, "damped_jacobi", "cg"), (256, "spai0", "bicgstab")]
    if args.big:
        plan.append((512, "spai0", "bicgstab"))
    if args.only:
        plan = [p for p in plan if p[0] == args.only]
    known = {"source": "AMGCL builtin<double> (oracle/_ref/libamgcl_ref.so), defaults, rhs=1, x0=0; "
                       "x_samples at np.linspace(0, rows-1, %d).astype(int64)" % NSAMPLES,
             "cases": []}
    if os.path.isfile(args.out):
        with open(args.out) as f:
            known = json.load(f)
    for n, relax, krylov in plan:
        case = run(n, relax, krylov)
        known["cases"] = [c for c in known["cases"]
                          if (c["n"], c["relax"], c["krylov"]) != (n, relax, krylov)] + [case]
        print({k: v for k, v in case.items() if k != "x_samples"}, flush=True)
        with open(args.out, "w") as f:
            json.dump(known, f, indent=1)


if __name__ == "__main__":
    main()
