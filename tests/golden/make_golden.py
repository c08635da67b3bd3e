#!/usr/bin/env python
"""Generate the committed golden fixtures from the REAL reference.

Runs AMGCL's builtin backend (oracle/_ref/libamgcl_ref.so, compiled in place
from /root/reference by oracle/Makefile) and stores its inputs and outputs as
small .npz files, so the C restatement (oracle/amg_oracle.c) and the CUDA path
can be pinned where the reference tree does not exist (the GPU box).

    python tests/golden/make_golden.py        # needs /root/reference

Fixtures (all FP64 / int64):
  poisson12_<relax>_<krylov>.npz   3-D Poisson 12^3, coarse_enough=100 -> 3 levels:
       every level operator (A, P, R), smoother diagonal, coarsest matrix, rhs,
       the reference's solution / iteration count / residual, one preconditioner
       application, the coarsest-level skyline-LU solve, and seeded (42) inputs
       with the builtin backend's outputs for each primitive.
  known_answers.json               iterations + residuals of the reference on
       16^3 .. 64^3 (run here) and the 128^3 / 256^3 figures of BASELINE.md.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from amgcl_b200 import poisson3d  # noqa: E402

CONFIGS = (("damped_jacobi", "cg"), ("spai0", "bicgstab"),
           ("spai0", "cg"), ("damped_jacobi", "bicgstab"))


def fixture(n, relax, krylov, coarse_enough):
    ptr, col, val, rhs = poisson3d(n)
    S = oracle.RefSolver(ptr, col, val, relax, krylov, coarse_enough=coarse_enough)
    x, iters, resid = S.solve(rhs)
    levels, coarse = S.hierarchy()
    out = {"n": n, "relax": relax, "krylov": krylov, "coarse_enough": coarse_enough,
           "nlevels": len(levels) + 1, "rhs": rhs, "x": x, "iters": iters, "resid": resid,
           "omega": 0.72 if relax == "damped_jacobi" else 1.0}
    for l, lv in enumerate(levels):
        for key in ("A", "P", "R"):
            for name, arr in zip(("ptr", "col", "val"), lv[key]):
                out["L%d_%s_%s" % (l, key, name)] = arr
        out["L%d_diag" % l] = lv["diag"]
    for name, arr in zip(("ptr", "col", "val"), coarse):
        out["C_%s" % name] = arr

    rng = np.random.default_rng(42)
    N = ptr.size - 1
    a, b, c = (rng.uniform(-1, 1, N) for _ in range(3))
    A0 = (ptr, col, val)
    r = oracle.ref()
    out.update({
        "in_a": a, "in_b": b, "in_c": c,
        "spmv_2_a_0": r.spmv(2.0, A0, N, a, 0.0, b),
        "spmv_2_a_m05_b": r.spmv(2.0, A0, N, a, -0.5, b),
        "residual_c_a": r.residual(c, A0, N, a),
        "dot_a_c": r.inner_product(a, c),
        "axpby_03_a_17_b": r.axpby(0.3, a, 1.7, b),
        "axpby_03_a_0_b": r.axpby(0.3, a, 0.0, b),
        "axpbypcz": r.axpbypcz(0.3, a, 1.7, b, -2.0, c),
        "axpbypcz_c0": r.axpbypcz(0.3, a, 1.7, b, 0.0, c),
        "vmul_072_a_b_1_c": r.vmul(0.72, a, b, 1.0, c),
        "vmul_1_a_b_0_c": r.vmul(1.0, a, b, 0.0, c),
        "precond_a": S.apply_precond(a),
    })
    # restriction / prolongation on level 0
    nc = levels[0]["R"][0].size - 1
    u = rng.uniform(-1, 1, nc)
    out["in_u"] = u
    out["restrict_a"] = r.spmv(1.0, levels[0]["R"], N, a, 0.0, np.zeros(nc))
    out["prolong_u_acc_b"] = r.spmv(1.0, levels[0]["P"], nc, u, 1.0, b)
    # coarsest-level direct solve (skyline LU)
    ncc = coarse[0].size - 1
    g = rng.uniform(-1, 1, ncc)
    out["in_g"] = g
    out["coarse_solve_g"] = S.coarse_solve(g)
    return out


def main():
    assert oracle.have_ref(), "needs the reference tree (/root/reference)"
    for relax, krylov in CONFIGS[:2]:
        fx = fixture(12, relax, krylov, 100)
        path = os.path.join(HERE, "poisson12_%s_%s.npz" % (relax, krylov))
        np.savez_compressed(path, **fx)
        print(path, "levels", fx["nlevels"], "iters", fx["iters"], "resid", fx["resid"])

    known = {"source": "AMGCL builtin<double> (oracle/_ref/libamgcl_ref.so), defaults, rhs=1, x0=0",
             "threads": oracle.ref().threads, "cases": []}
    for n in (16, 24, 32, 48, 64):
        ptr, col, val, rhs = poisson3d(n)
        for relax, krylov in CONFIGS:
            S = oracle.RefSolver(ptr, col, val, relax, krylov)
            x, iters, resid = S.solve(rhs)
            levels, coarse = S.hierarchy()
            known["cases"].append({
                "n": n, "relax": relax, "krylov": krylov, "iters": iters, "resid": resid,
                "x_first": float(x[0]), "x_mid": float(x[x.size // 2]),
                "x_norm2": float(np.linalg.norm(x)),
                "level_rows": [int(l["A"][0].size - 1) for l in levels] + [int(coarse[0].size - 1)],
                "level_nnz": [int(l["A"][0][-1]) for l in levels] + [int(coarse[0][-1])]})
            print(known["cases"][-1])
            S.close()
    # components that reach the backend only through its primitives
    known["primitive_only"] = []
    for n in (16, 32, 48):
        ptr, col, val, rhs = poisson3d(n)
        for relax, krylov in (("chebyshev", "cg"), ("damped_jacobi", "gmres"), ("spai0", "bicgstabl"),
                              ("ilu0", "bicgstab"), ("ilu0", "cg")):
            S = oracle.RefSolver(ptr, col, val, relax, krylov)
            x, iters, resid = S.solve(rhs)
            known["primitive_only"].append({
                "n": n, "relax": relax, "krylov": krylov, "iters": iters, "resid": resid,
                "x_first": float(x[0]), "x_mid": float(x[x.size // 2]),
                "x_norm2": float(np.linalg.norm(x))})
            print(known["primitive_only"][-1])
            S.close()
    # the reference's mixed-precision composition: amg<builtin<float>> under a builtin<double> solver
    known["mixed"] = []
    for n in (16, 32, 48, 64):
        ptr, col, val, rhs = poisson3d(n)
        for relax, krylov in CONFIGS:
            S = oracle.RefSolver(ptr, col, val, relax, krylov, precision="mixed")
            x, iters, resid = S.solve(rhs)
            known["mixed"].append({
                "n": n, "relax": relax, "krylov": krylov, "iters": iters, "resid": resid,
                "x_first": float(x[0]), "x_mid": float(x[x.size // 2]),
                "x_norm2": float(np.linalg.norm(x))})
            print("mixed", known["mixed"][-1])
            S.close()
    # BASELINE.md section 2 (measured by the survey with the same reference build)
    known["survey"] = [
        {"n": 128, "relax": "damped_jacobi", "krylov": "cg", "iters": 21, "resid": 6.07447143094944e-09},
        {"n": 128, "relax": "spai0", "krylov": "bicgstab", "iters": 10, "resid": 2.50964828306812e-09},
        {"n": 256, "relax": "damped_jacobi", "krylov": "cg", "iters": 27, "resid": 4.41192449561056e-09},
        {"n": 256, "relax": "spai0", "krylov": "bicgstab", "iters": 14, "resid": 9.25929720236484e-09},
    ]
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(known, f, indent=1)
    # files written by the reference's own writers (io/mm.hpp:349-420, io/binary.hpp:158-168)
    R = oracle.ref()
    ptr, col, val, rhs = poisson3d(5)
    R.mm_write_crs(os.path.join(HERE, "io_poisson5.mtx"), 125, ptr, col, val)
    R.bin_write_crs(os.path.join(HERE, "io_poisson5.bin"), ptr, col, val)
    R.mm_write_dense(os.path.join(HERE, "io_vec5.mtx"), np.sin(np.arange(125.0)))


if __name__ == "__main__":
    main()
