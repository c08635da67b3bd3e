#!/usr/bin/env python
"""Developer tool (GPU box): quick parity pass + kernel micro-benchmarks.

    python tools/gpu_check.py parity            # primitives + drop-in vs oracle / reference
    python tools/gpu_check.py spmv 256          # finest-level kernel sweep at n^3

Writes JSON lines to stdout; not part of the test-suite or of bench.py.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import amgcl_b200 as ab  # noqa: E402


def algorithmic_bytes(nrows, ncols, nnz, mode):
    """SURVEY.md section 8(d) figures (FP64 values, int32 indices)."""
    b = nnz * 12 + (nrows + 1) * 4 + ncols * 8 + nrows * 8
    if mode == "residual":
        b += nrows * 8
    elif mode == "spmv_acc":
        b += nrows * 8
    elif mode == "relax":
        b += nrows * 8 * 2       # rhs + diag (x_new write is the 'y' already counted)
    return b


def time_op(fn, reps=20, warm=3):
    st = torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record(st)
    for i in range(reps):
        fn()
        ev[i + 1].record(st)
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    return float(np.median(ts)), float(np.min(ts))


def cmd_parity():
    import oracle
    ctx = ab.Context(0)
    rng = np.random.default_rng(42)
    out = {}
    ptr, col, val, rhs = ab.poisson3d(24)
    n = ptr.size - 1
    A = ctx.csr(n, n, ptr, col, val)
    x = rng.uniform(-1, 1, n)
    y = rng.uniform(-1, 1, n)
    z = rng.uniform(-1, 1, n)
    dx, dy, dz = ctx.vector(x), ctx.vector(y), ctx.vector(z)
    o = oracle.c()
    ctx.spmv(2.0, A, dx, 0.0, dy)
    out["spmv"] = float(np.abs(dy.numpy() - o.spmv(2.0, (ptr, col, val), x, 0.0, y)).max())
    dy.upload(y)
    ctx.spmv(2.0, A, dx, -0.5, dy)
    out["spmv_acc"] = float(np.abs(dy.numpy() - o.spmv(2.0, (ptr, col, val), x, -0.5, y)).max())
    ctx.residual(dz, A, dx, dy)
    out["residual"] = float(np.abs(dy.numpy() - o.residual(z, (ptr, col, val), x)).max())
    out["dot"] = abs(ctx.dot(dx, dz) - o.inner_product(x, z))
    dy.upload(y)
    ctx.axpby(0.3, dx, 1.7, dy)
    out["axpby"] = float(np.abs(dy.numpy() - o.axpby(0.3, x, 1.7, y)).max())
    dz.upload(z)
    dy.upload(y)
    ctx.axpbypcz(0.3, dx, 1.7, dy, -2.0, dz)
    out["axpbypcz"] = float(np.abs(dz.numpy() - o.axpbypcz(0.3, x, 1.7, y, -2.0, z)).max())
    dz.upload(z)
    ctx.vmul(0.72, dx, dy, 1.0, dz)
    out["vmul"] = float(np.abs(dz.numpy() - o.vmul(0.72, x, y, 1.0, z)).max())
    # fused relax vs oracle
    d = o.relax_diag((ptr, col, val), "damped_jacobi")
    dd = ctx.vector(d)
    drhs = ctx.vector(rhs)
    dx.upload(x)
    tmp = ctx.vector(n)
    ctx.relax(A, drhs, dx, tmp, dd, 0.72)
    out["relax"] = float(np.abs(dx.numpy() - o.relax((ptr, col, val), rhs, x, d, 0.72)).max())
    ctx.clear(dx)
    ctx.relax(A, drhs, dx, tmp, dd, 0.72)
    out["relax_zero"] = float(np.abs(dx.numpy() - o.relax((ptr, col, val), rhs, np.zeros(n), d, 0.72)).max())
    for variant in (1,):
        ctx.set_option("spmv_variant", variant)
        dy.upload(y)
        dx.upload(x)
        ctx.spmv(2.0, A, dx, 0.0, dy)
        out["spmv_v%d" % variant] = float(np.abs(dy.numpy() - o.spmv(2.0, (ptr, col, val), x, 0.0, y)).max())
        ctx.set_option("spmv_variant", 0)
    print(json.dumps({"primitives_max_abs_err": out}))

    # ragged random matrices, all lane widths
    for lanes in (1, 2, 4, 8, 16, 32):
        ctx.set_option("lanes", lanes)
        nr, nc = 3001, 2500
        lens = rng.integers(0, 40, nr)
        lens[7] = 0
        lens[100] = 5000          # a row longer than a stage -> strided path
        lens[nr - 1] = 3
        p = np.zeros(nr + 1, dtype=np.int64)
        np.cumsum(lens, out=p[1:])
        c = rng.integers(0, nc, p[-1])
        v = rng.uniform(-1, 1, p[-1])
        M = ctx.csr(nr, nc, p, c, v)
        xx = rng.uniform(-1, 1, nc)
        errs = {}
        for variant in (0, 1):
            ctx.set_option("spmv_variant", variant)
            vx, vy = ctx.vector(xx), ctx.vector(nr)
            ctx.spmv(1.0, M, vx, 0.0, vy)
            ref = o.spmv(1.0, (p, c, v), xx, 0.0, np.zeros(nr))
            errs["v%d" % variant] = float(np.abs(vy.numpy() - ref).max() / np.abs(ref).max())
        ctx.set_option("spmv_variant", 0)
        print(json.dumps({"ragged_lanes": lanes, "plan": M.plan(), "rel_err": errs}))
    ctx.set_option("lanes", 0)

    # coarse solver
    S = oracle.RefSolver(*ab.poisson3d(16)[:3], coarse_enough=3000) if oracle.have_ref() else None
    pc, cc, vc, rc = ab.poisson3d(10)
    Cs = ctx.coarse(1000, pc, cc, vc)
    b = rng.uniform(-1, 1, 1000)
    vb, vx = ctx.vector(b), ctx.vector(1000)
    ctx.coarse_solve(Cs, vb, vx)
    xs = vx.numpy()
    r = b - o.spmv(1.0, (pc, cc, vc), xs, 0.0, np.zeros(1000))
    print(json.dumps({"coarse_rel_resid": float(np.linalg.norm(r) / np.linalg.norm(b))}))
    del S

    # drop-in vs reference
    for nn in (16, 32, 64):
        ptr, col, val, rhs = ab.poisson3d(nn)
        for relax, kry in (("damped_jacobi", "cg"), ("spai0", "bicgstab")):
            t0 = time.time()
            D = ab.DropinSolver(ptr, col, val, relax, kry, ctx=ctx)
            t1 = time.time()
            xg, itg, resg = D.solve(rhs)
            t2 = time.time()
            rec = {"n": nn, "relax": relax, "krylov": kry, "gpu_iters": itg, "gpu_resid": resg,
                   "setup_s": t1 - t0, "solve_s": t2 - t1}
            if oracle.have_ref():
                R = oracle.RefSolver(ptr, col, val, relax, kry)
                xr, itr, resr = R.solve(rhs)
                rec.update({"ref_iters": itr, "ref_resid": resr,
                            "x_rel_err": float(np.abs(xg - xr).max() / np.abs(xr).max())})
            print(json.dumps(rec))
            D.close()


def cmd_spmv(n):
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx = ab.Context(0, stream=side.cuda_stream)
    ptr, col, val, rhs = ab.poisson3d(n)
    nr = ptr.size - 1
    nnz = int(ptr[-1])
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, nr)
    d = np.full(nr, 1.0 / 6.0)
    configs = []
    for variant, nnz_cap, cps, stages in (
            (0, 2048, 0, 0), (1, 2048, 4, 2), (1, 2048, 5, 2), (1, 1536, 5, 2), (1, 2048, 3, 3),
            (1, 1536, 6, 2), (1, 2048, 4, 3)):
        configs.append((variant, nnz_cap, cps, stages))
    for variant, nnz_cap, cps, stages in configs:
        ctx.set_option("nnz_cap", nnz_cap)
        ctx.set_option("spmv_variant", variant)
        if variant == 1:
            ctx.set_option("ctas_per_sm", cps)
            ctx.set_option("stages", stages)
        A = ctx.csr(nr, nr, ptr, col, val)
        vx, vy, vf, vd, vt = ctx.vector(x), ctx.vector(nr), ctx.vector(rhs), ctx.vector(d), ctx.vector(nr)
        rec = {"n": n, "variant": variant, "nnz_cap": nnz_cap, "ctas_per_sm": cps, "stages": stages,
               "plan": A.plan()}
        for mode, fn in (("spmv", lambda: ctx.spmv(1.0, A, vx, 0.0, vy)),
                         ("residual", lambda: ctx.residual(vf, A, vx, vy)),
                         ("relax", lambda: ctx.relax(A, vf, vx, vt, vd, 0.72))):
            med, best = time_op(fn)
            gb = algorithmic_bytes(nr, nr, nnz, mode) / 1e9
            rec[mode] = {"ms": round(med, 4), "GBs": round(gb / (med * 1e-3), 1),
                         "best_GBs": round(gb / (best * 1e-3), 1)}
        print(json.dumps(rec), flush=True)
        del A, vx, vy, vf, vd, vt
    # vector kernels
    ctx.set_option("spmv_variant", 0)
    vx, vy, vz = ctx.vector(x), ctx.vector(x), ctx.vector(x)
    rec = {"n": n}
    for name, fn, nb in (("axpby", lambda: ctx.axpby(0.5, vx, 1.5, vy), 3),
                         ("axpbypcz", lambda: ctx.axpbypcz(0.5, vx, 1.5, vy, 0.25, vz), 4),
                         ("vmul", lambda: ctx.vmul(0.5, vx, vy, 1.0, vz), 4),
                         ("copy", lambda: ctx.copy(vx, vy), 2),
                         ("dot", lambda: ctx.dot(vx, vy), 2)):
        med, best = time_op(fn)
        gb = nb * nr * 8 / 1e9
        rec[name] = {"ms": round(med, 4), "GBs": round(gb / (med * 1e-3), 1)}
    print(json.dumps(rec), flush=True)


def cmd_levels(n):
    """Sweep lanes-per-row / stage size / ring shape on the REAL hierarchy operators
    (A_l, P_l, R_l of every level, taken from the reference-built hierarchy)."""
    import oracle
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx = ab.Context(0, stream=side.cuda_stream)
    ptr, col, val, rhs = ab.poisson3d(n)
    S = oracle.RefSolver(ptr, col, val, "damped_jacobi", "cg")
    levels, coarse = S.hierarchy()
    rng = np.random.default_rng(0)
    ops = []
    for l, lv in enumerate(levels):
        for key in ("A", "P", "R"):
            if l == 0 and key == "A":
                continue
            p_, c_, v_ = lv[key]
            nr = p_.size - 1
            nc = {"A": nr, "P": lv["R"][0].size - 1, "R": lv["A"][0].size - 1}[key]
            ops.append(("L%d_%s" % (l, key), nr, nc, p_, c_, v_))
    for name, nr, nc, p_, c_, v_ in ops:
        nnz = int(p_[-1])
        if nnz < 200000:
            continue
        x = rng.uniform(-1, 1, nc)
        gb = algorithmic_bytes(nr, nc, nnz, "spmv") / 1e9
        best = None
        cfgs = ((2048, 4, 2), (2048, 5, 2), (4096, 2, 2), (1536, 5, 2))
        lane_set = (1, 2, 4, 8, 16, 32)
        if os.environ.get("B200_SWEEP"):          # e.g. "1024:4:2,1024:2:2;2,4"
            c, l = os.environ["B200_SWEEP"].split(";")
            cfgs = tuple(tuple(int(v) for v in t.split(":")) for t in c.split(","))
            lane_set = tuple(int(v) for v in l.split(","))
        for lanes in lane_set:
            for nnz_cap, cps, stages in cfgs:
                ctx.set_option("lanes", lanes)
                ctx.set_option("nnz_cap", nnz_cap)
                ctx.set_option("ctas_per_sm", cps)
                ctx.set_option("stages", stages)
                A = ctx.csr(nr, nc, p_, c_, v_)
                vx, vy = ctx.vector(x), ctx.vector(nr)
                med, mn = time_op(lambda: ctx.spmv(1.0, A, vx, 0.0, vy), reps=10)
                rec = {"op": name, "rows": nr, "cols": nc, "nnz": nnz, "avg": round(nnz / nr, 1),
                       "lanes": lanes, "nnz_cap": nnz_cap, "cps": cps, "stages": stages,
                       "ms": round(med, 4), "GBs": round(gb / (med * 1e-3), 1)}
                print(json.dumps(rec), flush=True)
                if best is None or rec["GBs"] > best["GBs"]:
                    best = rec
                del A, vx, vy
        print(json.dumps({"BEST": best}), flush=True)


def cmd_unstructured():
    """BASELINE.json config #4 stand-in (poisson3Db.mtx is not available offline): synthetic
    unstructured SPD matrices with poisson3Db's row statistics; AMG + Krylov parity against the
    reference and an SpMV GB/s sweep over sizes from L2-resident to HBM-resident."""
    import oracle
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx = ab.Context(0, stream=side.cuda_stream)
    rng = np.random.default_rng(3)
    cases = [(1, "morton"), (4, "morton"), (16, "morton"), (64, "morton"), (16, "random")]
    # a real MatrixMarket / AMGCL-binary file (e.g. tutorial/1.poisson3Db's poisson3Db.mtx and
    # poisson3Db_b.mtx) when one is supplied: B200_MATRIX=path [B200_RHS=path]
    if os.environ.get("B200_MATRIX"):
        cases = [(None, os.environ["B200_MATRIX"])] + cases
    for mult, order in cases:
        t0 = time.time()
        if mult is None:
            from amgcl_b200 import io as bio
            if order.endswith(".bin"):
                n, ptr, col, val = bio.read_crs_binary(order)
            else:
                n, _, ptr, col, val = bio.read_mm(order)
            rp = os.environ.get("B200_RHS")
            rhs = (bio.read_dense_binary(rp) if rp and rp.endswith(".bin") else bio.read_mm(rp))[:, 0] \
                if rp else np.ones(n)
            name = "file %s" % os.path.basename(order)
        else:
            n = 85623 * mult
            ptr, col, val, rhs = ab.unstructured3d(n, order=order)
            name = "unstructured3d(n=%d, k=24, order=%s) [synthetic stand-in for poisson3Db]" % (n, order)
        nnz = int(ptr[-1])
        rec = {"matrix": name,
               "rows": n, "nnz": nnz, "nnz_per_row": round(nnz / n, 2), "generate_s": round(time.time() - t0, 1)}
        A = ctx.csr(n, n, ptr, col, val)
        x = rng.uniform(-1, 1, n)
        vx, vy, vf = ctx.vector(x), ctx.vector(n), ctx.vector(rhs)
        rec["plan"] = A.plan()
        for mode, fn in (("spmv", lambda: ctx.spmv(1.0, A, vx, 0.0, vy)),
                         ("residual", lambda: ctx.residual(vf, A, vx, vy))):
            med, best = time_op(fn)
            gb = algorithmic_bytes(n, n, nnz, mode) / 1e9
            rec[mode] = {"ms": round(med, 4), "GBs": round(gb / (med * 1e-3), 1)}
        rec["fits_l2"] = bool(nnz * 12 < 100e6)
        if mult is None or mult <= 4:
            for relax, kry in (("spai0", "bicgstab"),):
                D = ab.DropinSolver(ptr, col, val, relax, kry, ctx=ctx)
                t1 = time.time()
                xg, itg, resg = D.solve(rhs)
                key = "%s_%s" % (relax, kry)
                rec[key] = {"gpu_iters": itg, "gpu_resid": resg, "gpu_solve_s": round(time.time() - t1, 4)}
                if oracle.have_ref():
                    R = oracle.RefSolver(ptr, col, val, relax, kry)
                    t1 = time.time()
                    xr, itr, resr = R.solve(rhs)
                    rec[key].update({"ref_iters": itr, "ref_resid": resr,
                                     "ref_solve_s": round(time.time() - t1, 4),
                                     "x_rel_err": float(np.abs(xg - xr).max() / np.abs(xr).max())})
                    R.close()
                D.close()
        print(json.dumps(rec), flush=True)
        del A, vx, vy, vf


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if cmd == "parity":
        cmd_parity()
    elif cmd == "spmv":
        cmd_spmv(int(sys.argv[2]) if len(sys.argv) > 2 else 128)
    elif cmd == "unstructured":
        cmd_unstructured()
    elif cmd == "levels":
        cmd_levels(int(sys.argv[2]) if len(sys.argv) > 2 else 128)
