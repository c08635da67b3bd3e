#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py [n ...]

Every rank runs the same AMGCL program (drop-in) on a distributed context; rank 0 compares
iterations / residual / solution with the reference's known answers and, when shipped, with
the live reference."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 8) // max(world, 1))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import amgcl_b200 as ab  # noqa: E402


def make_ctx(min_rows, p2p=1):
    torch.cuda.set_device(local)
    ctx = ab.Context(local)
    if world > 1:
        box = [ab.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.set_option("p2p", p2p)
        ctx.dist_init(box[0], world, rank, min_rows)
    return ctx


def latency():
    """Per-call device time of the exchange primitives (tiny local work): halo exchange inside
    spmv on a partitioned 48^3 operator, and the all-reduced dot product."""
    ptr, col, val, rhs = ab.poisson3d(48)
    n = ptr.size - 1
    for p2p in (1, 0):
        torch.cuda.set_device(local)
        side = torch.cuda.Stream()
        torch.cuda.set_stream(side)
        ctx = make_ctx(n, p2p)
        ctx.set_stream(side.cuda_stream)
        A = ctx.csr(n, n, ptr, col, val)
        x, y = ctx.vector(rhs), ctx.vector(n)
        out = {"world": world, "p2p": ctx.dist_info()["p2p"]}
        for name, fn, reps in (("spmv_with_halo_us", lambda: ctx.spmv(1.0, A, x, 0.0, y), 300),
                               ("dot_us", lambda: ctx.dot(x, x), 300)):
            for _ in range(20):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            for _ in range(reps):
                fn()
            e1.record(side)
            torch.cuda.synchronize()
            out[name] = round(1e3 * e0.elapsed_time(e1) / reps, 2)
        if rank == 0:
            print(json.dumps(out), flush=True)
        del A, x, y
        ctx.close()
        dist.barrier()


def main():
    if world > 1:
        dist.init_process_group("gloo")     # plumbing only: id exchange + barriers
    if "--latency" in sys.argv:
        latency()
        dist.destroy_process_group()
        return
    sizes = [int(a) for a in sys.argv[1:]] or [32, 64]
    known = json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))["cases"]
    ok = True
    for n in sizes:
        ptr, col, val, rhs = ab.poisson3d(n)
        nrows = ptr.size - 1
        for mode, min_rows, p2p in (("finest", nrows, 1), ("all>=2000", 2000, 1), ("all>=2000", 2000, 0)):
            if world == 1 and p2p == 0:
                continue
            for relax, krylov in (("damped_jacobi", "cg"), ("spai0", "bicgstab")):
                ctx = make_ctx(min_rows, p2p)
                S = ab.DropinSolver(ptr, col, val, relax, krylov, ctx=ctx)
                x, it, res = S.solve(rhs)
                case = [c for c in known if (c["n"], c["relax"], c["krylov"]) == (n, relax, krylov)]
                rec = {"n": n, "world": world, "partitioned": mode, "relax": relax, "krylov": krylov,
                       "p2p": ctx.dist_info()["p2p"], "iters": it, "resid": res}
                if case:
                    c = case[0]
                    rec["ref_iters"] = c["iters"]
                    rec["resid_rel_diff"] = abs(res - c["resid"]) / c["resid"]
                    rec["x_norm_rel_diff"] = abs(np.linalg.norm(x) - c["x_norm2"]) / c["x_norm2"]
                    good = (it == c["iters"] and rec["resid_rel_diff"] < 1e-6 and
                            rec["x_norm_rel_diff"] < 1e-8)
                    rec["ok"] = bool(good)
                    ok = ok and good
                # true residual computed on the host
                import oracle
                r = rhs - oracle.c().spmv(1.0, (ptr, col, val), x, 0.0, np.zeros_like(x))
                rec["true_resid"] = float(np.linalg.norm(r) / np.linalg.norm(rhs))
                ok = ok and rec["true_resid"] < 2e-8
                if rank == 0:
                    print(json.dumps(rec), flush=True)
                S.close()
                ctx.close()
                if world > 1:
                    dist.barrier()
    # mixed precision (FP32 hierarchy under the FP64 Krylov solver) on the partitioned context:
    # against the reference's own mixed run (known_answers.json "mixed": iterations +-1, FP32-sized
    # solution tolerance, true FP64 residual)
    mixed = json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))["mixed"]
    for n in sizes:
        ptr, col, val, rhs = ab.poisson3d(n)
        for relax, krylov in (("damped_jacobi", "cg"), ("spai0", "bicgstab")):
            case = [c for c in mixed if (c["n"], c["relax"], c["krylov"]) == (n, relax, krylov)]
            if not case or world == 1:
                continue
            ctx = make_ctx(2000, 1)
            S = ab.DropinSolver(ptr, col, val, relax, krylov, ctx=ctx, precision="mixed")
            x, it, res = S.solve(rhs)
            import oracle
            r = rhs - oracle.c().spmv(1.0, (ptr, col, val), x, 0.0, np.zeros_like(x))
            c = case[0]
            rec = {"n": n, "world": world, "precision": "mixed", "relax": relax, "krylov": krylov,
                   "iters": it, "ref_iters": c["iters"], "resid": res,
                   "x_norm_rel_diff": abs(np.linalg.norm(x) - c["x_norm2"]) / c["x_norm2"],
                   "true_resid": float(np.linalg.norm(r) / np.linalg.norm(rhs))}
            rec["ok"] = bool(abs(it - c["iters"]) <= 1 and rec["x_norm_rel_diff"] < 1e-6 and rec["true_resid"] < 2e-8)
            ok = ok and rec["ok"]
            if rank == 0:
                print(json.dumps(rec), flush=True)
            S.close()
            ctx.close()
            dist.barrier()
    if rank == 0:
        print("DIST_CHECK", "PASS" if ok else "FAIL", flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
