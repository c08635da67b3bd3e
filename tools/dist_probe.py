#!/usr/bin/env python
"""Per-operator timing of the partitioned V-cycle operators (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29513 tools/dist_probe.py [n] [min_rows]

Builds the reference hierarchy of 3-D Poisson n^3 once (oracle/_ref), uploads its level operators
to a distributed context and times every kind of pass in a loop (CUDA events on the launching
stream, every rank reports): halo passes on A_l, restriction (with the gather of row shares when
the coarse level is replicated), prolongation, the all-reduced dot.  Test infrastructure."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // max(world, 1)))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import amgcl_b200 as ab  # noqa: E402
import oracle  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    min_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    reps = 200
    if world > 1:
        dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx = ab.Context(local, stream=side.cuda_stream)
    if world > 1:
        box = [ab.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.dist_init(box[0], world, rank, min_rows)
    ptr, col, val, rhs = ab.poisson3d(n)
    R = oracle.RefSolver(ptr, col, val, "damped_jacobi", "cg")
    levels, coarse = R.hierarchy()
    R.close()
    out = []

    def timed(name, fn, rows, nnz):
        for _ in range(10):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(reps):
            fn()
        e1.record(side)
        torch.cuda.synchronize()
        out.append({"op": name, "rows": rows, "nnz": nnz, "us": round(1e3 * e0.elapsed_time(e1) / reps, 2)})

    rng = np.random.default_rng(0)
    for l, lv in enumerate(levels[:3]):
        nf = lv["A"][0].size - 1
        nc = lv["R"][0].size - 1
        A = ctx.csr(nf, nf, *lv["A"])
        P = ctx.csr(nf, nc, *lv["P"])
        Rm = ctx.csr(nc, nf, *lv["R"])
        x, f, t, u = (ctx.vector(rng.uniform(-1, 1, nf)) for _ in range(4))
        d = ctx.vector(lv["diag"])
        xc, yc = ctx.vector(rng.uniform(-1, 1, nc)), ctx.vector(nc)
        nnzA, nnzP = int(lv["A"][0][-1]), int(lv["P"][0][-1])
        timed("L%d residual" % l, lambda: ctx.residual(f, A, x, t), nf, nnzA)
        timed("L%d relax" % l, lambda: ctx.relax(A, f, x, t, d, 0.72), nf, nnzA)
        timed("L%d spmv" % l, lambda: ctx.spmv(1.0, A, x, 0.0, t), nf, nnzA)
        timed("L%d restrict" % l, lambda: ctx.spmv(1.0, Rm, t, 0.0, yc), nc, nnzP)
        timed("L%d prolong" % l, lambda: ctx.spmv(1.0, P, xc, 1.0, u), nf, nnzP)
        timed("L%d dot" % l, lambda: ctx.dot(x, f), nf, 0)
        timed("L%d axpby" % l, lambda: ctx.axpby(0.5, x, 1.0, u), nf, 0)
        del A, P, Rm, x, f, t, u, d, xc, yc
    res = {"rank": rank, "world": world, "n": n, "min_rows": min_rows, "ops": out}
    if world > 1:
        allres = [None] * world
        dist.all_gather_object(allres, res)
    else:
        allres = [res]
    if rank == 0:
        names = [o["op"] for o in allres[0]["ops"]]
        print("%-14s %10s %10s  %s" % ("op", "rows", "nnz", "  ".join("rank%d us" % r for r in range(world))))
        for i, nm in enumerate(names):
            o = allres[0]["ops"][i]
            print("%-14s %10d %10d  %s" % (nm, o["rows"], o["nnz"],
                                           "  ".join("%8.2f" % allres[r]["ops"][i]["us"] for r in range(world))))
        print(json.dumps(allres))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
