#!/usr/bin/env python
"""A/B of the operator storage formats (offset-indexed / windowed columns) on the real solve (one GPU): for each option set, build
the drop-in solver, check the solution bits against the first set, time the solve and list the
device time of the big operators.

    python tools/window_ab.py [n] [solves]        # sets: see CONFIGS
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amgcl_b200 as ab  # noqa: E402

CONFIGS = [
    ("plain", {"window": 0, "offsets": 0, "patterns": 0}),
    ("offsets", {"window": 0, "offsets": 1, "patterns": 0}),
    ("patterns", {"window": 0, "offsets": 0, "patterns": 1}),
    ("patterns, nnz_cap 2560", {"window": 0, "offsets": 0, "patterns": 1, "nnz_cap": 2560}),
    ("patterns, 3 stages", {"window": 0, "offsets": 0, "patterns": 1, "stages": 3}),
    ("offsets, 3 CTAs x 3 stages", {"window": 0, "offsets": 1, "ctas_per_sm": 3, "stages": 3}),
    ("offsets, nnz_cap 3584 x 3 CTAs", {"window": 0, "offsets": 1, "nnz_cap": 3584, "ctas_per_sm": 3}),
    ("offsets, 5 CTAs", {"window": 0, "offsets": 1, "ctas_per_sm": 5}),
    ("window all", {"offsets": 0, "patterns": 0, "window": 1, "window_ratio": 75, "window_lanes": 15}),
    ("window P only", {"offsets": 0, "patterns": 0, "window": 1, "window_ratio": 50, "window_lanes": 15}),
    ("window +A1", {"offsets": 0, "patterns": 0, "window": 1, "window_ratio": 125, "window_lanes": 15, "window_gap": 1}),
]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    solves = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    which = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else range(len(CONFIGS))
    import torch
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx = ab.Context(0, stream=side.cuda_stream)

    def time_ms(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(side)
        fn()
        e1.record(side)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    ptr, col, val, rhs = ab.poisson3d(n)
    ref_hash = None
    for k in which:
        name, opts = CONFIGS[k]
        defaults = {"window_ratio": 75, "window_lanes": 15, "window_gap": 2, "ctas_per_sm": 4, "stages": 2,
                    "nnz_cap": 2048}
        for key, v in defaults.items():
            ctx.set_option(key, v)
        for key, v in opts.items():
            ctx.set_option(key, v)
        t0 = time.time()
        S = ab.DropinSolver(ptr, col, val, "damped_jacobi", "cg", ctx=ctx)
        setup = time.time() - t0
        x, it, res = S.solve(rhs)
        h = hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()[:16]
        if ref_hash is None:
            ref_hash = h
        S.upload_rhs(rhs)
        for _ in range(2):
            S.solve_resident()
        ts = []
        for _ in range(solves):
            t = time_ms(lambda: S.solve_resident())
            ts.append(t)
        ctx.profile_begin()
        S.solve_resident()
        prof = ctx.profile_end()
        big = sorted([p for p in prof if p["nnz"] >= 5000000], key=lambda p: -p["total_ms"])
        rec = {"config": name, "opts": opts, "setup_s": round(setup, 2), "iters": it, "resid": res,
               "x_sha": h, "same_bits_as_first": h == ref_hash, "solve_ms_median": round(float(np.median(ts)), 3),
               "solve_ms_min": round(float(np.min(ts)), 3),
               "ops": [{"rows": p["nrows"], "nnz": p["nnz"], "mode": p["mode"], "launches": p["launches"],
                        "avg_us": round(1e3 * p["total_ms"] / p["launches"], 1)} for p in big]}
        print(json.dumps(rec), flush=True)
        S.close()


if __name__ == "__main__":
    main()
