"""CPU, world_size 2 and 3, gloo: the N>1 host logic of the multi-GPU path.

Each rank takes its share of the operators from the library's own partition code
(b200_partition / b200_dist_split_i64 -- the same functions b200_csr_create_* uses on a
distributed context) and plays the device algorithm with numpy, moving data with the
collectives the CUDA path's NCCL transport issues (all_gather of the packed halo values, all_gather
of the row shares of a replicated result, all_reduce for the dot product).  A rank always keeps
WHOLE ROWS, so every product must equal the single-process one computed by the oracle exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import amgcl_b200 as ab
import oracle


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def csr_mv(ptr, col, val, x):
    y = np.zeros(ptr.size - 1)
    for r in range(ptr.size - 1):
        sl = slice(ptr[r], ptr[r + 1])
        y[r] = np.dot(val[sl], x[col[sl]])
    return y


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "poisson12_damped_jacobi_cg.npz"))
        A = tuple(fx["L0_A_%s" % k] for k in ("ptr", "col", "val"))
        P = tuple(fx["L0_P_%s" % k] for k in ("ptr", "col", "val"))
        R = tuple(fx["L0_R_%s" % k] for k in ("ptr", "col", "val"))
        n, nc = A[0].size - 1, R[0].size - 1
        x, u = fx["in_a"], fx["in_u"]
        B, lo, hi = ab.partition(n, world, rank)
        n_loc = hi - lo

        # ---- A x: pack boundary values, all-gather S per rank, gather from [x_loc | halo]
        sp = ab.dist_split("square", world, rank, n, n, *A)
        S = sp["slots"]
        assert sp["n_loc"] == n_loc and sp["ncols"] == n_loc + world * S
        seg = torch.zeros(max(S, 1), dtype=torch.float64)
        seg[:sp["send_idx"].size] = torch.from_numpy(x[lo:hi][sp["send_idx"]])
        halo = [torch.zeros(max(S, 1), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(halo, seg)
        xext = np.concatenate([x[lo:hi]] + [h.numpy()[:S] for h in halo])
        y_loc = csr_mv(sp["ptr"], sp["col"], sp["val"], xext)
        want = oracle.c().spmv(1.0, A, x, 0.0, np.zeros(n))
        err_a = np.abs(y_loc - want[lo:hi]).max()

        def halo_product(M, nr, ncol, v):
            """rows and columns partitioned: own rows, [local block of v | all-gathered halo]"""
            sp = ab.dist_split("halo", world, rank, nr, ncol, *M)
            _, clo, chi = ab.partition(ncol, world, rank)
            Sm = sp["slots"]
            assert sp["n_loc"] == chi - clo and sp["ncols"] == (chi - clo) + world * Sm
            seg = torch.zeros(max(Sm, 1), dtype=torch.float64)
            seg[:sp["send_idx"].size] = torch.from_numpy(v[clo:chi][sp["send_idx"]])
            hal = [torch.zeros(max(Sm, 1), dtype=torch.float64) for _ in range(world)]
            dist.all_gather(hal, seg)
            vext = np.concatenate([v[clo:chi]] + [h.numpy()[:Sm] for h in hal])
            return csr_mv(sp["ptr"], sp["col"], sp["val"], vext)

        # ---- P u between two partitioned levels: own fine rows, halo of the coarse vector
        _, plo, phi = ab.partition(n, world, rank)
        want = oracle.c().spmv(1.0, P, u, 0.0, np.zeros(n))
        err_p = np.abs(halo_product(P, n, nc, u) - want[plo:phi]).max()
        # ---- P u from a replicated (small) level: own rows, columns untouched, no exchange
        spp = ab.dist_split("replicated", world, rank, n, nc, *P)
        assert spp["slots"] == 0 and spp["ncols"] == nc
        err_p = max(err_p, np.abs(csr_mv(spp["ptr"], spp["col"], spp["val"], u) - want[plo:phi]).max())

        # ---- R t between two partitioned levels: own coarse rows, halo of the fine vector
        Bc, rlo, rhi = ab.partition(nc, world, rank)
        want = oracle.c().spmv(1.0, R, x, 0.0, np.zeros(nc))
        mine = halo_product(R, nc, n, x)
        err_r = np.abs(mine - want[rlo:rhi]).max()
        # ---- R t onto a replicated level: the row shares are all-gathered to every rank
        share = torch.zeros(Bc, dtype=torch.float64)
        share[:mine.size] = torch.from_numpy(mine)
        shares = [torch.zeros(Bc, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(shares, share)
        full = np.concatenate([t.numpy() for t in shares])[:nc]
        err_r = max(err_r, np.abs(full - want).max())

        # ---- <x, x>: local + all_reduce
        d = torch.tensor([float(np.dot(x[lo:hi], x[lo:hi]))], dtype=torch.float64)
        dist.all_reduce(d)
        err_d = abs(d.item() - oracle.c().inner_product(x, x))
        q.put((rank, err_a, err_p, err_r, err_d, int(S)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_operators_match_single_process(world):
    port = free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(world))
    for rank, ea, ep, er, ed, S in got:
        assert ea < 1e-13 and ep < 1e-13 and er < 1e-13 and ed < 1e-11, (rank, ea, ep, er, ed)
        # 12x12 planes of boundary values: one per neighbour (interior ranks have two)
        assert S == (144 if world == 2 else 288)


def test_partition_arithmetic():
    for n, P in ((1728, 2), (1728, 8), (10, 4), (3, 2), (16777216, 8), (2029083, 8)):
        blocks = [ab.partition(n, P, r) for r in range(P)]
        B = blocks[0][0]
        assert B % 4 == 0 and B * P >= n
        assert blocks[0][1] == 0 and blocks[-1][2] == n
        for r in range(P - 1):
            assert blocks[r][2] == blocks[r + 1][1]
        assert all(hi - lo <= B for _, lo, hi in blocks)


def test_split_square_reassembles_to_the_global_matrix():
    ptr, col, val, rhs = ab.poisson3d(7)
    n = ptr.size - 1
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, n)
    want = oracle.c().spmv(1.0, (ptr, col, val), x, 0.0, np.zeros(n))
    for P in (1, 2, 3, 5):
        parts = [ab.dist_split("square", P, r, n, n, ptr, col, val) for r in range(P)]
        S = parts[0]["slots"]
        assert all(p["slots"] == S for p in parts)
        bounds = [ab.partition(n, P, r) for r in range(P)]
        halo = np.zeros(P * S)
        for r, p in enumerate(parts):
            lo = bounds[r][1]
            halo[r * S:r * S + p["send_idx"].size] = x[lo + p["send_idx"]]
        got = np.zeros(n)
        for r, p in enumerate(parts):
            _, lo, hi = bounds[r]
            got[lo:hi] = csr_mv(p["ptr"], p["col"], p["val"], np.concatenate([x[lo:hi], halo]))
        assert np.abs(got - want).max() < 1e-13
