"""GPU parity of the windowed operator format (csr_kernels.cuh WIN, window.cuh): the same
operator run through the windowed kernel and through the plain one must give the same bits --
the window only changes where x is read from, not the arithmetic -- and both must agree with
the oracle."""
import numpy as np
import pytest

import amgcl_b200 as ab
import oracle
from conftest import rel_err
from test_gpu_primitives import _f32csr, _f32vec

pytestmark = pytest.mark.gpu


@pytest.fixture()
def wctx(ctx):
    """Every operator that qualifies is windowed, whatever its size."""
    ctx.set_option("window_min_nnz", 0)
    ctx.set_option("window", 1)
    ctx.set_option("offsets", 0)          # (an operator that also qualifies for a compressed
    ctx.set_option("patterns", 0)         #  column format is stored that way)
    yield ctx
    ctx.set_option("window_min_nnz", 1000000)
    ctx.set_option("window_ratio", 75)
    ctx.set_option("window", 0)
    ctx.set_option("offsets", 1)
    ctx.set_option("patterns", 1)
    ctx.set_option("lanes", 0)


def banded(nr, nc, per_row, seed, spread=1):
    """Rows of `per_row` entries around the diagonal position (blocks gather from one band)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(max(1, per_row - 3), per_row + 4, nr)
    lens[::97] = 0                                           # some empty rows
    ptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    col = np.empty(ptr[-1], dtype=np.int64)
    for i in range(nr):
        c0 = i * (nc - 1) // max(1, nr - 1)
        cand = np.arange(max(0, c0 - 2 * per_row * spread), min(nc, c0 + 2 * per_row * spread + 1), spread)
        col[ptr[i]:ptr[i + 1]] = np.sort(rng.choice(cand, lens[i], replace=False))
    val = rng.uniform(-1, 1, ptr[-1])
    return ptr, col, val


def both(ctx, fn):
    """Run fn() with the windowed kernels, then with the plain ones; return both results."""
    ctx.set_option("window", 1)
    a = fn()
    ctx.set_option("window", 0)
    b = fn()
    ctx.set_option("window", 1)
    return a, b


@pytest.mark.parametrize("per_row,lanes", [(6, 1), (30, 2), (50, 4), (100, 8)])
@pytest.mark.parametrize("shape", [(3001, 3001), (2500, 4001)])
def test_windowed_kernels_give_the_bits_of_the_plain_ones(wctx, per_row, lanes, shape):
    ctx = wctx
    o = oracle.c()
    nr, nc = shape
    ptr, col, val = banded(nr, nc, per_row, seed=per_row + nr)
    A = ctx.csr(nr, nc, ptr, col, val)
    assert A.plan()["lanes"] == lanes
    w = A.window()
    assert w["windowed"] and 0 < w["total_slots"] < 0.75 * col.size
    rng = np.random.default_rng(1)
    x, y, f = rng.uniform(-1, 1, nc), rng.uniform(-1, 1, nr), rng.uniform(-1, 1, nr)
    vx, vf = ctx.vector(x), ctx.vector(f)

    def spmv(beta):
        vy = ctx.vector(y)
        ctx.spmv(1.5, A, vx, beta, vy)
        return vy.numpy()
    for beta in (0.0, -0.25):
        a, b = both(ctx, lambda: spmv(beta))
        assert np.array_equal(a, b)
        assert rel_err(a, o.spmv(1.5, (ptr, col, val), x, beta, y)) < 1e-12

    def resid():
        vr = ctx.vector(nr)
        ctx.residual(vf, A, vx, vr)
        return vr.numpy()
    a, b = both(ctx, resid)
    assert np.array_equal(a, b)
    assert rel_err(a, o.residual(f, (ptr, col, val), x)) < 1e-12

    if nr == nc:
        d = rng.uniform(0.1, 1.0, nr)
        vd = ctx.vector(d)

        def sweep(zero):
            vxx, vt = ctx.vector(x), ctx.vector(nr)
            if zero:
                ctx.clear(vxx)
            ctx.relax(A, vf, vxx, vt, vd, 0.72)          # (from zero: postponed ...)
            vr = ctx.vector(nr)
            ctx.residual(vf, A, vxx, vr)                  # (... and fused into this residual)
            return np.concatenate([vxx.numpy(), vr.numpy()])
        for zero in (False, True):
            a, b = both(ctx, lambda: sweep(zero))
            assert np.array_equal(a, b)
        x1 = x + 0.72 * d * o.residual(f, (ptr, col, val), x)
        assert rel_err(sweep(False)[:nr], x1) < 1e-12

        # the streaming pass that also leaves scalars behind (CG: q = A p with <q, p>)
        def step():
            K = ab.Krylov(ctx, nr)                    # (fresh scalars: no history from the other run)
            vp, vq, vxx, vr = ctx.vector(x), ctx.vector(nr), ctx.vector(y), ctx.vector(f)
            K.cg_direction(vf, vf, vp)
            K.cg_step(A, vp, vq, vxx, vr)
            s = K.scalars()
            K.close()
            return np.concatenate([vq.numpy(), vxx.numpy(), vr.numpy(), [s["qp"], s["alpha"], s["rr"]]])
        a, b = both(ctx, step)
        assert np.array_equal(a, b)


def test_windowed_mixed_precision_combinations(wctx):
    """FP32 operator on FP32 / FP64 vectors: every combination the mixed hierarchy launches."""
    ctx = wctx
    n = 4000
    ptr, col, val = banded(n, n, 30, seed=11)
    A32 = _f32csr(ctx, n, n, ptr, col, val)
    assert A32.window()["windowed"]
    rng = np.random.default_rng(2)
    x, f, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    d = rng.uniform(0.1, 1.0, n).astype(np.float32)

    def run():
        out = []
        vx, vf = ctx.vector(x), ctx.vector(f)
        fx, ff = _f32vec(ctx, x), _f32vec(ctx, f)
        vy = ctx.vector(y); ctx.spmv(1.0, A32, vx, 0.5, vy); out.append(vy.numpy())          # FD
        fy = _f32vec(ctx, y); ctx.spmv(1.0, A32, fx, 0.5, fy); out.append(fy.numpy32())       # FF
        vz = ctx.vector(y); ctx.spmv(1.0, A32, fx, 1.0, vz); out.append(vz.numpy())           # FFD
        vr = ctx.vector(n); ctx.residual(vf, A32, vx, vr); out.append(vr.numpy())             # FD
        fr = _f32vec(ctx, np.zeros(n)); ctx.residual(vf, A32, vx, fr); out.append(fr.numpy32())   # FDF
        fr2 = _f32vec(ctx, np.zeros(n)); ctx.residual(ff, A32, fx, fr2); out.append(fr2.numpy32())  # FF
        fd, ft = _f32vec(ctx, d), _f32vec(ctx, np.zeros(n))
        fxx = _f32vec(ctx, x); ctx.relax(A32, ff, fxx, ft, fd, 0.72); out.append(fxx.numpy32())     # FF sweep
        vxx = ctx.vector(x); ctx.relax(A32, vf, vxx, ft, fd, 0.72); out.append(vxx.numpy())         # FD sweep
        return np.concatenate([np.asarray(v, dtype=np.float64) for v in out])
    a, b = both(ctx, run)
    assert np.array_equal(a, b)


def test_blocks_cut_on_upload_still_compute_the_same(wctx):
    """Windows that do not fit force the upload to cut row blocks: more blocks, same result."""
    ctx = wctx
    ctx.set_option("window_ratio", 1000)
    ptr, col, val = banded(6000, 90000, 6, seed=4, spread=8)
    nr, nc = 6000, 90000
    ctx.set_option("window", 0)
    Aplain = ctx.csr(nr, nc, ptr, col, val)
    ctx.set_option("window", 1)
    A = ctx.csr(nr, nc, ptr, col, val)
    assert A.window()["windowed"] and not Aplain.window()["windowed"]
    assert A.plan()["blocks"] > Aplain.plan()["blocks"]
    rng = np.random.default_rng(6)
    x = rng.uniform(-1, 1, nc)
    vx, vy, vz = ctx.vector(x), ctx.vector(nr), ctx.vector(nr)
    ctx.spmv(1.0, A, vx, 0.0, vy)
    ctx.spmv(1.0, Aplain, vx, 0.0, vz)
    assert np.array_equal(vy.numpy(), vz.numpy())
    assert rel_err(vy.numpy(), oracle.c().spmv(1.0, (ptr, col, val), x, 0.0, np.zeros(nr))) < 1e-12


@pytest.mark.parametrize("relax,krylov,precision", [("damped_jacobi", "cg", "f64"), ("spai0", "bicgstab", "f64"),
                                                     ("damped_jacobi", "cg", "mixed")])
def test_solver_is_bit_transparent_to_the_windowed_format(wctx, known_answers, relax, krylov, precision):
    """The whole drop-in solve (hierarchy uploaded windowed where it qualifies) against the same
    solve with plain operators: same iterations, same solution bits; and the reference's
    iteration count."""
    ctx = wctx
    n = 32
    ptr, col, val, rhs = ab.poisson3d(n)
    res = []
    for window in (1, 0):
        ctx.set_option("window", window)
        S = ab.DropinSolver(ptr, col, val, relax, krylov, ctx=ctx, precision=precision)
        x, it, r = S.solve(rhs)
        res.append((x, it, r))
        S.close()
    ctx.set_option("window", 0)
    assert res[0][1] == res[1][1] and np.array_equal(res[0][0], res[1][0])
    if precision == "f64":
        case = [c for c in known_answers["cases"] if (c["n"], c["relax"], c["krylov"]) == (n, relax, krylov)][0]
        assert res[0][1] == case["iters"]
