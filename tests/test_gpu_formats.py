"""GPU parity of the compressed column formats (csr_kernels.cuh FMT_PATTERN / FMT_OFFSET;
patterns.cuh, offsets.cuh): the same operator streamed pattern-indexed (no per-entry columns),
offset-indexed (8-bit columns) and with plain int32 columns must give the same bits -- only
where the column number comes from changes -- and agree with the oracle."""
import numpy as np
import pytest

import amgcl_b200 as ab
import oracle
from conftest import rel_err
from test_gpu_primitives import _f32csr, _f32vec
from test_offsets import diag_matrix

pytestmark = pytest.mark.gpu

FORMATS = ["patterns", "offsets"]


@pytest.fixture(params=FORMATS)
def fctx(ctx, request):
    """(ctx, format): every operator that qualifies for `format` is stored that way, whatever
    its size; the other compressed format is off."""
    fmt = request.param
    other = [f for f in FORMATS if f != fmt][0]
    ctx.set_option(fmt + "_min_nnz", 0)
    ctx.set_option(fmt, 1)
    ctx.set_option(other, 0)
    yield ctx, fmt
    for f in FORMATS:
        ctx.set_option(f + "_min_nnz", 1000000)
        ctx.set_option(f, 1)


def stored(A, fmt):
    return A.patterns()["pattern_indexed"] if fmt == "patterns" else A.offsets()["offset_indexed"]


def both(ctx, fmt, fn):
    """fn() with the compressed format, then with plain columns."""
    ctx.set_option(fmt, 1)
    a = fn()
    ctx.set_option(fmt, 0)
    b = fn()
    ctx.set_option(fmt, 1)
    return a, b


# (lanes per row, offsets, probability an entry is kept)
STENCILS = {
    1: ([-900, -30, -1, 0, 1, 30, 900], 0.9),                                                   # 7-point, ragged
    2: ([d + 30 * j + 900 * k for k in (-1, 0, 1) for j in (-1, 0, 1) for d in (-1, 0, 1)], 1.0),  # 27-point
}


@pytest.mark.parametrize("lanes", [1, 2])
def test_compressed_formats_give_the_bits_of_plain_columns(fctx, lanes):
    ctx, fmt = fctx
    o = oracle.c()
    nr = nc = 5001
    offs, keep = STENCILS[lanes]
    ptr, col, val = diag_matrix(nr, nc, offs, seed=lanes + nr, keep=keep)
    A = ctx.csr(nr, nc, ptr, col, val)
    assert A.plan()["lanes"] == lanes and stored(A, fmt)
    rng = np.random.default_rng(1)
    x, y, f = rng.uniform(-1, 1, nc), rng.uniform(-1, 1, nr), rng.uniform(-1, 1, nr)
    vx, vf = ctx.vector(x), ctx.vector(f)

    def spmv(beta):
        vy = ctx.vector(y)
        ctx.spmv(1.5, A, vx, beta, vy)
        return vy.numpy()
    for beta in (0.0, -0.25):
        a, b = both(ctx, fmt, lambda: spmv(beta))
        assert np.array_equal(a, b)
        assert rel_err(a, o.spmv(1.5, (ptr, col, val), x, beta, y)) < 1e-12

    def resid():
        vr = ctx.vector(nr)
        ctx.residual(vf, A, vx, vr)
        return vr.numpy()
    a, b = both(ctx, fmt, resid)
    assert np.array_equal(a, b)
    assert rel_err(a, o.residual(f, (ptr, col, val), x)) < 1e-12

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
        a, b = both(ctx, fmt, lambda: sweep(zero))
        assert np.array_equal(a, b)
    x1 = x + 0.72 * d * o.residual(f, (ptr, col, val), x)
    assert rel_err(sweep(False)[:nr], x1) < 1e-12

    # the streaming pass that also leaves scalars behind (CG: q = A p with <q, p>)
    def step():
        K = ab.Krylov(ctx, nr)                        # (fresh scalars: no history from the other run)
        vp, vq, vxx, vr = ctx.vector(x), ctx.vector(nr), ctx.vector(y), ctx.vector(f)
        K.cg_direction(vf, vf, vp)
        K.cg_step(A, vp, vq, vxx, vr)
        s = K.scalars()
        K.close()
        return np.concatenate([vq.numpy(), vxx.numpy(), vr.numpy(), [s["qp"], s["alpha"], s["rr"]]])
    a, b = both(ctx, fmt, step)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("lanes,width", [(1, 9), (2, 30), (4, 50)])
def test_rectangular_operators_all_lane_widths(fctx, lanes, width):
    """A band that never leaves the matrix (every row has the same pattern), 1 / 2 / 4 lanes."""
    ctx, fmt = fctx
    nr, nc = 4000, 4000 + 3 * width
    offs = [3 * k for k in range(width)]
    ptr, col, val = diag_matrix(nr, nc, offs, seed=width, keep=1.0)
    A = ctx.csr(nr, nc, ptr, col, val)
    assert A.plan()["lanes"] == lanes and stored(A, fmt)
    rng = np.random.default_rng(3)
    x, y = rng.uniform(-1, 1, nc), rng.uniform(-1, 1, nr)
    vx = ctx.vector(x)

    def spmv():
        vy = ctx.vector(y)
        ctx.spmv(1.0, A, vx, 2.0, vy)
        return vy.numpy()
    a, b = both(ctx, fmt, spmv)
    assert np.array_equal(a, b)
    assert rel_err(a, oracle.c().spmv(1.0, (ptr, col, val), x, 2.0, y)) < 1e-12


def test_mixed_precision_combinations(fctx):
    """FP32 operator on FP32 / FP64 vectors: every combination the mixed hierarchy launches."""
    ctx, fmt = fctx
    n = 6000
    ptr, col, val = diag_matrix(n, n, STENCILS[2][0], seed=11, keep=1.0)
    A32 = _f32csr(ctx, n, n, ptr, col, val)
    assert stored(A32, fmt)
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
    a, b = both(ctx, fmt, run)
    assert np.array_equal(a, b)


def test_operators_that_do_not_qualify_stay_plain(fctx):
    ctx, fmt = fctx
    rng = np.random.default_rng(9)
    nr = nc = 3000
    ptr = np.arange(nr + 1, dtype=np.int64) * 8
    col = np.sort(rng.integers(0, nc, (nr, 8)), axis=1).ravel()
    val = rng.uniform(-1, 1, col.size)
    A = ctx.csr(nr, nc, ptr, col, val)
    assert not stored(A, fmt)
    x = rng.uniform(-1, 1, nc)
    vx, vy = ctx.vector(x), ctx.vector(nr)
    ctx.spmv(1.0, A, vx, 0.0, vy)
    assert rel_err(vy.numpy(), oracle.c().spmv(1.0, (ptr, col, val), x, 0.0, np.zeros(nr))) < 1e-12


@pytest.mark.parametrize("relax,krylov,precision", [("damped_jacobi", "cg", "f64"), ("spai0", "bicgstab", "f64"),
                                                     ("damped_jacobi", "cg", "mixed")])
def test_solver_is_bit_transparent_to_the_format(fctx, known_answers, relax, krylov, precision):
    """The whole drop-in solve with the finest operator compressed against the same solve with
    plain columns: same iterations, same solution bits; and the reference's iteration count."""
    ctx, fmt = fctx
    n = 32
    ptr, col, val, rhs = ab.poisson3d(n)
    res = []
    for on in (1, 0):
        ctx.set_option(fmt, on)
        S = ab.DropinSolver(ptr, col, val, relax, krylov, ctx=ctx, precision=precision)
        x, it, r = S.solve(rhs)
        res.append((x, it, r))
        S.close()
    ctx.set_option(fmt, 1)
    assert res[0][1] == res[1][1] and np.array_equal(res[0][0], res[1][0])
    if precision == "f64":
        case = [c for c in known_answers["cases"] if (c["n"], c["relax"], c["krylov"]) == (n, relax, krylov)][0]
        assert res[0][1] == case["iters"]
