"""GPU parity of the offset-indexed column format (csr_kernels.cuh FMT_OFFSET, offsets.cuh):
the same operator streamed with 8-bit column indices and with plain int32 columns must give
the same bits -- only where the column number comes from changes -- and both must agree with
the oracle."""
import numpy as np
import pytest

import amgcl_b200 as ab
import oracle
from conftest import rel_err
from test_gpu_primitives import _f32csr, _f32vec
from test_offsets import diag_matrix

pytestmark = pytest.mark.gpu


@pytest.fixture()
def octx(ctx):
    """Every operator that qualifies is offset-indexed, whatever its size."""
    ctx.set_option("offsets_min_nnz", 0)
    ctx.set_option("offsets", 1)
    yield ctx
    ctx.set_option("offsets_min_nnz", 1000000)
    ctx.set_option("offsets", 1)


def both(ctx, fn):
    """fn() with the offset-indexed kernels, then with the plain ones."""
    ctx.set_option("offsets", 1)
    a = fn()
    ctx.set_option("offsets", 0)
    b = fn()
    ctx.set_option("offsets", 1)
    return a, b


STENCILS = {
    1: [-900, -30, -1, 0, 1, 30, 900],                                              # 7-point
    2: [d + 30 * j + 900 * k for k in (-1, 0, 1) for j in (-1, 0, 1) for d in (-1, 0, 1)],   # 27-point
    4: list(range(-26, 27)),                                                        # band of 53
}


@pytest.mark.parametrize("lanes", [1, 2, 4])
@pytest.mark.parametrize("shape", [(5001, 5001), (4000, 5203)])
def test_offset_indexed_kernels_give_the_bits_of_the_plain_ones(octx, lanes, shape):
    ctx = octx
    o = oracle.c()
    nr, nc = shape
    ptr, col, val = diag_matrix(nr, nc, STENCILS[lanes], seed=lanes + nr, keep=0.9)
    A = ctx.csr(nr, nc, ptr, col, val)
    assert A.plan()["lanes"] == lanes
    assert A.offsets()["offset_indexed"] and A.offsets()["count"] <= len(STENCILS[lanes])
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
        K = ab.Krylov(ctx, nr)

        def step():
            vp, vq, vxx, vr = ctx.vector(x), ctx.vector(nr), ctx.vector(y), ctx.vector(f)
            K.cg_direction(vf, vf, vp)
            K.cg_step(A, vp, vq, vxx, vr)
            s = K.scalars()
            return np.concatenate([vq.numpy(), vxx.numpy(), vr.numpy(), [s["qp"], s["alpha"], s["rr"]]])
        a, b = both(ctx, step)
        assert np.array_equal(a, b)
        K.close()


def test_offset_indexed_mixed_precision_combinations(octx):
    """FP32 operator on FP32 / FP64 vectors: every combination the mixed hierarchy launches."""
    ctx = octx
    n = 6000
    ptr, col, val = diag_matrix(n, n, STENCILS[2], seed=11)
    A32 = _f32csr(ctx, n, n, ptr, col, val)
    assert A32.offsets()["offset_indexed"]
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


def test_operators_with_many_offsets_stay_plain(octx):
    ctx = octx
    rng = np.random.default_rng(9)
    nr = nc = 3000
    lens = np.full(nr, 8)
    ptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    col = np.sort(rng.integers(0, nc, (nr, 8)), axis=1).ravel()
    A = ctx.csr(nr, nc, ptr, col, rng.uniform(-1, 1, col.size))
    assert not A.offsets()["offset_indexed"]


@pytest.mark.parametrize("relax,krylov,precision", [("damped_jacobi", "cg", "f64"), ("spai0", "bicgstab", "f64"),
                                                     ("damped_jacobi", "cg", "mixed")])
def test_solver_is_bit_transparent_to_the_offset_format(octx, known_answers, relax, krylov, precision):
    """The whole drop-in solve with the finest operator offset-indexed against the same solve
    with plain columns: same iterations, same solution bits; and the reference's iteration
    count."""
    ctx = octx
    n = 32
    ptr, col, val, rhs = ab.poisson3d(n)
    res = []
    for on in (1, 0):
        ctx.set_option("offsets", on)
        S = ab.DropinSolver(ptr, col, val, relax, krylov, ctx=ctx, precision=precision)
        x, it, r = S.solve(rhs)
        res.append((x, it, r))
        S.close()
    ctx.set_option("offsets", 1)
    assert res[0][1] == res[1][1] and np.array_equal(res[0][0], res[1][0])
    if precision == "f64":
        case = [c for c in known_answers["cases"] if (c["n"], c["relax"], c["krylov"]) == (n, relax, krylov)][0]
        assert res[0][1] == case["iters"]
