"""GPU tests of the fused Krylov steps (C ABI section "Krylov steps") and of the
solver::cg / solver::bicgstab specialisations for backend::b200 that call them.

Step level: one iteration body against numpy on seeded inputs (tolerance TOL_PRIMITIVE-class:
the only differences are summation order and FMA contraction).  Solver level: the fused
solvers against the reference's own call sequence on the same backend (option
"fused_krylov" = 0) and against the known answers of the reference."""
import numpy as np
import pytest
import scipy.sparse as sp

import amgcl_b200 as ab
from conftest import rel_err, TOL_RESID_REL, TOL_SOLUTION

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _csr(ptr, col, val):
    n = ptr.size - 1
    return sp.csr_matrix((val, col, ptr), shape=(n, n))


@pytest.fixture()
def system(ctx):
    ptr, col, val, _ = ab.poisson3d(20)
    A = ctx.csr(ptr.size - 1, ptr.size - 1, ptr, col, val)
    return A, _csr(ptr, col, val)


def test_residual_with_norm_and_zero_guess_shortcut(ctx, system):
    A, M = system
    n = M.shape[0]
    rng = np.random.default_rng(1)
    f, x0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    K = ab.Krylov(ctx, n)
    rhs, x, r = ctx.vector(f), ctx.vector(x0), ctx.vector(n)
    rr = K.residual(rhs, A, x, r)
    want = f - M @ x0
    assert rel_err(r.numpy(), want) < TOL
    assert abs(rr - want @ want) <= TOL * (want @ want)
    # x known to be zero: no pass over A, r = rhs exactly
    ctx.clear(x)
    before = ctx.launches
    rr = K.residual(rhs, A, x, r)
    assert ctx.launches - before == 1
    assert np.array_equal(r.numpy(), f)
    assert abs(rr - f @ f) <= TOL * (f @ f)
    K.close()


def test_cg_iteration_matches_numpy(ctx, system):
    """Two iterations of cg.hpp:180-198 with a diagonal 'preconditioner' supplied by the test."""
    A, M = system
    n = M.shape[0]
    rng = np.random.default_rng(2)
    f = rng.uniform(-1, 1, n)
    d = 1.0 / M.diagonal()
    K = ab.Krylov(ctx, n)
    rhs, x, r, s, p, q = (ctx.vector(v) for v in (f, np.zeros(n), np.zeros(n), np.zeros(n),
                                                 np.full(n, np.nan), np.zeros(n)))
    dv = ctx.vector(d)
    rr = K.residual(rhs, A, x, r)
    xr, rn = np.zeros(n), f.copy()
    pn, rho_prev = None, None
    for it in range(3):
        ctx.vmul(1.0, dv, r, 0.0, s)               # s = M^-1 r
        sn = d * rn
        rho = rn @ sn
        pn = sn.copy() if it == 0 else sn + (rho / rho_prev) * pn
        qn = M @ pn
        alpha = rho / (qn @ pn)
        xr = xr + alpha * pn
        rn = rn - alpha * qn
        rho_prev = rho
        K.cg_direction(r, s, p)                    # p = NaN before the first call: must not be read
        rr = K.cg_step(A, p, q, x, r)
        sc = K.scalars()
        assert rel_err(p.numpy(), pn) < TOL
        assert rel_err(q.numpy(), qn) < TOL
        assert rel_err(x.numpy(), xr) < TOL
        assert rel_err(r.numpy(), rn) < 1e-11
        assert abs(rr - rn @ rn) <= 1e-11 * (rn @ rn)
        assert abs(sc["rho"] - rho) <= TOL * abs(rho)
        assert abs(sc["qp"] - qn @ pn) <= TOL * abs(qn @ pn)
        assert abs(sc["alpha"] - alpha) <= 1e-11 * abs(alpha)
    K.close()


def test_bicgstab_iteration_matches_numpy(ctx):
    """Two iterations of bicgstab.hpp:198-236 (right preconditioning, T = D^-1 p)."""
    ptr, col, val, _ = ab.poisson3d(16, convection=0.7)       # non-symmetric
    n = ptr.size - 1
    A, M = ctx.csr(n, n, ptr, col, val), _csr(ptr, col, val)
    rng = np.random.default_rng(3)
    f, x0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    d = 1.0 / M.diagonal()
    K = ab.Krylov(ctx, n)
    rhs, x = ctx.vector(f), ctx.vector(x0)
    r, p, v, s, t, rh, T = (ctx.vector(n) for _ in range(7))
    dv = ctx.vector(d)
    K.residual(rhs, A, x, r)
    K.bicg_start(r, rh)
    rn = f - M @ x0
    rhn, xn = rn.copy(), x0.copy()
    assert np.array_equal(rh.numpy(), r.numpy())
    rho_prev = alpha = omega = None
    pn = vn = None
    for it in range(3):
        rho = rn @ rhn
        if it == 0:
            pn = rn.copy()
        else:
            beta = (rho * alpha) / (rho_prev * omega)
            pn = rn - beta * omega * vn + beta * pn
        Tn = d * pn
        vn = M @ Tn
        alpha = rho / (rhn @ vn)
        xn = xn + alpha * Tn
        sn = rn - alpha * vn
        K.bicg_direction(r, v, p)
        ctx.vmul(1.0, dv, p, 0.0, T)
        ss = K.bicg_step_s(A, rh, T, v, r, s, x)
        sc = K.scalars()
        assert rel_err(p.numpy(), pn) < 1e-11
        assert rel_err(v.numpy(), vn) < 1e-11
        assert rel_err(s.numpy(), sn) < 1e-10
        assert abs(ss - sn @ sn) <= 1e-10 * (sn @ sn)
        assert abs(sc["rho"] - rho) <= 1e-10 * abs(rho)
        assert abs(sc["alpha"] - alpha) <= 1e-10 * abs(alpha)
        Tn = d * sn
        tn = M @ Tn
        omega = (tn @ sn) / (tn @ tn)
        xn = xn + omega * Tn
        rn = sn - omega * tn
        rho_prev = rho
        ctx.vmul(1.0, dv, s, 0.0, T)
        rr = K.bicg_step_r(A, rh, T, t, s, r, x)
        sc = K.scalars()
        assert rel_err(t.numpy(), tn) < 1e-10
        assert rel_err(x.numpy(), xn) < 1e-10
        assert rel_err(r.numpy(), rn) < 1e-9
        assert abs(rr - rn @ rn) <= 1e-9 * (rn @ rn)
        assert abs(sc["omega"] - omega) <= 1e-10 * abs(omega)
        assert abs(sc["rho_next"] - rn @ rhn) <= 1e-9 * max(abs(rn @ rhn), rn @ rn)
    K.close()


def test_smoother_sweep_leaves_the_product_for_the_next_dot(ctx, system):
    """With a Krylov workspace of the operator's size alive, b200_relax leaves <rhs, x_new> in
    the scalar table; b200_dot returns it without a launch while both operands are untouched
    and recomputes as soon as one of them has been written."""
    A, M = system
    n = M.shape[0]
    rng = np.random.default_rng(4)
    f, x0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    d = 1.0 / M.diagonal()
    K = ab.Krylov(ctx, n)
    rhs, x, tmp, dv = ctx.vector(f), ctx.vector(x0), ctx.vector(n), ctx.vector(d)
    ctx.relax(A, rhs, x, tmp, dv, 0.72)
    ctx.flush()
    xn = x0 + 0.72 * d * (f - M @ x0)
    before = ctx.launches
    got = ctx.dot(rhs, x)
    assert ctx.launches == before                      # taken from the table
    assert abs(got - f @ xn) <= TOL * abs(f @ xn)
    assert abs(ctx.dot(x, rhs) - f @ xn) <= TOL * abs(f @ xn)
    ctx.axpby(1.0, rhs, 2.0, x)                         # x changes: the cached product is stale
    before = ctx.launches
    got = ctx.dot(rhs, x)
    assert ctx.launches == before + 1
    want = f @ (f + 2.0 * xn)
    assert abs(got - want) <= TOL * abs(want)
    K.close()
    # no workspace of this size: nothing extra is computed, dot launches its own kernel
    ctx.relax(A, rhs, x, tmp, dv, 0.72)
    ctx.flush()                                         # (a small operator: the sweep was deferred)
    before = ctx.launches
    ctx.dot(rhs, x)
    assert ctx.launches == before + 1


@pytest.mark.parametrize("relax,krylov", [("damped_jacobi", "cg"), ("spai0", "bicgstab"),
                                          ("spai0", "cg"), ("damped_jacobi", "bicgstab")])
@pytest.mark.parametrize("precision", ["f64", "mixed"])
def test_fused_solver_vs_reference_sequence(ctx, relax, krylov, precision):
    """solver::cg / bicgstab specialisations (fused steps) against the primary templates (the
    reference's sequence of primitives) on the same backend: same iteration count, residual
    within TOL_RESID_REL, solution within TOL_SOLUTION; and fewer launches."""
    n = 40
    ptr, col, val, rhs = ab.poisson3d(n)
    out = {}
    for fused in (0, 1):
        ctx.set_option("fused_krylov", fused)
        try:
            S = ab.DropinSolver(ptr, col, val, relax, krylov, ctx=ctx, precision=precision)
            S.solve(rhs)                               # warm-up (lazy allocations)
            before = ctx.launches
            x, it, res = S.solve(rhs)
            out[fused] = (x, it, res, ctx.launches - before)
            S.close()
        finally:
            ctx.set_option("fused_krylov", 1)
    (x0, it0, res0, l0), (x1, it1, res1, l1) = out[0], out[1]
    assert it1 == it0
    tol = TOL_RESID_REL
    if krylov == "bicgstab":                                    # BiCGStab amplifies rounding
        tol = 1e-3 if precision == "mixed" else 1e-4
    assert abs(res1 - res0) <= tol * res0
    assert rel_err(x1, x0) < (1e-6 if precision == "mixed" else TOL_SOLUTION)
    assert l1 < l0
    per_iter_saved = (l0 - l1) / it0
    assert per_iter_saved >= (3.5 if krylov == "cg" else 6.5), (l0, l1, it0)


@pytest.mark.parametrize("n", [32, 64])
@pytest.mark.parametrize("relax,krylov", [("damped_jacobi", "cg"), ("spai0", "bicgstab")])
def test_fused_solver_known_answers_and_sync_count(ctx, known_answers, n, relax, krylov):
    case = [c for c in known_answers["cases"]
            if (c["n"], c["relax"], c["krylov"]) == (n, relax, krylov)][0]
    ptr, col, val, rhs = ab.poisson3d(n)
    S = ab.DropinSolver(ptr, col, val, relax, krylov, ctx=ctx)
    x, iters, resid = S.solve(rhs)
    assert iters == case["iters"]
    assert abs(resid - case["resid"]) <= TOL_RESID_REL * case["resid"]
    assert abs(np.linalg.norm(x) - case["x_norm2"]) <= TOL_SOLUTION * case["x_norm2"]
    S.close()


def test_fused_solver_nonzero_guess_and_zero_rhs(ctx):
    n = 24
    ptr, col, val, rhs = ab.poisson3d(n)
    rng = np.random.default_rng(5)
    x0 = rng.uniform(-1, 1, rhs.size)
    M = _csr(ptr, col, val)
    for krylov in ("cg", "bicgstab"):
        S = ab.DropinSolver(ptr, col, val, "spai0", krylov, ctx=ctx)
        x, it, res = S.solve(rhs, x0)
        assert np.linalg.norm(rhs - M @ x) <= 2e-8 * np.linalg.norm(rhs)
        x, it, res = S.solve(np.zeros_like(rhs), x0)        # zero rhs: x = 0, 0 iterations
        assert it == 0 and not x.any()
        S.close()
