"""CPU: pin the C restatement (oracle/amg_oracle.c) to the real reference.

(1) against the committed golden vectors written by the reference itself
    (tests/golden/make_golden.py), always;
(2) against oracle/_ref/libamgcl_ref.so directly, whenever it is present.
"""
import numpy as np
import pytest

import oracle
from conftest import rel_err, TOL_PRIMITIVE, TOL_RESID_REL, TOL_SOLUTION


def test_primitives_match_golden(golden):
    o = oracle.c()
    g = golden
    A0 = g.levels[0]["A"]
    a, b, c = g["in_a"], g["in_b"], g["in_c"]
    assert rel_err(o.spmv(2.0, A0, a, 0.0, b), g["spmv_2_a_0"]) < TOL_PRIMITIVE
    assert rel_err(o.spmv(2.0, A0, a, -0.5, b), g["spmv_2_a_m05_b"]) < TOL_PRIMITIVE
    assert rel_err(o.residual(c, A0, a), g["residual_c_a"]) < TOL_PRIMITIVE
    assert abs(o.inner_product(a, c) - float(g["dot_a_c"])) < 1e-12
    assert rel_err(o.axpby(0.3, a, 1.7, b), g["axpby_03_a_17_b"]) < TOL_PRIMITIVE
    assert rel_err(o.axpbypcz(0.3, a, 1.7, b, -2.0, c), g["axpbypcz"]) < TOL_PRIMITIVE
    assert rel_err(o.vmul(0.72, a, b, 1.0, c), g["vmul_072_a_b_1_c"]) < TOL_PRIMITIVE
    assert rel_err(o.vmul(1.0, a, b, 0.0, c), g["vmul_1_a_b_0_c"]) < TOL_PRIMITIVE
    assert rel_err(o.spmv(1.0, g.levels[0]["R"], a, 0.0, np.zeros(g.levels[0]["R"][0].size - 1)),
                   g["restrict_a"]) < TOL_PRIMITIVE
    assert rel_err(o.spmv(1.0, g.levels[0]["P"], g["in_u"], 1.0, b), g["prolong_u_acc_b"]) < TOL_PRIMITIVE


def test_zero_coefficient_never_reads_output(golden):
    """beta == 0 / b == 0 / c == 0: the output may hold NaNs (builtin.hpp:1197,1224,1253)."""
    o = oracle.c()
    g = golden
    a, b = g["in_a"], g["in_b"]
    nan = np.full_like(a, np.nan)
    assert rel_err(o.axpby(0.3, a, 0.0, nan), g["axpby_03_a_0_b"]) < TOL_PRIMITIVE
    assert rel_err(o.axpbypcz(0.3, a, 1.7, b, 0.0, nan), g["axpbypcz_c0"]) < TOL_PRIMITIVE
    assert rel_err(o.vmul(1.0, a, b, 0.0, nan), g["vmul_1_a_b_0_c"]) < TOL_PRIMITIVE
    assert rel_err(o.spmv(2.0, g.levels[0]["A"], a, 0.0, nan), g["spmv_2_a_0"]) < TOL_PRIMITIVE


def test_smoother_diagonals_match_golden(golden):
    o = oracle.c()
    for lv in golden.levels:
        d = o.relax_diag(lv["A"], golden.relax)
        assert rel_err(d, lv["diag"]) < 1e-15


def test_vcycle_and_coarse_solve_match_golden(golden):
    H = oracle.Hierarchy(golden.levels, golden.coarse, golden.omega)
    # dense LU (oracle) vs skyline LU after Cuthill-McKee (reference): same factorisation
    # up to ordering, so agreement is to rounding of a well conditioned small system
    assert rel_err(H.coarse_solve(golden["in_g"]), golden["coarse_solve_g"]) < 1e-11
    assert rel_err(H.apply(golden["in_a"]), golden["precond_a"]) < 1e-11


def test_solve_matches_golden(golden):
    H = oracle.Hierarchy(golden.levels, golden.coarse, golden.omega)
    x, iters, resid, hist = H.solve(golden["rhs"], golden.krylov)
    assert iters == int(golden["iters"])
    assert abs(resid - float(golden["resid"])) <= TOL_RESID_REL * float(golden["resid"])
    assert rel_err(x, golden["x"]) < TOL_SOLUTION
    assert hist.size == iters and hist[-1] == pytest.approx(resid)
    assert np.all(np.diff(hist) < 0) or golden.krylov == "bicgstab"


def test_zero_rhs_returns_zero_solution(golden):
    """cg.hpp:162-169 / bicgstab.hpp:169-176: ||rhs|| == 0 -> x = 0, 0 iterations."""
    H = oracle.Hierarchy(golden.levels, golden.coarse, golden.omega)
    x, iters, resid, _ = H.solve(np.zeros(H.n), golden.krylov, x0=np.ones(H.n))
    assert iters == 0 and resid == 0.0 and not x.any()


needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("n", [8, 16, 32])
@pytest.mark.parametrize("relax,krylov", [("damped_jacobi", "cg"), ("spai0", "bicgstab"),
                                          ("spai0", "cg"), ("damped_jacobi", "bicgstab")])
def test_oracle_tracks_live_reference(n, relax, krylov, known_answers):
    from amgcl_b200 import poisson3d
    ptr, col, val, rhs = poisson3d(n)
    S = oracle.RefSolver(ptr, col, val, relax, krylov, coarse_enough=(50 if n == 8 else -1))
    xr, itr, resr = S.solve(rhs)
    levels, coarse = S.hierarchy()
    H = oracle.Hierarchy(levels, coarse, 0.72 if relax == "damped_jacobi" else 1.0)
    xo, ito, reso, _ = H.solve(rhs, krylov)
    assert ito == itr
    assert abs(reso - resr) <= TOL_RESID_REL * resr
    assert rel_err(xo, xr) < TOL_SOLUTION
    for case in known_answers["cases"]:
        if (case["n"], case["relax"], case["krylov"]) == (n, relax, krylov):
            assert case["iters"] == itr
            assert abs(case["resid"] - resr) <= 1e-9 * resr   # thread-count wobble only
            assert case["level_rows"] == [l["A"][0].size - 1 for l in levels] + [coarse[0].size - 1]


@needs_ref
def test_live_reference_primitives(golden):
    """The golden vectors are reproducible from the reference build in this container."""
    r = oracle.ref()
    A0 = golden.levels[0]["A"]
    N = A0[0].size - 1
    a, b, c = golden["in_a"], golden["in_b"], golden["in_c"]
    assert rel_err(r.spmv(2.0, A0, N, a, -0.5, b), golden["spmv_2_a_m05_b"]) < 1e-15
    assert rel_err(r.residual(c, A0, N, a), golden["residual_c_a"]) < 1e-15
    assert abs(r.inner_product(a, c) - float(golden["dot_a_c"])) < 1e-12
    assert rel_err(r.relax_diag(A0, golden.relax), golden.levels[0]["diag"]) == 0.0
