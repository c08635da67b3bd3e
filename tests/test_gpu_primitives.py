"""GPU parity tests proper: every C-ABI primitive against the oracle on identical
inputs -- the committed golden vectors written by the real reference, and seeded
random cases covering the edge cases the CSR kernels have to survive."""
import numpy as np
import pytest

import amgcl_b200 as ab
import oracle
from conftest import rel_err, TOL_PRIMITIVE

pytestmark = pytest.mark.gpu


def up(ctx, A):
    ptr, col, val = A
    return ctx.csr(ptr.size - 1, int(col.max()) + 1 if col.size else 0, ptr, col, val)


@pytest.mark.parametrize("variant", [0, 1])
def test_golden_level0_primitives(ctx, golden, variant):
    ctx.set_option("spmv_variant", variant)
    try:
        g = golden
        ptr, col, val = g.levels[0]["A"]
        n = ptr.size - 1
        A = ctx.csr(n, n, ptr, col, val)
        a, b, c = g["in_a"], g["in_b"], g["in_c"]
        va, vb, vc = ctx.vector(a), ctx.vector(b), ctx.vector(c)

        ctx.spmv(2.0, A, va, 0.0, vb)
        assert rel_err(vb.numpy(), g["spmv_2_a_0"]) < TOL_PRIMITIVE
        vb.upload(b)
        ctx.spmv(2.0, A, va, -0.5, vb)
        assert rel_err(vb.numpy(), g["spmv_2_a_m05_b"]) < TOL_PRIMITIVE
        vr = ctx.vector(n)
        ctx.residual(vc, A, va, vr)
        assert rel_err(vr.numpy(), g["residual_c_a"]) < TOL_PRIMITIVE
        assert abs(ctx.dot(va, vc) - float(g["dot_a_c"])) < 1e-12
        vb.upload(b)
        ctx.axpby(0.3, va, 1.7, vb)
        assert rel_err(vb.numpy(), g["axpby_03_a_17_b"]) < TOL_PRIMITIVE
        vb.upload(b)
        ctx.axpbypcz(0.3, va, 1.7, vb, -2.0, vc)
        assert rel_err(vc.numpy(), g["axpbypcz"]) < TOL_PRIMITIVE
        vc.upload(c)
        ctx.vmul(0.72, va, vb, 1.0, vc)
        assert rel_err(vc.numpy(), g["vmul_072_a_b_1_c"]) < TOL_PRIMITIVE
        vc.upload(c)
        ctx.vmul(1.0, va, vb, 0.0, vc)
        assert rel_err(vc.numpy(), g["vmul_1_a_b_0_c"]) < TOL_PRIMITIVE
    finally:
        ctx.set_option("spmv_variant", 1)


def test_golden_transfer_operators(ctx, golden):
    g = golden
    lv = g.levels[0]
    n = lv["A"][0].size - 1
    nc = lv["R"][0].size - 1
    R = ctx.csr(nc, n, *lv["R"])
    P = ctx.csr(n, nc, *lv["P"])
    va, vu, vb = ctx.vector(g["in_a"]), ctx.vector(g["in_u"]), ctx.vector(g["in_b"])
    vf = ctx.vector(nc)
    ctx.spmv(1.0, R, va, 0.0, vf)                     # restriction  (amg.hpp:540)
    assert rel_err(vf.numpy(), g["restrict_a"]) < TOL_PRIMITIVE
    ctx.spmv(1.0, P, vu, 1.0, vb)                     # prolongation (amg.hpp:545)
    assert rel_err(vb.numpy(), g["prolong_u_acc_b"]) < TOL_PRIMITIVE


def test_zero_coefficient_never_reads_output(ctx, golden):
    """Outputs whose coefficient is zero may hold NaNs (matrix_ops.hpp:65,73;
    builtin.hpp:1197,1224,1253): the kernels must not touch them."""
    g = golden
    ptr, col, val = g.levels[0]["A"]
    n = ptr.size - 1
    A = ctx.csr(n, n, ptr, col, val)
    nan = np.full(n, np.nan)
    va, vb = ctx.vector(g["in_a"]), ctx.vector(g["in_b"])
    out = ctx.vector(nan)
    ctx.spmv(2.0, A, va, 0.0, out)
    assert rel_err(out.numpy(), g["spmv_2_a_0"]) < TOL_PRIMITIVE
    out.upload(nan)
    ctx.axpby(0.3, va, 0.0, out)
    assert rel_err(out.numpy(), g["axpby_03_a_0_b"]) < TOL_PRIMITIVE
    out.upload(nan)
    ctx.axpbypcz(0.3, va, 1.7, vb, 0.0, out)
    assert rel_err(out.numpy(), g["axpbypcz_c0"]) < TOL_PRIMITIVE
    out.upload(nan)
    ctx.vmul(1.0, va, vb, 0.0, out)
    assert rel_err(out.numpy(), g["vmul_1_a_b_0_c"]) < TOL_PRIMITIVE


@pytest.mark.parametrize("fuse", [1, 0])
def test_smoother_sweep_matches_oracle(ctx, golden, fuse):
    """b200_relax == residual + vmul of the reference (damped_jacobi.hpp:108-109,
    spai0.hpp:91-92), fused or not, and with the x == 0 shortcut."""
    o = oracle.c()
    ctx.set_option("fuse_relax", fuse)
    try:
        for lv in golden.levels:
            ptr, col, val = lv["A"]
            n = ptr.size - 1
            rng = np.random.default_rng(n)
            A = ctx.csr(n, n, ptr, col, val)
            rhs, x0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
            vr, vx, vt, vd = ctx.vector(rhs), ctx.vector(x0), ctx.vector(n), ctx.vector(lv["diag"])
            ctx.relax(A, vr, vx, vt, vd, golden.omega)
            want = o.relax(lv["A"], rhs, x0, lv["diag"], golden.omega)
            assert rel_err(vx.numpy(), want) < TOL_PRIMITIVE
            ctx.relax(A, vr, vx, vt, vd, golden.omega)            # second sweep (storage swapped)
            want = o.relax(lv["A"], rhs, want, lv["diag"], golden.omega)
            assert rel_err(vx.numpy(), want) < TOL_PRIMITIVE
            ctx.clear(vx)                                          # lazy zero -> shortcut
            ctx.relax(A, vr, vx, vt, vd, golden.omega)
            want = o.relax(lv["A"], rhs, np.zeros(n), lv["diag"], golden.omega)
            assert rel_err(vx.numpy(), want) < TOL_PRIMITIVE
    finally:
        ctx.set_option("fuse_relax", 1)


def test_lazy_clear_semantics(ctx):
    n = 1000
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, n)
    va, vz = ctx.vector(a), ctx.vector(a)
    ctx.clear(vz)
    assert not vz.numpy().any()
    assert ctx.dot(va, vz) == 0.0
    ctx.axpby(2.0, va, 1.0, vz)           # z = 2a + 0
    assert rel_err(vz.numpy(), 2 * a) < 1e-16
    ctx.clear(vz)
    ctx.copy(vz, va)                      # copies the zero
    assert not va.numpy().any()
    fresh = ctx.vector(n)                 # create_vector is zero filled
    assert not fresh.numpy().any()
    ctx.clear(vz)
    assert vz.data_ptr() != 0 and not vz.numpy().any()


def test_coarse_solver_matches_reference_lu(ctx, golden):
    ptr, col, val = golden.coarse
    n = ptr.size - 1
    S = ctx.coarse(n, ptr, col, val)
    vg, vx = ctx.vector(golden["in_g"]), ctx.vector(n)
    ctx.coarse_solve(S, vg, vx)
    assert rel_err(vx.numpy(), golden["coarse_solve_g"]) < 1e-11


def test_coarse_solver_needs_pivoting(ctx):
    # a system whose natural-order elimination hits a zero pivot
    A = np.array([[0.0, 2.0, 1.0], [1.0, 0.0, 3.0], [4.0, 1.0, 0.0]])
    import scipy.sparse as sp
    M = sp.csr_matrix(A)
    S = ctx.coarse(3, M.indptr, M.indices, M.data)
    b = np.array([1.0, -2.0, 0.5])
    vb, vx = ctx.vector(b), ctx.vector(3)
    ctx.coarse_solve(S, vb, vx)
    assert rel_err(vx.numpy(), np.linalg.solve(A, b)) < 1e-13
    with pytest.raises(ab.B200Error):
        ctx.coarse(2, np.array([0, 2, 4]), np.array([0, 1, 0, 1]), np.array([1.0, 2.0, 2.0, 4.0]))


@pytest.mark.parametrize("lanes", [0, 1, 2, 4, 8, 16, 32])
@pytest.mark.parametrize("variant", [0, 1])
def test_ragged_matrix_all_lane_widths(ctx, lanes, variant):
    """Empty rows, a row longer than a stage, a short last quad, rectangular shape."""
    o = oracle.c()
    rng = np.random.default_rng(lanes * 7 + variant)
    nr, nc = 3001, 2500
    lens = rng.integers(0, 40, nr)
    lens[:9] = 0
    lens[100] = 7000
    lens[nr - 1] = 3
    ptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    col = rng.integers(0, nc, ptr[-1])
    val = rng.uniform(-1, 1, ptr[-1])
    x, y, f = rng.uniform(-1, 1, nc), rng.uniform(-1, 1, nr), rng.uniform(-1, 1, nr)
    ctx.set_option("lanes", lanes)
    ctx.set_option("spmv_variant", variant)
    try:
        A = ctx.csr(nr, nc, ptr, col, val)
        assert A.plan()["long_blocks"] >= 1
        vx, vy, vf = ctx.vector(x), ctx.vector(y), ctx.vector(f)
        ctx.spmv(1.5, A, vx, -0.25, vy)
        assert rel_err(vy.numpy(), o.spmv(1.5, (ptr, col, val), x, -0.25, y)) < 1e-12
        ctx.residual(vf, A, vx, vy)
        assert rel_err(vy.numpy(), o.residual(f, (ptr, col, val), x)) < 1e-12
    finally:
        ctx.set_option("lanes", 0)
        ctx.set_option("spmv_variant", 1)


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (4, 4), (5, 3), (257, 255), (1025, 64)])
def test_small_and_odd_shapes(ctx, shape):
    o = oracle.c()
    nr, nc = shape
    rng = np.random.default_rng(nr * 31 + nc)
    import scipy.sparse as sp
    M = sp.random(nr, nc, density=min(1.0, 6.0 / nc), random_state=1, format="csr")
    M.sort_indices()
    ptr, col, val = M.indptr.astype(np.int64), M.indices.astype(np.int64), M.data
    x, y = rng.uniform(-1, 1, nc), rng.uniform(-1, 1, nr)
    A = ctx.csr(nr, nc, ptr, col, val)
    vx, vy = ctx.vector(x), ctx.vector(y)
    ctx.spmv(1.0, A, vx, 2.0, vy)
    assert rel_err(vy.numpy(), o.spmv(1.0, (ptr, col, val), x, 2.0, y)) < 1e-13 or not val.size
    # int32 host indices take the other entry point
    A32 = ctx.csr(nr, nc, ptr.astype(np.int32), col.astype(np.int32), val)
    vy.upload(y)
    ctx.spmv(1.0, A32, vx, 2.0, vy)
    assert rel_err(vy.numpy(), o.spmv(1.0, (ptr, col, val), x, 2.0, y)) < 1e-13 or not val.size


@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 256, 257, 100003])
def test_vector_ops_odd_lengths(ctx, n):
    o = oracle.c()
    rng = np.random.default_rng(n)
    a, b, c = (rng.uniform(-1, 1, n) for _ in range(3))
    va, vb, vc = ctx.vector(a), ctx.vector(b), ctx.vector(c)
    ctx.axpby(-1.25, va, 0.5, vb)
    want_b = o.axpby(-1.25, a, 0.5, b)
    assert n == 0 or rel_err(vb.numpy(), want_b) < 1e-15
    ctx.axpbypcz(2.0, va, -1.0, vb, 0.125, vc)
    want_c = o.axpbypcz(2.0, a, -1.0, want_b, 0.125, c)
    assert n == 0 or rel_err(vc.numpy(), want_c) < 1e-15
    got = ctx.dot(va, vc)
    want = o.inner_product(a, want_c)
    assert abs(got - want) <= 1e-13 * max(1.0, np.abs(a * want_c).sum())
    ctx.copy(vc, va)
    assert n == 0 or np.array_equal(va.numpy(), vc.numpy())


def test_argument_errors(ctx):
    ptr, col, val, rhs = ab.poisson3d(4)
    A = ctx.csr(64, 64, ptr, col, val)
    good, bad = ctx.vector(64), ctx.vector(63)
    with pytest.raises(ab.B200Error):
        ctx.spmv(1.0, A, bad, 0.0, good)
    with pytest.raises(ab.B200Error):
        ctx.spmv(1.0, A, good, 0.0, good)          # aliasing
    with pytest.raises(ab.B200Error):
        ctx.axpby(1.0, good, 1.0, bad)
    with pytest.raises(ab.B200Error):
        ctx.csr(64, 10, ptr, col, val)              # column index out of range
    with pytest.raises(ab.B200Error):
        ctx.set_option("no_such_option", 1)


def test_dot_is_deterministic_and_compensated(ctx):
    n = 1 << 22
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, n) * 1e8
    a[::2] = -a[1::2]                      # massive cancellation
    a[0] += 1.0
    ones = np.ones(n)
    va, vo = ctx.vector(a), ctx.vector(ones)
    r1 = ctx.dot(va, vo)
    r2 = ctx.dot(va, vo)
    assert r1 == r2
    import math
    assert abs(r1 - math.fsum(a)) < 1e-6


def _f32vec(ctx, a):
    """FP32 device vector from a numpy array (C ABI: b200_vec_create_f32 / upload_f32)."""
    import ctypes
    L = ab.lib()
    L.b200_vec_create_f32.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
    L.b200_vec_upload_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.b200_vec_download_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    v = ab.Vector.__new__(ab.Vector)
    v.ctx, v.h = ctx, ctypes.c_void_p()
    a = np.ascontiguousarray(a, dtype=np.float32)
    v.n = a.size
    assert L.b200_vec_create_f32(ctx.h, v.n, ctypes.byref(v.h)) == 0
    assert L.b200_vec_upload_f32(v.h, a.ctypes.data, v.n) == 0

    def numpy32():
        out = np.empty(v.n, dtype=np.float32)
        assert L.b200_vec_download_f32(v.h, out.ctypes.data, v.n) == 0, L.b200_last_error()
        return out
    v.numpy32 = numpy32
    return v


def _f32csr(ctx, nrows, ncols, ptr, col, val):
    import ctypes
    L = ab.lib()
    L.b200_csr_create_i64_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    A = ab.Csr.__new__(ab.Csr)
    A.ctx, A.h = ctx, ctypes.c_void_p()
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float32)
    assert L.b200_csr_create_i64_f32(ctx.h, nrows, ncols, ptr.ctypes.data, col.ctypes.data,
                                     val.ctypes.data, ctypes.byref(A.h)) == 0, L.b200_last_error()
    A.nrows, A.ncols, A.nnz = nrows, ncols, int(ptr[-1])
    return A


def test_mixed_precision_primitives(ctx, golden):
    """Every precision combination the mixed composition produces, against a numpy model that
    rounds where the kernels round (row sums live in the output's element type)."""
    g = golden
    ptr, col, val = g.levels[0]["A"]
    n = ptr.size - 1
    import scipy.sparse as sp
    A64 = sp.csr_matrix((val, col, ptr), shape=(n, n))
    A32 = _f32csr(ctx, n, n, ptr, col, val)
    a, b, c = g["in_a"], g["in_b"], g["in_c"]
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    Af = sp.csr_matrix((val.astype(np.float32), col, ptr), shape=(n, n))
    scale = np.abs(A64).dot(np.abs(a)).max()

    # FP32 operator on FP64 vectors (the Krylov solver's q = A p and r = f - A x)
    va, vy = ctx.vector(a), ctx.vector(n)
    ctx.spmv(1.0, A32, va, 0.0, vy)
    assert np.abs(vy.numpy() - Af.astype(np.float64).dot(a)).max() < 1e-13 * scale
    vc, vr = ctx.vector(c), ctx.vector(n)
    ctx.residual(vc, A32, va, vr)
    assert np.abs(vr.numpy() - (c - Af.astype(np.float64).dot(a))).max() < 1e-13 * scale

    # all FP32 (levels below the finest)
    fa, fy = _f32vec(ctx, a32), _f32vec(ctx, np.zeros(n))
    ctx.spmv(1.0, A32, fa, 0.0, fy)
    want = Af.astype(np.float64).dot(a32.astype(np.float64))
    assert np.abs(fy.numpy32() - want).max() < 5e-6 * scale

    # finest-level residual: FP64 rhs and x, FP32 result
    ft = _f32vec(ctx, np.zeros(n))
    ctx.residual(vc, A32, va, ft)
    assert np.abs(ft.numpy32() - (c - Af.astype(np.float64).dot(a))).max() < 5e-6 * scale

    # prolongation of an FP32 correction into the FP64 iterate
    lv = g.levels[0]
    nc = lv["R"][0].size - 1
    P32 = _f32csr(ctx, n, nc, *lv["P"])
    fu = _f32vec(ctx, g["in_u"])
    vb = ctx.vector(b)
    ctx.spmv(1.0, P32, fu, 1.0, vb)
    Pf = sp.csr_matrix((lv["P"][2].astype(np.float32), lv["P"][1], lv["P"][0]), shape=(n, nc))
    want = b + Pf.astype(np.float64).dot(g["in_u"].astype(np.float32).astype(np.float64))
    assert np.abs(vb.numpy() - want).max() < 1e-13 * max(1.0, np.abs(want).max())

    # smoother sweep: FP32 operator + diagonal + scratch, FP64 rhs and iterate (fused and not)
    d32 = lv["diag"].astype(np.float32)
    for fuse in (1, 0):
        ctx.set_option("fuse_relax", fuse)
        try:
            vx, vrhs = ctx.vector(a), ctx.vector(c)
            fd, ftmp = _f32vec(ctx, d32), _f32vec(ctx, np.zeros(n))
            ctx.relax(A32, vrhs, vx, ftmp, fd, g.omega)
            t = c - Af.astype(np.float64).dot(a)
            want = a + (g.omega * d32.astype(np.float64)) * t
            assert np.abs(vx.numpy() - want).max() < 2e-6 * max(1.0, np.abs(want).max())
            ctx.clear(vx)
            ctx.relax(A32, vrhs, vx, ftmp, fd, g.omega)       # x == 0 shortcut, mixed
            want0 = (g.omega * d32.astype(np.float64)) * c
            assert np.abs(vx.numpy() - want0).max() < 2e-6 * max(1.0, np.abs(want0).max())
        finally:
            ctx.set_option("fuse_relax", 1)

    # unsupported mixes are refused, not silently converted
    with pytest.raises(ab.B200Error):
        ctx.axpby(1.0, fa, 1.0, va)
    A64d = ctx.csr(n, n, ptr, col, val)
    with pytest.raises(ab.B200Error):
        ctx.spmv(1.0, A64d, fa, 0.0, vy)


def test_wrapping_torch_memory(ctx):
    """b200_vec_wrap: the primitives run on caller-owned device memory (here torch tensors on
    torch's current stream); the fused sweep copies back instead of swapping storage."""
    import ctypes
    import torch
    o = oracle.c()
    ptr, col, val, rhs = ab.poisson3d(10)
    n = ptr.size - 1
    rng = np.random.default_rng(11)
    xh, fh = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    dev = torch.device("cuda", 0)
    tx = torch.tensor(xh, device=dev)
    tf = torch.tensor(fh, device=dev)
    ty = torch.zeros(n, dtype=torch.float64, device=dev)
    L = ab.lib()

    def wrap(t):
        v = ab.Vector.__new__(ab.Vector)
        v.ctx, v.h, v.n = ctx, ctypes.c_void_p(), t.numel()
        assert L.b200_vec_wrap(ctx.h, ctypes.c_void_p(t.data_ptr()), v.n, ctypes.byref(v.h)) == 0
        return v

    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        A = ctx.csr(n, n, ptr, col, val)
        vx, vf, vy = wrap(tx), wrap(tf), wrap(ty)
        ctx.residual(vf, A, vx, vy)
        torch.cuda.synchronize()
        assert rel_err(ty.cpu().numpy(), o.residual(fh, (ptr, col, val), xh)) < TOL_PRIMITIVE
        d = o.relax_diag((ptr, col, val), "damped_jacobi")
        vd = ctx.vector(d)
        before = tx.data_ptr()
        ctx.relax(A, vf, vx, vy, vd, 0.72)             # x is external: result must land in tx
        torch.cuda.synchronize()
        assert tx.data_ptr() == before
        assert rel_err(tx.cpu().numpy(), o.relax((ptr, col, val), fh, xh, d, 0.72)) < TOL_PRIMITIVE
        assert abs(ctx.dot(vx, vx) - float((tx * tx).sum().item())) < 1e-9
    finally:
        ctx.set_stream(None)


# ---------------------------------------------------------------------------
# recorded call sequences (b200_graph_*)
# ---------------------------------------------------------------------------
def _two_grid_sequence(ctx, A, vr, vx, vt, vd, vy, omega):
    """clear -> sweep (x == 0 shortcut) -> sweep (storage swap) -> residual -> axpby: the
    pieces of a V-cycle that carry host-side state."""
    ctx.clear(vx)
    ctx.relax(A, vr, vx, vt, vd, omega)
    ctx.relax(A, vr, vx, vt, vd, omega)
    ctx.residual(vr, A, vx, vy)
    ctx.axpby(0.5, vx, 2.0, vy)


def test_graph_replay_matches_direct_execution(ctx, golden):
    """A recorded sequence replays bit-identically, only from the vector state it was
    recorded in, and leaves the same state behind as the direct calls."""
    lv = golden.levels[0]
    ptr, col, val = lv["A"]
    n = ptr.size - 1
    rng = np.random.default_rng(11)
    A = ctx.csr(n, n, ptr, col, val)
    rhs = rng.uniform(-1, 1, n)
    vr, vx, vt, vd, vy = ctx.vector(rhs), ctx.vector(n), ctx.vector(n), ctx.vector(lv["diag"]), ctx.vector(n)

    _two_grid_sequence(ctx, A, vr, vx, vt, vd, vy, golden.omega)        # direct
    want_x, want_y = vx.numpy(), vy.numpy()

    launches0 = ctx.launches
    assert ctx.graph_begin()
    _two_grid_sequence(ctx, A, vr, vx, vt, vd, vy, golden.omega)        # recorded, runs at graph_end
    g = ctx.graph_end()
    per_run = ctx.launches - launches0
    assert np.array_equal(vx.numpy(), want_x) and np.array_equal(vy.numpy(), want_y)
    info = g.info()
    assert info["kernels"] == per_run >= 4 and info["nodes"] >= info["kernels"] and not info["stale"]

    # the sequence swapped x <-> tmp once: the state differs from the recorded one, so the
    # graph must refuse; one direct run swaps back, then it replays
    assert not g.launch()
    assert np.array_equal(vx.numpy(), want_x)
    _two_grid_sequence(ctx, A, vr, vx, vt, vd, vy, golden.omega)
    for _ in range(2):
        vy.upload(np.full(n, np.nan))
        launches1 = ctx.launches
        assert g.launch()
        assert ctx.launches - launches1 == per_run
        assert np.array_equal(vx.numpy(), want_x) and np.array_equal(vy.numpy(), want_y)
        assert not g.launch()                                          # parity flipped again
        _two_grid_sequence(ctx, A, vr, vx, vt, vd, vy, golden.omega)
    assert g.info()["replays"] == 2

    # a new right-hand side flows through the replay (pointers are baked in, data is not)
    rhs2 = rng.uniform(-1, 1, n)
    vr.upload(rhs2)
    assert g.launch()
    got_x, got_y = vx.numpy(), vy.numpy()
    _two_grid_sequence(ctx, A, vr, vx, vt, vd, vy, golden.omega)        # parity back
    _two_grid_sequence(ctx, A, vr, vx, vt, vd, vy, golden.omega)        # same state as the replay had
    assert np.array_equal(vx.numpy(), got_x) and np.array_equal(vy.numpy(), got_y)

    # option changes and destroyed operands make the graph stale
    ctx.set_option("fuse_relax", 0)
    try:
        assert g.info()["stale"] and not g.launch()
    finally:
        ctx.set_option("fuse_relax", 1)
    g.close()


def test_graph_recording_rejects_host_synchronous_calls_and_aborts_cleanly(ctx, golden):
    lv = golden.levels[0]
    ptr, col, val = lv["A"]
    n = ptr.size - 1
    rng = np.random.default_rng(12)
    A = ctx.csr(n, n, ptr, col, val)
    rhs, x0 = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    vr, vx, vt, vd = ctx.vector(rhs), ctx.vector(x0), ctx.vector(n), ctx.vector(lv["diag"])
    launches0 = ctx.launches
    assert ctx.graph_begin()
    ctx.relax(A, vr, vx, vt, vd, golden.omega)          # swaps storage on the host side
    ctx.clear(vt)
    for bad in (lambda: ctx.dot(vr, vx), lambda: vx.numpy(), lambda: ctx.vector(n), lambda: ctx.sync()):
        with pytest.raises(ab.B200Error, match="recorded"):
            bad()
    with pytest.raises(ab.B200Error):
        ctx.graph_begin()                               # no nesting
    ctx.graph_abort()
    assert ctx.launches == launches0
    # nothing ran and the state is the one before graph_begin
    assert np.array_equal(vx.numpy(), x0)
    ctx.relax(A, vr, vx, vt, vd, golden.omega)
    want = oracle.c().relax(lv["A"], rhs, x0, lv["diag"], golden.omega)
    assert rel_err(vx.numpy(), want) < TOL_PRIMITIVE

    # contexts that cannot record say so instead of failing
    ctx.set_option("cycle_graph", 0)
    try:
        assert not ctx.graph_begin()
    finally:
        ctx.set_option("cycle_graph", 1)
    ctx.profile_begin()
    try:
        assert not ctx.graph_begin()
    finally:
        ctx.profile_end()


# ---------------------------------------------------------------------------
# index lists: Backend::gather / Backend::scatter (cuda.hpp:548-577)
# ---------------------------------------------------------------------------
def test_gather_scatter(ctx):
    rng = np.random.default_rng(21)
    n, m = 10007, 3001
    src = rng.uniform(-1, 1, n)
    idx = rng.permutation(n)[:m]                      # distinct, unordered
    I = ctx.index(idx, n)
    vs, vg = ctx.vector(src), ctx.vector(m)
    ctx.gather(I, vs, vg)
    assert np.array_equal(vg.numpy(), src[idx])
    assert np.array_equal(ctx.gather_host(I, vs), src[idx])
    base = rng.uniform(-1, 1, n)
    vd = ctx.vector(base)
    ctx.scatter(I, vg, vd)
    want = base.copy()
    want[idx] = src[idx]
    assert np.array_equal(vd.numpy(), want)
    # a lazily cleared destination is materialised before the partial overwrite
    ctx.clear(vd)
    ctx.scatter(I, vg, vd)
    want = np.zeros(n)
    want[idx] = src[idx]
    assert np.array_equal(vd.numpy(), want)
    # ... and a cleared source gathers zeros
    ctx.clear(vs)
    ctx.gather(I, vs, vg)
    assert not vg.numpy().any()
    # empty list, bad index, size mismatch
    E = ctx.index(np.zeros(0, dtype=np.int64), n)
    ctx.gather(E, vs, ctx.vector(0))
    with pytest.raises(ab.B200Error, match="out of range"):
        ctx.index([0, n], n)
    with pytest.raises(ab.B200Error, match="sizes"):
        ctx.gather(I, vg, vs)


def test_gather_scatter_through_the_cpp_backend(ctx):
    """Backend::gather / Backend::scatter of include/amgcl/backend/b200.hpp."""
    import ctypes
    D = ab.dropin_lib()
    rng = np.random.default_rng(22)
    n, m = 5000, 777
    src = rng.uniform(-1, 1, n)
    idx = np.ascontiguousarray(rng.permutation(n)[:m], dtype=np.int64)
    g1, g2, sc = np.empty(m), np.empty(m), np.empty(n)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = D.dropin_gather_scatter(ctx.h, n, p(src), m, p(idx), -7.0, p(g1), p(g2), p(sc))
    assert rc == 0, D.dropin_last_error()
    assert np.array_equal(g1, src[idx]) and np.array_equal(g2, src[idx])
    want = np.full(n, -7.0)
    want[idx] = src[idx]
    assert np.array_equal(sc, want)
