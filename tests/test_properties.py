"""Property-based tests (hypothesis) of the host-side logic behind the C ABI: the row-block
plan every CSR kernel walks, the multi-GPU partition / operator splitting, and the file
formats.  Pure CPU: nothing here needs a device."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

import amgcl_b200 as ab
from amgcl_b200 import io as bio
from test_capi import plan, check_plan

SETTINGS = dict(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])


def csr_mv(ptr, col, val, x):
    y = np.zeros(ptr.size - 1)
    row = np.repeat(np.arange(ptr.size - 1), np.diff(ptr))
    np.add.at(y, row, val * x[col])
    return y


@st.composite
def row_lengths(draw):
    n = draw(st.integers(0, 600))
    kind = draw(st.sampled_from(["short", "mixed", "long_tail", "empty_rows"]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    if kind == "short":
        lens = rng.integers(0, 9, n)
    elif kind == "mixed":
        lens = rng.integers(0, 200, n)
    elif kind == "long_tail":
        lens = rng.integers(0, 30, n)
        for _ in range(min(n, 3)):
            lens[rng.integers(0, n)] = rng.integers(500, 9000)
    else:
        lens = rng.integers(0, 40, n) * (rng.uniform(size=n) < 0.3)
    return lens.astype(np.int64)


@settings(**SETTINGS)
@given(row_lengths(), st.sampled_from([0, 1, 2, 4, 8, 16, 32]), st.sampled_from([256, 512, 2048, 6144]))
def test_row_block_plan_invariants(lens, lanes, cap):
    """b200_plan_i64: the blocks tile the rows exactly once in order, every block starts on a
    multiple of four rows (16-byte aligned ptr slice for the TMA copy), holds at most rows_cap
    rows and at most nnz_cap non-zeros unless it is a single over-long quad, and the block's
    first non-zero is ptr[first row]."""
    ptr = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    blk, ln, rows_cap, nlong = plan(ptr, lanes, cap)
    assert ln in (1, 2, 4, 8, 16, 32) and (lanes == 0 or ln == lanes)
    assert rows_cap % 4 == 0 and 4 <= rows_cap <= 1024
    if lens.size == 0:
        assert blk.shape[0] == 1 and nlong == 0
        return
    assert nlong == check_plan(ptr, blk, rows_cap, cap)
    # greedy packing: a block could not have taken the next quad as well
    rows = blk[:, 0].astype(np.int64)
    for b in range(rows.size - 2):
        nxt_end = min(rows[b + 1] + 4, lens.size)
        fits_rows = nxt_end - rows[b] <= rows_cap
        fits_nnz = ptr[nxt_end] - ptr[rows[b]] <= cap
        assert not (fits_rows and fits_nnz), "block %d stopped early" % b


@settings(**SETTINGS)
@given(st.integers(0, 5000), st.integers(1, 16))
def test_partition_is_a_uniform_block_cover(n, P):
    """b200_partition: one uniform block size (multiple of 4) for every rank, contiguous,
    covering [0, n) exactly; trailing ranks may be short or empty."""
    parts = [ab.partition(n, P, r) for r in range(P)]
    B = parts[0][0]
    assert B % 4 == 0 and B * P >= n and (n == 0 or B >= 4)
    assert all(p[0] == B for p in parts)
    assert parts[0][1] == 0 and parts[-1][2] == n
    for r in range(P):
        _, lo, hi = parts[r]
        assert lo == min(r * B, n) and hi == min((r + 1) * B, n)


@st.composite
def sparse_matrix(draw, square=False):
    nr = draw(st.integers(1, 60))
    nc = nr if square else draw(st.integers(1, 60))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    dens = draw(st.sampled_from([0.02, 0.1, 0.4]))
    rng = np.random.default_rng(seed)
    mask = rng.uniform(size=(nr, nc)) < dens
    if square:
        mask |= np.eye(nr, dtype=bool)
    dense = np.where(mask, rng.uniform(-1, 1, (nr, nc)), 0.0)
    ptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=ptr[1:])
    col = np.nonzero(mask)[1].astype(np.int64)
    val = dense[mask]
    return nr, nc, ptr, col, val, dense


@settings(**SETTINGS)
@given(sparse_matrix(square=True), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_split_square_reproduces_the_global_product(m, P, seed):
    """b200_dist_split_i64(kind=square): local columns first, remote columns renumbered into
    the halo (rank-major slots); owner ranks publish x[send_idx]; the pieces together give A x."""
    n, _, ptr, col, val, dense = m
    x = np.random.default_rng(seed).uniform(-1, 1, n)
    parts = [ab.dist_split("square", P, r, n, n, ptr, col, val) for r in range(P)]
    bounds = [ab.partition(n, P, r) for r in range(P)]
    S = parts[0]["slots"]
    assert all(p["slots"] == S for p in parts)
    halo = np.zeros(P * S)
    for r, p in enumerate(parts):
        lo, hi = bounds[r][1], bounds[r][2]
        assert p["send_idx"].size <= S and (p["send_idx"].size == 0 or p["send_idx"].max() < hi - lo)
        halo[r * S:r * S + p["send_idx"].size] = x[lo + p["send_idx"]]
    got = np.zeros(n)
    for r, p in enumerate(parts):
        _, lo, hi = bounds[r]
        assert p["nrows"] == hi - lo and p["n_loc"] == hi - lo
        if hi > lo:
            got[lo:hi] = csr_mv(p["ptr"], p["col"], p["val"], np.concatenate([x[lo:hi], halo]))
    assert np.allclose(got, dense @ x, rtol=0, atol=1e-12)


@settings(**SETTINGS)
@given(sparse_matrix(), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_split_rectangular_operators(m, P, seed):
    """Any rectangular operator (P_l, R_l): a rank keeps whole rows of its block of the row
    partition.  'replicated': the input vector is replicated, columns stay global; 'halo': the
    input is partitioned like the columns -- local columns + halo slots filled from every
    rank's send list.  Either way the rows reassemble to the global product exactly."""
    nr, nc, ptr, col, val, dense = m
    rng = np.random.default_rng(seed)
    u = rng.uniform(-1, 1, nc)
    rows = [ab.partition(nr, P, r) for r in range(P)]
    cols = [ab.partition(nc, P, r) for r in range(P)]
    got = np.zeros(nr)
    for r in range(P):
        p = ab.dist_split("replicated", P, r, nr, nc, ptr, col, val)
        _, lo, hi = rows[r]
        assert p["nrows"] == hi - lo and p["ncols"] == nc and p["slots"] == 0
        if hi > lo:
            got[lo:hi] = csr_mv(p["ptr"], p["col"], p["val"], u)
    assert np.allclose(got, dense @ u, rtol=0, atol=1e-12)
    parts = [ab.dist_split("halo", P, r, nr, nc, ptr, col, val) for r in range(P)]
    S = parts[0]["slots"]
    assert all(p["slots"] == S for p in parts)
    halo = np.zeros(P * S)
    for r, p in enumerate(parts):
        _, clo, chi = cols[r]
        assert p["n_loc"] == chi - clo and p["ncols"] == (chi - clo) + P * S
        if p["send_idx"].size:
            assert p["send_idx"].max() < chi - clo
        halo[r * S:r * S + p["send_idx"].size] = u[clo + p["send_idx"]]
    got = np.zeros(nr)
    for r, p in enumerate(parts):
        _, lo, hi = rows[r]
        _, clo, chi = cols[r]
        assert p["nrows"] == hi - lo
        if hi > lo:
            got[lo:hi] = csr_mv(p["ptr"], p["col"], p["val"], np.concatenate([u[clo:chi], halo]))
    assert np.allclose(got, dense @ u, rtol=0, atol=1e-12)


@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.too_slow,
                                                                   HealthCheck.function_scoped_fixture])
@given(sparse_matrix(), st.integers(0, 2 ** 31 - 1))
def test_file_formats_round_trip(tmp_path_factory, m, seed):
    nr, nc, ptr, col, val, dense = m
    d = tmp_path_factory.mktemp("io")
    # MatrixMarket keeps 20 significant digits: exact for doubles
    pm = str(d / "a.mtx")
    bio.write_mm(pm, nc, ptr, col, val)
    n2, m2, p2, c2, v2 = bio.read_mm(pm)
    assert (n2, m2) == (nr, nc)
    assert np.array_equal(p2, ptr) and np.array_equal(c2, col) and np.array_equal(v2, val)
    # a random strip of rows
    rng = np.random.default_rng(seed)
    lo = int(rng.integers(0, nr + 1))
    hi = int(rng.integers(lo, nr + 1))
    n3, m3, p3, c3, v3 = bio.read_mm(pm, rows=(lo, hi))
    assert n3 == hi - lo and np.array_equal(p3, ptr[lo:hi + 1] - ptr[lo])
    assert np.array_equal(c3, col[ptr[lo]:ptr[hi]]) and np.array_equal(v3, val[ptr[lo]:ptr[hi]])
    pb = str(d / "a.bin")
    bio.write_crs_binary(pb, ptr, col, val)
    n4, p4, c4, v4 = bio.read_crs_binary(pb, rows=(lo, hi))
    assert n4 == hi - lo and np.array_equal(p4, p3) and np.array_equal(c4, c3) and np.array_equal(v4, v3)
    pd = str(d / "d.bin")
    bio.write_dense_binary(pd, dense)
    assert np.array_equal(bio.read_dense_binary(pd, rows=(lo, hi)), dense[lo:hi])
    pdm = str(d / "d.mtx")
    bio.write_mm(pdm, dense)
    assert np.array_equal(bio.read_mm(pdm), dense)


# ---- column formats (patterns.cuh / offsets.cuh / window.cuh): whatever the matrix, a planner
# either declines or returns a format that reproduces every column ---------------------------

@st.composite
def structured_matrix(draw):
    """Entries on a few diagonals (structured-grid like), randomly thinned, possibly rectangular,
    or fully random columns."""
    nr = draw(st.integers(1, 700))
    nc = draw(st.integers(1, 900))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    kind = draw(st.sampled_from(["diagonals", "few_diagonals_full", "random"]))
    rng = np.random.default_rng(seed)
    if kind == "random":
        lens = rng.integers(0, 12, nr)
        ptr = np.zeros(nr + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        col = np.concatenate([np.sort(rng.integers(0, nc, k)) for k in lens] + [np.zeros(0, dtype=np.int64)])
        return nr, nc, ptr, col.astype(np.int64)
    ndiag = draw(st.integers(1, 40 if kind == "diagonals" else 6))
    offs = np.unique(rng.integers(-nr, nc, ndiag))
    keep = 1.0 if kind == "few_diagonals_full" else draw(st.sampled_from([0.3, 0.7, 0.95]))
    rows = np.repeat(np.arange(nr, dtype=np.int64), offs.size)
    cols = rows + np.tile(offs, nr)
    ok = (cols >= 0) & (cols < nc) & (rng.uniform(size=cols.size) < keep)
    rows, cols = rows[ok], cols[ok]
    ptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=nr), out=ptr[1:])
    return nr, nc, ptr, cols


@settings(**SETTINGS)
@given(structured_matrix())
def test_pattern_and_offset_formats_reproduce_the_columns(m):
    from test_offsets import decode as decode_offsets
    from test_patterns import decode as decode_patterns
    nr, nc, ptr, col = m
    rows = np.repeat(np.arange(nr, dtype=np.int64), np.diff(ptr))
    ndist = np.unique(col - rows).size if col.size else 0
    o = ab.offset_plan(nr, nc, ptr, col)
    if col.size == 0:
        assert o is None
    else:
        assert (o is not None) == (ndist <= 256)
    if o is not None:
        assert o["count"] == ndist and (decode_offsets(o, ptr) == col).all()
    p = ab.pattern_plan(nr, nc, ptr, col)
    if p is not None:
        assert p["count"] <= 256 and p["total"] <= 1024
        assert (decode_patterns(p, ptr) == col).all()
        # the patterns are distinct and together exactly as long as `total`
        pats = {tuple(p["off"][p["start"][k]:p["start"][k + 1]]) for k in range(p["count"])}
        assert len(pats) == p["count"] and p["start"][p["count"]] == p["total"]
    else:
        # declined: really more than 256 patterns, or more than 1024 offsets in all
        tuples = {tuple(col[ptr[r]:ptr[r + 1]] - r) for r in range(nr)}
        assert col.size == 0 or len(tuples) > 256 or sum(len(t) for t in tuples) > 1024


@settings(**SETTINGS)
@given(structured_matrix(), st.sampled_from([256, 512, 1400]), st.sampled_from([1, 2]))
def test_window_format_reproduces_the_columns(m, slot_cap, gap):
    from test_window import _decode
    nr, nc, ptr, col = m
    w = ab.window_plan(nr, nc, ptr, col, slot_cap=slot_cap, max_ratio=100000, gap=gap)
    if w is None:
        return                      # (declined: a quad of rows alone does not fit the caps)
    got, rows = _decode(w, nc)
    assert rows == nr and (got == col).all()
    assert w["max_slots"] <= slot_cap and w["max_runs"] <= 126
