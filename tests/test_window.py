"""CPU checks of the windowed operator format (amgcl_b200/csrc/window.cuh through
b200_window_plan_i64): what the upload stores must reproduce the matrix's columns exactly."""
import numpy as np
import pytest

import amgcl_b200 as ab


def _decode(w, ncols):
    """Columns of every entry, rebuilt from the runs + 16-bit window slots; checks the format's
    invariants on the way."""
    blk, runs, c16 = w["blk"], w["runs"], w["col16"]
    first = runs[:, 0].astype(np.int64)
    length = (runs[:, 1] & 0xffff).astype(np.int64)
    slot = (runs[:, 1].astype(np.int64) & 0xffffffff) >> 16
    assert (first % 4 == 0).all() and (length >= 1).all() and (length <= 64).all()
    assert (first + length <= ncols).all()
    out = np.full(c16.size, -1, dtype=np.int64)
    rows_seen = 0
    for r0, r1, e0, e1, q0, q1 in blk:
        assert r0 == rows_seen and r0 % 4 == 0 and r1 > r0
        rows_seen = r1
        win = np.full(w["max_slots"] + 64, -1, dtype=np.int64)
        off = 0
        for q in range(q0, q1):
            assert slot[q] == off                      # runs are laid one after the other
            win[off:off + length[q]] = first[q] + np.arange(length[q])
            off += length[q]
        assert off <= w["max_slots"] and q1 - q0 <= w["max_runs"]
        out[e0:e1] = win[c16[e0:e1]]
    return out, rows_seen


@pytest.mark.parametrize("n", [12, 20])
def test_poisson_windows_reproduce_the_columns(n):
    ptr, col, val, rhs = ab.poisson3d(n)
    nr = ptr.size - 1
    w = ab.window_plan(nr, nr, ptr, col)
    assert w is not None
    got, rows = _decode(w, nr)
    assert rows == nr and (got == col).all()
    # a 7-point stencil block reads about five lines of x: far fewer slots than entries
    assert (w["runs"][:, 1] & 0xffff).sum() < 0.75 * col.size


def test_blocks_are_cut_until_their_window_fits():
    """Rows that gather from many scattered places: with a small cap the planner has to cut the
    default blocks; the result still reproduces every column."""
    rng = np.random.default_rng(3)
    nr, nc = 4000, 60000
    lens = rng.integers(3, 9, nr)
    ptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    base = (np.arange(nr) * (nc - 4000) // nr)
    col = np.concatenate([np.sort(rng.choice(np.arange(b, b + 4000, 8), k, replace=False)) for b, k in zip(base, lens)])
    plain = ab.window_plan(nr, nc, ptr, col, slot_cap=8000, max_ratio=1000)
    cut = ab.window_plan(nr, nc, ptr, col, slot_cap=512, max_ratio=1000)
    assert plain is not None and cut is not None
    assert cut["blk"].shape[0] > plain["blk"].shape[0] and cut["max_slots"] <= 512
    for w in (plain, cut):
        got, rows = _decode(w, nc)
        assert rows == nr and (got == col).all()


def test_operators_without_reuse_do_not_qualify():
    """A restriction-like operator (every row gathers its own set of columns): the windows
    would be as large as the entry list, the operator stays in the plain format."""
    rng = np.random.default_rng(5)
    nr, nc = 2000, 200000
    lens = np.full(nr, 30)
    ptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    col = np.sort(rng.integers(0, nc, (nr, 30)), axis=1).ravel()
    assert ab.window_plan(nr, nc, ptr, col) is None


def test_last_run_is_clamped_to_the_matrix_width():
    """ncols not a multiple of four: the run that covers the last sector must stop at ncols."""
    nr, nc = 8, 10
    ptr = np.arange(nr + 1, dtype=np.int64) * 2
    col = np.tile(np.array([0, 9]), nr).astype(np.int64)
    w = ab.window_plan(nr, nc, ptr, col, max_ratio=1000)
    assert w is not None
    got, rows = _decode(w, nc)
    assert rows == nr and (got == col).all()
