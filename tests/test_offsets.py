"""CPU checks of the offset-indexed column format (amgcl_b200/csrc/offsets.cuh through
b200_offset_plan_i64): row + table[index] must reproduce every column."""
import numpy as np

import amgcl_b200 as ab


def diag_matrix(nr, nc, offsets, seed, keep=0.85):
    """Entries on the given diagonals (col - row in `offsets`), each kept with probability
    `keep` (ragged rows, some empty), columns sorted within a row."""
    rng = np.random.default_rng(seed)
    offsets = np.sort(np.asarray(offsets, dtype=np.int64))
    rows = np.repeat(np.arange(nr, dtype=np.int64), offsets.size)
    cols = rows + np.tile(offsets, nr)
    ok = (cols >= 0) & (cols < nc) & (rng.uniform(size=cols.size) < keep)
    rows, cols = rows[ok], cols[ok]
    ptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=nr), out=ptr[1:])
    val = rng.uniform(-1, 1, cols.size)
    return ptr, cols, val


def decode(o, ptr):
    rows = np.repeat(np.arange(ptr.size - 1, dtype=np.int64), np.diff(ptr))
    return rows + o["tab"][o["idx8"]].astype(np.int64)


def test_poisson_has_seven_offsets():
    ptr, col, val, rhs = ab.poisson3d(12)
    n = ptr.size - 1
    o = ab.offset_plan(n, n, ptr, col)
    assert o is not None and o["count"] == 7
    assert list(o["tab"][:7]) == [-144, -12, -1, 0, 1, 12, 144]
    assert (decode(o, ptr) == col).all()


def test_ragged_rows_and_rectangular_shapes():
    for nr, nc, offs in ((3001, 3001, [-700, -50, -1, 0, 1, 50, 700]),
                         (2000, 2600, list(range(-13, 14)) + [300, 600]),
                         (2600, 2000, [-600, -3, 0, 5, 9])):
        ptr, col, val = diag_matrix(nr, nc, offs, seed=nr)
        o = ab.offset_plan(nr, nc, ptr, col)
        assert o is not None and o["count"] <= len(offs)
        assert (np.diff(o["tab"][:o["count"]]) > 0).all()
        assert (decode(o, ptr) == col).all()


def test_exactly_256_offsets_qualify_and_257_do_not():
    nr = 1500
    for count, want in ((256, True), (257, False)):
        offs = np.arange(count) * 3 - 300
        ptr, col, val = diag_matrix(nr, nr + 600, offs, seed=count, keep=1.0)
        o = ab.offset_plan(nr, nr + 600, ptr, col)
        assert (o is not None) == want
        if o:
            assert o["count"] == 256 and (decode(o, ptr) == col).all()
