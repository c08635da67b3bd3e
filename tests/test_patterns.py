"""CPU checks of the pattern-indexed row format (amgcl_b200/csrc/patterns.cuh through
b200_pattern_plan_i64): row + pattern[k] must reproduce the column of every entry."""
import numpy as np

import amgcl_b200 as ab
from test_offsets import diag_matrix


def decode(o, ptr):
    nr = ptr.size - 1
    lens = np.diff(ptr)
    rows = np.repeat(np.arange(nr, dtype=np.int64), lens)
    k = np.arange(ptr[-1]) - np.repeat(ptr[:-1], lens)
    return rows + o["off"][np.repeat(o["start"][o["pid"]].astype(np.int64), lens) + k]


def test_poisson_has_27_row_patterns():
    """Interior rows + every combination of truncated directions: 3^3 patterns, whatever n."""
    for n in (6, 12):
        ptr, col, val, rhs = ab.poisson3d(n)
        nr = ptr.size - 1
        o = ab.pattern_plan(nr, nr, ptr, col)
        assert o is not None and o["count"] == 27 and o["total"] == 135
        assert (decode(o, ptr) == col).all()
        assert o["start"][o["count"]] == o["total"]


def test_ragged_rows_empty_rows_and_rectangular_shapes():
    for nr, nc, offs, keep in ((3001, 3001, [-700, -50, -1, 0, 1, 50, 700], 0.85),
                               (2000, 2600, [0, 3, 4, 90, 300, 600], 0.4),
                               (2600, 2000, [-600, -3, 0, 5, 9], 1.0)):
        ptr, col, val = diag_matrix(nr, nc, offs, seed=nr, keep=keep)
        assert (np.diff(ptr) == 0).any() == (keep == 0.4)         # (the second case has empty rows)
        o = ab.pattern_plan(nr, nc, ptr, col)
        assert o is not None and o["count"] <= 2 ** len(offs)
        assert (decode(o, ptr) == col).all()


def test_too_many_patterns_do_not_qualify():
    ptr, col, val = diag_matrix(4000, 4000, list(range(-13, 14)), seed=1, keep=0.8)
    assert ab.pattern_plan(4000, 4000, ptr, col) is None
    # few patterns, but longer than the table
    ptr, col, val = diag_matrix(3000, 6000, list(range(0, 1100)), seed=2, keep=1.0)
    assert ab.pattern_plan(3000, 6000, ptr, col) is None


def test_a_ranks_block_of_the_partitioned_operator_still_qualifies():
    """Multi-GPU: a rank keeps whole rows of the finest operator with the halo columns
    renumbered to [n_loc + owner * slots + position) (dist.cuh).  A face of the slab maps to one
    constant offset, so the block has a few more patterns than the whole operator, not many."""
    n = 16
    ptr, col, val, rhs = ab.poisson3d(n)
    nr = ptr.size - 1
    for nranks in (2, 4, 8):
        for rank in range(nranks):
            sp = ab.dist_split("halo", nranks, rank, nr, nr, ptr, col, val)
            p = ab.pattern_plan(sp["nrows"], sp["ncols"], sp["ptr"], sp["col"])
            assert p is not None and p["count"] <= 2 * 27 and p["total"] <= 1024
            assert (decode(p, sp["ptr"]) == sp["col"]).all()
            o = ab.offset_plan(sp["nrows"], sp["ncols"], sp["ptr"], sp["col"])
            assert o is not None and o["count"] <= 9
