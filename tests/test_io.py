"""MatrixMarket / AMGCL-binary readers (amgcl_b200/io.py) against the reference's own
io/mm.hpp and io/binary.hpp (through oracle/_ref when it is available), scipy.io, and
committed fixtures written by the reference."""
import os

import numpy as np
import pytest

import amgcl_b200 as ab
from amgcl_b200 import io as bio
import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def random_crs(n, m, density, seed, sort=True):
    rng = np.random.default_rng(seed)
    rows = []
    ptr = [0]
    col, val = [], []
    for i in range(n):
        k = rng.binomial(m, density)
        c = rng.choice(m, size=k, replace=False)
        if sort:
            c.sort()
        col.extend(c.tolist())
        val.extend(rng.uniform(-1, 1, k).tolist())
        ptr.append(len(col))
    return (np.array(ptr, np.int64), np.array(col, np.int64), np.array(val, np.float64))


def same_crs(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


def test_mm_sparse_round_trip_and_scipy(tmp_path):
    import scipy.io
    import scipy.sparse as sp
    ptr, col, val = random_crs(37, 23, 0.2, 1)
    path = str(tmp_path / "a.mtx")
    bio.write_mm(path, 23, ptr, col, val)
    n, m, p2, c2, v2 = bio.read_mm(path)
    assert (n, m) == (37, 23) and same_crs((ptr, col, val), (p2, c2, v2))
    A = sp.csr_matrix(scipy.io.mmread(path))
    A.sort_indices()
    assert np.array_equal(A.indptr, ptr) and np.array_equal(A.indices, col) and np.array_equal(A.data, val)
    # a file written by scipy (different float formatting, comment lines)
    path2 = str(tmp_path / "b.mtx")
    scipy.io.mmwrite(path2, sp.csr_matrix((val, col, ptr), shape=(37, 23)), comment="written by scipy",
                     precision=17)
    n, m, p3, c3, v3 = bio.read_mm(path2)
    assert same_crs((ptr, col), (p3, c3)) and np.allclose(v3, val, rtol=1e-15, atol=0)


def test_mm_symmetric_storage_is_expanded(tmp_path):
    path = str(tmp_path / "s.mtx")
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n% lower triangle\n3 3 4\n"
                "1 1 2.0\n2 1 -1.0\n3 2 -1.5\n3 3 4.0\n")
    n, m, ptr, col, val = bio.read_mm(path)
    dense = np.zeros((3, 3))
    for i in range(3):
        dense[i, col[ptr[i]:ptr[i + 1]]] = val[ptr[i]:ptr[i + 1]]
    assert np.array_equal(dense, [[2, -1, 0], [-1, 0, -1.5], [0, -1.5, 4]])
    assert all(np.all(np.diff(col[ptr[i]:ptr[i + 1]]) > 0) for i in range(3))     # rows sorted
    # integer data, unsorted general entries, a row block
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n4 5 5\n"
                "4 5 7\n1 3 2\n1 1 1\n3 2 -3\n4 1 9\n")
    n, m, ptr, col, val = bio.read_mm(path)
    assert (n, m) == (4, 5) and ptr.tolist() == [0, 2, 2, 3, 5]
    assert col.tolist() == [0, 2, 1, 0, 4] and val.tolist() == [1, 2, -3, 9, 7]
    n, m, ptr, col, val = bio.read_mm(path, rows=(2, 4))
    assert (n, m) == (2, 5) and ptr.tolist() == [0, 1, 3] and col.tolist() == [1, 0, 4]


def test_mm_rejects_what_the_reference_rejects(tmp_path):
    path = str(tmp_path / "bad.mtx")
    for banner in ("%%MatrixMarket matrix coordinate pattern general",
                   "%%MatrixMarket matrix coordinate real hermitian",
                   "%%MatrixMarket vector coordinate real general",
                   "%MatrixMarket matrix coordinate real general"):
        with open(path, "w") as f:
            f.write(banner + "\n1 1 1\n1 1 1.0\n")
        with pytest.raises(ValueError):
            bio.read_mm(path)
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n")
    with pytest.raises(ValueError):
        bio.read_mm(path)


def test_mm_dense_and_binary_round_trips(tmp_path):
    rng = np.random.default_rng(2)
    a = rng.uniform(-1, 1, (7, 3))
    path = str(tmp_path / "d.mtx")
    bio.write_mm(path, a)
    assert np.array_equal(bio.read_mm(path), a)
    assert np.array_equal(bio.read_mm(path, rows=(2, 5)), a[2:5])
    v = rng.uniform(-1, 1, 11)
    bio.write_mm(path, v)
    assert np.array_equal(bio.read_mm(path), v.reshape(-1, 1))

    ptr, col, val = random_crs(50, 50, 0.1, 3)
    pb = str(tmp_path / "A.bin")
    bio.write_crs_binary(pb, ptr, col, val)
    assert os.path.getsize(pb) == 8 + 8 * ptr.size + 16 * col.size
    n, p2, c2, v2 = bio.read_crs_binary(pb)
    assert n == 50 and same_crs((ptr, col, val), (p2, c2, v2))
    n, p3, c3, v3 = bio.read_crs_binary(pb, rows=(10, 30))
    assert n == 20 and np.array_equal(p3, ptr[10:31] - ptr[10])
    assert np.array_equal(c3, col[ptr[10]:ptr[30]]) and np.array_equal(v3, val[ptr[10]:ptr[30]])
    # unsorted rows are sorted on the way in (binary.hpp:117-122)
    up, uc, uv = random_crs(20, 40, 0.3, 4, sort=False)
    bio.write_crs_binary(pb, up, uc, uv)
    n, p4, c4, v4 = bio.read_crs_binary(pb)
    for i in range(20):
        o = np.argsort(uc[up[i]:up[i + 1]], kind="stable")
        assert np.array_equal(c4[p4[i]:p4[i + 1]], uc[up[i]:up[i + 1]][o])
        assert np.array_equal(v4[p4[i]:p4[i + 1]], uv[up[i]:up[i + 1]][o])
    pd = str(tmp_path / "d.bin")
    bio.write_dense_binary(pd, a)
    assert np.array_equal(bio.read_dense_binary(pd), a)
    assert np.array_equal(bio.read_dense_binary(pd, rows=(1, 4)), a[1:4])
    with pytest.raises(ValueError):
        bio.read_dense_binary(pd, rows=(3, 9))


def test_committed_fixtures_written_by_the_reference():
    """tests/golden/io_* were written by the reference's mm_write / io::write
    (tests/golden/make_golden.py) from the 5^3 Poisson matrix."""
    ptr, col, val, rhs = ab.poisson3d(5)
    n, m, p, c, v = bio.read_mm(os.path.join(GOLD, "io_poisson5.mtx"))
    assert (n, m) == (125, 125) and same_crs((ptr, col, val), (p, c, v))
    n, p, c, v = bio.read_crs_binary(os.path.join(GOLD, "io_poisson5.bin"))
    assert n == 125 and same_crs((ptr, col, val), (p, c, v))
    x = bio.read_mm(os.path.join(GOLD, "io_vec5.mtx"))
    assert x.shape == (125, 1) and np.array_equal(x[:, 0], np.sin(np.arange(125.0)))


@pytest.mark.skipif(not oracle.have_ref(), reason="reference build (oracle/_ref) not available")
def test_against_the_reference_readers_and_writers(tmp_path):
    R = oracle.ref()
    ptr, col, val = random_crs(60, 45, 0.15, 7, sort=False)
    sp_, sc_, sv_ = ptr, *bio._sort_rows(ptr, col, val)
    # our writer -> reference reader ; reference writer -> our reader
    mine = str(tmp_path / "mine.mtx")
    bio.write_mm(mine, 45, ptr, col, val)
    n, m, p, c, v = R.mm_read_crs(mine)
    assert (n, m) == (60, 45) and same_crs((sp_, sc_, sv_), (p, c, v))
    theirs = str(tmp_path / "theirs.mtx")
    R.mm_write_crs(theirs, 45, ptr, col, val)
    assert same_crs(bio.read_mm(theirs)[2:], (sp_, sc_, sv_))
    # row strips with global columns, through both readers
    for rows in ((0, 60), (13, 41), (59, 60), (20, 20)):
        a = bio.read_mm(theirs, rows=rows)
        b = R.mm_read_crs(theirs, rows=rows)
        assert a[:2] == b[:2] and same_crs(a[2:], b[2:])
    # symmetric file through both
    sym = str(tmp_path / "sym.mtx")
    with open(sym, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n4 4 6\n"
                "1 1 4\n2 1 -1\n2 2 4\n4 2 -2\n3 3 4\n4 4 5\n")
    for rows in ((-1, -1), (1, 3)):
        a = bio.read_mm(sym, rows=None if rows[0] < 0 else rows)
        b = R.mm_read_crs(sym, rows=rows)
        assert a[:2] == b[:2] and same_crs(a[2:], b[2:])
    # dense
    rng = np.random.default_rng(8)
    d = rng.uniform(-1, 1, (9, 4))
    dm = str(tmp_path / "d.mtx")
    R.mm_write_dense(dm, d)
    assert np.array_equal(bio.read_mm(dm), d)
    bio.write_mm(dm, d)
    assert np.array_equal(R.mm_read_dense(dm), d)
    assert np.array_equal(R.mm_read_dense(dm, rows=(2, 7)), bio.read_mm(dm, rows=(2, 7)))
    # binary
    bn = str(tmp_path / "a.bin")
    R.bin_write_crs(bn, ptr, col, val)
    assert open(bn, "rb").read() == (bio.write_crs_binary(str(tmp_path / "b.bin"), ptr, col, val),
                                     open(str(tmp_path / "b.bin"), "rb").read())[1]
    for rows in ((-1, -1), (5, 22)):
        a = bio.read_crs_binary(bn, rows=None if rows[0] < 0 else rows)
        b = R.bin_read_crs(bn, rows=rows)
        assert a[0] == b[0] and same_crs(a[1:], b[1:])
    db = str(tmp_path / "d.bin")
    bio.write_dense_binary(db, d)
    assert np.array_equal(R.bin_read_dense(db), d)
    assert np.array_equal(R.bin_read_dense(db, rows=(3, 8)), d[3:8])
