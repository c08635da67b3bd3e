"""Readers / writers for the two file formats the reference's examples feed to a solver
(SURVEY section 8(f) rank 4): MatrixMarket text (amgcl/io/mm.hpp) and AMGCL's raw binary
CRS / dense files (amgcl/io/binary.hpp, produced by examples/mm2bin.cpp).

Host-side only: the result is the (ptr, col, val) triple ``DropinSolver`` and
``Context.csr`` take.  Behaviour mirrors the reference readers:

* MatrixMarket (mm.hpp:52-340): ``matrix coordinate|array  real|integer  general|symmetric``;
  complex and pattern files are rejected like the reference rejects what it does not know;
  symmetric storage is expanded (mm.hpp:201-206); duplicates are kept as separate entries;
  every row is sorted by column (mm.hpp:233); ``rows=(beg, end)`` reads a block of rows with
  row numbers rebased to the block and GLOBAL column numbers (mm.hpp:134-135, the MPI
  examples' strip reader).  Dense arrays are column-major in the file and row-major in memory
  (mm.hpp:263-300).
* Binary CRS (binary.hpp:69-123; writer examples/mm2bin.cpp:22-31): ``size_t rows``,
  ``ptrdiff_t ptr[rows+1]``, ``ptrdiff_t col[nnz]``, ``double val[nnz]``, native byte order.
* Binary dense (binary.hpp:125-157; mm2bin.cpp:37-44): ``size_t n, m``, ``double v[n*m]`` row-major.
"""
import numpy as np

__all__ = ["read_mm", "write_mm", "read_crs_binary", "write_crs_binary",
           "read_dense_binary", "write_dense_binary"]


class FormatError(ValueError):
    pass


def _sort_rows(ptr, col, val):
    """Sort every row by column (stable, so duplicate entries keep their file order)."""
    nrows = ptr.size - 1
    row = np.repeat(np.arange(nrows, dtype=np.int64), np.diff(ptr))
    order = np.lexsort((col, row))
    return col[order], val[order]


def _coo_to_crs(nrows, row, col, val):
    counts = np.bincount(row, minlength=nrows).astype(np.int64)
    ptr = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    order = np.lexsort((col, row))
    return ptr, col[order].astype(np.int64), val[order].astype(np.float64)


def read_mm(path, rows=None):
    """Read a MatrixMarket file.

    Sparse (``coordinate``) files return ``(nrows, ncols, ptr, col, val)`` with int64 indices;
    dense (``array``) files return a float64 ndarray of shape (nrows, ncols).
    ``rows=(beg, end)`` restricts either to that block of rows."""
    with open(path, "r") as f:
        banner = f.readline().split()
        if len(banner) < 5 or banner[0] != "%%MatrixMarket":
            raise FormatError("MatrixMarket format error (no banner)")
        if banner[1].lower() != "matrix":
            raise FormatError("MatrixMarket format error (not a matrix)")
        coord, dtype, storage = (s.lower() for s in banner[2:5])
        if storage not in ("general", "symmetric"):
            raise FormatError("unsupported storage type")
        if coord not in ("coordinate", "array"):
            raise FormatError("MatrixMarket format error (unsupported coordinate type)")
        if dtype not in ("real", "integer"):
            raise FormatError("unsupported data type: " + dtype)
        symmetric = storage == "symmetric"
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        if not line:
            raise FormatError("MatrixMarket format error (unexpected eof)")
        sizes = line.split()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")            # "input contained no data" for empty matrices
            body = np.loadtxt(f, dtype=np.float64, ndmin=2, comments="%")

    if coord == "array":
        nrows, ncols = int(sizes[0]), int(sizes[1])
        if symmetric:
            raise FormatError("symmetric dense arrays are not supported")
        data = body.reshape(-1)
        if data.size != nrows * ncols:
            raise FormatError("MatrixMarket format error (wrong number of values)")
        dense = data.reshape(ncols, nrows).T.copy()          # file is column-major
        if rows is not None:
            beg, end = _row_range(rows, nrows)
            dense = dense[beg:end].copy()
        return dense

    nrows, ncols, nnz = int(sizes[0]), int(sizes[1]), int(sizes[2])
    if body.size == 0:
        body = body.reshape(0, 3)
    if body.shape[0] != nnz or body.shape[1] != 3:
        raise FormatError("MatrixMarket format error (entry count / columns)")
    i = body[:, 0].astype(np.int64) - 1
    j = body[:, 1].astype(np.int64) - 1
    v = body[:, 2]
    if nnz and (i.min() < 0 or j.min() < 0 or i.max() >= nrows or j.max() >= ncols):
        raise FormatError("MatrixMarket format error (index out of range)")
    if symmetric:
        off = i != j
        i, j, v = np.concatenate([i, j[off]]), np.concatenate([j, i[off]]), np.concatenate([v, v[off]])
    beg, end = _row_range(rows, nrows)
    if (beg, end) != (0, nrows):
        keep = (i >= beg) & (i < end)
        i, j, v = i[keep] - beg, j[keep], v[keep]
    ptr, col, val = _coo_to_crs(end - beg, i, j, v)
    return end - beg, ncols, ptr, col, val


def _row_range(rows, nrows):
    if rows is None:
        return 0, nrows
    beg, end = rows
    beg = 0 if beg is None or beg < 0 else int(beg)
    end = nrows if end is None or end < 0 else int(end)
    if not (0 <= beg <= end <= nrows):
        raise ValueError("Wrong subset of rows is requested")
    return beg, end


def write_mm(path, *args):
    """``write_mm(path, dense)`` (1-D or 2-D array) or ``write_mm(path, ncols, ptr, col, val)``;
    values are written with the reference's 20 significant digits (mm.hpp:340-343)."""
    if len(args) == 1:
        a = np.asarray(args[0], dtype=np.float64)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        with open(path, "w") as f:
            f.write("%%MatrixMarket matrix array real general\n")
            f.write("%d %d\n" % a.shape)
            np.savetxt(f, a.T.reshape(-1), fmt="%.20e")
        return
    ncols, ptr, col, val = args
    ptr = np.asarray(ptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.asarray(val, dtype=np.float64)
    nrows = ptr.size - 1
    row = np.repeat(np.arange(nrows, dtype=np.int64), np.diff(ptr))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write("%d %d %d\n" % (nrows, int(ncols), col.size))
        for r, c, v in zip(row + 1, col + 1, val):
            f.write("%d %d %.20e\n" % (r, c, v))


def read_crs_binary(path, rows=None):
    """``(nrows, ptr, col, val)`` from an AMGCL binary CRS file; ``rows=(beg, end)`` reads a strip
    (pointer array rebased to 0, global columns), exactly binary.hpp:69-123."""
    with open(path, "rb") as f:
        head = np.fromfile(f, dtype=np.uint64, count=1)
        if head.size != 1:
            raise FormatError("File I/O error")
        n = int(head[0])
        beg, end = _row_range(rows, n)
        f.seek(8 + beg * 8)
        ptr = np.fromfile(f, dtype=np.int64, count=end - beg + 1)
        f.seek(8 + n * 8)
        total = np.fromfile(f, dtype=np.int64, count=1)
        if ptr.size != end - beg + 1 or total.size != 1:
            raise FormatError("File I/O error")
        nnz = int(total[0])
        first = int(ptr[0])
        ptr = ptr - first
        cnt = int(ptr[-1])
        col_beg = 8 + (n + 1) * 8
        f.seek(col_beg + first * 8)
        col = np.fromfile(f, dtype=np.int64, count=cnt)
        f.seek(col_beg + nnz * 8 + first * 8)
        val = np.fromfile(f, dtype=np.float64, count=cnt)
        if col.size != cnt or val.size != cnt:
            raise FormatError("File I/O error")
    col, val = _sort_rows(ptr, col, val)
    return end - beg, ptr, col, val


def write_crs_binary(path, ptr, col, val):
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    with open(path, "wb") as f:
        np.array([ptr.size - 1], dtype=np.uint64).tofile(f)
        ptr.tofile(f)
        np.ascontiguousarray(col, dtype=np.int64).tofile(f)
        np.ascontiguousarray(val, dtype=np.float64).tofile(f)


def read_dense_binary(path, rows=None):
    with open(path, "rb") as f:
        head = np.fromfile(f, dtype=np.uint64, count=2)
        if head.size != 2:
            raise FormatError("File I/O error")
        n, m = int(head[0]), int(head[1])
        beg, end = _row_range(rows, n)
        f.seek(16 + beg * m * 8)
        v = np.fromfile(f, dtype=np.float64, count=(end - beg) * m)
        if v.size != (end - beg) * m:
            raise FormatError("File I/O error")
    return v.reshape(end - beg, m)


def write_dense_binary(path, a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    with open(path, "wb") as f:
        np.array(a.shape, dtype=np.uint64).tofile(f)
        a.tofile(f)
