// ref_builtin.cpp -- TEST INFRASTRUCTURE (oracle side), not product code.
//
// Compiles the REAL reference, in place, from /root/reference: AMGCL's builtin
// (OpenMP) backend driving make_solver<amg<builtin, smoothed_aggregation,
// damped_jacobi|spai0>, cg|bicgstab>.  It is the ground truth every parity test
// compares against, the pin for the C restatement in amg_oracle.c, and the CPU
// baseline of bench.py ("kind": "reference").  Built by oracle/Makefile into
// oracle/_ref/libamgcl_ref.so; no reference source is copied into this repo.
//
// A thin "recording" backend derived from builtin<double> intercepts the
// Backend::copy_matrix / copy_vector / create_solver calls amg makes while
// moving each level to the backend (amg.hpp:351-417), which is exactly the
// stream of operators our b200 backend receives; this exposes the hierarchy
// (A, P, R, smoother diagonal per level, coarsest matrix) to the tests without
// touching the reference's private members.
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>
#include <omp.h>

#include <amgcl/backend/builtin.hpp>
#include <amgcl/adapter/crs_tuple.hpp>
#include <amgcl/make_solver.hpp>
#include <amgcl/amg.hpp>
#include <amgcl/coarsening/smoothed_aggregation.hpp>
#include <amgcl/relaxation/damped_jacobi.hpp>
#include <amgcl/relaxation/spai0.hpp>
#include <amgcl/solver/cg.hpp>
#include <amgcl/solver/bicgstab.hpp>
#include <amgcl/relaxation/chebyshev.hpp>
#include <amgcl/relaxation/ilu0.hpp>
#include <amgcl/solver/gmres.hpp>
#include <amgcl/solver/bicgstabl.hpp>
#include <amgcl/solver/skyline_lu.hpp>
#include <amgcl/io/mm.hpp>
#include <amgcl/io/binary.hpp>
#include "tests/sample_problem.hpp"

namespace {

typedef amgcl::backend::builtin<double> Builtin;
typedef Builtin::matrix HostMatrix;
typedef amgcl::backend::numa_vector<double> HostVector;

struct Recorder {
    std::vector<std::shared_ptr<HostMatrix>> matrices;   // A0,P0,R0,A1,P1,R1,...
    std::vector<std::shared_ptr<HostVector>> diagonals;  // one per smoothed level
    std::shared_ptr<HostMatrix> coarse;                  // coarsest-level matrix
    std::shared_ptr<Builtin::direct_solver> coarse_solver;
};

struct RecBackend : Builtin {
    struct params {
        Recorder *rec;
        params(Recorder *rec = 0) : rec(rec) {}
    };
    static std::string name() { return "builtin(recording)"; }

    static std::shared_ptr<matrix> copy_matrix(std::shared_ptr<matrix> A, const params &p) {
        if (p.rec) p.rec->matrices.push_back(A);
        return A;
    }
    template <class T>
    static std::shared_ptr<amgcl::backend::numa_vector<T>>
    copy_vector(std::shared_ptr<amgcl::backend::numa_vector<T>> x, const params &p) {
        if (p.rec) p.rec->diagonals.push_back(x);
        return x;
    }
    template <class T>
    static std::shared_ptr<amgcl::backend::numa_vector<T>>
    copy_vector(const std::vector<T> &x, const params &) {
        return std::make_shared<amgcl::backend::numa_vector<T>>(x);
    }
    static std::shared_ptr<vector> create_vector(size_t size, const params &) {
        return std::make_shared<vector>(size);
    }
    static std::shared_ptr<direct_solver> create_solver(std::shared_ptr<matrix> A, const params &p) {
        auto s = std::make_shared<direct_solver>(*A);
        if (p.rec) { p.rec->coarse = A; p.rec->coarse_solver = s; }
        return s;
    }
};

thread_local std::string g_error;

struct SolverBase {
    virtual ~SolverBase() {}
    virtual std::tuple<size_t, double> solve(const HostVector &f, HostVector &x) = 0;
    virtual void apply_precond(const HostVector &f, HostVector &x) = 0;
    virtual std::string report() const = 0;
};

template <template <class> class Relax, template <class, class> class Krylov>
struct MixedImpl : SolverBase {
    // the reference's mixed-precision composition (tutorial/1.poisson3Db/poisson3Db.cpp:45-51)
    typedef amgcl::make_solver<
        amgcl::amg<amgcl::backend::builtin<float>, amgcl::coarsening::smoothed_aggregation, Relax>,
        Krylov<amgcl::backend::builtin<double>, amgcl::solver::detail::default_inner_product>
        > Solver;
    std::unique_ptr<Solver> S;
    MixedImpl(size_t n, const int64_t *ptr, const int64_t *col, const double *val, double tol,
              int maxiter, int coarse_enough)
    {
        typename Solver::params prm;
        prm.solver.tol = tol;
        prm.solver.maxiter = maxiter;
        if (coarse_enough >= 0) prm.precond.coarse_enough = coarse_enough;
        auto A = std::make_tuple(n,
                amgcl::make_iterator_range(ptr, ptr + n + 1),
                amgcl::make_iterator_range(col, col + ptr[n]),
                amgcl::make_iterator_range(val, val + ptr[n]));
        S.reset(new Solver(A, prm));
    }
    std::tuple<size_t, double> solve(const HostVector &f, HostVector &x) override { return (*S)(f, x); }
    void apply_precond(const HostVector &f, HostVector &x) override { S->precond().apply(f, x); }
    std::string report() const override { std::ostringstream os; os << *S; return os.str(); }
};

template <template <class> class Relax, template <class, class> class Krylov>
struct SolverImpl : SolverBase {
    typedef amgcl::make_solver<
        amgcl::amg<RecBackend, amgcl::coarsening::smoothed_aggregation, Relax>,
        Krylov<RecBackend, amgcl::solver::detail::default_inner_product>
        > Solver;
    std::unique_ptr<Solver> S;

    SolverImpl(size_t n, const int64_t *ptr, const int64_t *col, const double *val, double tol,
               int maxiter, int coarse_enough, const RecBackend::params &bprm)
    {
        typename Solver::params prm;
        prm.solver.tol = tol;
        prm.solver.maxiter = maxiter;
        if (coarse_enough >= 0) prm.precond.coarse_enough = coarse_enough;
        auto A = std::make_tuple(n,
                amgcl::make_iterator_range(ptr, ptr + n + 1),
                amgcl::make_iterator_range(col, col + ptr[n]),
                amgcl::make_iterator_range(val, val + ptr[n]));
        S.reset(new Solver(A, prm, bprm));
    }
    std::tuple<size_t, double> solve(const HostVector &f, HostVector &x) override { return (*S)(f, x); }
    void apply_precond(const HostVector &f, HostVector &x) override { S->precond().apply(f, x); }
    std::string report() const override { std::ostringstream os; os << *S; return os.str(); }
};

struct Handle {
    size_t n;
    Recorder rec;
    std::unique_ptr<SolverBase> solver;
};

// non-owning CRS view over caller arrays
HostMatrix view(int64_t n, int64_t m, const int64_t *ptr, const int64_t *col, const double *val) {
    HostMatrix A;
    A.nrows = n; A.ncols = m; A.nnz = ptr[n];
    A.ptr = const_cast<int64_t *>(reinterpret_cast<const int64_t *>(ptr));
    A.col = const_cast<int64_t *>(reinterpret_cast<const int64_t *>(col));
    A.val = const_cast<double *>(val);
    A.own_data = false;
    return A;
}

struct VecView {   // minimal builtin-vector concept over a raw pointer
    typedef double value_type;
    double *p; size_t n;
    size_t size() const { return n; }
    double &operator[](size_t i) { return p[i]; }
    const double &operator[](size_t i) const { return p[i]; }
    double *data() { return p; }
    const double *data() const { return p; }
};

} // namespace

namespace amgcl { namespace backend {
template <> struct is_builtin_vector<VecView> : std::true_type {};
}}

static_assert(sizeof(ptrdiff_t) == sizeof(int64_t), "64-bit host expected");

extern "C" {

const char *ref_last_error() { return g_error.c_str(); }
int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

int ref_create(int64_t n, const int64_t *ptr, const int64_t *col, const double *val, int relax,
               int krylov, double tol, int maxiter, int coarse_enough, void **out)
{
    try {
        std::unique_ptr<Handle> h(new Handle());
        h->n = (size_t)n;
        RecBackend::params bprm(&h->rec);
        using namespace amgcl;
        if (relax == 0 && krylov == 0)
            h->solver.reset(new SolverImpl<relaxation::damped_jacobi, solver::cg>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm));
        else if (relax == 0 && krylov == 1)
            h->solver.reset(new SolverImpl<relaxation::damped_jacobi, solver::bicgstab>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm));
        else if (relax == 1 && krylov == 0)
            h->solver.reset(new SolverImpl<relaxation::spai0, solver::cg>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm));
        else if (relax == 1 && krylov == 1)
            h->solver.reset(new SolverImpl<relaxation::spai0, solver::bicgstab>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm));
        else if (relax == 2 && krylov == 0)
            h->solver.reset(new SolverImpl<relaxation::chebyshev, solver::cg>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm));
        else if (relax == 0 && krylov == 2)
            h->solver.reset(new SolverImpl<relaxation::damped_jacobi, solver::gmres>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm));
        else if (relax == 1 && krylov == 3)
            h->solver.reset(new SolverImpl<relaxation::spai0, solver::bicgstabl>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm));
        // ILU(0) smoother: on a non-builtin backend (RecBackend here, backend::cuda / b200 on the
        // GPU) its triangular solves are damped Jacobi sweeps (relaxation/detail/ilu_solve.hpp:97-113)
        else if (relax == 3 && krylov == 1)
            h->solver.reset(new SolverImpl<relaxation::ilu0, solver::bicgstab>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm));
        else if (relax == 3 && krylov == 0)
            h->solver.reset(new SolverImpl<relaxation::ilu0, solver::cg>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm));
        else { g_error = "unknown relax/krylov selector"; return -1; }
        *out = h.release();
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

int ref_create_mixed(int64_t n, const int64_t *ptr, const int64_t *col, const double *val, int relax,
                     int krylov, double tol, int maxiter, int coarse_enough, void **out)
{
    try {
        std::unique_ptr<Handle> h(new Handle());
        h->n = (size_t)n;
        using namespace amgcl;
        if (relax == 0 && krylov == 0)
            h->solver.reset(new MixedImpl<relaxation::damped_jacobi, solver::cg>(n, ptr, col, val, tol, maxiter, coarse_enough));
        else if (relax == 1 && krylov == 1)
            h->solver.reset(new MixedImpl<relaxation::spai0, solver::bicgstab>(n, ptr, col, val, tol, maxiter, coarse_enough));
        else if (relax == 1 && krylov == 0)
            h->solver.reset(new MixedImpl<relaxation::spai0, solver::cg>(n, ptr, col, val, tol, maxiter, coarse_enough));
        else if (relax == 0 && krylov == 1)
            h->solver.reset(new MixedImpl<relaxation::damped_jacobi, solver::bicgstab>(n, ptr, col, val, tol, maxiter, coarse_enough));
        else { g_error = "unknown relax/krylov selector"; return -1; }
        *out = h.release();
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

void ref_destroy(void *handle) { delete static_cast<Handle *>(handle); }

int ref_solve(void *handle, const double *rhs, double *x, int64_t *iters, double *resid)
{
    Handle *h = static_cast<Handle *>(handle);
    try {
        HostVector f(rhs, rhs + h->n), xx(x, x + h->n);
        size_t it; double r;
        std::tie(it, r) = h->solver->solve(f, xx);
        std::memcpy(x, xx.data(), h->n * sizeof(double));
        *iters = (int64_t)it; *resid = r;
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

// As ref_solve, timing solve() alone ("time solve(rhs, x) only"): the
// host vectors are built -- first-touched in parallel by numa_vector's constructor -- before
// the clock starts, the copy back happens after it stops.
int ref_solve_timed(void *handle, const double *rhs, double *x, int64_t *iters, double *resid,
                    double *seconds)
{
    Handle *h = static_cast<Handle *>(handle);
    try {
        HostVector f(rhs, rhs + h->n), xx(x, x + h->n);
        size_t it; double r;
        const double t0 = omp_get_wtime();
        std::tie(it, r) = h->solver->solve(f, xx);
        *seconds = omp_get_wtime() - t0;
        std::memcpy(x, xx.data(), h->n * sizeof(double));
        *iters = (int64_t)it; *resid = r;
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

int ref_apply_precond(void *handle, const double *f_in, double *x)
{
    Handle *h = static_cast<Handle *>(handle);
    try {
        HostVector f(f_in, f_in + h->n), xx(h->n);
        h->solver->apply_precond(f, xx);
        std::memcpy(x, xx.data(), h->n * sizeof(double));
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

int64_t ref_report(void *handle, char *buf, int64_t size)
{
    Handle *h = static_cast<Handle *>(handle);
    const std::string s = h->solver->report();
    if (buf && size > 0) {
        const size_t m = std::min<size_t>(s.size(), (size_t)size - 1);
        std::memcpy(buf, s.data(), m);
        buf[m] = 0;
    }
    return (int64_t)s.size() + 1;
}

// ---- hierarchy introspection ------------------------------------------------
// Smoothed levels: 0 .. ref_nlevels-2 carry (A, P, R, diag); the last level is
// the coarsest one and carries only A (solved directly).
int ref_nlevels(void *handle)
{
    Handle *h = static_cast<Handle *>(handle);
    return (int)h->rec.diagonals.size() + (h->rec.coarse ? 1 : 0);
}

static std::shared_ptr<HostMatrix> pick(Handle *h, int lvl, int which)
{
    const int smoothed = (int)h->rec.diagonals.size();
    if (lvl < smoothed) {
        // matrices: A0 P0 R0 A1 P1 R1 ...; the last smoothed level may lack P,R
        // when the hierarchy ends without a direct coarse solve
        const size_t idx = (size_t)lvl * 3 + which;
        if (idx < h->rec.matrices.size()) return h->rec.matrices[idx];
        return std::shared_ptr<HostMatrix>();
    }
    if (lvl == smoothed && which == 0) return h->rec.coarse;
    return std::shared_ptr<HostMatrix>();
}

int ref_level_info(void *handle, int lvl, int which, int64_t *rows, int64_t *cols, int64_t *nnz)
{
    auto A = pick(static_cast<Handle *>(handle), lvl, which);
    if (!A) { g_error = "no such level operator"; return -1; }
    *rows = A->nrows; *cols = A->ncols; *nnz = A->nnz;
    return 0;
}

int ref_level_matrix(void *handle, int lvl, int which, int64_t *ptr, int64_t *col, double *val)
{
    auto A = pick(static_cast<Handle *>(handle), lvl, which);
    if (!A) { g_error = "no such level operator"; return -1; }
    for (size_t i = 0; i <= A->nrows; ++i) ptr[i] = A->ptr[i];
    for (size_t e = 0; e < A->nnz; ++e) { col[e] = A->col[e]; val[e] = A->val[e]; }
    return 0;
}

int ref_level_diag(void *handle, int lvl, double *d)
{
    Handle *h = static_cast<Handle *>(handle);
    if (lvl < 0 || lvl >= (int)h->rec.diagonals.size()) { g_error = "no such level"; return -1; }
    const HostVector &v = *h->rec.diagonals[lvl];
    std::memcpy(d, v.data(), v.size() * sizeof(double));
    return 0;
}

// coarsest-level direct solve with the reference's skyline LU (skyline_lu.hpp:179-200)
int ref_coarse_solve(void *handle, const double *rhs, double *x)
{
    Handle *h = static_cast<Handle *>(handle);
    if (!h->rec.coarse_solver) { g_error = "hierarchy has no direct coarse solver"; return -1; }
    const size_t n = h->rec.coarse->nrows;
    HostVector f(rhs, rhs + n), xx(n);
    (*h->rec.coarse_solver)(f, xx);
    std::memcpy(x, xx.data(), n * sizeof(double));
    return 0;
}

// ---- the builtin backend's primitives on caller data -------------------------
// (backend/detail/matrix_ops.hpp:47-115, backend/builtin.hpp:1081-1321)
void ref_spmv(int64_t n, int64_t m, const int64_t *ptr, const int64_t *col, const double *val,
              double alpha, const double *x, double beta, double *y)
{
    HostMatrix A = view(n, m, ptr, col, val);
    VecView X{const_cast<double *>(x), (size_t)m}, Y{y, (size_t)n};
    amgcl::backend::spmv(alpha, A, X, beta, Y);
}

void ref_residual(int64_t n, int64_t m, const int64_t *ptr, const int64_t *col, const double *val,
                  const double *f, const double *x, double *r)
{
    HostMatrix A = view(n, m, ptr, col, val);
    VecView F{const_cast<double *>(f), (size_t)n}, X{const_cast<double *>(x), (size_t)m}, R{r, (size_t)n};
    amgcl::backend::residual(F, A, X, R);
}

double ref_inner_product(int64_t n, const double *x, const double *y)
{
    VecView X{const_cast<double *>(x), (size_t)n}, Y{const_cast<double *>(y), (size_t)n};
    return amgcl::backend::inner_product(X, Y);
}

void ref_axpby(int64_t n, double a, const double *x, double b, double *y)
{
    VecView X{const_cast<double *>(x), (size_t)n}, Y{y, (size_t)n};
    amgcl::backend::axpby(a, X, b, Y);
}

void ref_axpbypcz(int64_t n, double a, const double *x, double b, const double *y, double c, double *z)
{
    VecView X{const_cast<double *>(x), (size_t)n}, Y{const_cast<double *>(y), (size_t)n}, Z{z, (size_t)n};
    amgcl::backend::axpbypcz(a, X, b, Y, c, Z);
}

void ref_vmul(int64_t n, double a, const double *x, const double *y, double b, double *z)
{
    VecView X{const_cast<double *>(x), (size_t)n}, Y{const_cast<double *>(y), (size_t)n}, Z{z, (size_t)n};
    amgcl::backend::vmul(a, X, Y, b, Z);
}

// smoother diagonals straight from the reference constructors
// (damped_jacobi.hpp:92 -> builtin.hpp:753-773 ; spai0.hpp:60-82)
void ref_relax_diag(int64_t n, const int64_t *ptr, const int64_t *col, const double *val, int relax, double *d)
{
    HostMatrix A = view(n, n, ptr, col, val);
    if (relax == 0) {
        auto dia = amgcl::backend::diagonal(A, true);
        std::memcpy(d, dia->data(), (size_t)n * sizeof(double));
    } else {
        amgcl::relaxation::spai0<Builtin> S(A, amgcl::relaxation::spai0<Builtin>::params(), Builtin::params());
        std::memcpy(d, S.M->data(), (size_t)n * sizeof(double));
    }
}

// ---- the reference's file readers / writers (io/mm.hpp, io/binary.hpp) --------------------
// Two-phase reads: call with null arrays for the sizes, then with buffers.
int ref_mm_read_crs(const char *path, int64_t row_beg, int64_t row_end, int64_t *rows, int64_t *cols,
                    int64_t *nnz, int64_t *ptr, int64_t *col, double *val)
{
    try {
        amgcl::io::mm_reader mm(path);
        if (!mm.is_sparse()) { g_error = "not a sparse file"; return -1; }
        std::vector<ptrdiff_t> p, c;
        std::vector<double> v;
        size_t n, m;
        std::tie(n, m) = mm(p, c, v, row_beg, row_end);
        *rows = (int64_t)n; *cols = (int64_t)m; *nnz = (int64_t)c.size();
        if (ptr) std::copy(p.begin(), p.end(), ptr);
        if (col) std::copy(c.begin(), c.end(), col);
        if (val) std::copy(v.begin(), v.end(), val);
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

int ref_mm_read_dense(const char *path, int64_t row_beg, int64_t row_end, int64_t *rows, int64_t *cols,
                      double *data)
{
    try {
        amgcl::io::mm_reader mm(path);
        if (mm.is_sparse()) { g_error = "not a dense file"; return -1; }
        std::vector<double> v;
        size_t n, m;
        std::tie(n, m) = mm(v, row_beg, row_end);
        *rows = (int64_t)n; *cols = (int64_t)m;
        if (data) std::copy(v.begin(), v.end(), data);
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

int ref_mm_write_crs(const char *path, int64_t n, int64_t m, const int64_t *ptr, const int64_t *col,
                     const double *val)
{
    try {
        HostMatrix A = view(n, m, ptr, col, val);
        amgcl::io::mm_write(path, A);
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

int ref_mm_write_dense(const char *path, const double *data, int64_t rows, int64_t cols)
{
    try {
        amgcl::io::mm_write(path, data, (size_t)rows, (size_t)cols);
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

int ref_bin_read_crs(const char *path, int64_t row_beg, int64_t row_end, int64_t *rows, int64_t *nnz,
                     int64_t *ptr, int64_t *col, double *val)
{
    try {
        size_t n;
        std::vector<ptrdiff_t> p, c;
        std::vector<double> v;
        amgcl::io::read_crs(path, n, p, c, v, row_beg, row_end);
        *rows = (int64_t)p.size() - 1; *nnz = (int64_t)c.size();
        if (ptr) std::copy(p.begin(), p.end(), ptr);
        if (col) std::copy(c.begin(), c.end(), col);
        if (val) std::copy(v.begin(), v.end(), val);
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

// the writer of examples/mm2bin.cpp:22-31 / :37-44
int ref_bin_write_crs(const char *path, int64_t n, const int64_t *ptr, const int64_t *col, const double *val)
{
    try {
        std::ofstream f(path, std::ios::binary);
        size_t rows = (size_t)n;
        std::vector<ptrdiff_t> p(ptr, ptr + n + 1), c(col, col + ptr[n]);
        std::vector<double> v(val, val + ptr[n]);
        bool ok = amgcl::io::write(f, rows) && amgcl::io::write(f, p) && amgcl::io::write(f, c) &&
                  amgcl::io::write(f, v);
        if (!ok) { g_error = "File I/O error"; return -1; }
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

int ref_bin_read_dense(const char *path, int64_t row_beg, int64_t row_end, int64_t *rows, int64_t *cols,
                       double *data)
{
    try {
        size_t n, m;
        std::vector<double> v;
        amgcl::io::read_dense(path, n, m, v, row_beg, row_end);
        *rows = (int64_t)(m ? v.size() / m : 0); *cols = (int64_t)m;
        if (data) std::copy(v.begin(), v.end(), data);
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

// the reference's own test-matrix generator (tests/sample_problem.hpp:11-82); two-phase
int64_t ref_sample_problem(int64_t n, double anisotropy, int64_t *ptr, int64_t *col, double *val, double *rhs)
{
    std::vector<double> v, f;
    std::vector<int64_t> c, p;
    sample_problem((ptrdiff_t)n, v, c, p, f, anisotropy);
    if (ptr) std::copy(p.begin(), p.end(), ptr);
    if (col) std::copy(c.begin(), c.end(), col);
    if (val) std::copy(v.begin(), v.end(), val);
    if (rhs) std::copy(f.begin(), f.end(), rhs);
    return (int64_t)c.size();
}

} // extern "C"
