// dropin.cpp -- the drop-in proof: UNMODIFIED amgcl::make_solver / amg / cg /
// bicgstab (compiled from the reference headers) instantiated on
// amgcl::backend::b200<double>.  Everything numerical at solve time runs in
// libamgcl_b200.so; AMGCL contributes the host-side setup (smoothed
// aggregation coarsening, Galerkin products) and the host control flow of the
// V-cycle and the Krylov loop, exactly as with its own cuda backend
// (tutorial/1.poisson3Db/poisson3Db_cuda.cu:51-87).
//
// Exposed as a small C API so bench.py / tests can drive it through ctypes.
// Build: see amgcl_b200/build.py (needs the AMGCL headers on the include path;
// the resulting .so is self-contained apart from libamgcl_b200.so).
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include <omp.h>

#include <amgcl/backend/b200.hpp>
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

namespace {

typedef amgcl::backend::b200<double> Backend;

thread_local std::string g_error;

struct SolverBase {
    virtual ~SolverBase() {}
    virtual std::tuple<size_t, double> solve(const Backend::vector &f, Backend::vector &x) = 0;
    virtual void apply_precond(const Backend::vector &f, Backend::vector &x) = 0;
    virtual std::string report() const = 0;
    virtual size_t bytes() const = 0;
    virtual void graph_stats(size_t &ngraphs, size_t &kernels, size_t &replays) const {
        ngraphs = kernels = replays = 0;
    }
};

typedef amgcl::backend::b200<float> BackendF32;   // hierarchy of a mixed-precision solver

// Graph = true wraps the hierarchy in preconditioner::b200_cycle_graph (one CUDA graph launch
// per V-cycle); Graph = false is the unmodified reference composition.
template <class AMG, bool Graph> struct precond_of { typedef AMG type; };
template <class AMG> struct precond_of<AMG, true> { typedef amgcl::preconditioner::b200_cycle_graph<AMG> type; };

template <class P> void stats_of(const P &, size_t &g, size_t &k, size_t &r) { g = k = r = 0; }
template <class AMG> void stats_of(const amgcl::preconditioner::b200_cycle_graph<AMG> &p,
                                   size_t &g, size_t &k, size_t &r) { p.graph_stats(g, k, r); }

template <template <class> class Relax, template <class, class> class Krylov, class PBackend = Backend,
          bool Graph = false>
struct SolverImpl : SolverBase {
    typedef amgcl::amg<PBackend, amgcl::coarsening::smoothed_aggregation, Relax> AMG;
    typedef amgcl::make_solver<
        typename precond_of<AMG, Graph>::type,
        Krylov<Backend, amgcl::solver::detail::default_inner_product>
        > Solver;

    std::unique_ptr<Solver> S;

    SolverImpl(size_t n, const int64_t *ptr, const int64_t *col, const double *val,
               double tol, int maxiter, int coarse_enough, const Backend::params &bprm)
    {
        typename Solver::params prm;
        prm.solver.tol = tol;
        prm.solver.maxiter = maxiter;
        if (coarse_enough >= 0) prm.precond.coarse_enough = coarse_enough;
        auto A = std::make_tuple(
                n,
                amgcl::make_iterator_range(ptr, ptr + n + 1),
                amgcl::make_iterator_range(col, col + ptr[n]),
                amgcl::make_iterator_range(val, val + ptr[n]));
        S.reset(new Solver(A, prm, bprm));
    }

    std::tuple<size_t, double> solve(const Backend::vector &f, Backend::vector &x) override {
        return (*S)(f, x);
    }
    void apply_precond(const Backend::vector &f, Backend::vector &x) override {
        S->precond().apply(f, x);
    }
    std::string report() const override {
        std::ostringstream os;
        os << *S;
        return os.str();
    }
    size_t bytes() const override { return amgcl::backend::bytes(*S); }
    void graph_stats(size_t &g, size_t &k, size_t &r) const override { stats_of(S->precond(), g, k, r); }
};

// the four BASELINE.json combinations with the V-cycle replayed as a CUDA graph
template <class PBackend>
SolverBase *make_graphed(int relax, int krylov, size_t n, const int64_t *ptr, const int64_t *col,
                         const double *val, double tol, int maxiter, int coarse_enough,
                         const Backend::params &bprm) {
    using namespace amgcl;
    if (relax == 0 && krylov == 0)
        return new SolverImpl<relaxation::damped_jacobi, solver::cg, PBackend, true>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm);
    if (relax == 0 && krylov == 1)
        return new SolverImpl<relaxation::damped_jacobi, solver::bicgstab, PBackend, true>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm);
    if (relax == 1 && krylov == 0)
        return new SolverImpl<relaxation::spai0, solver::cg, PBackend, true>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm);
    if (relax == 1 && krylov == 1)
        return new SolverImpl<relaxation::spai0, solver::bicgstab, PBackend, true>(n, ptr, col, val, tol, maxiter, coarse_enough, bprm);
    return nullptr;
}

struct Handle {
    size_t n;
    Backend::params bprm;
    std::unique_ptr<SolverBase> solver;
    std::shared_ptr<Backend::vector> f, x;
};

} // namespace

extern "C" {

const char *dropin_last_error() { return g_error.c_str(); }

// OpenMP threads used by AMGCL's host-side setup (launchers such as torchrun export
// OMP_NUM_THREADS=1, which would serialise the coarsening).
void dropin_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int dropin_num_threads() { return omp_get_max_threads(); }

// relax: 0 = damped_jacobi, 1 = spai0 ; krylov: 0 = cg, 1 = bicgstab
// ctx: a b200_ctx_t or NULL for the library default
int dropin_create(void *ctx, int64_t n, const int64_t *ptr, const int64_t *col, const double *val,
                  int relax, int krylov, double tol, int maxiter, int coarse_enough, void **out)
{
    try {
        std::unique_ptr<Handle> h(new Handle());
        h->n = (size_t)n;
        h->bprm = Backend::params(static_cast<b200_ctx_t>(ctx));
        using namespace amgcl;
        if (relax == 0 && krylov == 0)
            h->solver.reset(new SolverImpl<relaxation::damped_jacobi, solver::cg>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else if (relax == 0 && krylov == 1)
            h->solver.reset(new SolverImpl<relaxation::damped_jacobi, solver::bicgstab>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else if (relax == 1 && krylov == 0)
            h->solver.reset(new SolverImpl<relaxation::spai0, solver::cg>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else if (relax == 1 && krylov == 1)
            h->solver.reset(new SolverImpl<relaxation::spai0, solver::bicgstab>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        // combinations that run on the backend purely through its primitives
        else if (relax == 2 && krylov == 0)
            h->solver.reset(new SolverImpl<relaxation::chebyshev, solver::cg>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else if (relax == 0 && krylov == 2)
            h->solver.reset(new SolverImpl<relaxation::damped_jacobi, solver::gmres>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else if (relax == 1 && krylov == 3)
            h->solver.reset(new SolverImpl<relaxation::spai0, solver::bicgstabl>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        // ILU(0): setup on the host, triangular solves as damped Jacobi sweeps built from
        // residual / axpby / vmul (relaxation/detail/ilu_solve.hpp:97-113)
        else if (relax == 3 && krylov == 1)
            h->solver.reset(new SolverImpl<relaxation::ilu0, solver::bicgstab>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else if (relax == 3 && krylov == 0)
            h->solver.reset(new SolverImpl<relaxation::ilu0, solver::cg>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else {
            g_error = "unknown relax/krylov selector";
            return -1;
        }
        h->f = Backend::create_vector(n, h->bprm);
        h->x = Backend::create_vector(n, h->bprm);
        *out = h.release();
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// Mixed precision (tutorial/1.poisson3Db/poisson3Db.cpp:45-51): FP32 hierarchy
// (amg<backend::b200<float>>) under an FP64 Krylov solver (backend::b200<double>).
int dropin_create_mixed(void *ctx, int64_t n, const int64_t *ptr, const int64_t *col, const double *val,
                        int relax, int krylov, double tol, int maxiter, int coarse_enough, void **out)
{
    try {
        std::unique_ptr<Handle> h(new Handle());
        h->n = (size_t)n;
        h->bprm = Backend::params(static_cast<b200_ctx_t>(ctx));
        using namespace amgcl;
        if (relax == 0 && krylov == 0)
            h->solver.reset(new SolverImpl<relaxation::damped_jacobi, solver::cg, BackendF32>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else if (relax == 1 && krylov == 1)
            h->solver.reset(new SolverImpl<relaxation::spai0, solver::bicgstab, BackendF32>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else if (relax == 1 && krylov == 0)
            h->solver.reset(new SolverImpl<relaxation::spai0, solver::cg, BackendF32>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else if (relax == 0 && krylov == 1)
            h->solver.reset(new SolverImpl<relaxation::damped_jacobi, solver::bicgstab, BackendF32>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm));
        else {
            g_error = "unknown relax/krylov selector";
            return -1;
        }
        h->f = Backend::create_vector(n, h->bprm);
        h->x = Backend::create_vector(n, h->bprm);
        *out = h.release();
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// As dropin_create / dropin_create_mixed (mixed != 0), with the hierarchy wrapped in
// amgcl::preconditioner::b200_cycle_graph.  GMRES exercises many (rhs, x) pairs per solve.
int dropin_create_graph(void *ctx, int64_t n, const int64_t *ptr, const int64_t *col, const double *val,
                        int relax, int krylov, int mixed, double tol, int maxiter, int coarse_enough,
                        void **out)
{
    try {
        std::unique_ptr<Handle> h(new Handle());
        h->n = (size_t)n;
        h->bprm = Backend::params(static_cast<b200_ctx_t>(ctx));
        SolverBase *sb = nullptr;
        if (!mixed && relax == 0 && krylov == 2)
            sb = new SolverImpl<amgcl::relaxation::damped_jacobi, amgcl::solver::gmres, Backend, true>(n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm);
        else if (mixed)
            sb = make_graphed<BackendF32>(relax, krylov, n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm);
        else
            sb = make_graphed<Backend>(relax, krylov, n, ptr, col, val, tol, maxiter, coarse_enough, h->bprm);
        if (!sb) {
            g_error = "unknown relax/krylov selector";
            return -1;
        }
        h->solver.reset(sb);
        h->f = Backend::create_vector(n, h->bprm);
        h->x = Backend::create_vector(n, h->bprm);
        *out = h.release();
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// recorded graphs, kernels per replay (first graph), replays so far
int dropin_graph_stats(void *handle, int64_t *ngraphs, int64_t *kernels, int64_t *replays)
{
    Handle *h = static_cast<Handle *>(handle);
    size_t g, k, r;
    h->solver->graph_stats(g, k, r);
    if (ngraphs) *ngraphs = (int64_t)g;
    if (kernels) *kernels = (int64_t)k;
    if (replays) *replays = (int64_t)r;
    return 0;
}

void dropin_destroy(void *handle) { delete static_cast<Handle *>(handle); }

// End-to-end call with HOST buffers: H2D(rhs), solve from x0 = x_host (in/out), D2H(x).
int dropin_solve(void *handle, const double *rhs_host, double *x_host, int64_t *iters, double *resid)
{
    Handle *h = static_cast<Handle *>(handle);
    try {
        h->f->upload(rhs_host);
        h->x->upload(x_host);
        size_t it; double r;
        std::tie(it, r) = h->solver->solve(*h->f, *h->x);
        h->x->download(x_host);
        *iters = (int64_t)it; *resid = r;
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// End-to-end call as the reference's tutorial does it (poisson3Db_cuda.cu:83-87): the
// right-hand side comes from the host, the initial guess x0 = 0 is created on the device,
// the solution goes back to the host.  On a multi-GPU context every rank moves the rows it owns
// (rhs rows in, solution rows into their place in x_host), like a row-distributed MPI program.
int dropin_solve_zero_guess(void *handle, const double *rhs_host, double *x_host, int64_t *iters,
                            double *resid)
{
    Handle *h = static_cast<Handle *>(handle);
    try {
        h->f->upload(rhs_host);
        amgcl::backend::clear(*h->x);
        size_t it; double r;
        std::tie(it, r) = h->solver->solve(*h->f, *h->x);
        h->x->download_local(x_host);
        *iters = (int64_t)it; *resid = r;
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// Device-resident variant: rhs uploaded beforehand, x0 = 0, x stays on the device.
int dropin_upload_rhs(void *handle, const double *rhs_host)
{
    Handle *h = static_cast<Handle *>(handle);
    try { h->f->upload(rhs_host); return 0; }
    catch (const std::exception &e) { g_error = e.what(); return -1; }
}

int dropin_solve_resident(void *handle, int64_t *iters, double *resid)
{
    Handle *h = static_cast<Handle *>(handle);
    try {
        amgcl::backend::clear(*h->x);
        size_t it; double r;
        std::tie(it, r) = h->solver->solve(*h->f, *h->x);
        *iters = (int64_t)it; *resid = r;
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

int dropin_download_x(void *handle, double *x_host)
{
    Handle *h = static_cast<Handle *>(handle);
    try { h->x->download(x_host); return 0; }
    catch (const std::exception &e) { g_error = e.what(); return -1; }
}

// One application of the AMG preconditioner x = M^-1 f (a single V-cycle), host buffers.
int dropin_apply_precond(void *handle, const double *f_host, double *x_host)
{
    Handle *h = static_cast<Handle *>(handle);
    try {
        h->f->upload(f_host);
        h->solver->apply_precond(*h->f, *h->x);
        h->x->download(x_host);
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// Text report of solver + hierarchy (amg.hpp:561-598).  Returns the needed size.
int64_t dropin_report(void *handle, char *buf, int64_t size)
{
    Handle *h = static_cast<Handle *>(handle);
    const std::string s = h->solver->report();
    if (buf && size > 0) {
        const size_t m = std::min<size_t>(s.size(), (size_t)size - 1);
        memcpy(buf, s.data(), m);
        buf[m] = 0;
    }
    return (int64_t)s.size() + 1;
}

int64_t dropin_bytes(void *handle) { return (int64_t)static_cast<Handle *>(handle)->solver->bytes(); }

// Backend::gather / Backend::scatter exercised through the C++ binding (cuda.hpp:548-577):
// g[k] = src[I[k]] on the device and straight to the host; then dst (size n, pre-filled with
// `fill`) receives dst[I[k]] = g[k].
int dropin_gather_scatter(void *ctx, int64_t n, const double *src_host, int64_t m, const int64_t *I,
                          double fill, double *gathered_dev_path, double *gathered_host_path,
                          double *scattered)
{
    try {
        Backend::params bprm(static_cast<b200_ctx_t>(ctx));
        std::vector<ptrdiff_t> idx(I, I + m);
        Backend::gather G((size_t)n, idx, bprm);
        Backend::scatter S((size_t)n, idx, bprm);
        Backend::vector src(src_host, (size_t)n, bprm), g((size_t)m, bprm);
        G(src, g);
        g.download(gathered_dev_path);
        std::vector<double> vals((size_t)m);
        G(src, vals);
        std::copy(vals.begin(), vals.end(), gathered_host_path);
        std::vector<double> init((size_t)n, fill);
        Backend::vector dst(init.data(), (size_t)n, bprm);
        S(g, dst);
        dst.download(scattered);
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

} // extern "C"
