// examples/poisson_b200.cpp -- the reference tutorial program
// (tutorial/1.poisson3Db/poisson3Db_cuda.cu:51-87) with the backend typedef switched to
// amgcl::backend::b200<double>.  Plain C++: no nvcc, no CUDA headers; the CUDA code lives in
// libamgcl_b200.so behind the C ABI.
//
//   g++ -std=c++17 -O2 -fopenmp -DAMGCL_NO_BOOST -I<amgcl> -I<repo>/include \
//       examples/poisson_b200.cpp -L<repo>/amgcl_b200/lib -lamgcl_b200 \
//       -Wl,-rpath,<repo>/amgcl_b200/lib -o poisson_b200
//   ./poisson_b200 [n=64] [cg|bicgstab] [graph]
//
// (tests/test_capi.py compiles and links it whenever the AMGCL headers are available.)
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <tuple>
#include <vector>

#include <amgcl/backend/b200.hpp>
#include <amgcl/adapter/crs_tuple.hpp>
#include <amgcl/make_solver.hpp>
#include <amgcl/amg.hpp>
#include <amgcl/coarsening/smoothed_aggregation.hpp>
#include <amgcl/relaxation/spai0.hpp>
#include <amgcl/solver/cg.hpp>
#include <amgcl/solver/bicgstab.hpp>

typedef amgcl::backend::b200<double> Backend;
typedef amgcl::amg<Backend, amgcl::coarsening::smoothed_aggregation, amgcl::relaxation::spai0> AMG;

// 7-point Poisson problem on an n^3 grid, as tests/sample_problem.hpp builds it
static size_t poisson(ptrdiff_t n, std::vector<ptrdiff_t> &ptr, std::vector<ptrdiff_t> &col,
                      std::vector<double> &val, std::vector<double> &rhs)
{
    const ptrdiff_t n3 = n * n * n;
    ptr.assign(1, 0);
    for (ptrdiff_t k = 0, idx = 0; k < n; ++k)
        for (ptrdiff_t j = 0; j < n; ++j)
            for (ptrdiff_t i = 0; i < n; ++i, ++idx) {
                if (k > 0)     { col.push_back(idx - n * n); val.push_back(-1.0); }
                if (j > 0)     { col.push_back(idx - n);     val.push_back(-1.0); }
                if (i > 0)     { col.push_back(idx - 1);     val.push_back(-1.0); }
                col.push_back(idx); val.push_back(6.0);
                if (i + 1 < n) { col.push_back(idx + 1);     val.push_back(-1.0); }
                if (j + 1 < n) { col.push_back(idx + n);     val.push_back(-1.0); }
                if (k + 1 < n) { col.push_back(idx + n * n); val.push_back(-1.0); }
                ptr.push_back((ptrdiff_t)col.size());
            }
    rhs.assign(n3, 1.0);
    return (size_t)n3;
}

template <class Solver>
static int run(size_t rows, const std::vector<ptrdiff_t> &ptr, const std::vector<ptrdiff_t> &col,
               const std::vector<double> &val, const std::vector<double> &rhs)
{
    Backend::params bprm;                       // default context on the current device
    Solver solve(std::tie(rows, ptr, col, val), typename Solver::params(), bprm);
    std::cout << solve << std::endl;

    auto f = Backend::copy_vector(rhs, bprm);
    auto x = Backend::create_vector(rows, bprm);

    size_t iters;
    double error;
    std::tie(iters, error) = solve(*f, *x);

    std::vector<double> x_host(rows);
    amgcl::backend::copy(*x, x_host);
    std::cout << "Iterations: " << iters << "\nError:      " << error
              << "\nx[0]:       " << x_host[0] << std::endl;
    return error < 1e-6 ? 0 : 1;
}

int main(int argc, char *argv[])
{
    const ptrdiff_t n = argc > 1 ? std::atol(argv[1]) : 64;
    const bool bicg  = argc > 2 && !std::strcmp(argv[2], "bicgstab");
    const bool graph = argc > 3 && !std::strcmp(argv[3], "graph");

    std::vector<ptrdiff_t> ptr, col;
    std::vector<double> val, rhs;
    const size_t rows = poisson(n, ptr, col, val, rhs);

    try {
        if (graph) {        // every V-cycle replayed as one CUDA graph launch
            typedef amgcl::preconditioner::b200_cycle_graph<AMG> GAMG;
            if (bicg) return run<amgcl::make_solver<GAMG, amgcl::solver::bicgstab<Backend>>>(rows, ptr, col, val, rhs);
            return run<amgcl::make_solver<GAMG, amgcl::solver::cg<Backend>>>(rows, ptr, col, val, rhs);
        }
        if (bicg) return run<amgcl::make_solver<AMG, amgcl::solver::bicgstab<Backend>>>(rows, ptr, col, val, rhs);
        return run<amgcl::make_solver<AMG, amgcl::solver::cg<Backend>>>(rows, ptr, col, val, rhs);
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 2;
    }
}
