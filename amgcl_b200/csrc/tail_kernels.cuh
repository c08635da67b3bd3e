// tail_kernels.cuh -- the coarse tail of the V-cycle as ONE kernel.
//
// amg::cycle (amg.hpp:514-553) issues ~5 backend calls per level; on the coarse levels
// (<= ~10^5 rows) each of them is a kernel that runs for 2-4 us but costs the 8-9 us a
// dependent launch takes to start, so at 256^3 the levels below 10^5 rows cost ~14 launches =
// ~110 us per cycle for ~20 us of work, and at 64^3 they ARE the cycle.  AMGCL's host control
// flow stays unmodified: the C ABI DEFERS calls whose operator is small (api_tail.cu) into a
// command list and, as soon as a call that cannot be deferred arrives, runs the whole list as
// one persistent kernel -- one CTA per SM, commands separated by a device-wide barrier instead
// of a kernel boundary.
//
// Bit-compatibility: every command evaluates exactly the arithmetic of the stand-alone kernel
// it replaces -- same lanes per row, same per-lane entry order, same shuffle tree, same
// epilogue (csr_kernels.cuh: compute_staged / store_row; relax_zero_kernel;
// coarse_kernels.cuh: coarse_gemv_kernel) -- so a solve with and without the tail kernel gives
// the same bits (tests/test_gpu_solver.py::test_coarse_tail_is_bit_transparent).
//
// Memory model: vectors written by one command are read by later commands of the same kernel
// on other SMs, so vector reads go to L2 (ld.global.cg); only data no command writes (matrix
// arrays, smoother diagonal, dense inverse) use the read-only path.
#pragma once
#include "common.cuh"
#include "csr_kernels.cuh"

namespace b200 {

enum { TAIL_CSR = 0, TAIL_RELAX_ZERO = 1, TAIL_GEMV = 2 };
constexpr int kTailMaxCmds = 28;

struct TailCmd {
    int           op;       // TAIL_*
    int           mode;     // TAIL_CSR: MODE_*
    int           nrows;    // rows / vector length / dense dimension
    int           lanes;    // TAIL_CSR: lanes per row
    const int    *ptr;
    const int    *col;
    const double *val;      // TAIL_GEMV: dense inverse, row-major
    const double *x;        // gathered vector / GEMV right-hand side
    const double *f;
    const double *d;
    double       *y;
    double        alpha, beta;
};

struct TailArgs {
    int           n;
    unsigned int *bar;      // [2]: arrival counter, generation
    TailCmd       cmd[kTailMaxCmds];
};

// device-wide barrier between two commands (all CTAs are co-resident: cooperative launch)
__device__ __forceinline__ void tail_barrier(unsigned int *bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int gen = *reinterpret_cast<volatile unsigned int *>(bar + 1);
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            *reinterpret_cast<volatile unsigned int *>(bar) = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*reinterpret_cast<volatile unsigned int *>(bar + 1) == gen) { }
        }
        __threadfence();
    }
    __syncthreads();
}

// one CSR pass, L lanes per row: the arithmetic of compute_staged<MODE, L> + store_row<MODE>
template <int L>
__device__ __forceinline__ void tail_csr(const TailCmd &c) {
    const int gid     = (blockIdx.x * kThreads + threadIdx.x) / L;
    const int lane    = threadIdx.x % L;
    const int ngroups = gridDim.x * kThreads / L;
    for (int base = 0; base < c.nrows; base += ngroups) {
        const int  r     = base + gid;
        const bool valid = r < c.nrows;
        double sum = 0.0;
        if (valid) {
            const int beg = __ldg(c.ptr + r), end = __ldg(c.ptr + r + 1);
            int e = beg + lane;
            if (L >= 16) {
                // wide groups: the first entry of a lane is a plain product (csr_kernels.cuh)
                if (e < end) { sum = __ldg(c.val + e) * __ldcg(c.x + __ldg(c.col + e)); e += L; }
            }
            for (; e < end; e += L) sum = fma(__ldg(c.val + e), __ldcg(c.x + __ldg(c.col + e)), sum);
        }
        if (L > 1) {
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        }
        if (valid && lane == 0) {
            double out;
            if (c.mode == MODE_SPMV) {
                out = c.alpha * sum;
            } else if (c.mode == MODE_SPMV_ACC) {
                out = c.alpha * sum + c.beta * __ldcg(c.y + r);
            } else if (c.mode == MODE_RESID) {
                out = __ldcg(c.f + r) - sum;
            } else {
                const double t = __ldcg(c.f + r) - sum;
                const double w = c.alpha * __ldg(c.d + r);
                out = fma(w, t, __ldcg(c.x + r));
            }
            c.y[r] = out;
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) coarse_tail_kernel(const TailArgs a) {
    for (int k = 0; k < a.n; ++k) {
        const TailCmd &c = a.cmd[k];
        if (c.op == TAIL_CSR) {
            switch (c.lanes) {
            case 1:  tail_csr<1>(c); break;
            case 2:  tail_csr<2>(c); break;
            case 4:  tail_csr<4>(c); break;
            case 8:  tail_csr<8>(c); break;
            case 16: tail_csr<16>(c); break;
            default: tail_csr<32>(c); break;
            }
        } else if (c.op == TAIL_RELAX_ZERO) {
            // x = (omega*d).*rhs (relax_zero_kernel)
            const int stride = gridDim.x * kThreads;
            for (int i = blockIdx.x * kThreads + threadIdx.x; i < c.nrows; i += stride)
                c.y[i] = fma(c.alpha * __ldg(c.d + i), __ldcg(c.f + i), 0.0);
        } else {
            // x = Ainv * rhs: one warp per row (coarse_gemv_kernel)
            const int lane   = threadIdx.x & 31;
            const int nwarps = gridDim.x * (kThreads / 32);
            for (int row = (blockIdx.x * kThreads + threadIdx.x) >> 5; row < c.nrows; row += nwarps) {
                const double *rowp = c.val + (size_t)row * c.nrows;
                double s = 0.0;
                for (int j = lane; j < c.nrows; j += 32) s = fma(__ldg(rowp + j), __ldcg(c.x + j), s);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) c.y[row] = s;
            }
        }
        if (k + 1 < a.n) tail_barrier(a.bar);
    }
}

} // namespace b200
