// coarse_kernels.cuh -- coarsest-level direct solve kept on the device.
//
// The reference cuda backend solves the coarsest system on the HOST every
// cycle (device->host copy, serial skyline LU sweeps, host->device copy:
// amgcl/backend/cuda.hpp:61-84, amgcl/solver/skyline_lu.hpp:179-200), which
// stalls the stream once per V-cycle.  Here the n x n inverse (n <=
// coarse_enough = 3000, skyline_lu.hpp:93-95) is formed once at setup by
// Gauss-Jordan elimination with partial pivoting on the augmented matrix
// [A | I], and each cycle applies it with one dense GEMV that never leaves HBM/L2.
#pragma once
#include "common.cuh"

namespace b200 {

// scatter CSR entries into the left half of the augmented row-major matrix M[n][2n]
// and set the right half to the identity (M was zero-filled)
__global__ void coarse_scatter_kernel(int n, const int *__restrict__ ptr,
                                      const int *__restrict__ col,
                                      const double *__restrict__ val, double *M) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double *row = M + (size_t)r * 2 * n;
    for (int e = ptr[r]; e < ptr[r + 1]; ++e) row[col[e]] += val[e];
    row[n + r] = 1.0;
}

// one CTA: p = argmax_{i>=k} |M[i][k]| ; piv[0] = p ; pivval[0] = M[p][k]
__global__ void __launch_bounds__(kThreads)
coarse_pivot_kernel(int n, int k, const double *M, int *piv, double *pivval) {
    __shared__ double sv[kThreads];
    __shared__ int    si[kThreads];
    double best = -1.0;
    int    bi   = k;
    for (int i = k + threadIdx.x; i < n; i += kThreads) {
        const double a = fabs(M[(size_t)i * 2 * n + k]);
        if (a > best) { best = a; bi = i; }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double b = sv[threadIdx.x + o];
            const int    j = si[threadIdx.x + o];
            // ties resolved towards the smaller row index: deterministic
            if (b > sv[threadIdx.x] || (b == sv[threadIdx.x] && j < si[threadIdx.x])) {
                sv[threadIdx.x] = b;
                si[threadIdx.x] = j;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        piv[0]    = si[0];
        pivval[0] = M[(size_t)si[0] * 2 * n + k];
    }
}

// Save column k of every row (the elimination multipliers) as it will be after
// rows k and p are exchanged; done in its own launch so it never races with the
// row exchange below.
__global__ void __launch_bounds__(kThreads)
coarse_colk_kernel(int n, int k, const double *M, const int *piv, double *colk) {
    const int p = piv[0];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int src = j;
    if (j == k) src = p; else if (j == p) src = k;
    colk[j] = M[(size_t)src * 2 * n + k];
}
// exchange rows k and p, scale the new row k by 1/pivot
__global__ void __launch_bounds__(kThreads)
coarse_swap_scale_kernel(int n, int k, double *M, const int *piv, const double *pivval) {
    const int p   = piv[0];
    const double inv = 1.0 / pivval[0];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 2 * n) return;
    double *rk = M + (size_t)k * 2 * n;
    double *rp = M + (size_t)p * 2 * n;
    const double vk = rk[j], vp = rp[j];
    rk[j] = vp * inv;
    if (p != k) rp[j] = vk;
}

// M[i][:] -= colk[i] * M[k][:] for every row i != k.  Columns < k of row k are
// already zero, so only columns [k, 2n) can change (row exchanges move the
// identity's non-zeros anywhere in the right half).
__global__ void __launch_bounds__(kThreads)
coarse_eliminate_kernel(int n, int k, double *M, const double *__restrict__ colk) {
    const int j = k + blockIdx.x * blockDim.x + threadIdx.x;   // column
    if (j >= 2 * n) return;
    const double pk = M[(size_t)k * 2 * n + j];
    const int rows_per = (n + gridDim.y - 1) / gridDim.y;
    const int i0 = blockIdx.y * rows_per;
    const int i1 = min(n, i0 + rows_per);
    for (int i = i0; i < i1; ++i) {
        if (i == k) continue;
        const double m = colk[i];
        if (m != 0.0) M[(size_t)i * 2 * n + j] = fma(-m, pk, M[(size_t)i * 2 * n + j]);
    }
}

// copy the right half of M into the dense inverse
__global__ void coarse_extract_kernel(int n, const double *__restrict__ M, double *Ainv) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const size_t r = idx / n, c = idx % n;
    Ainv[idx] = M[r * 2 * n + n + c];
}

// x[0:nloc) = Ainv[row0 : row0+nloc, :] * rhs : one warp per row, coalesced row reads,
// shuffle reduction.  (row0, nloc) select this rank's rows when the coarsest level is
// itself partitioned; single GPU: row0 = 0, nloc = n.
template <class T>
__global__ void __launch_bounds__(kThreads)
coarse_gemv_kernel(int n, int row0, int nloc, const double *__restrict__ Ainv,
                   const T *__restrict__ rhs, T *__restrict__ x) {
    ptx::pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nloc) return;
    const double *row = Ainv + (size_t)(row0 + warp) * n;
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma(row[j], (double)rhs[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) x[warp] = (T)s;
}

} // namespace b200
