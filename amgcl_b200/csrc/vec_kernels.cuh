// vec_kernels.cuh -- BLAS-1 style kernels of the Krylov loop and the cycle.
//
//   axpby      y = a*x + b*y            builtin.hpp:1185-1209 / cuda.hpp:676-716
//   axpbypcz   z = a*x + b*y + c*z      builtin.hpp:1211-1236 / cuda.hpp:718-765
//   vmul       z = a*x.*y + b*z         builtin.hpp:1238-1265 / cuda.hpp:767-807
//   dot        sum_i x_i*y_i            builtin.hpp:1099-1183 / cuda.hpp:662-674
//
// All are pure HBM streams: 16-byte (LDG.128 / STG.128) accesses, two independent
// vector loads in flight per stream per thread, grid sized to a multiple of the
// SM count.  As in the reference, an output whose coefficient is zero is never
// read (it may hold uninitialised memory, i.e. NaNs).
#pragma once
#include "common.cuh"

namespace b200 {

// ---- generic element-wise driver ---------------------------------------------
// F: double operator()(double x, double y, double z) ; NIN = streams read (1..3);
// the output aliases the last input stream when RMW is set.
template <class F, bool READ_Y, bool READ_Z>
__global__ void __launch_bounds__(kThreads)
ew_kernel(size_t n, F f, const double *x, const double *y, const double *z_in,
          double *out, bool vec_ok) {
    const size_t tid    = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (vec_ok) {
        const size_t n2 = n >> 1;
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        const double2 *y2 = reinterpret_cast<const double2 *>(y);
        const double2 *z2 = reinterpret_cast<const double2 *>(z_in);
        double2 *o2 = reinterpret_cast<double2 *>(out);
        size_t i = tid;
        for (; i + stride < n2; i += 2 * stride) {
            const size_t j = i + stride;
            double2 xa = x2[i], xb = x2[j];
            double2 ya = make_double2(0, 0), yb = ya, za = ya, zb = ya;
            if (READ_Y) { ya = y2[i]; yb = y2[j]; }
            if (READ_Z) { za = z2[i]; zb = z2[j]; }
            o2[i] = make_double2(f(xa.x, ya.x, za.x), f(xa.y, ya.y, za.y));
            o2[j] = make_double2(f(xb.x, yb.x, zb.x), f(xb.y, yb.y, zb.y));
        }
        if (i < n2) {
            double2 xa = x2[i];
            double2 ya = make_double2(0, 0), za = ya;
            if (READ_Y) ya = y2[i];
            if (READ_Z) za = z2[i];
            o2[i] = make_double2(f(xa.x, ya.x, za.x), f(xa.y, ya.y, za.y));
        }
        if ((n & 1) && tid == 0) {
            const size_t k = n - 1;
            out[k] = f(x[k], READ_Y ? y[k] : 0.0, READ_Z ? z_in[k] : 0.0);
        }
    } else {
        for (size_t i = tid; i < n; i += stride)
            out[i] = f(x[i], READ_Y ? y[i] : 0.0, READ_Z ? z_in[i] : 0.0);
    }
}

// functors: arithmetic written with the reference's association
struct AxF      { double a;       __device__ double operator()(double x, double, double) const { return a * x; } };
struct AxpbyF   { double a, b;    __device__ double operator()(double x, double y, double) const { return a * x + b * y; } };
struct AxpbyZF  { double a, b;    __device__ double operator()(double x, double y, double) const { return a * x + b * y; } };
struct AxpbypczF{ double a, b, c; __device__ double operator()(double x, double y, double z) const { return a * x + b * y + c * z; } };
struct VmulF    { double a;       __device__ double operator()(double x, double y, double) const { return a * x * y; } };
struct VmulAccF { double a, b;    __device__ double operator()(double x, double y, double z) const { return a * x * y + b * z; } };
struct CopyF    {                 __device__ double operator()(double x, double, double) const { return x; } };

// ---- dot product: one kernel, deterministic, compensated -----------------------
// Each thread accumulates its grid-strided products with Kahan compensation
// (the reference's builtin backend does the same per OpenMP thread,
// builtin.hpp:1143-1181); the CTA reduces through warp shuffles and shared
// memory; the last CTA to finish (ticket counter) adds the per-CTA partials in
// index order and writes the scalar straight into mapped pinned host memory.
__global__ void __launch_bounds__(kThreads)
dot_kernel(size_t n, const double *__restrict__ x, const double *__restrict__ y,
           double *partial, unsigned int *ticket, double *result, bool vec_ok) {
    const size_t tid    = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    double s = 0.0, c = 0.0;
    auto acc = [&](double p) {
        const double d = p - c;
        const double t = s + d;
        c = (t - s) - d;
        s = t;
    };
    if (vec_ok) {
        const size_t n2 = n >> 1;
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        const double2 *y2 = reinterpret_cast<const double2 *>(y);
        size_t i = tid;
        for (; i + 3 * stride < n2; i += 4 * stride) {
            const double2 xa = x2[i], ya = y2[i];
            const double2 xb = x2[i + stride], yb = y2[i + stride];
            const double2 xc = x2[i + 2 * stride], yc = y2[i + 2 * stride];
            const double2 xd = x2[i + 3 * stride], yd = y2[i + 3 * stride];
            acc(xa.x * ya.x); acc(xa.y * ya.y);
            acc(xb.x * yb.x); acc(xb.y * yb.y);
            acc(xc.x * yc.x); acc(xc.y * yc.y);
            acc(xd.x * yd.x); acc(xd.y * yd.y);
        }
        for (; i < n2; i += stride) {
            const double2 xa = x2[i], ya = y2[i];
            acc(xa.x * ya.x); acc(xa.y * ya.y);
        }
        if ((n & 1) && tid == 0) acc(x[n - 1] * y[n - 1]);
    } else {
        for (size_t i = tid; i < n; i += stride) acc(x[i] * y[i]);
    }

    __shared__ double warp_sum[kThreads / 32];
    __shared__ bool   is_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) b += warp_sum[w];
        partial[blockIdx.x] = b;
        __threadfence();
        const unsigned int done = atomicAdd(ticket, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fixed-order tree over the per-CTA partials (gridDim.x <= kDotMaxBlocks)
        double v = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += kThreads)
            v += __ldcg(partial + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) tot += warp_sum[w];
            *result = tot;
            *ticket = 0;            // ready for the next call on this stream
        }
    }
}

} // namespace b200
