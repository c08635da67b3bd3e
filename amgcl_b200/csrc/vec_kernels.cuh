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
//
// Templated over the element types (FP64 default; FP32 and the FP32->FP64 mixes of
// AMGCL's mixed-precision composition).  Arithmetic is done in the OUTPUT's type,
// as the reference's templates do.
#pragma once
#include "common.cuh"

namespace b200 {

template <class T> struct Vec16;
template <> struct Vec16<double> { typedef double2 type; static constexpr int N = 2; };
template <> struct Vec16<float>  { typedef float4  type; static constexpr int N = 4; };

__device__ __forceinline__ double vec_elem(const double2 &v, int i) { return i ? v.y : v.x; }
__device__ __forceinline__ float  vec_elem(const float4 &v, int i) {
    return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}
__device__ __forceinline__ void vec_set(double2 &v, int i, double s) { if (i) v.y = s; else v.x = s; }
__device__ __forceinline__ void vec_set(float4 &v, int i, float s) {
    if (i == 0) v.x = s; else if (i == 1) v.y = s; else if (i == 2) v.z = s; else v.w = s;
}

// 16-byte path: all streams share the output's element type (decided on the host)
template <class F, bool READ_Y, bool READ_Z, class T>
__device__ __forceinline__ void ew_vector_path(size_t n, F f, const T *x, const T *y, const T *z_in,
                                               T *out, size_t tid, size_t stride) {
    typedef typename Vec16<T>::type V;
    constexpr int N = Vec16<T>::N;
    const size_t nv = n / N;
    const V *xv = reinterpret_cast<const V *>(x);
    const V *yv = reinterpret_cast<const V *>(y);
    const V *zv = reinterpret_cast<const V *>(z_in);
    V *ov = reinterpret_cast<V *>(out);
    size_t i = tid;
    for (; i + stride < nv; i += 2 * stride) {
        const size_t j = i + stride;
        V xa = xv[i], xb = xv[j];
        V ya = xa, yb = xb, za = xa, zb = xb, oa, ob;
        if (READ_Y) { ya = yv[i]; yb = yv[j]; }
        if (READ_Z) { za = zv[i]; zb = zv[j]; }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            vec_set(oa, k, f(vec_elem(xa, k), vec_elem(ya, k), vec_elem(za, k)));
            vec_set(ob, k, f(vec_elem(xb, k), vec_elem(yb, k), vec_elem(zb, k)));
        }
        ov[i] = oa;
        ov[j] = ob;
    }
    if (i < nv) {
        V xa = xv[i];
        V ya = xa, za = xa, oa;
        if (READ_Y) ya = yv[i];
        if (READ_Z) za = zv[i];
#pragma unroll
        for (int k = 0; k < N; ++k)
            vec_set(oa, k, f(vec_elem(xa, k), vec_elem(ya, k), vec_elem(za, k)));
        ov[i] = oa;
    }
    // tail (n not a multiple of the vector width)
    for (size_t k = nv * N + tid; k < n; k += stride)
        out[k] = f(x[k], READ_Y ? y[k] : (T)0, READ_Z ? z_in[k] : (T)0);
}

// ---- generic element-wise driver ---------------------------------------------
// F: TO operator()(TO x, TO y, TO z); mixed element types take the scalar loop.
template <class F, bool READ_Y, bool READ_Z, class TX, class TY, class TZ, class TO>
__global__ void __launch_bounds__(kThreads)
ew_kernel(size_t n, F f, const TX *x, const TY *y, const TZ *z_in, TO *out, bool vec_ok) {
    ptx::pdl_wait();
    const size_t tid    = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < n; i += stride)
        out[i] = f((TO)x[i], READ_Y ? (TO)y[i] : (TO)0, READ_Z ? (TO)z_in[i] : (TO)0);
}
template <class F, bool READ_Y, bool READ_Z, class T>
__global__ void __launch_bounds__(kThreads)
ew_kernel_same(size_t n, F f, const T *x, const T *y, const T *z_in, T *out, bool vec_ok) {
    ptx::pdl_wait();
    const size_t tid    = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (vec_ok) {
        ew_vector_path<F, READ_Y, READ_Z, T>(n, f, x, y, z_in, out, tid, stride);
    } else {
        for (size_t i = tid; i < n; i += stride)
            out[i] = f(x[i], READ_Y ? y[i] : (T)0, READ_Z ? z_in[i] : (T)0);
    }
}

// functors: arithmetic written with the reference's association, in the output's type
template <class T> struct AxF       { T a;       __device__ T operator()(T x, T, T) const { return a * x; } };
template <class T> struct AxpbyF    { T a, b;    __device__ T operator()(T x, T y, T) const { return a * x + b * y; } };
template <class T> struct AxpbypczF { T a, b, c; __device__ T operator()(T x, T y, T z) const { return a * x + b * y + c * z; } };
template <class T> struct VmulF     { T a;       __device__ T operator()(T x, T y, T) const { return a * x * y; } };
template <class T> struct VmulAccF  { T a, b;    __device__ T operator()(T x, T y, T z) const { return a * x * y + b * z; } };
template <class T> struct CopyF     {            __device__ T operator()(T x, T, T) const { return x; } };

// ---- dot product: one kernel, deterministic, compensated -----------------------
// Each thread accumulates its grid-strided products with Kahan compensation in FP64
// (the reference's builtin backend does the same per OpenMP thread,
// builtin.hpp:1143-1181); the CTA reduces through warp shuffles and shared
// memory; the last CTA to finish (ticket counter) adds the per-CTA partials in
// index order and writes the scalar straight into mapped pinned host memory.
template <class T>
__global__ void __launch_bounds__(kThreads)
dot_kernel(size_t n, const T *__restrict__ x, const T *__restrict__ y,
           double *partial, unsigned int *ticket, double *result, bool vec_ok) {
    ptx::pdl_wait();
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
        typedef typename Vec16<T>::type V;
        constexpr int N = Vec16<T>::N;
        const size_t nv = n / N;
        const V *xv = reinterpret_cast<const V *>(x);
        const V *yv = reinterpret_cast<const V *>(y);
        size_t i = tid;
        for (; i + 3 * stride < nv; i += 4 * stride) {
            const V xa = xv[i], ya = yv[i];
            const V xb = xv[i + stride], yb = yv[i + stride];
            const V xc = xv[i + 2 * stride], yc = yv[i + 2 * stride];
            const V xd = xv[i + 3 * stride], yd = yv[i + 3 * stride];
#pragma unroll
            for (int k = 0; k < N; ++k) acc((double)vec_elem(xa, k) * (double)vec_elem(ya, k));
#pragma unroll
            for (int k = 0; k < N; ++k) acc((double)vec_elem(xb, k) * (double)vec_elem(yb, k));
#pragma unroll
            for (int k = 0; k < N; ++k) acc((double)vec_elem(xc, k) * (double)vec_elem(yc, k));
#pragma unroll
            for (int k = 0; k < N; ++k) acc((double)vec_elem(xd, k) * (double)vec_elem(yd, k));
        }
        for (; i < nv; i += stride) {
            const V xa = xv[i], ya = yv[i];
#pragma unroll
            for (int k = 0; k < N; ++k) acc((double)vec_elem(xa, k) * (double)vec_elem(ya, k));
        }
        for (size_t k = nv * N + tid; k < n; k += stride) acc((double)x[k] * (double)y[k]);
    } else {
        for (size_t i = tid; i < n; i += stride) acc((double)x[i] * (double)y[i]);
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

// ---- index lists (Backend::gather / Backend::scatter, cuda.hpp:548-577) --------------------
// dst[k] = src[idx[k]]: the index stream and the output are coalesced, the gathered side goes
// through the read-only path
template <class T>
__global__ void __launch_bounds__(kThreads)
gather_kernel(size_t n, const int *__restrict__ idx, const T *__restrict__ src, T *__restrict__ dst) {
    ptx::pdl_wait();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        dst[k] = __ldg(src + idx[k]);
}
// dst[idx[k]] = src[k] (indices are expected to be distinct, as thrust::scatter requires)
template <class T>
__global__ void __launch_bounds__(kThreads)
scatter_kernel(size_t n, const int *__restrict__ idx, const T *__restrict__ src, T *__restrict__ dst) {
    ptx::pdl_wait();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride)
        dst[idx[k]] = src[k];
}

} // namespace b200
