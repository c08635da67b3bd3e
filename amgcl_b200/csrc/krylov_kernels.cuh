// krylov_kernels.cuh -- the vector half of a Krylov iteration as fused passes.
//
// The reference issues the vector updates and inner products of CG / BiCGStab as separate
// backend calls with the scalars travelling through the host (solver/cg.hpp:180-198:
// dot; axpby | copy; spmv; dot; axpby; axpby; dot -- 3 host syncs per iteration;
// solver/bicgstab.hpp:198-236: 6 syncs).  Here each group of updates that share operands is
// ONE pass that reads every vector once, takes its coefficients from the device scalar table
// (reduce.cuh) -- so the quotient rho/<q,p> is formed on the device by the kernel that needs
// it -- and leaves the inner products of what it just wrote in the table:
//
//   cg_direction      p = s + (rho/rho_prev) p                       cg.hpp:186-189
//   cg_update         x += alpha p ; r -= alpha q ; <r,r>            cg.hpp:193-198, alpha = rho/<q,p>
//   bicg_direction    p = r + beta (p - omega v)                     bicgstab.hpp:204-208
//   bicg_update_s     x += alpha T ; s = r - alpha v ; <s,s>         bicgstab.hpp:212-222
//   bicg_update_r     x += omega T ; r = s - omega t ; <r,r>, <r,rh> bicgstab.hpp:226-236, 200
//   dot / norm        <x,y> (,<x,z>) , <x,x>                         interface.hpp:356-371
//
// Element arithmetic uses the same functors as the stand-alone axpby / axpbypcz kernels
// (vec_kernels.cuh), i.e. the reference's association a*x + b*y (+ c*z).  Products are
// accumulated per thread with Kahan compensation like dot_kernel / builtin.hpp:1143-1181.
#pragma once
#include "common.cuh"
#include "reduce.cuh"
#include "vec_kernels.cuh"

namespace b200 {

template <int NIN, int NOUT>
struct FusedArgs {
    const double *in[NIN];
    double       *out[NOUT > 0 ? NOUT : 1];
};

// F provides:   static constexpr int NIN, NOUT, NRED;
//               __device__ void prepare(bool leader);      (reads the device scalars)
//               __device__ void apply(const double *in, double *out, double *prod) const;
template <class F, int UNR>
__global__ void __launch_bounds__(kThreads)
fused_vec_kernel(size_t n, F f, FusedArgs<F::NIN, F::NOUT> a, RedOut ro, bool vec_ok) {
    constexpr int NIN = F::NIN, NOUT = F::NOUT, NRED = F::NRED;
    ptx::pdl_wait();
    f.prepare(blockIdx.x == 0 && threadIdx.x == 0);
    const size_t tid    = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    double s[NRED > 0 ? NRED : 1], c[NRED > 0 ? NRED : 1];
#pragma unroll
    for (int k = 0; k < (NRED > 0 ? NRED : 1); ++k) { s[k] = 0.0; c[k] = 0.0; }
    auto element = [&](const double *in, double *out) {
        double prod[NRED > 0 ? NRED : 1];
        f.apply(in, out, prod);
#pragma unroll
        for (int k = 0; k < NRED; ++k) {
            const double d = prod[k] - c[k];
            const double t = s[k] + d;
            c[k] = (t - s[k]) - d;
            s[k] = t;
        }
    };
    if (vec_ok) {
        const size_t nv = n / 2;
        size_t i = tid;
        for (; i + (UNR - 1) * stride < nv; i += UNR * stride) {
            double2 v[UNR][NIN];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int k = 0; k < NIN; ++k)
                    v[u][k] = reinterpret_cast<const double2 *>(a.in[k])[i + u * stride];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                double ia[NIN], ib[NIN], oa[NOUT > 0 ? NOUT : 1], ob[NOUT > 0 ? NOUT : 1];
#pragma unroll
                for (int k = 0; k < NIN; ++k) { ia[k] = v[u][k].x; ib[k] = v[u][k].y; }
                element(ia, oa);
                element(ib, ob);
#pragma unroll
                for (int k = 0; k < NOUT; ++k)
                    reinterpret_cast<double2 *>(a.out[k])[i + u * stride] = make_double2(oa[k], ob[k]);
            }
        }
        for (; i < nv; i += stride) {
            double ia[NIN], ib[NIN], oa[NOUT > 0 ? NOUT : 1], ob[NOUT > 0 ? NOUT : 1];
#pragma unroll
            for (int k = 0; k < NIN; ++k) {
                const double2 v = reinterpret_cast<const double2 *>(a.in[k])[i];
                ia[k] = v.x; ib[k] = v.y;
            }
            element(ia, oa);
            element(ib, ob);
#pragma unroll
            for (int k = 0; k < NOUT; ++k)
                reinterpret_cast<double2 *>(a.out[k])[i] = make_double2(oa[k], ob[k]);
        }
        for (size_t j = nv * 2 + tid; j < n; j += stride) {
            double ia[NIN], oa[NOUT > 0 ? NOUT : 1];
#pragma unroll
            for (int k = 0; k < NIN; ++k) ia[k] = a.in[k][j];
            element(ia, oa);
#pragma unroll
            for (int k = 0; k < NOUT; ++k) a.out[k][j] = oa[k];
        }
    } else {
        for (size_t j = tid; j < n; j += stride) {
            double ia[NIN], oa[NOUT > 0 ? NOUT : 1];
#pragma unroll
            for (int k = 0; k < NIN; ++k) ia[k] = a.in[k][j];
            element(ia, oa);
#pragma unroll
            for (int k = 0; k < NOUT; ++k) a.out[k][j] = oa[k];
        }
    }
    if (NRED > 0) red_finish<(NRED > 0 ? NRED : 1)>(ro, s);
}

// a derived scalar goes to the device table and its host mirror
struct SaveSlot {
    double *d, *h;
    // (one thread per launch; the system fence orders the host copy before the launch's
    // "ready" word, which another CTA releases later)
    __device__ void put(double v) const { *d = v; if (h) { *h = v; __threadfence_system(); } }
};

// ---- reductions only --------------------------------------------------------------------
struct NormF {                      // <x,x>
    static constexpr int NIN = 1, NOUT = 0, NRED = 1;
    __device__ void prepare(bool) {}
    __device__ void apply(const double *in, double *, double *prod) const { prod[0] = in[0] * in[0]; }
};
struct DotF {                       // <x,y>
    static constexpr int NIN = 2, NOUT = 0, NRED = 1;
    __device__ void prepare(bool) {}
    __device__ void apply(const double *in, double *, double *prod) const { prod[0] = in[0] * in[1]; }
};
struct Dot2F {                      // <x,y>, <x,z>
    static constexpr int NIN = 3, NOUT = 0, NRED = 2;
    __device__ void prepare(bool) {}
    __device__ void apply(const double *in, double *, double *prod) const {
        prod[0] = in[0] * in[1];
        prod[1] = in[0] * in[2];
    }
};

struct CopyNormF {                  // y = x ; <x,x>   (r = rhs - A*0)
    static constexpr int NIN = 1, NOUT = 1, NRED = 1;
    __device__ void prepare(bool) {}
    __device__ void apply(const double *in, double *out, double *prod) const {
        out[0] = in[0];
        prod[0] = in[0] * in[0];
    }
};

// ---- CG (solver/cg.hpp:180-198) ---------------------------------------------------------
// p = s + beta p, beta = rho / rho_prev (first iteration: p = s).  in: s, p ; out: p.
struct CgDirectionF {
    static constexpr int NIN = 2, NOUT = 1, NRED = 0;
    const double *rho, *rho_prev;
    SaveSlot rho_save;              // the NEXT call's rho_prev (a different slot than rho_prev)
    int first;
    double beta;
    __device__ void prepare(bool leader) {
        const double r1 = *rho;
        beta = first ? 0.0 : r1 / *rho_prev;
        if (leader) rho_save.put(r1);
    }
    __device__ void apply(const double *in, double *out, double *) const {
        out[0] = first ? in[0] : AxpbyF<double>{1.0, beta}(in[0], in[1], 0.0);
    }
};
// alpha = rho / <q,p> ; x = alpha p + x ; r = -alpha q + r ; <r,r>.  in: p, q, x, r ; out: x, r.
struct CgUpdateF {
    static constexpr int NIN = 4, NOUT = 2, NRED = 1;
    const double *rho, *qp;
    SaveSlot alpha_save;
    double alpha;
    __device__ void prepare(bool leader) {
        alpha = *rho / *qp;
        if (leader) alpha_save.put(alpha);
    }
    __device__ void apply(const double *in, double *out, double *prod) const {
        out[0] = AxpbyF<double>{alpha, 1.0}(in[0], in[2], 0.0);
        out[1] = AxpbyF<double>{-alpha, 1.0}(in[1], in[3], 0.0);
        prod[0] = out[1] * out[1];
    }
};

// ---- BiCGStab, right preconditioning (solver/bicgstab.hpp:198-236) ------------------------
// p = r - beta*omega v + beta p, beta = (rho*alpha)/(rho_prev*omega) (first: p = r).
// in: r, v, p ; out: p.
struct BicgDirectionF {
    static constexpr int NIN = 3, NOUT = 1, NRED = 0;
    const double *rho, *rho_prev, *alpha, *omega;
    SaveSlot rho_save;
    int first;
    double b, c;
    __device__ void prepare(bool leader) {
        const double r1 = *rho;
        b = 0.0; c = 0.0;
        if (!first) {
            const double om = *omega;
            const double beta = (r1 * *alpha) / (*rho_prev * om);
            b = -beta * om;
            c = beta;
        }
        if (leader) rho_save.put(r1);
    }
    __device__ void apply(const double *in, double *out, double *) const {
        out[0] = first ? in[0] : AxpbypczF<double>{1.0, b, c}(in[0], in[1], in[2]);
    }
};
// alpha = rho / <rh,v> ; x = alpha T + x ; s = r - alpha v ; <s,s>.  in: T, x, r, v ; out: x, s.
struct BicgUpdateSF {
    static constexpr int NIN = 4, NOUT = 2, NRED = 1;
    const double *rho, *rhv;
    SaveSlot alpha_save;
    double alpha;
    __device__ void prepare(bool leader) {
        alpha = *rho / *rhv;
        if (leader) alpha_save.put(alpha);
    }
    __device__ void apply(const double *in, double *out, double *prod) const {
        out[0] = AxpbyF<double>{alpha, 1.0}(in[0], in[1], 0.0);
        out[1] = AxpbyF<double>{1.0, -alpha}(in[2], in[3], 0.0);
        prod[0] = out[1] * out[1];
    }
};
// omega = <t,s> / <t,t> ; x = omega T + x ; r = s - omega t ; <r,r> ; <r,rh>.
// in: T, x, s, t, rh ; out: x, r.
struct BicgUpdateRF {
    static constexpr int NIN = 5, NOUT = 2, NRED = 2;
    const double *ts, *tt;
    SaveSlot omega_save;
    double omega;
    __device__ void prepare(bool leader) {
        omega = *ts / *tt;
        if (leader) omega_save.put(omega);
    }
    __device__ void apply(const double *in, double *out, double *prod) const {
        out[0] = AxpbyF<double>{omega, 1.0}(in[0], in[1], 0.0);
        out[1] = AxpbyF<double>{1.0, -omega}(in[2], in[3], 0.0);
        prod[0] = out[1] * out[1];
        prod[1] = out[1] * in[4];
    }
};

} // namespace b200
