// reduce.cuh -- scalar results produced INSIDE the kernels that stream the data.
//
// The Krylov loop of the reference (solver/cg.hpp:180-198, solver/bicgstab.hpp:198-236) asks
// for an inner product right after the kernel that wrote one of its operands: q = A p then
// <q, p>; r -= alpha q then <r, r>; the V-cycle's last smoother sweep then <r, s>.  On a
// bandwidth-bound device the separate reduction re-reads both vectors and costs a launch and a
// host round trip.  Here every kernel that can produce such a scalar carries a RedOut: each
// thread accumulates its products while the data is in registers, red_finish() reduces them
// per CTA (shuffle + shared memory, fixed order), the last CTA to arrive (ticket) adds the
// per-CTA partials in index order -- run-to-run deterministic for a fixed grid -- and leaves
// the scalar in DEVICE memory for the next kernel to consume (plus a mapped host mirror the
// host may read after a stream synchronize).  Scalars never have to visit the host between
// the kernels of one Krylov iteration.
//
// Multi-GPU (one process per GPU): the finishing CTA also does the all-reduce.  It stores the
// rank's partial into slot [parity][slot][rank] of every peer's exchange buffer (plain stores
// over NVLink into CUDA-IPC mapped memory), releases a flag per peer, waits for the peers'
// flags and adds the P partials in rank order -- bitwise identical on every rank, no extra
// kernel, no NCCL call (replaces mpi/inner_product.hpp:53-62's MPI_Allreduce).
#pragma once
#include "common.cuh"

namespace b200 {

// (kMaxRed scalars per launch, kScalSlots table slots: common.cuh)

// layout of one rank's scalar exchange buffer (peer-mapped)
struct ScalExchange {
    unsigned long long flag[2][kScalSlots][kMaxRanks];
    double             val[2][kScalSlots][kMaxRanks];
};

struct RedOut {
    double       *partial;            // [nred * gridDim.x] per-CTA partial sums (scratch)
    unsigned int *ticket;             // self-resetting arrival counter
    double       *dev[kMaxRed];       // where each scalar goes (device)
    double       *host[kMaxRed];      // mapped host mirror (nullptr: none) ...
    unsigned long long *host_seq[kMaxRed];   // ... and its "ready" word: set to seq[k] after the value,
                                      // so the host can poll instead of synchronising the stream
    int           nred;
    // multi-GPU: all-reduce inside the finishing CTA
    int           nranks, rank;
    ScalExchange *const *peers;       // device array [nranks] of the ranks' exchange buffers
    int           slot[kMaxRed];
    unsigned long long seq[kMaxRed];  // sequence number of this use of the slot (parity = seq & 1)
};

__device__ __forceinline__ void red_st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long red_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of EVERY CTA as the last thing the kernel does.  v[k] is the thread's
// contribution to scalar k (k < o.nred <= NRED).
template <int NRED>
__device__ __forceinline__ void red_finish(const RedOut &o, double (&v)[NRED]) {
    __shared__ double red_warp[NRED][kThreads / 32];
    __shared__ double red_tot[NRED];
    __shared__ bool   red_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
        double s = v[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) red_warp[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NRED; ++k) {
            if (k >= o.nred) break;
            double b = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) b += red_warp[k][w];
            o.partial[(size_t)k * gridDim.x + blockIdx.x] = b;
        }
        __threadfence();
        const unsigned int done = atomicAdd(o.ticket, 1u);
        red_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!red_last) return;
    __threadfence();
    // fixed-order tree over the per-CTA partials
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
        if (k >= o.nred) break;
        double s = 0.0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += kThreads)
            s += __ldcg(o.partial + (size_t)k * gridDim.x + i);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        __syncthreads();                     // red_warp free (previous k / first phase)
        if (lane == 0) red_warp[k][warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) tot += red_warp[k][w];
            red_tot[k] = tot;
        }
    }
    __syncthreads();
    if (o.nranks > 1) {
        // all-reduce over the ranks: thread q talks to peer q
        const int q = threadIdx.x;
#pragma unroll
        for (int k = 0; k < NRED; ++k) {
            if (k >= o.nred) break;
            const int par = (int)(o.seq[k] & 1ull);
            if (q < o.nranks) {
                ScalExchange *pq = o.peers[q];
                pq->val[par][o.slot[k]][o.rank] = red_tot[k];
                // (release at system scope orders the store above before the flag)
                red_st_release_sys(&pq->flag[par][o.slot[k]][o.rank], o.seq[k]);
                const ScalExchange *mine = o.peers[o.rank];
                while (red_ld_acquire_sys(&mine->flag[par][o.slot[k]][q]) < o.seq[k]) { }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const ScalExchange *mine = o.peers[o.rank];
#pragma unroll
            for (int k = 0; k < NRED; ++k) {
                if (k >= o.nred) break;
                const int par = (int)(o.seq[k] & 1ull);
                double tot = 0.0;
                for (int r = 0; r < o.nranks; ++r)
                    tot += __ldcv(&mine->val[par][o.slot[k]][r]);
                red_tot[k] = tot;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NRED; ++k) {
            if (k >= o.nred) break;
            *o.dev[k] = red_tot[k];
            if (o.host[k]) {
                *o.host[k] = red_tot[k];
                red_st_release_sys(o.host_seq[k], o.seq[k]);
            }
        }
        *o.ticket = 0;                       // ready for the next launch on this stream
    }
}

} // namespace b200
