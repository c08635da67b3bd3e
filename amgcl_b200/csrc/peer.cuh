// peer.cuh -- exchange kernels over NVLink peer memory (multi-GPU, one process per GPU).
//
// The collectives of the partitioned solve are tiny (a halo plane, one coarse
// vector, one scalar) and sit on the critical path ~200 times per solve, so their
// cost is latency, not bandwidth: an NCCL call costs ~20-25 us each on B200/NVSwitch
// (measured, profiles/r1_bench_n2_nccl_*.json).  Here every rank maps its peers'
// exchange buffers once (CUDA IPC) and the data moves with plain stores from our own
// kernels:
//
//   push_kernel        producer side: writes this rank's contribution straight into
//                      each consumer's buffer over NVLink (st.global on mapped peer
//                      pointers), then -- last CTA done, after a system-scope fence --
//                      releases one 64-bit flag per consumer (st.release.sys).
//   wait_kernel        consumer side: spins (ld.acquire.sys) until the flags of all
//                      the producers it depends on carry the expected sequence number.
//   reduce_sum_kernel  consumer side of a reduction: waits like wait_kernel, then adds
//                      the P staged partial vectors in rank order (deterministic,
//                      unlike a ring all-reduce) into the destination.
//
// Buffers are double-buffered by the parity of a per-object sequence number; in an
// SPMD program that suffices against write-after-read: a rank can only reach
// exchange k+2 after it consumed exchange k+1, which its peers only produce after
// their own exchange-k consumers finished (stream order).
#pragma once
#include "common.cuh"

namespace b200 {

// (kMaxRanks, kFlagStride, kFlagBytes: common.cuh)

struct PeerTargets {
    double             *data[kMaxRanks];   // where my contribution goes in peer q (nullptr: skip)
    unsigned long long *flag[kMaxRanks];   // flag to release in peer q           (nullptr: skip)
};

struct WaitList {
    const unsigned long long *flag[kMaxRanks];   // local flags to wait on (nullptr: skip)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Push `count` doubles (src[idx[i]] if idx, else src[i]) to every target; per target the
// source window may be shifted: target q receives src[q*seg_stride + i] (reduce-scatter).
// Contiguous sources move as 16-byte stores, four per thread in flight (NVLink stores are
// posted; what matters is bytes in flight per SM).  grid-stride: any grid size works.
__global__ void __launch_bounds__(kThreads)
push_kernel(int64_t count, const double *__restrict__ src, const int *__restrict__ idx,
            PeerTargets tgt, int nranks, int64_t seg_stride /* 0: same window for all */,
            unsigned int *ticket, unsigned long long seq) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    if (idx) {
        for (int64_t i = tid; i < count; i += nthr) {
            const double v = src[idx[i]];
#pragma unroll 1
            for (int q = 0; q < nranks; ++q)
                if (tgt.data[q]) tgt.data[q][i] = v;
        }
    } else {
        const int64_t n2 = count >> 1;
#pragma unroll 1
        for (int q = 0; q < nranks; ++q) {
            if (!tgt.data[q]) continue;
            const double *sq = src + (int64_t)q * seg_stride;
            double *dq = tgt.data[q];
            if ((reinterpret_cast<uintptr_t>(sq) | reinterpret_cast<uintptr_t>(dq)) & 15) {
                // a window that starts on an odd element (e.g. rank * n_coarse): 8-byte stores
                for (int64_t i = tid; i < count; i += nthr) dq[i] = sq[i];
                continue;
            }
            const double2 *s2 = reinterpret_cast<const double2 *>(sq);
            double2 *d2 = reinterpret_cast<double2 *>(dq);
            int64_t i = tid;
            for (; i + 3 * nthr < n2; i += 4 * nthr) {
                const double2 a = s2[i], b = s2[i + nthr], c = s2[i + 2 * nthr], d = s2[i + 3 * nthr];
                d2[i] = a; d2[i + nthr] = b; d2[i + 2 * nthr] = c; d2[i + 3 * nthr] = d;
            }
            for (; i < n2; i += nthr) d2[i] = s2[i];
            if ((count & 1) && tid == 0)
                tgt.data[q][count - 1] = src[(int64_t)q * seg_stride + count - 1];
        }
    }
    __threadfence_system();                  // my peer stores are visible system-wide ...
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(ticket, 1u);   // ... before the ticket moves
        last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < nranks && tgt.flag[threadIdx.x])
            st_release_sys(tgt.flag[threadIdx.x], seq);
        if (threadIdx.x == 0) *ticket = 0;
    }
}

__global__ void wait_kernel(WaitList w, int nranks, unsigned long long seq) {
    const int q = threadIdx.x;
    if (q < nranks && w.flag[q]) {
        while (ld_acquire_sys(w.flag[q]) < seq) { __nanosleep(20); }
    }
}

// dst[i] = sum_r staged[r*stride + i] (rank order), after all flags arrived.
// If host_out is set (dot product) the single result also goes to mapped host memory.
__global__ void __launch_bounds__(kThreads)
reduce_sum_kernel(int64_t count, const double *staged, int64_t stride, int nranks, WaitList w,
                  unsigned long long seq, double *dst, double *host_out) {
    if (threadIdx.x < nranks && w.flag[threadIdx.x]) {
        while (ld_acquire_sys(w.flag[threadIdx.x]) < seq) { __nanosleep(20); }
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        double s = 0.0;
#pragma unroll 1
        for (int r = 0; r < nranks; ++r) s += __ldcv(staged + (int64_t)r * stride + i);
        dst[i] = s;
        if (host_out && i == 0) *host_out = s;
    }
}


// ---- device helper: pack this rank's boundary values into its halo segment ----------
__global__ void __launch_bounds__(kThreads)
halo_pack_kernel(int64_t count, const int *__restrict__ send_idx, const double *__restrict__ x,
                 double *__restrict__ segment) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) segment[i] = x[send_idx[i]];
}

} // namespace b200
