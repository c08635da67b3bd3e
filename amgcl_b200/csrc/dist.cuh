// dist.cuh -- row-partitioned operators across GPUs (one process per GPU, SPMD).
//
// Model: amgcl/mpi/distributed_matrix.hpp:51-557 (row-block partition, local +
// remote columns, halo exchange before the product) -- but the transport is
// NCCL over NVLink on device buffers, not host-staged MPI (distributed_matrix.hpp:
// 249-273), and the partitioning is transparent under the C ABI: every rank
// runs the same AMGCL program on the same (replicated) host hierarchy, and
// b200_csr_create_* / b200_vec_create keep only this rank's share of every
// object whose dimension reaches the distribution threshold.
//
//   level vectors   n >= threshold: uniform contiguous blocks of B = ceil4(n / P) rows;
//                   smaller: replicated -- every rank holds (and computes) the whole vector,
//                   as mpi::amg consolidates small levels (mpi/amg.hpp:430-465)
//   operators       every rank keeps WHOLE ROWS (split_rows below), so each row sum is formed
//                   on one GPU in the reference's order: results equal the single-GPU ones bit
//                   for bit, only the inner products see a different summation order.
//                   rows partitioned, columns partitioned  (A_l, P_l, R_l between partitioned
//                       levels): local columns + halo slots; the halo is the packed boundary
//                       values of every rank (S doubles per rank) in a P*S buffer the kernel
//                       gathers from directly -- no unpack step
//                   rows partitioned, columns replicated   (P_l from a small level): no exchange
//                   rows replicated,  columns partitioned  (R_l onto a small level): each rank
//                       computes a share of the rows (halo as above) and the shares are
//                       all-gathered into the replicated result
//                   both replicated: every rank computes everything, no exchange
//   inner products  partitioned vectors: reduced and all-reduced inside the producing kernel
//                   (reduce.cuh); replicated vectors: local
//
// This file holds the pure host logic (partition arithmetic, matrix splitting;
// exported through b200_dist_split_* so it can be tested on CPU with gloo) and
// the lazily dlopen'ed NCCL entry points.
#pragma once
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstring>

namespace b200 {

// ---------------------------------------------------------------------------
// NCCL, resolved at first use (single-GPU users never need libnccl)
// ---------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;

    bool load() {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = std::string("cannot load libnccl: ") + dlerror();
            return false;
        }
#define B200_NCCL_SYM(field, name)                                             \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));            \
    if (!field) { error = std::string("missing NCCL symbol ") + name; return false; }
        B200_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        B200_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        B200_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        B200_NCCL_SYM(AllGather, "ncclAllGather");
        B200_NCCL_SYM(AllReduce, "ncclAllReduce");
        B200_NCCL_SYM(Reduce, "ncclReduce");
        B200_NCCL_SYM(ReduceScatter, "ncclReduceScatter");
        B200_NCCL_SYM(Broadcast, "ncclBroadcast");
        B200_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef B200_NCCL_SYM
        return true;
    }
};

inline NcclApi &nccl() {
    static NcclApi api;
    return api;
}

// ---------------------------------------------------------------------------
// partition arithmetic
// ---------------------------------------------------------------------------
struct Partition {
    int64_t n = 0;      // global size
    int64_t B = 0;      // uniform block (multiple of 4); rank p owns [p*B, min(n,(p+1)*B))
    int     P = 1;
    Partition() {}
    Partition(int64_t n_, int P_) : n(n_), P(P_) {
        B = (n_ + P_ - 1) / P_;
        B = (B + 3) & ~int64_t(3);
        if (B < 4) B = 4;
    }
    int64_t lo(int p) const { return std::min<int64_t>(n, (int64_t)p * B); }
    int64_t hi(int p) const { return std::min<int64_t>(n, (int64_t)(p + 1) * B); }
    int64_t count(int p) const { return hi(p) - lo(p); }
    int owner(int64_t i) const { return (int)(i / B); }
};

// ---------------------------------------------------------------------------
// matrix splitting (pure host; indices of the input are Ptr / Col typed)
// ---------------------------------------------------------------------------
struct SplitMatrix {
    int64_t nrows = 0, ncols = 0;         // shape of the local matrix handed to the kernels
    std::vector<int64_t> ptr;             // [nrows+1]
    std::vector<int64_t> col;             // remapped columns
    int64_t val_offset = 0;               // values are val[val_offset .. ) of the input (always a
    bool    val_contiguous = true;        //   contiguous slice: a rank keeps whole rows)
    std::vector<double>  val;             // (unused; kept for the host view of b200_dist_split_*)
    // column side
    int64_t S = 0;                        // send-list slots per rank (max over ranks)
    std::vector<int64_t> send_idx;        // indices (local to my block of x) the others gather, in slot order
    int64_t n_loc = 0;                    // length of my block of x (columns < n_loc are local)
    std::vector<unsigned char> dep;       // [P*P] dep[p*P+o] = rows of p reference columns owned by o != p
};

// One rule for every operator of the hierarchy (A_l, P_l, R_l): a rank keeps WHOLE ROWS --
// the rows [rows.lo(rank), rows.hi(rank)) -- so every row sum is formed on one GPU in the
// reference's entry order, exactly as on a single GPU.  If the vector the operator is applied
// to is partitioned (cols_dist), a column owned by this rank (cols partition) becomes the local
// index c - cols.lo(rank); a column owned by rank o becomes n_loc + o*S + slot, where slot is
// the column's position in o's send list (the columns of o's block that ANY other rank
// references, in ascending order) -- the kernel reads those from the all-gathered halo buffer.
// If the vector is replicated (!cols_dist) columns are left alone.
// (model: mpi/distributed_matrix.hpp:388-435 splits A into A_loc / A_rem the same way.)
template <class Ptr, class Col>
static void split_rows(const Partition &rows, const Partition &cols, bool cols_dist, int rank,
                       const Ptr *ptr, const Col *col, SplitMatrix &out) {
    const int P = rows.P;
    const int64_t rlo = rows.lo(rank), rhi = rows.hi(rank), nr = rhi - rlo;
    const int64_t e0 = nr ? (int64_t)ptr[rlo] : 0;
    const int64_t e1 = nr ? (int64_t)ptr[rhi] : 0;
    out.nrows = nr;
    out.ptr.resize((size_t)nr + 1);
    for (int64_t r = rlo; r <= rhi && nr; ++r) out.ptr[(size_t)(r - rlo)] = (int64_t)ptr[r] - e0;
    if (!nr) out.ptr[0] = 0;
    out.col.resize((size_t)(e1 - e0));
    out.val_contiguous = true;
    out.val_offset = e0;
    out.send_idx.clear();
    out.dep.assign((size_t)P * P, 0);
    if (!cols_dist) {
        out.S = 0;
        out.n_loc = cols.n;
        out.ncols = cols.n;
        for (int64_t e = e0; e < e1; ++e) out.col[(size_t)(e - e0)] = (int64_t)col[e];
        return;
    }
    const int64_t nc = cols.n;
    // mark[c] = 1 if some row not owned by owner(c) references column c; dep[p][o]
    std::vector<unsigned char> mark((size_t)nc, 0);
    for (int p = 0; p < P; ++p) {
        const int64_t lo = rows.lo(p), hi = rows.hi(p);
        const int64_t clo = cols.lo(p), chi = cols.hi(p);
        unsigned char *drow = out.dep.data() + (size_t)p * P;
        const int64_t b = lo < hi ? (int64_t)ptr[lo] : 0, e = lo < hi ? (int64_t)ptr[hi] : 0;
        for (int64_t k = b; k < e; ++k) {
            const int64_t c = (int64_t)col[k];
            if (c < clo || c >= chi) {
                mark[(size_t)c] = 1;
                drow[cols.owner(c)] = 1;
            }
        }
    }
    // slot of every marked column inside its owner's send list, S = longest list
    std::vector<int64_t> slot((size_t)nc, -1);
    int64_t S = 0;
    for (int p = 0; p < P; ++p) {
        int64_t k = 0;
        for (int64_t c = cols.lo(p); c < cols.hi(p); ++c)
            if (mark[(size_t)c]) slot[(size_t)c] = k++;
        S = std::max(S, k);
    }
    S = (S + 1) & ~int64_t(1);            // keep every rank's segment 16-byte aligned
    const int64_t clo = cols.lo(rank), chi = cols.hi(rank), n_loc = chi - clo;
    out.S = S;
    out.n_loc = n_loc;
    out.ncols = n_loc + (int64_t)P * S;
    for (int64_t c = clo; c < chi; ++c)
        if (mark[(size_t)c]) out.send_idx.push_back(c - clo);
    for (int64_t e = e0; e < e1; ++e) {
        const int64_t c = (int64_t)col[e];
        if (c >= clo && c < chi) out.col[(size_t)(e - e0)] = c - clo;
        else out.col[(size_t)(e - e0)] = n_loc + (int64_t)cols.owner(c) * S + slot[(size_t)c];
    }
}

} // namespace b200
