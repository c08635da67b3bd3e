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
//   level vectors   uniform contiguous blocks of B = ceil4(n / P) rows
//   A_l (square)    own rows; local columns -> [0, n_loc), remote columns ->
//                   n_loc + owner*S + position in the owner's send list.  The
//                   halo is ONE in-place ncclAllGather of every rank's packed
//                   boundary values (S doubles per rank) into a P*S buffer the
//                   kernel gathers from directly -- no unpack step.
//   P_l (prolong)   own rows, global coarse columns; the coarse vector arrives by
//                   ncclAllGather (coarse level distributed) or ncclBroadcast
//                   (coarse level lives on rank 0).
//   R_l (restrict)  all rows, own columns only; partial sums leave by
//                   ncclReduceScatter / ncclReduce.
//   inner products  local kernel + ncclAllReduce of one double.
//   anything smaller than the threshold lives on rank 0 only; other ranks hold
//   "ghost" handles whose operations are no-ops.
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
    int64_t val_offset = 0;               // values are val[val_offset .. ) of the input when
    bool    val_contiguous = true;        //   contiguous, else `val` below holds a copy
    std::vector<double>  val;
    // square operators only
    int64_t S = 0;                        // send-list slots per rank (max over ranks)
    std::vector<int64_t> send_idx;        // local indices this rank contributes, in slot order
    int64_t n_loc = 0;
};

// A_l: rows of `rank`; columns -> local / halo slots.
template <class Ptr, class Col>
static void split_square(const Partition &part, int rank, const Ptr *ptr, const Col *col,
                         SplitMatrix &out) {
    const int64_t n = part.n;
    const int P = part.P;
    // mark[c] = 1 if some row not owned by owner(c) references column c
    std::vector<unsigned char> mark((size_t)n, 0);
    for (int p = 0; p < P; ++p) {
        const int64_t lo = part.lo(p), hi = part.hi(p);
        for (int64_t r = lo; r < hi; ++r)
            for (int64_t e = (int64_t)ptr[r]; e < (int64_t)ptr[r + 1]; ++e) {
                const int64_t c = (int64_t)col[e];
                if (c < lo || c >= hi) mark[(size_t)c] = 1;
            }
    }
    // slot of every marked column inside its owner's send list, S = longest list
    std::vector<int64_t> slot((size_t)n, -1);
    int64_t S = 0;
    for (int p = 0; p < P; ++p) {
        int64_t k = 0;
        for (int64_t c = part.lo(p); c < part.hi(p); ++c)
            if (mark[(size_t)c]) slot[(size_t)c] = k++;
        S = std::max(S, k);
    }
    S = (S + 1) & ~int64_t(1);            // keep every rank's segment 16-byte aligned
    const int64_t lo = part.lo(rank), hi = part.hi(rank), n_loc = hi - lo;
    out.S = S;
    out.n_loc = n_loc;
    out.nrows = n_loc;
    out.ncols = n_loc + (int64_t)P * S;
    out.send_idx.clear();
    for (int64_t c = lo; c < hi; ++c)
        if (mark[(size_t)c]) out.send_idx.push_back(c - lo);
    out.ptr.resize((size_t)n_loc + 1);
    const int64_t e0 = n_loc ? (int64_t)ptr[lo] : 0;
    const int64_t e1 = n_loc ? (int64_t)ptr[hi] : 0;
    out.col.resize((size_t)(e1 - e0));
    for (int64_t r = lo; r <= hi && n_loc; ++r) out.ptr[(size_t)(r - lo)] = (int64_t)ptr[r] - e0;
    if (!n_loc) out.ptr[0] = 0;
    for (int64_t e = e0; e < e1; ++e) {
        const int64_t c = (int64_t)col[e];
        if (c >= lo && c < hi) out.col[(size_t)(e - e0)] = c - lo;
        else out.col[(size_t)(e - e0)] = n_loc + (int64_t)part.owner(c) * S + slot[(size_t)c];
    }
    out.val_contiguous = true;
    out.val_offset = e0;
}

// P_l: rows of `rank` (fine partition), columns stay global coarse indices.
template <class Ptr, class Col>
static void split_prolong(const Partition &fine, int rank, int64_t ncols, const Ptr *ptr,
                          const Col *col, SplitMatrix &out) {
    const int64_t lo = fine.lo(rank), hi = fine.hi(rank), n_loc = hi - lo;
    out.nrows = n_loc;
    out.ncols = ncols;
    out.n_loc = n_loc;
    out.ptr.resize((size_t)n_loc + 1);
    const int64_t e0 = n_loc ? (int64_t)ptr[lo] : 0;
    const int64_t e1 = n_loc ? (int64_t)ptr[hi] : 0;
    for (int64_t r = lo; r <= hi && n_loc; ++r) out.ptr[(size_t)(r - lo)] = (int64_t)ptr[r] - e0;
    if (!n_loc) out.ptr[0] = 0;
    out.col.resize((size_t)(e1 - e0));
    for (int64_t e = e0; e < e1; ++e) out.col[(size_t)(e - e0)] = (int64_t)col[e];
    out.val_contiguous = true;
    out.val_offset = e0;
}

// R_l: all rows, only the columns `rank` owns (fine partition), remapped to local.
template <class Ptr, class Col>
static void split_restrict(const Partition &fine, int rank, int64_t nrows, const Ptr *ptr,
                           const Col *col, const double *val, SplitMatrix &out) {
    const int64_t lo = fine.lo(rank), hi = fine.hi(rank);
    out.nrows = nrows;
    out.ncols = hi - lo;
    out.n_loc = hi - lo;
    out.ptr.assign((size_t)nrows + 1, 0);
    for (int64_t r = 0; r < nrows; ++r) {
        int64_t k = 0;
        for (int64_t e = (int64_t)ptr[r]; e < (int64_t)ptr[r + 1]; ++e) {
            const int64_t c = (int64_t)col[e];
            k += (c >= lo && c < hi);
        }
        out.ptr[(size_t)r + 1] = out.ptr[(size_t)r] + k;
    }
    const int64_t nnz = out.ptr[(size_t)nrows];
    out.col.resize((size_t)nnz);
    out.val.resize((size_t)nnz);
    out.val_contiguous = false;
    int64_t k = 0;
    for (int64_t r = 0; r < nrows; ++r)
        for (int64_t e = (int64_t)ptr[r]; e < (int64_t)ptr[r + 1]; ++e) {
            const int64_t c = (int64_t)col[e];
            if (c >= lo && c < hi) {
                out.col[(size_t)k] = c - lo;
                out.val[(size_t)k] = val[e];
                ++k;
            }
        }
}

// dep[p][o] = rows owned by p (row partition) reference columns owned by o (column
// partition), p != o for square operators.  Identical on every rank (global matrix).
template <class Ptr, class Col>
static void dependency_matrix(const Partition &rows, const Partition &cols, const Ptr *ptr,
                              const Col *col, std::vector<unsigned char> &dep) {
    const int P = rows.P;
    dep.assign((size_t)P * P, 0);
    for (int p = 0; p < P; ++p) {
        unsigned char *row = dep.data() + (size_t)p * P;
        for (int64_t r = rows.lo(p); r < rows.hi(p); ++r)
            for (int64_t e = (int64_t)ptr[r]; e < (int64_t)ptr[r + 1]; ++e)
                row[cols.owner((int64_t)col[e])] = 1;
    }
}

} // namespace b200
