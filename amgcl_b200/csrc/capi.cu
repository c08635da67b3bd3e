// capi.cu -- implementation of the C ABI declared in include/amgcl_b200.h.
//
// Host-side logic only: argument checking, the row-block plan, lazy-clear
// bookkeeping and kernel launches.  There is deliberately no CPU fallback: every
// entry point needs a CUDA device and reports B200_ECUDA otherwise.
#include "common.cuh"
#include "csr_kernels.cuh"
#include "vec_kernels.cuh"
#include "coarse_kernels.cuh"
#include "dist.cuh"
#include "peer.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
namespace b200 {

static void peer_release(b200_ctx_t ctx, void *local, void **peers);
static int peer_alloc(b200_ctx_t ctx, size_t bytes, void **local, void **peers);

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

int cuda_fail(cudaError_t rc, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s at %s:%d", (int)rc,
             cudaGetErrorString(rc), what, file, line);
    g_last_error = buf;
    cudaGetLastError();   // clear the sticky-less error state
    return rc == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;
}

// device guard: every entry point runs with the context's device current
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr size_t kProfMaxPairs = 1 << 17;

// Bracket one launch with a pair of events on the launching stream (profiling only).
struct ProfScope {
    b200_ctx_t ctx;
    bool on = false;
    size_t ev = 0;
    int64_t nrows, ncols, nnz;
    int mode;
    ProfScope(b200_ctx_t c, int mode_, int64_t nr, int64_t nc, int64_t nz)
        : ctx(c), nrows(nr), ncols(nc), nnz(nz), mode(mode_) {
        if (!ctx->profiling || ctx->prof_recs.size() >= kProfMaxPairs) return;
        while (ctx->prof_events.size() < ctx->prof_used + 2) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            ctx->prof_events.push_back(e);
        }
        ev = ctx->prof_used;
        if (cudaEventRecord(ctx->prof_events[ev], ctx->stream) != cudaSuccess) return;
        on = true;
    }
    ~ProfScope() {
        if (!on) return;
        if (cudaEventRecord(ctx->prof_events[ev + 1], ctx->stream) != cudaSuccess) return;
        ctx->prof_used += 2;
        ctx->prof_recs.push_back({nrows, ncols, nnz, mode, ev});
    }
};

// Lazy clear bookkeeping ------------------------------------------------------
static int materialize(b200_vec_t v) {
    if (v->zero_pending) {
        if (v->len) {
            ProfScope prof(v->ctx, B200_PROF_MEMSET, (int64_t)v->len, 1, 0);
            B200_CUDA(cudaMemsetAsync(v->ptr, 0, v->len * v->esz, v->ctx->stream));
        }
        v->zero_pending = false;
    }
    return B200_OK;
}
// pointer for reading (or read-modify-write)
static int rd(b200_vec_t v, const double **p) {
    int rc = materialize(v);
    *p = v->ptr;
    return rc;
}
// pointer for a full overwrite
static double *wr(b200_vec_t v) {
    v->zero_pending = false;
    return v->ptr;
}
// typed views (FP32 vectors keep their floats behind the same pointer)
template <class T> static inline T *tp(double *p) { return reinterpret_cast<T *>(p); }
template <class T> static inline const T *tp(const double *p) { return reinterpret_cast<const T *>(p); }
static inline bool all64(std::initializer_list<b200_vec_t> vs) {
    for (b200_vec_t v : vs) if (v->dtype != B200_F64) return false;
    return true;
}
static inline bool all32(std::initializer_list<b200_vec_t> vs) {
    for (b200_vec_t v : vs) if (v->dtype != B200_F32) return false;
    return true;
}

// CUDA-graph recording -----------------------------------------------------------
// The library keeps two pieces of host-side state per vector that decide WHICH kernels run and
// on WHICH addresses: the storage pointer (b200_relax trades x's storage with tmp's) and the
// lazy-clear flag.  A recorded graph bakes both in, so it remembers the state every object it
// touched had on entry (the graph may only be replayed from exactly that state) and the state
// the recorded calls left behind (applied after each replay).
struct GraphSlot {
    double **slot;      // &vec->ptr or &csr->scratch64
    bool    *zp;        // &vec->zero_pending (nullptr for operator scratch)
    double  *p0; bool z0;   // on entry
    double  *p1; bool z1;   // on exit
};
} // namespace b200

struct b200_graph_s {
    b200_ctx_t ctx = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    std::vector<b200::GraphSlot> slots;
    uint64_t destroy_epoch = 0, option_epoch = 0;
    uint64_t launches0 = 0;     // ctx->launches when recording started
    uint64_t launches = 0;      // kernels in the graph
    size_t   nodes = 0;
    uint64_t replays = 0;
};

namespace b200 {

static void touch_slot(b200_ctx_t ctx, double **slot, bool *zp) {
    b200_graph_s *g = ctx->recording;
    for (const GraphSlot &s : g->slots)
        if (s.slot == slot) return;
    g->slots.push_back({slot, zp, *slot, zp ? *zp : false, nullptr, false});
}
static inline void touch(b200_ctx_t ctx, std::initializer_list<b200_vec_t> vs) {
    if (!ctx->recording) return;
    for (b200_vec_t v : vs) {
        touch_slot(ctx, &v->ptr, &v->zero_pending);
        v->in_graph = true;
    }
}

static int grid_for(const b200_ctx_t ctx, size_t n_items, int per_thread_items) {
    // enough CTAs to cover the range once, capped at 8 CTAs per SM (2048 threads)
    size_t want = (n_items + (size_t)kThreads * per_thread_items - 1) /
                  ((size_t)kThreads * per_thread_items);
    size_t cap = (size_t)ctx->sm_count * 8;
    if (want < 1) want = 1;
    return (int)std::min(want, cap);
}

} // namespace b200

using namespace b200;

#define CHECK_CTX(ctx) B200_REQUIRE((ctx) != nullptr, "null context")
#define B200_NCCL(call)                                                        \
    do {                                                                       \
        ncclResult_t rc__ = (call);                                            \
        if (rc__ != ncclSuccess)                                               \
            return fail(B200_ENCCL, std::string("NCCL error in " #call ": ") + \
                                        nccl().GetErrorString(rc__));          \
    } while (0)
static inline ncclComm_t comm_of(b200_ctx_t ctx) { return static_cast<ncclComm_t>(ctx->comm); }
static inline bool same_layout(b200_vec_t a, b200_vec_t b) {
    return a->n == b->n && a->kind == b->kind && a->len == b->len;
}
#define B200_REQUIRE_F64_DIST(ctx, what)                                                    \
    B200_REQUIRE(!(ctx)->dist, what ": FP32 objects are not supported on a distributed context")
#define NOT_RECORDING(ctx, what)                                                        \
    B200_REQUIRE(!(ctx)->recording, what ": not allowed while a graph is being recorded")
#define GUARD(ctx)                                                             \
    DeviceGuard guard__((ctx)->device);                                        \
    if (!guard__.ok) return fail(B200_ECUDA, "cudaSetDevice failed")

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
extern "C" const char *b200_last_error(void) { return g_last_error.c_str(); }

extern "C" const char *b200_version(void) { return "amgcl_b200 0.1.0 sm_100a"; }

extern "C" int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int b200_ctx_create(int device, b200_ctx_t *out) {
    B200_REQUIRE(out != nullptr, "null output pointer");
    *out = nullptr;
    int ndev = 0;
    B200_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(B200_EINVAL, "no such CUDA device");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(B200_ECUDA, "cudaSetDevice failed");

    b200_ctx_s *ctx = new (std::nothrow) b200_ctx_s();
    if (!ctx) return fail(B200_ENOMEM, "out of host memory");
    ctx->device = device;
    if (const char *e = getenv("B200_PDL")) ctx->opt_pdl = atoi(e) ? 1 : 0;
    if (const char *e = getenv("B200_CYCLE_GRAPH")) ctx->opt_cycle_graph = atoi(e) ? 1 : 0;
    if (const char *e = getenv("B200_GRAPH_PDL")) ctx->opt_graph_pdl = atoi(e) ? 1 : 0;
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        delete ctx;
        return fail(B200_ECUDA, "amgcl_b200 needs an sm_100a (Blackwell B200) device");
    }
    B200_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    B200_CUDA(cudaMalloc(&ctx->dot_partial, kDotMaxBlocks * sizeof(double)));
    B200_CUDA(cudaMalloc(&ctx->dot_ticket, sizeof(unsigned int)));
    B200_CUDA(cudaMemset(ctx->dot_ticket, 0, sizeof(unsigned int)));
    B200_CUDA(cudaHostAlloc(&ctx->dot_result_h, 8 * sizeof(double), cudaHostAllocMapped));
    B200_CUDA(cudaHostGetDevicePointer(&ctx->dot_result_d, ctx->dot_result_h, 0));
    B200_CUDA(cudaMalloc(&ctx->dot_dev, 2 * sizeof(double)));
    *out = ctx;
    return B200_OK;
}

extern "C" int b200_ctx_destroy(b200_ctx_t ctx) {
    if (!ctx) return B200_OK;
    GUARD(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->dot_partial) cudaFree(ctx->dot_partial);
    if (ctx->dot_ticket) cudaFree(ctx->dot_ticket);
    if (ctx->dot_result_h) cudaFreeHost(ctx->dot_result_h);
    if (ctx->dot_dev) cudaFree(ctx->dot_dev);
    if (ctx->push_ticket) cudaFree(ctx->push_ticket);
    if (ctx->ipc_dev) cudaFree(ctx->ipc_dev);
    if (ctx->dot_pb_local) {
        for (int q = 0; q < ctx->nranks; ++q)
            if (q != ctx->rank && ctx->dot_pb_peer[q]) cudaIpcCloseMemHandle(ctx->dot_pb_peer[q]);
        cudaFree(ctx->dot_pb_local);
    }
    for (void *ptr : ctx->deferred_free) cudaFree(ptr);
    if (ctx->comm && nccl().handle) nccl().CommDestroy(comm_of(ctx));
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    delete ctx;
    return B200_OK;
}

extern "C" int b200_ctx_default(b200_ctx_t *out) {
    B200_REQUIRE(out != nullptr, "null output pointer");
    static std::mutex mtx;
    static b200_ctx_t def = nullptr;
    std::lock_guard<std::mutex> lock(mtx);
    if (!def) {
        int dev = 0;
        B200_CUDA(cudaGetDevice(&dev));
        int rc = b200_ctx_create(dev, &def);
        if (rc != B200_OK) return rc;
    }
    *out = def;
    return B200_OK;
}

extern "C" int b200_ctx_set_stream(b200_ctx_t ctx, void *cuda_stream) {
    CHECK_CTX(ctx);
    NOT_RECORDING(ctx, "stream change");
    ctx->option_epoch++;
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return B200_OK;
}

extern "C" int b200_ctx_get_stream(b200_ctx_t ctx, void **cuda_stream) {
    CHECK_CTX(ctx);
    B200_REQUIRE(cuda_stream != nullptr, "null output pointer");
    *cuda_stream = ctx->stream;
    return B200_OK;
}

extern "C" int b200_ctx_device(b200_ctx_t ctx, int *device) {
    CHECK_CTX(ctx);
    B200_REQUIRE(device != nullptr, "null output pointer");
    *device = ctx->device;
    return B200_OK;
}

extern "C" int b200_ctx_sync(b200_ctx_t ctx) {
    CHECK_CTX(ctx);
    NOT_RECORDING(ctx, "sync");
    GUARD(ctx);
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_ctx_launch_count(b200_ctx_t ctx, uint64_t *count) {
    CHECK_CTX(ctx);
    B200_REQUIRE(count != nullptr, "null output pointer");
    *count = ctx->launches;
    return B200_OK;
}

extern "C" int b200_ctx_reset_launch_count(b200_ctx_t ctx) {
    CHECK_CTX(ctx);
    ctx->launches = 0;
    return B200_OK;
}

extern "C" int b200_profile_begin(b200_ctx_t ctx) {
    CHECK_CTX(ctx);
    NOT_RECORDING(ctx, "profiling");
    ctx->prof_used = 0;
    ctx->prof_recs.clear();
    ctx->profiling = true;
    return B200_OK;
}

extern "C" int b200_profile_end(b200_ctx_t ctx, b200_profile_entry *out, int64_t capacity,
                                int64_t *count) {
    CHECK_CTX(ctx);
    B200_REQUIRE(count != nullptr, "null output pointer");
    GUARD(ctx);
    ctx->profiling = false;
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<b200_profile_entry> agg;
    for (const auto &r : ctx->prof_recs) {
        float ms = 0.f;
        B200_CUDA(cudaEventElapsedTime(&ms, ctx->prof_events[r.ev], ctx->prof_events[r.ev + 1]));
        b200_profile_entry *hit = nullptr;
        for (auto &a : agg)
            if (a.nrows == r.nrows && a.ncols == r.ncols && a.nnz == r.nnz && a.mode == r.mode) {
                hit = &a;
                break;
            }
        if (!hit) {
            agg.push_back({r.nrows, r.ncols, r.nnz, r.mode, 0, 0.0, 1e30});
            hit = &agg.back();
        }
        hit->launches += 1;
        hit->total_ms += ms;
        if (ms < hit->min_ms) hit->min_ms = ms;
    }
    ctx->prof_recs.clear();
    ctx->prof_used = 0;
    *count = (int64_t)agg.size();
    if (out) {
        const int64_t m = std::min<int64_t>(capacity, (int64_t)agg.size());
        for (int64_t i = 0; i < m; ++i) out[i] = agg[(size_t)i];
    }
    return B200_OK;
}

// ---------------------------------------------------------------------------
// multi-GPU
// ---------------------------------------------------------------------------
extern "C" int b200_nccl_unique_id(char *id, size_t size) {
    B200_REQUIRE(id != nullptr && size >= sizeof(ncclUniqueId), "id buffer must hold 128 bytes");
    if (!nccl().load()) return fail(B200_ENCCL, nccl().error);
    ncclUniqueId uid;
    B200_NCCL(nccl().GetUniqueId(&uid));
    memset(id, 0, size);
    memcpy(id, &uid, sizeof(uid));
    return B200_OK;
}

extern "C" int b200_dist_init(b200_ctx_t ctx, const char *id, size_t size, int nranks, int rank,
                              int64_t dist_min_rows) {
    CHECK_CTX(ctx);
    B200_REQUIRE(id != nullptr && size >= sizeof(ncclUniqueId), "id buffer must hold 128 bytes");
    B200_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    B200_REQUIRE(dist_min_rows >= 1, "dist_min_rows must be positive");
    B200_REQUIRE(!ctx->dist, "context is already distributed");
    GUARD(ctx);
    if (!nccl().load()) return fail(B200_ENCCL, nccl().error);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    B200_NCCL(nccl().CommInitRank(&comm, nranks, uid, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    ctx->dist_min_rows = dist_min_rows;
    ctx->dist = true;
    B200_CUDA(cudaMalloc(&ctx->push_ticket, sizeof(unsigned int)));
    B200_CUDA(cudaMemset(ctx->push_ticket, 0, sizeof(unsigned int)));
    B200_CUDA(cudaMalloc(&ctx->ipc_dev, (size_t)kMaxRanks * sizeof(cudaIpcMemHandle_t)));

    // peer-memory exchange: try to map a small buffer of every peer; agree collectively
    ctx->p2p = false;
    if (ctx->opt_p2p && nranks > 1 && nranks <= kMaxRanks) {
        int ok = peer_alloc(ctx, kFlagBytes + 2 * 256, &ctx->dot_pb_local, ctx->dot_pb_peer) == B200_OK;
        int *flag_d = reinterpret_cast<int *>(ctx->ipc_dev);
        B200_CUDA(cudaMemcpy(flag_d, &ok, sizeof(int), cudaMemcpyHostToDevice));
        B200_NCCL(nccl().AllReduce(flag_d, flag_d, 1, ncclInt, ncclMin, comm, ctx->stream));
        B200_CUDA(cudaMemcpyAsync(&ok, flag_d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->p2p = ok != 0;
        if (!ctx->p2p) cudaGetLastError();
    }
    return B200_OK;
}

extern "C" int b200_dist_info(b200_ctx_t ctx, int *rank, int *nranks, int64_t *dist_min_rows,
                              int *p2p) {
    CHECK_CTX(ctx);
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    if (dist_min_rows) *dist_min_rows = ctx->dist ? ctx->dist_min_rows : 0;
    if (p2p) *p2p = ctx->p2p ? 1 : 0;
    return B200_OK;
}

// ---- pure host view of the partitioning (no device, no NCCL): for CPU tests ------------
struct b200_split_s {
    SplitMatrix m;
    std::vector<double> val;
    int kind;
};

extern "C" int b200_dist_split_i64(int kind, int nranks, int rank, int64_t nrows, int64_t ncols,
                                   const int64_t *ptr, const int64_t *col, const double *val,
                                   b200_split_t *out) {
    B200_REQUIRE(out != nullptr && ptr != nullptr, "null argument");
    B200_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    B200_REQUIRE(kind >= 1 && kind <= 3, "kind must be 1 (square), 2 (prolong) or 3 (restrict)");
    b200_split_s *sp = new (std::nothrow) b200_split_s();
    if (!sp) return fail(B200_ENOMEM, "out of host memory");
    sp->kind = kind;
    if (kind == B200_CK_SQUARE) {
        if (nrows != ncols) { delete sp; return fail(B200_EINVAL, "square operator expected"); }
        split_square(Partition(nrows, nranks), rank, ptr, col, sp->m);
    } else if (kind == B200_CK_PROLONG) {
        split_prolong(Partition(nrows, nranks), rank, ncols, ptr, col, sp->m);
    } else {
        split_restrict(Partition(ncols, nranks), rank, nrows, ptr, col, val, sp->m);
    }
    const int64_t nnz = sp->m.ptr.back();
    if (sp->m.val_contiguous) sp->val.assign(val + sp->m.val_offset, val + sp->m.val_offset + nnz);
    else sp->val = sp->m.val;
    *out = sp;
    return B200_OK;
}

extern "C" int b200_split_info(b200_split_t sp, int64_t *nrows, int64_t *ncols, int64_t *nnz,
                               int64_t *n_loc, int64_t *slots, int64_t *n_send) {
    B200_REQUIRE(sp != nullptr, "null argument");
    if (nrows) *nrows = sp->m.nrows;
    if (ncols) *ncols = sp->m.ncols;
    if (nnz) *nnz = sp->m.ptr.back();
    if (n_loc) *n_loc = sp->m.n_loc;
    if (slots) *slots = sp->m.S;
    if (n_send) *n_send = (int64_t)sp->m.send_idx.size();
    return B200_OK;
}

extern "C" int b200_split_copy(b200_split_t sp, int64_t *ptr, int64_t *col, double *val,
                               int64_t *send_idx) {
    B200_REQUIRE(sp != nullptr, "null argument");
    if (ptr) memcpy(ptr, sp->m.ptr.data(), sp->m.ptr.size() * sizeof(int64_t));
    if (col) memcpy(col, sp->m.col.data(), sp->m.col.size() * sizeof(int64_t));
    if (val) memcpy(val, sp->val.data(), sp->val.size() * sizeof(double));
    if (send_idx) memcpy(send_idx, sp->m.send_idx.data(), sp->m.send_idx.size() * sizeof(int64_t));
    return B200_OK;
}

extern "C" int b200_split_destroy(b200_split_t sp) {
    delete sp;
    return B200_OK;
}

extern "C" int b200_partition(int64_t n, int nranks, int rank, int64_t *block, int64_t *lo,
                              int64_t *hi) {
    B200_REQUIRE(n >= 0 && nranks >= 1 && rank >= 0 && rank < nranks, "bad argument");
    const Partition part(n, nranks);
    if (block) *block = part.B;
    if (lo) *lo = part.lo(rank);
    if (hi) *hi = part.hi(rank);
    return B200_OK;
}

static int64_t *option_slot(b200_ctx_t ctx, const char *key) {
    if (!key) return nullptr;
    if (!strcmp(key, "spmv_variant")) return &ctx->opt_spmv_variant;
    if (!strcmp(key, "fuse_relax")) return &ctx->opt_fuse_relax;
    if (!strcmp(key, "zero_shortcut")) return &ctx->opt_zero_shortcut;
    if (!strcmp(key, "nnz_cap")) return &ctx->opt_nnz_cap;
    if (!strcmp(key, "lanes")) return &ctx->opt_lanes;
    if (!strcmp(key, "ctas_per_sm")) return &ctx->opt_ctas_per_sm;
    if (!strcmp(key, "stages")) return &ctx->opt_stages;
    if (!strcmp(key, "p2p")) return &ctx->opt_p2p;
    if (!strcmp(key, "pdl")) return &ctx->opt_pdl;
    if (!strcmp(key, "cycle_graph")) return &ctx->opt_cycle_graph;
    if (!strcmp(key, "graph_pdl")) return &ctx->opt_graph_pdl;
    return nullptr;
}

extern "C" int b200_ctx_set_option(b200_ctx_t ctx, const char *key, int64_t value) {
    CHECK_CTX(ctx);
    int64_t *slot = option_slot(ctx, key);
    if (!slot) return fail(B200_EINVAL, std::string("unknown option: ") + (key ? key : "(null)"));
    if (slot == &ctx->opt_nnz_cap) {
        if (value < 256 || value > kNnzCapMax || (value % 8))
            return fail(B200_EINVAL, "nnz_cap must be a multiple of 8 in [256, 6144]");
    } else if (slot == &ctx->opt_lanes) {
        if (value != 0 && (value < 1 || value > 32 || (value & (value - 1))))
            return fail(B200_EINVAL, "lanes must be 0 (auto) or a power of two <= 32");
    } else if (slot == &ctx->opt_stages) {
        if (value < 1 || value > 8) return fail(B200_EINVAL, "stages must be in [1, 8]");
    } else if (slot == &ctx->opt_ctas_per_sm) {
        if (value < 1 || value > 8) return fail(B200_EINVAL, "ctas_per_sm must be in [1, 8]");
    } else if (slot == &ctx->opt_spmv_variant) {
        if (value < 0 || value > 1) return fail(B200_EINVAL, "spmv_variant must be 0 or 1");
    }
    B200_REQUIRE(!ctx->recording, "options cannot change while a graph is being recorded");
    if (*slot != value) ctx->option_epoch++;      // recorded graphs were built with the old value
    *slot = value;
    return B200_OK;
}

extern "C" int b200_ctx_get_option(b200_ctx_t ctx, const char *key, int64_t *value) {
    CHECK_CTX(ctx);
    B200_REQUIRE(value != nullptr, "null output pointer");
    int64_t *slot = option_slot(ctx, key);
    if (!slot) return fail(B200_EINVAL, std::string("unknown option: ") + (key ? key : "(null)"));
    *value = *slot;
    return B200_OK;
}

// ---------------------------------------------------------------------------
// vectors
// ---------------------------------------------------------------------------
static int vec_create_typed(b200_ctx_t ctx, size_t n, int dtype, b200_vec_t *out) {
    CHECK_CTX(ctx);
    NOT_RECORDING(ctx, "vector creation");
    B200_REQUIRE(out != nullptr, "null output pointer");
    *out = nullptr;
    if (dtype == B200_F32) B200_REQUIRE_F64_DIST(ctx, "b200_vec_create_f32");
    GUARD(ctx);
    b200_vec_s *v = new (std::nothrow) b200_vec_s();
    if (!v) return fail(B200_ENOMEM, "out of host memory");
    v->ctx = ctx;
    v->n = n;
    v->owned = true;
    v->dtype = dtype;
    v->esz = dtype == B200_F32 ? sizeof(float) : sizeof(double);
    if (ctx->dist && (int64_t)n >= ctx->dist_min_rows) {
        const Partition part((int64_t)n, ctx->nranks);
        v->kind = B200_VK_DIST;
        v->off = (size_t)part.lo(ctx->rank);
        v->len = (size_t)part.count(ctx->rank);
        v->cap = (size_t)part.B;
    } else if (ctx->dist && ctx->rank != 0) {
        v->kind = B200_VK_GHOST;      // the object lives on rank 0; operations here are no-ops
        v->len = 0;
        v->cap = 0;
        v->zero_pending = false;
        *out = v;
        return B200_OK;
    } else {
        v->kind = B200_VK_LOCAL;
        v->len = n;
        v->cap = n;
    }
    // +2 doubles of padding so 16-byte vector accesses of the tail stay in bounds
    cudaError_t rc = cudaMalloc(&v->ptr, (v->cap + 4) * v->esz);
    if (rc != cudaSuccess) {
        delete v;
        return cuda_fail(rc, "cudaMalloc(vector)", __FILE__, __LINE__);
    }
    if (v->cap > v->len) {
        // padding of a partial block takes part in collectives: keep it zero
        rc = cudaMemsetAsync(v->ptr + v->len, 0, (v->cap - v->len) * sizeof(double), ctx->stream);
        if (rc != cudaSuccess) {
            cudaFree(v->ptr);
            delete v;
            return cuda_fail(rc, "cudaMemsetAsync(vector padding)", __FILE__, __LINE__);
        }
    }
    v->zero_pending = true;   // logically zero; memset only if somebody looks
    *out = v;
    return B200_OK;
}

extern "C" int b200_vec_create(b200_ctx_t ctx, size_t n, b200_vec_t *out) {
    return vec_create_typed(ctx, n, B200_F64, out);
}

extern "C" int b200_vec_create_f32(b200_ctx_t ctx, size_t n, b200_vec_t *out) {
    return vec_create_typed(ctx, n, B200_F32, out);
}

extern "C" int b200_vec_dtype(b200_vec_t v, int *dtype) {
    B200_REQUIRE(v && dtype, "null argument");
    *dtype = v->dtype;
    return B200_OK;
}

extern "C" int b200_vec_wrap(b200_ctx_t ctx, double *device_ptr, size_t n, b200_vec_t *out) {
    CHECK_CTX(ctx);
    B200_REQUIRE(out != nullptr, "null output pointer");
    B200_REQUIRE(device_ptr != nullptr || n == 0, "null device pointer");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(device_ptr) & 7) == 0, "device pointer not 8-byte aligned");
    B200_REQUIRE(!ctx->dist, "b200_vec_wrap is not available on a distributed context");
    b200_vec_s *v = new (std::nothrow) b200_vec_s();
    if (!v) return fail(B200_ENOMEM, "out of host memory");
    v->ctx = ctx;
    v->n = n;
    v->len = n;
    v->cap = n;
    v->ptr = device_ptr;
    v->owned = false;
    v->zero_pending = false;
    *out = v;
    return B200_OK;
}

extern "C" int b200_vec_destroy(b200_vec_t v) {
    if (!v) return B200_OK;
    if (v->in_graph) v->ctx->destroy_epoch++;      // recorded graphs that refer to it are dead
    if (b200_graph_s *g = v->ctx->recording) {
        // (e.g. a garbage-collected handle of the host language)  The recording may already
        // use the storage: it is released after the recorded calls have run.
        for (size_t i = 0; i < g->slots.size();)
            if (g->slots[i].slot == &v->ptr) g->slots.erase(g->slots.begin() + i);
            else ++i;
        if (v->owned && v->ptr) v->ctx->graph_deferred.push_back(v->ptr);
        delete v;
        return B200_OK;
    }
    GUARD(v->ctx);
    if (v->owned && v->ptr) {
        // cudaFree synchronises the device, so no kernel can still be using it
        cudaFree(v->ptr);
    }
    delete v;
    return B200_OK;
}

extern "C" int b200_vec_size(b200_vec_t v, size_t *n) {
    B200_REQUIRE(v && n, "null argument");
    *n = v->n;
    return B200_OK;
}

extern "C" int b200_vec_bytes(b200_vec_t v, size_t *bytes) {
    B200_REQUIRE(v && bytes, "null argument");
    *bytes = v->len * v->esz;
    return B200_OK;
}

extern "C" int b200_vec_data(b200_vec_t v, double **device_ptr) {
    B200_REQUIRE(v && device_ptr, "null argument");
    B200_REQUIRE(v->dtype == B200_F64, "b200_vec_data: FP64 vectors only");
    GUARD(v->ctx);
    int rc = materialize(v);
    *device_ptr = v->ptr;
    return rc;
}

extern "C" int b200_vec_upload_f32(b200_vec_t v, const float *host, size_t n) {
    B200_REQUIRE(v && (host || n == 0), "null argument");
    NOT_RECORDING(v->ctx, "host transfer");
    B200_REQUIRE(n == v->n, "size mismatch in vector upload");
    B200_REQUIRE(v->dtype == B200_F32 && v->kind == B200_VK_LOCAL, "upload_f32: FP32 local vector expected");
    GUARD(v->ctx);
    if (n) {
        B200_CUDA(cudaMemcpyAsync(wr(v), host, n * sizeof(float), cudaMemcpyHostToDevice, v->ctx->stream));
        B200_CUDA(cudaStreamSynchronize(v->ctx->stream));
    }
    v->zero_pending = false;
    return B200_OK;
}

extern "C" int b200_vec_download_f32(b200_vec_t v, float *host, size_t n) {
    B200_REQUIRE(v && (host || n == 0), "null argument");
    NOT_RECORDING(v->ctx, "host transfer");
    B200_REQUIRE(n == v->n, "size mismatch in vector download");
    B200_REQUIRE(v->dtype == B200_F32 && v->kind == B200_VK_LOCAL, "download_f32: FP32 local vector expected");
    GUARD(v->ctx);
    if (!n) return B200_OK;
    int rc = materialize(v);
    if (rc) return rc;
    B200_CUDA(cudaMemcpyAsync(host, v->ptr, n * sizeof(float), cudaMemcpyDeviceToHost, v->ctx->stream));
    B200_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return B200_OK;
}

extern "C" int b200_vec_upload(b200_vec_t v, const double *host, size_t n) {
    B200_REQUIRE(v && (host || n == 0), "null argument");
    NOT_RECORDING(v->ctx, "host transfer");
    B200_REQUIRE(n == v->n, "size mismatch in vector upload");
    B200_REQUIRE(v->dtype == B200_F64, "b200_vec_upload: FP64 vector expected (use _f32)");
    GUARD(v->ctx);
    if (v->kind == B200_VK_GHOST) return B200_OK;
    if (v->len) {
        // distributed: every rank is handed the full host vector and keeps its block
        B200_CUDA(cudaMemcpyAsync(wr(v), host + v->off, v->len * sizeof(double),
                                  cudaMemcpyHostToDevice, v->ctx->stream));
        B200_CUDA(cudaStreamSynchronize(v->ctx->stream));
    }
    v->zero_pending = false;
    return B200_OK;
}

extern "C" int b200_vec_download(b200_vec_t v, double *host, size_t n) {
    B200_REQUIRE(v && (host || n == 0), "null argument");
    NOT_RECORDING(v->ctx, "host transfer");
    B200_REQUIRE(n == v->n, "size mismatch in vector download");
    B200_REQUIRE(v->dtype == B200_F64, "b200_vec_download: FP64 vector expected (use _f32)");
    b200_ctx_t ctx = v->ctx;
    GUARD(ctx);
    if (!n) return B200_OK;
    if (v->kind == B200_VK_GHOST) {          // lives on rank 0 only
        memset(host, 0, n * sizeof(double));
        return B200_OK;
    }
    if (v->kind == B200_VK_DIST) {
        // every rank receives the complete vector: all-gather the blocks, then one D2H copy
        int rc = materialize(v);
        if (rc) return rc;
        double *full = nullptr;
        B200_CUDA(cudaMalloc(&full, (size_t)ctx->nranks * v->cap * sizeof(double)));
        ncclResult_t nrc = nccl().AllGather(v->ptr, full, v->cap, ncclDouble, comm_of(ctx), ctx->stream);
        if (nrc != ncclSuccess) {
            cudaFree(full);
            return fail(B200_ENCCL, std::string("ncclAllGather: ") + nccl().GetErrorString(nrc));
        }
        cudaError_t crc = cudaMemcpyAsync(host, full, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (crc == cudaSuccess) crc = cudaStreamSynchronize(ctx->stream);
        cudaFree(full);
        if (crc != cudaSuccess) return cuda_fail(crc, "download(distributed)", __FILE__, __LINE__);
        return B200_OK;
    }
    if (v->zero_pending) {          // nothing to fetch: the vector is zero
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        memset(host, 0, n * sizeof(double));
        return B200_OK;
    }
    B200_CUDA(cudaMemcpyAsync(host, v->ptr, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

// ---------------------------------------------------------------------------
// matrices
// ---------------------------------------------------------------------------
namespace b200 {

static int choose_lanes(double avg) {
    // measured on the B200 with the operators of a 192^3 Poisson hierarchy
    // (tools/gpu_check.py levels): few lanes per row keep many rows -- and so many
    // independent x-gathers -- in flight per CTA; wide groups only pay off for long rows
    if (avg <= 12.0) return 1;
    if (avg <= 40.0) return 2;
    if (avg <= 64.0) return 4;
    if (avg <= 160.0) return 8;
    if (avg <= 320.0) return 16;
    return 32;
}

// Row-block plan (pure host logic, also exported as b200_plan_i64 for tests):
// consecutive rows, starting at a multiple of four, are packed greedily while
// they fit `rows_cap` rows and `nnz_cap` non-zeros.  A block that still exceeds
// nnz_cap (a single quad of very long rows) is counted as "long" and is handled
// by the strided path of the kernels.
struct RowBlockPlan {
    int lanes = 1, rows_cap = 256, nnz_cap = 2048;
    int64_t nlong = 0;
    std::vector<int2> blk;   // {first row, first non-zero}; last entry = {nrows, nnz}
};

template <class Ptr>
static void build_plan(int64_t nrows, const Ptr *ptr, int lanes_opt, int nnz_cap,
                       RowBlockPlan &plan) {
    const int64_t nnz = nrows ? (int64_t)ptr[nrows] : 0;
    const double avg = nrows ? (double)nnz / (double)nrows : 0.0;
    const int lanes = lanes_opt ? lanes_opt : choose_lanes(avg);
    const int groups = kThreads / lanes;
    // rows per block: a multiple of the number of row groups, sized so a typical
    // block fills the stage
    int k = 1;
    if (avg > 0.0) k = (int)std::floor((double)nnz_cap / (avg * groups));
    k = std::max(1, std::min(k, kRowsCapMax / groups));
    int rows_cap = std::min(kRowsCapMax, groups * k);
    rows_cap = std::max(4, rows_cap & ~3);
    plan.lanes = lanes; plan.rows_cap = rows_cap; plan.nnz_cap = nnz_cap; plan.nlong = 0;
    plan.blk.clear();
    plan.blk.reserve((size_t)(nnz / std::max(1, nnz_cap / 2) + nrows / rows_cap + 16));
    int64_t r = 0;
    while (r < nrows) {
        const int64_t r0 = r;
        const int64_t e0 = (int64_t)ptr[r0];
        // always take the first quad, then grow quad by quad while it fits
        int64_t r1 = std::min<int64_t>(nrows, r0 + 4);
        while (r1 < nrows) {
            const int64_t rn = std::min<int64_t>(nrows, r1 + 4);
            if (rn - r0 > rows_cap) break;
            if ((int64_t)ptr[rn] - e0 > nnz_cap) break;
            r1 = rn;
        }
        if ((int64_t)ptr[r1] - e0 > nnz_cap) ++plan.nlong;
        plan.blk.push_back(make_int2((int)r0, (int)e0));
        r = r1;
    }
    plan.blk.push_back(make_int2((int)nrows, (int)nnz));
}

// Upload one CSR matrix exactly as the kernels will see it (indices narrowed to int32,
// row-block plan built).  Single-GPU matrices come straight through here; the
// distributed kinds hand in the local part produced by dist.cuh.
template <class Ptr, class Col, class Val>
static int csr_upload(b200_ctx_t ctx, int64_t nrows, int64_t ncols, const Ptr *ptr,
                      const Col *col, const Val *val, b200_csr_t *out) {
    CHECK_CTX(ctx);
    B200_REQUIRE(out != nullptr, "null output pointer");
    *out = nullptr;
    B200_REQUIRE(nrows >= 0 && ncols >= 0, "negative matrix dimension");
    B200_REQUIRE(ptr != nullptr, "null row pointer array");
    const int64_t imax = std::numeric_limits<int32_t>::max();
    if (nrows >= imax - 8 || ncols >= imax) return fail(B200_ERANGE, "matrix dimension exceeds int32");
    B200_REQUIRE(ptr[0] == 0, "ptr[0] must be 0");
    const int64_t nnz = (int64_t)ptr[nrows];
    if (nnz < 0) return fail(B200_EINVAL, "negative number of non-zeros");
    if (nnz >= imax - 8) return fail(B200_ERANGE, "number of non-zeros exceeds int32");
    B200_REQUIRE(nnz == 0 || (col != nullptr && val != nullptr), "null col/val array");
    GUARD(ctx);

    // ---- narrow indices, validate ------------------------------------------
    std::vector<int32_t> hptr((size_t)nrows + 1), hcol((size_t)nnz);
    for (int64_t i = 0; i <= nrows; ++i) {
        const int64_t p = (int64_t)ptr[i];
        if (i && p < (int64_t)ptr[i - 1]) return fail(B200_EINVAL, "row pointers not monotone");
        hptr[(size_t)i] = (int32_t)p;
    }
    for (int64_t e = 0; e < nnz; ++e) {
        const int64_t c = (int64_t)col[e];
        if (c < 0 || c >= ncols) return fail(B200_EINVAL, "column index out of range");
        hcol[(size_t)e] = (int32_t)c;
    }

    // ---- row-block plan -------------------------------------------------------
    RowBlockPlan plan;
    build_plan(nrows, hptr.data(), (int)ctx->opt_lanes, (int)ctx->opt_nnz_cap, plan);
    const int lanes = plan.lanes, rows_cap = plan.rows_cap, nnz_cap = plan.nnz_cap;
    const int64_t nlong = plan.nlong;
    std::vector<int2> &blk = plan.blk;
    const int64_t nblocks = (int64_t)blk.size() - 1;

    // ---- upload ---------------------------------------------------------------------
    b200_csr_s *A = new (std::nothrow) b200_csr_s();
    if (!A) return fail(B200_ENOMEM, "out of host memory");
    A->ctx = ctx; A->nrows = nrows; A->ncols = ncols; A->nnz = nnz;
    A->gl_rows = nrows; A->gl_cols = ncols; A->gl_nnz = nnz;
    A->dtype = std::is_same<Val, float>::value ? B200_F32 : B200_F64;
    A->lanes = lanes; A->rows_cap = rows_cap; A->nnz_cap = nnz_cap;
    A->nblocks = nblocks; A->nlong = nlong;
    // padding: bulk copies round sizes up to 16 bytes
    const size_t ptr_bytes = ((size_t)nrows + 1 + 8) * sizeof(int);
    const size_t col_bytes = ((size_t)nnz + 8) * sizeof(int);
    const size_t val_bytes = ((size_t)nnz + 8) * sizeof(Val);
    const size_t blk_bytes = ((size_t)nblocks + 1) * sizeof(int2);
    auto cleanup = [&]() {
        if (A->ptr) cudaFree(A->ptr);
        if (A->col) cudaFree(A->col);
        if (A->val) cudaFree(A->val);
        if (A->blk) cudaFree(A->blk);
        delete A;
    };
#define CSR_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t rc__ = (call);                                             \
        if (rc__ != cudaSuccess) {                                             \
            cleanup();                                                         \
            return cuda_fail(rc__, #call, __FILE__, __LINE__);                 \
        }                                                                      \
    } while (0)
    CSR_CUDA(cudaMalloc(&A->ptr, ptr_bytes));
    CSR_CUDA(cudaMalloc(&A->col, col_bytes));
    CSR_CUDA(cudaMalloc(&A->val, val_bytes));
    CSR_CUDA(cudaMalloc(&A->blk, blk_bytes));
    CSR_CUDA(cudaMemsetAsync(A->ptr, 0, ptr_bytes, ctx->stream));
    CSR_CUDA(cudaMemsetAsync(A->col, 0, col_bytes, ctx->stream));
    CSR_CUDA(cudaMemsetAsync(A->val, 0, val_bytes, ctx->stream));
    CSR_CUDA(cudaMemcpyAsync(A->ptr, hptr.data(), ((size_t)nrows + 1) * sizeof(int),
                             cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) {
        CSR_CUDA(cudaMemcpyAsync(A->col, hcol.data(), (size_t)nnz * sizeof(int),
                                 cudaMemcpyHostToDevice, ctx->stream));
        CSR_CUDA(cudaMemcpyAsync(A->val, val, (size_t)nnz * sizeof(Val),
                                 cudaMemcpyHostToDevice, ctx->stream));
    }
    CSR_CUDA(cudaMemcpyAsync(A->blk, blk.data(), blk_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CSR_CUDA(cudaStreamSynchronize(ctx->stream));   // host staging buffers die here
#undef CSR_CUDA
    A->bytes = ptr_bytes + col_bytes + val_bytes + blk_bytes;
    *out = A;
    return B200_OK;
}

static void csr_free(b200_csr_t A) {
    if (!A) return;
    if (A->ptr) cudaFree(A->ptr);
    if (A->col) cudaFree(A->col);
    if (A->val) cudaFree(A->val);
    if (A->blk) cudaFree(A->blk);
    if (A->send_idx) cudaFree(A->send_idx);
    if (A->blk_halo) cudaFree(A->blk_halo);
    if (A->halo_owned) cudaFree(A->halo_owned);
    if (A->cbuf) cudaFree(A->cbuf);
    if (A->scratch64) cudaFree(A->scratch64);
    if (A->pb_local) peer_release(A->ctx, A->pb_local, A->pb_peer);
    delete A;
}

// The public constructor: on a distributed context decide from the shape which
// kind of operator this is (see dist.cuh) and keep only this rank's share.
template <class Ptr, class Col>
static int csr_create_f32(b200_ctx_t ctx, int64_t nrows, int64_t ncols, const Ptr *ptr,
                          const Col *col, const float *val, b200_csr_t *out) {
    CHECK_CTX(ctx);
    NOT_RECORDING(ctx, "matrix creation");
    B200_REQUIRE_F64_DIST(ctx, "b200_csr_create_*_f32");
    return csr_upload(ctx, nrows, ncols, ptr, col, val, out);
}

template <class Ptr, class Col>
static int csr_create(b200_ctx_t ctx, int64_t nrows, int64_t ncols, const Ptr *ptr,
                      const Col *col, const double *val, b200_csr_t *out) {
    CHECK_CTX(ctx);
    NOT_RECORDING(ctx, "matrix creation");
    B200_REQUIRE(out != nullptr, "null output pointer");
    *out = nullptr;
    if (!ctx->dist) return csr_upload(ctx, nrows, ncols, ptr, col, val, out);

    B200_REQUIRE(nrows >= 0 && ncols >= 0 && ptr != nullptr && ptr[0] == 0, "bad CSR input");
    const int64_t nnz = (int64_t)ptr[nrows];
    B200_REQUIRE(nnz == 0 || (col != nullptr && val != nullptr), "null col/val array");
    const int64_t T = ctx->dist_min_rows;
    const bool rd = nrows >= T, cd = ncols >= T;
    int kind = B200_CK_LOCAL;
    if (rd && cd && nrows == ncols) kind = B200_CK_SQUARE;
    else if (rd && (!cd || nrows > ncols)) kind = B200_CK_PROLONG;
    else if (cd && (!rd || ncols > nrows)) kind = B200_CK_RESTRICT;

    if (kind == B200_CK_LOCAL) {
        if (ctx->rank == 0) return csr_upload(ctx, nrows, ncols, ptr, col, val, out);
        b200_csr_s *G = new (std::nothrow) b200_csr_s();     // ghost: lives on rank 0
        if (!G) return fail(B200_ENOMEM, "out of host memory");
        G->ctx = ctx; G->kind = B200_CK_GHOST;
        G->gl_rows = nrows; G->gl_cols = ncols; G->gl_nnz = nnz;
        *out = G;
        return B200_OK;
    }
    for (int64_t e = 0; e < nnz; ++e)
        if ((int64_t)col[e] < 0 || (int64_t)col[e] >= ncols)
            return fail(B200_EINVAL, "column index out of range");
    GUARD(ctx);

    SplitMatrix sp;
    const int P = ctx->nranks, rank = ctx->rank;
    if (kind == B200_CK_SQUARE) split_square(Partition(nrows, P), rank, ptr, col, sp);
    else if (kind == B200_CK_PROLONG) split_prolong(Partition(nrows, P), rank, ncols, ptr, col, sp);
    else split_restrict(Partition(ncols, P), rank, nrows, ptr, col, val, sp);

    b200_csr_t A = nullptr;
    const double *lval = sp.val_contiguous ? val + sp.val_offset : sp.val.data();
    int64_t kernel_cols = sp.ncols;
    if (kind == B200_CK_PROLONG && cd) kernel_cols = Partition(ncols, P).B * P;   // gathered blocks
    int rc = csr_upload(ctx, sp.nrows, kernel_cols, sp.ptr.data(), sp.col.data(), lval, &A);
    if (rc) return rc;
    A->kind = kind;
    A->gl_rows = nrows; A->gl_cols = ncols; A->gl_nnz = nnz;
    A->n_loc = sp.n_loc;
#define DCSR_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t rc__ = (call);                                             \
        if (rc__ != cudaSuccess) {                                             \
            csr_free(A);                                                       \
            return cuda_fail(rc__, #call, __FILE__, __LINE__);                 \
        }                                                                      \
    } while (0)
    if (kind == B200_CK_SQUARE) {
        A->S = sp.S;
        A->n_send = (int64_t)sp.send_idx.size();
        std::vector<int32_t> idx(sp.send_idx.begin(), sp.send_idx.end());
        DCSR_CUDA(cudaMalloc(&A->send_idx, std::max<size_t>(1, idx.size()) * sizeof(int)));
        DCSR_CUDA(cudaMalloc(&A->halo_owned, std::max<size_t>(2, (size_t)(P * sp.S)) * sizeof(double)));
        A->halo = A->halo_owned;
        DCSR_CUDA(cudaMemsetAsync(A->halo, 0, std::max<size_t>(2, (size_t)(P * sp.S)) * sizeof(double), ctx->stream));
        if (!idx.empty())
            DCSR_CUDA(cudaMemcpyAsync(A->send_idx, idx.data(), idx.size() * sizeof(int),
                                      cudaMemcpyHostToDevice, ctx->stream));
        A->bytes += idx.size() * sizeof(int) + (size_t)(P * sp.S) * sizeof(double);
        // which row blocks touch the halo (the same plan csr_upload just built)
        RowBlockPlan plan;
        build_plan(sp.nrows, sp.ptr.data(), A->lanes, A->nnz_cap, plan);
        std::vector<unsigned char> bh((size_t)std::max<int64_t>(1, A->nblocks), 0);
        if ((int64_t)plan.blk.size() - 1 == A->nblocks) {
            for (int64_t b = 0; b < A->nblocks; ++b) {
                const int64_t e0 = plan.blk[(size_t)b].y, e1 = plan.blk[(size_t)b + 1].y;
                for (int64_t e = e0; e < e1 && !bh[(size_t)b]; ++e)
                    if (sp.col[(size_t)e] >= sp.n_loc) bh[(size_t)b] = 1;
            }
        } else {
            std::fill(bh.begin(), bh.end(), 1);      // cannot happen; be safe: every block waits
        }
        DCSR_CUDA(cudaMalloc(&A->blk_halo, bh.size()));
        DCSR_CUDA(cudaMemcpyAsync(A->blk_halo, bh.data(), bh.size(), cudaMemcpyHostToDevice, ctx->stream));
        DCSR_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        // coarse-side buffer: gathered input of P, partial sums of R
        const bool coarse_dist = (kind == B200_CK_PROLONG) ? cd : rd;
        const int64_t nc = (kind == B200_CK_PROLONG) ? ncols : nrows;
        A->coarse_dist = coarse_dist;
        A->coarse_B = coarse_dist ? Partition(nc, P).B : nc;
        A->cbuf_n = coarse_dist ? A->coarse_B * P : nc;
        DCSR_CUDA(cudaMalloc(&A->cbuf, ((size_t)A->cbuf_n + 2) * sizeof(double)));
        DCSR_CUDA(cudaMemsetAsync(A->cbuf, 0, ((size_t)A->cbuf_n + 2) * sizeof(double), ctx->stream));
        A->bytes += (size_t)A->cbuf_n * sizeof(double);
    }
    DCSR_CUDA(cudaStreamSynchronize(ctx->stream));
#undef DCSR_CUDA

    // who exchanges with whom (identical on every rank: derived from the global matrix)
    {
        std::vector<unsigned char> dep;
        const int me = rank;
        if (kind == B200_CK_SQUARE) {
            const Partition part(nrows, P);
            dependency_matrix(part, part, ptr, col, dep);
            for (int q = 0; q < P; ++q) {
                A->need_from[q] = q != me && dep[(size_t)me * P + q];
                A->needed_by[q] = q != me && dep[(size_t)q * P + me];
            }
        } else if (kind == B200_CK_PROLONG) {
            // consumer p (fine rows) needs the coarse entries owned by o
            const Partition fine(nrows, P);
            const Partition coarse = A->coarse_dist ? Partition(ncols, P) : Partition(ncols, 1);
            if (A->coarse_dist) {
                dependency_matrix(fine, coarse, ptr, col, dep);
                for (int q = 0; q < P; ++q) {
                    A->need_from[q] = dep[(size_t)me * P + q];
                    A->needed_by[q] = dep[(size_t)q * P + me];
                }
            } else {
                for (int q = 0; q < P; ++q) {
                    const bool has = (int64_t)ptr[fine.hi(q)] > (int64_t)ptr[fine.lo(q)];
                    if (me == 0) A->needed_by[q] = q != 0 && has;
                    if (q == 0) A->need_from[0] = me != 0 && (int64_t)ptr[fine.hi(me)] > (int64_t)ptr[fine.lo(me)];
                }
            }
        } else {
            // producer r (fine columns) contributes to the coarse rows owned by o
            const Partition fine(ncols, P);
            if (A->coarse_dist) {
                const Partition coarse(nrows, P);
                dependency_matrix(coarse, fine, ptr, col, dep);     // dep[o][r]
                for (int q = 0; q < P; ++q) {
                    A->need_from[q] = dep[(size_t)me * P + q];       // r = q contributes to me
                    A->needed_by[q] = dep[(size_t)q * P + me];       // I contribute to owner q
                }
            } else {
                std::vector<unsigned char> has((size_t)P, 0);
                for (int64_t e = 0; e < nnz; ++e) has[(size_t)fine.owner((int64_t)col[e])] = 1;
                A->needed_by[0] = has[(size_t)me];
                if (me == 0)
                    for (int q = 0; q < P; ++q) A->need_from[q] = has[(size_t)q];
            }
        }
    }
    if (ctx->p2p) {
        size_t half = 16;
        if (kind == B200_CK_SQUARE) half = (size_t)P * (size_t)A->S * sizeof(double);
        else if (kind == B200_CK_PROLONG) half = (size_t)A->cbuf_n * sizeof(double);
        else if (A->coarse_dist) half = (size_t)P * (size_t)A->coarse_B * sizeof(double);
        else if (rank == 0) half = (size_t)P * (size_t)nrows * sizeof(double);
        half = (half + 255) & ~size_t(255);
        A->pb_half = half;
        // layout of the buffer a producer writes INTO (the owner's): equal to mine except for
        // a restriction onto a rank-0-only level, where only rank 0 holds the staging area
        A->pb_half_owner = half;
        if (kind == B200_CK_RESTRICT && !A->coarse_dist)
            A->pb_half_owner = (((size_t)P * (size_t)nrows * sizeof(double)) + 255) & ~size_t(255);
        int rc2 = peer_alloc(ctx, kFlagBytes + 2 * half, &A->pb_local, A->pb_peer);
        if (rc2) {
            csr_free(A);
            return rc2;
        }
        A->bytes += kFlagBytes + 2 * half;
    }
    *out = A;
    return B200_OK;
}

// Launch with programmatic stream serialization (PDL) when enabled: the kernel may be
// scheduled while its predecessor drains and orders itself with griddepcontrol.wait.
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(b200_ctx_t ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block,
                              size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed =
        (ctx->opt_pdl && (!ctx->recording || ctx->opt_graph_pdl)) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ---- launch one streaming pass over A ------------------------------------------------
// P = precision combination (csr_kernels.cuh).  Only FP64 carries the multi-GPU halo path
// and the one-block-per-CTA cross-check variant; the mixed-precision combinations use the
// persistent ring only.
template <int MODE, int L, bool HALO, class P>
static int launch_csr_LH(b200_ctx_t ctx, b200_csr_t A, const CsrArgsT<P> &args) {
    constexpr bool fp64 = std::is_same<P, PrecDD>::value;
    const StageLayout lay = stage_layout(A->rows_cap, A->nnz_cap, (int)sizeof(typename P::TV));
    if (fp64 && ctx->opt_spmv_variant == 0) {
        const int smem = kHeaderBytes + lay.bytes;
        static bool attr_set[64] = {};   // per instantiation and device
        if (!attr_set[ctx->device & 63]) {
            B200_CUDA(cudaFuncSetAttribute(csr_block_kernel<MODE, L, HALO, PrecDD>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[ctx->device & 63] = true;
        }
        // (only reachable with P == PrecDD)
        csr_block_kernel<MODE, L, HALO, PrecDD><<<(unsigned)A->nblocks, kThreads, smem, ctx->stream>>>(
            *reinterpret_cast<const CsrArgsT<PrecDD> *>(&args));
    } else {
        int stages = (int)ctx->opt_stages;
        const int max_smem = 227 * 1024;
        const int per_cta_budget = max_smem / (int)ctx->opt_ctas_per_sm - 1024;
        while (stages > 1 && kHeaderBytes + stages * lay.bytes > per_cta_budget) --stages;
        const int smem = kHeaderBytes + stages * lay.bytes;
        static bool attr_set[64] = {};
        if (!attr_set[ctx->device & 63]) {
            B200_CUDA(cudaFuncSetAttribute(csr_ring_kernel<MODE, L, HALO, P>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
            attr_set[ctx->device & 63] = true;
        }
        const int64_t cap = (int64_t)ctx->sm_count * ctx->opt_ctas_per_sm;
        const unsigned grid = (unsigned)std::min<int64_t>(A->nblocks, cap);
        B200_CUDA(launch_pdl(ctx, csr_ring_kernel<MODE, L, HALO, P>, dim3(grid), dim3(kThreads), (size_t)smem,
                             args, stages));
    }
    B200_CHECK_LAUNCH();
    ctx->launches++;
    return B200_OK;
}

template <int MODE, int L, class P>
static int launch_csr_L(b200_ctx_t ctx, b200_csr_t A, const CsrArgsT<P> &args) {
    if (std::is_same<P, PrecDD>::value && args.xh)
        return launch_csr_LH<MODE, L, true, PrecDD>(ctx, A, *reinterpret_cast<const CsrArgsT<PrecDD> *>(&args));
    return launch_csr_LH<MODE, L, false, P>(ctx, A, args);
}

template <int MODE, class P>
static int launch_csr(b200_ctx_t ctx, b200_csr_t A, const CsrArgsT<P> &args) {
    if (A->nblocks == 0) return B200_OK;
    if (ctx->recording) A->in_graph = true;
    ProfScope prof(ctx, MODE, A->nrows, A->ncols, A->nnz);
    switch (A->lanes) {
    case 1:  return launch_csr_L<MODE, 1>(ctx, A, args);
    case 2:  return launch_csr_L<MODE, 2>(ctx, A, args);
    case 4:  return launch_csr_L<MODE, 4>(ctx, A, args);
    case 8:  return launch_csr_L<MODE, 8>(ctx, A, args);
    case 16: return launch_csr_L<MODE, 16>(ctx, A, args);
    default: return launch_csr_L<MODE, 32>(ctx, A, args);
    }
}

template <class P>
static CsrArgsT<P> base_args_t(b200_csr_t A) {
    CsrArgsT<P> a;
    memset(&a, 0, sizeof(a));
    a.ptr = A->ptr; a.col = A->col; a.val = static_cast<const typename P::TV *>(A->val); a.blk = A->blk;
    a.nrows = (int)A->nrows; a.nblocks = (int)A->nblocks;
    a.rows_cap = A->rows_cap; a.nnz_cap = A->nnz_cap;
    return a;
}
static CsrArgs base_args(b200_csr_t A) { return base_args_t<PrecDD>(A); }

// ---- element-wise launch helpers ----------------------------------------------------------
// all streams of one element type T: 16-byte vector path when aligned
template <class F, bool RY, bool RZ, class T>
static int launch_ew(b200_ctx_t ctx, size_t n, F f, const T *x, const T *y, const T *z, T *out) {
    if (n == 0) return B200_OK;
    const bool vec_ok = aligned16(x) && aligned16(out) && (!RY || aligned16(y)) &&
                        (!RZ || aligned16(z));
    const int grid = grid_for(ctx, n, 16 / (int)sizeof(T) * 2);
    ProfScope prof(ctx, B200_PROF_VECTOR + (RY ? 1 : 0) + (RZ ? 1 : 0), (int64_t)n, 1, 0);
    B200_CUDA(launch_pdl(ctx, ew_kernel_same<F, RY, RZ, T>, dim3(grid), dim3(kThreads), 0, n, f, x, y, z, out,
                         vec_ok));
    B200_CHECK_LAUNCH();
    ctx->launches++;
    return B200_OK;
}
// mixed element types (FP32 inputs accumulated into an FP64 vector, precision-changing copy)
template <class F, bool RY, bool RZ, class TX, class TY, class TZ, class TO>
static int launch_ew_mixed(b200_ctx_t ctx, size_t n, F f, const TX *x, const TY *y, const TZ *z,
                           TO *out) {
    if (n == 0) return B200_OK;
    const int grid = grid_for(ctx, n, 2);
    ProfScope prof(ctx, B200_PROF_VECTOR + (RY ? 1 : 0) + (RZ ? 1 : 0), (int64_t)n, 1, 0);
    B200_CUDA(launch_pdl(ctx, ew_kernel<F, RY, RZ, TX, TY, TZ, TO>, dim3(grid), dim3(kThreads), 0, n, f, x, y,
                         z, out, false));
    B200_CHECK_LAUNCH();
    ctx->launches++;
    return B200_OK;
}

} // namespace b200

namespace b200 {

// Collective: every rank allocates `bytes` (zero filled) and maps the allocations of all
// its peers through CUDA IPC.  peers[rank] is the local pointer.
static int peer_alloc(b200_ctx_t ctx, size_t bytes, void **local, void **peers) {
    bytes = (bytes + 255) & ~size_t(255);
    B200_CUDA(cudaMalloc(local, bytes));
    B200_CUDA(cudaMemsetAsync(*local, 0, bytes, ctx->stream));
    cudaIpcMemHandle_t mine;
    B200_CUDA(cudaIpcGetMemHandle(&mine, *local));
    char *stage = static_cast<char *>(ctx->ipc_dev);
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    B200_CUDA(cudaMemcpyAsync(stage + ctx->rank * hs, &mine, hs, cudaMemcpyHostToDevice, ctx->stream));
    B200_NCCL(nccl().AllGather(stage + ctx->rank * hs, stage, hs, ncclChar, comm_of(ctx), ctx->stream));
    std::vector<cudaIpcMemHandle_t> all((size_t)ctx->nranks);
    B200_CUDA(cudaMemcpyAsync(all.data(), stage, hs * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < ctx->nranks; ++q) {
        if (q == ctx->rank) { peers[q] = *local; continue; }
        B200_CUDA(cudaIpcOpenMemHandle(&peers[q], all[(size_t)q], cudaIpcMemLazyEnablePeerAccess));
    }
    return B200_OK;
}

static void peer_release(b200_ctx_t ctx, void *local, void **peers) {
    if (!local) return;
    for (int q = 0; q < ctx->nranks; ++q)
        if (q != ctx->rank && peers[q]) cudaIpcCloseMemHandle(peers[q]);
    // peers may still have this allocation mapped: keep it until the context dies
    ctx->deferred_free.push_back(local);
}

static inline unsigned long long *flag_at(void *base, int parity, int slot) {
    return reinterpret_cast<unsigned long long *>(base) + parity * kFlagStride + slot;
}
static inline double *data_at(void *base, int parity, size_t half_bytes) {
    return reinterpret_cast<double *>(static_cast<char *>(base) + kFlagBytes + (size_t)parity * half_bytes);
}

static int launch_push(b200_ctx_t ctx, int64_t count, const double *src, const int *idx,
                       const PeerTargets &tgt, int64_t seg_stride, unsigned long long seq) {
    // indexed (halo) pushes: one value per thread; contiguous ones: 8 doubles per thread;
    // never more than 4 CTAs per SM -- the grid-stride loops cover the rest
    const int64_t per_cta = idx ? kThreads : (int64_t)kThreads * 8;
    const int64_t want = std::max<int64_t>(1, (count + per_cta - 1) / per_cta);
    const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)ctx->sm_count * 4);
    push_kernel<<<grid, kThreads, 0, ctx->stream>>>(count, src, idx, tgt, ctx->nranks, seg_stride,
                                                    ctx->push_ticket, seq);
    B200_CHECK_LAUNCH();
    ctx->launches++;
    return B200_OK;
}

// SQUARE operators: make every rank's boundary values of x visible in A->halo.
// One pack kernel + one in-place ncclAllGather (S doubles per rank) on the stream.
static int halo_exchange(b200_ctx_t ctx, b200_csr_t A, const double *x, CsrArgs &a) {
    a.xh = A->halo; a.nloc = (int)A->n_loc;
    if (A->S == 0) return B200_OK;
    ProfScope prof(ctx, B200_PROF_COMM, A->n_send, ctx->nranks, 0);
    if (ctx->p2p) {
        // push my boundary values straight into the halo buffers of the ranks that gather
        // them, release their flags, then wait for the ranks I gather from
        const int par = (int)(A->seq & 1);
        const unsigned long long seq = ++A->seq;
        PeerTargets tgt;
        WaitList w;
        bool any_wait = false;
        for (int q = 0; q < kMaxRanks; ++q) { tgt.data[q] = nullptr; tgt.flag[q] = nullptr; w.flag[q] = nullptr; }
        for (int q = 0; q < ctx->nranks; ++q) {
            if (q == ctx->rank) continue;
            if (A->needed_by[q]) {
                tgt.data[q] = data_at(A->pb_peer[q], par, A->pb_half) + (size_t)ctx->rank * A->S;
                tgt.flag[q] = flag_at(A->pb_peer[q], par, ctx->rank);
            }
            if (A->need_from[q]) { w.flag[q] = flag_at(A->pb_local, par, q); any_wait = true; }
        }
        int rc = launch_push(ctx, A->n_send, x, A->send_idx, tgt, 0, seq);
        if (rc) return rc;
        A->halo = data_at(A->pb_local, par, A->pb_half);   // what the kernel gathers from
        a.xh = A->halo;
        if (any_wait) {
            // no separate wait launch: the consumer kernel starts on its interior rows at
            // once and only blocks that gather remote columns poll the flags
            unsigned int mask = 0;
            for (int q = 0; q < ctx->nranks; ++q)
                if (w.flag[q]) mask |= 1u << q;
            a.blk_halo = A->blk_halo;
            a.wait_flags = flag_at(A->pb_local, par, 0);
            a.wait_mask = mask;
            a.wait_seq = seq;
        }
        return B200_OK;
    }
    double *mine = A->halo + (size_t)ctx->rank * A->S;
    if (A->n_send) {
        const unsigned grid = (unsigned)((A->n_send + kThreads - 1) / kThreads);
        halo_pack_kernel<<<grid, kThreads, 0, ctx->stream>>>(A->n_send, A->send_idx, x, mine);
        B200_CHECK_LAUNCH();
        ctx->launches++;
    }
    B200_NCCL(nccl().AllGather(mine, A->halo, (size_t)A->S, ncclDouble, comm_of(ctx), ctx->stream));
    return B200_OK;
}

// PROLONG operators: bring the coarse vector to every rank; returns the pointer to gather from.
static int coarse_to_all(b200_ctx_t ctx, b200_csr_t A, b200_vec_t xc, const double **px) {
    ProfScope prof(ctx, B200_PROF_COMM, A->gl_cols, ctx->nranks, 1);
    if (ctx->p2p) {
        const int par = (int)(A->seq & 1);
        const unsigned long long seq = ++A->seq;
        PeerTargets tgt;
        WaitList w;
        bool any_wait = false;
        for (int q = 0; q < kMaxRanks; ++q) { tgt.data[q] = nullptr; tgt.flag[q] = nullptr; w.flag[q] = nullptr; }
        if (A->coarse_dist) {
            B200_REQUIRE(xc->kind == B200_VK_DIST && (int64_t)xc->cap == A->coarse_B,
                         "prolongation: coarse vector is not partitioned like the operator");
            int rc = materialize(xc);
            if (rc) return rc;
            for (int q = 0; q < ctx->nranks; ++q) {
                if (A->needed_by[q]) {
                    tgt.data[q] = data_at(A->pb_peer[q], par, A->pb_half) + (size_t)ctx->rank * A->coarse_B;
                    tgt.flag[q] = flag_at(A->pb_peer[q], par, ctx->rank);
                }
                if (A->need_from[q]) { w.flag[q] = flag_at(A->pb_local, par, q); any_wait = true; }
            }
            rc = launch_push(ctx, A->coarse_B, xc->ptr, nullptr, tgt, 0, seq);
            if (rc) return rc;
            *px = data_at(A->pb_local, par, A->pb_half);
        } else if (ctx->rank == 0) {
            B200_REQUIRE(xc->kind == B200_VK_LOCAL, "prolongation: coarse vector must live on rank 0");
            int rc = materialize(xc);
            if (rc) return rc;
            bool any = false;
            for (int q = 1; q < ctx->nranks; ++q)
                if (A->needed_by[q]) {
                    tgt.data[q] = data_at(A->pb_peer[q], par, A->pb_half);
                    tgt.flag[q] = flag_at(A->pb_peer[q], par, 0);
                    any = true;
                }
            if (any) {
                rc = launch_push(ctx, A->gl_cols, xc->ptr, nullptr, tgt, 0, seq);
                if (rc) return rc;
            }
            *px = xc->ptr;
        } else {
            if (A->need_from[0]) { w.flag[0] = flag_at(A->pb_local, par, 0); any_wait = true; }
            *px = data_at(A->pb_local, par, A->pb_half);
        }
        if (any_wait) {
            wait_kernel<<<1, 32, 0, ctx->stream>>>(w, ctx->nranks, seq);
            B200_CHECK_LAUNCH();
            ctx->launches++;
        }
        return B200_OK;
    }
    if (A->coarse_dist) {
        B200_REQUIRE(xc->kind == B200_VK_DIST && (int64_t)xc->cap == A->coarse_B,
                     "prolongation: coarse vector is not partitioned like the operator");
        int rc = materialize(xc);
        if (rc) return rc;
        B200_NCCL(nccl().AllGather(xc->ptr, A->cbuf, (size_t)A->coarse_B, ncclDouble, comm_of(ctx), ctx->stream));
        *px = A->cbuf;
        return B200_OK;
    }
    if (ctx->rank == 0) {
        B200_REQUIRE(xc->kind == B200_VK_LOCAL, "prolongation: coarse vector must live on rank 0");
        int rc = materialize(xc);
        if (rc) return rc;
        B200_NCCL(nccl().Broadcast(xc->ptr, xc->ptr, (size_t)A->gl_cols, ncclDouble, 0, comm_of(ctx), ctx->stream));
        *px = xc->ptr;
    } else {
        B200_NCCL(nccl().Broadcast(A->cbuf, A->cbuf, (size_t)A->gl_cols, ncclDouble, 0, comm_of(ctx), ctx->stream));
        *px = A->cbuf;
    }
    return B200_OK;
}

// RESTRICT operators: combine the per-rank partial sums in A->cbuf into the coarse vector.
static int partials_to_coarse(b200_ctx_t ctx, b200_csr_t A, b200_vec_t yc) {
    ProfScope prof(ctx, B200_PROF_COMM, A->gl_rows, ctx->nranks, 2);
    if (ctx->p2p) {
        // every rank stores its partial sums for owner q directly into q's staging area;
        // the owner adds the staged partials in rank order (deterministic)
        const int par = (int)(A->seq & 1);
        const unsigned long long seq = ++A->seq;
        PeerTargets tgt;
        WaitList w;
        for (int q = 0; q < kMaxRanks; ++q) { tgt.data[q] = nullptr; tgt.flag[q] = nullptr; w.flag[q] = nullptr; }
        const int64_t seg = A->coarse_dist ? A->coarse_B : A->gl_rows;     // staged entries per source
        const int nowners = A->coarse_dist ? ctx->nranks : 1;
        bool any = false;
        for (int q = 0; q < nowners; ++q)
            if (A->needed_by[q]) {
                tgt.data[q] = data_at(A->pb_peer[q], par, A->pb_half_owner) + (size_t)ctx->rank * seg;
                tgt.flag[q] = flag_at(A->pb_peer[q], par, ctx->rank);
                any = true;
            }
        if (any) {
            int rc = launch_push(ctx, seg, A->cbuf, nullptr, tgt, A->coarse_dist ? seg : 0, seq);
            if (rc) return rc;
        }
        const bool owner = A->coarse_dist || ctx->rank == 0;
        if (owner) {
            if (A->coarse_dist)
                B200_REQUIRE(yc->kind == B200_VK_DIST && (int64_t)yc->cap == A->coarse_B,
                             "restriction: coarse vector is not partitioned like the operator");
            else
                B200_REQUIRE(yc->kind == B200_VK_LOCAL, "restriction: coarse vector must live on rank 0");
            for (int q = 0; q < ctx->nranks; ++q)
                if (A->need_from[q]) w.flag[q] = flag_at(A->pb_local, par, q);
            const int64_t count = (int64_t)yc->len;
            if (count) {
                const unsigned grid = (unsigned)((count + kThreads - 1) / kThreads);
                reduce_sum_kernel<<<grid, kThreads, 0, ctx->stream>>>(
                    count, data_at(A->pb_local, par, A->pb_half), seg, ctx->nranks, w, seq, wr(yc), nullptr);
                B200_CHECK_LAUNCH();
                ctx->launches++;
            }
            yc->zero_pending = false;
        }
        return B200_OK;
    }
    if (A->coarse_dist) {
        B200_REQUIRE(yc->kind == B200_VK_DIST && (int64_t)yc->cap == A->coarse_B,
                     "restriction: coarse vector is not partitioned like the operator");
        B200_NCCL(nccl().ReduceScatter(A->cbuf, wr(yc), (size_t)A->coarse_B, ncclDouble, ncclSum,
                                       comm_of(ctx), ctx->stream));
        return B200_OK;
    }
    double *dst = A->cbuf;
    if (ctx->rank == 0) {
        B200_REQUIRE(yc->kind == B200_VK_LOCAL, "restriction: coarse vector must live on rank 0");
        dst = wr(yc);
    }
    B200_NCCL(nccl().Reduce(A->cbuf, dst, (size_t)A->gl_rows, ncclDouble, ncclSum, 0, comm_of(ctx), ctx->stream));
    return B200_OK;
}

} // namespace b200

extern "C" int b200_csr_create_i64(b200_ctx_t ctx, int64_t nrows, int64_t ncols,
                                   const int64_t *ptr, const int64_t *col, const double *val,
                                   b200_csr_t *A) {
    return csr_create(ctx, nrows, ncols, ptr, col, val, A);
}

extern "C" int b200_csr_create_i32(b200_ctx_t ctx, int64_t nrows, int64_t ncols,
                                   const int32_t *ptr, const int32_t *col, const double *val,
                                   b200_csr_t *A) {
    return csr_create(ctx, nrows, ncols, ptr, col, val, A);
}

extern "C" int b200_csr_create_i64_f32(b200_ctx_t ctx, int64_t nrows, int64_t ncols,
                                       const int64_t *ptr, const int64_t *col, const float *val,
                                       b200_csr_t *A) {
    return csr_create_f32(ctx, nrows, ncols, ptr, col, val, A);
}

extern "C" int b200_csr_create_i32_f32(b200_ctx_t ctx, int64_t nrows, int64_t ncols,
                                       const int32_t *ptr, const int32_t *col, const float *val,
                                       b200_csr_t *A) {
    return csr_create_f32(ctx, nrows, ncols, ptr, col, val, A);
}

extern "C" int b200_csr_dtype(b200_csr_t A, int *dtype) {
    B200_REQUIRE(A && dtype, "null argument");
    *dtype = A->dtype;
    return B200_OK;
}

extern "C" int b200_plan_i64(int64_t nrows, const int64_t *ptr, int lanes, int nnz_cap,
                             int32_t *blk_out, int64_t blk_capacity, int64_t *nblocks,
                             int *lanes_out, int *rows_cap_out, int64_t *nlong_out) {
    B200_REQUIRE(nrows >= 0 && ptr != nullptr && nblocks != nullptr, "bad argument");
    B200_REQUIRE(nnz_cap >= 256 && nnz_cap <= kNnzCapMax && nnz_cap % 8 == 0, "bad nnz_cap");
    B200_REQUIRE(lanes == 0 || (lanes >= 1 && lanes <= 32 && !(lanes & (lanes - 1))), "bad lanes");
    RowBlockPlan plan;
    build_plan(nrows, ptr, lanes, nnz_cap, plan);
    *nblocks = (int64_t)plan.blk.size() - 1;
    if (lanes_out) *lanes_out = plan.lanes;
    if (rows_cap_out) *rows_cap_out = plan.rows_cap;
    if (nlong_out) *nlong_out = plan.nlong;
    if (blk_out) {
        if ((int64_t)plan.blk.size() > blk_capacity)
            return fail(B200_EINVAL, "plan output buffer too small");
        for (size_t i = 0; i < plan.blk.size(); ++i) {
            blk_out[2 * i] = plan.blk[i].x;
            blk_out[2 * i + 1] = plan.blk[i].y;
        }
    }
    return B200_OK;
}

extern "C" int b200_csr_destroy(b200_csr_t A) {
    if (!A) return B200_OK;
    NOT_RECORDING(A->ctx, "matrix destruction");
    if (A->in_graph) A->ctx->destroy_epoch++;
    GUARD(A->ctx);
    csr_free(A);
    return B200_OK;
}

extern "C" int b200_csr_rows(b200_csr_t A, size_t *n) {
    B200_REQUIRE(A && n, "null argument");
    *n = (size_t)A->gl_rows;
    return B200_OK;
}
extern "C" int b200_csr_cols(b200_csr_t A, size_t *n) {
    B200_REQUIRE(A && n, "null argument");
    *n = (size_t)A->gl_cols;
    return B200_OK;
}
extern "C" int b200_csr_nonzeros(b200_csr_t A, size_t *n) {
    B200_REQUIRE(A && n, "null argument");
    *n = (size_t)A->gl_nnz;
    return B200_OK;
}
extern "C" int b200_csr_bytes(b200_csr_t A, size_t *bytes) {
    B200_REQUIRE(A && bytes, "null argument");
    *bytes = A->bytes;
    return B200_OK;
}
extern "C" int b200_csr_plan(b200_csr_t A, int *lanes_per_row, int64_t *n_blocks,
                             int64_t *n_long_blocks) {
    B200_REQUIRE(A, "null argument");
    if (lanes_per_row) *lanes_per_row = A->lanes;
    if (n_blocks) *n_blocks = A->nblocks;
    if (n_long_blocks) *n_long_blocks = A->nlong;
    return B200_OK;
}

// ---------------------------------------------------------------------------
// primitives
// ---------------------------------------------------------------------------
namespace b200 {

template <class P>
static int spmv_local(b200_ctx_t ctx, double alpha, b200_csr_t A, b200_vec_t x, double beta,
                      b200_vec_t y) {
    CsrArgsT<P> a = base_args_t<P>(A);
    const double *px;
    int rc = rd(x, &px);
    if (rc) return rc;
    a.x = tp<typename P::TX>(px);
    a.alpha = alpha; a.beta = beta;
    if (beta == 0.0 || y->zero_pending) {
        a.y = tp<typename P::TY>(wr(y));
        return launch_csr<MODE_SPMV>(ctx, A, a);
    }
    a.y = tp<typename P::TY>(y->ptr);
    return launch_csr<MODE_SPMV_ACC>(ctx, A, a);
}

template <class P>
static int residual_local(b200_ctx_t ctx, b200_vec_t f, b200_csr_t A, b200_vec_t x, b200_vec_t r) {
    CsrArgsT<P> a = base_args_t<P>(A);
    const double *px, *pf;
    int rc = rd(x, &px);
    if (rc) return rc;
    rc = rd(f, &pf);
    if (rc) return rc;
    a.x = tp<typename P::TX>(px);
    a.f = tp<typename P::TF>(pf);
    a.y = tp<typename P::TY>((f == r) ? r->ptr : wr(r));
    return launch_csr<MODE_RESID>(ctx, A, a);
}

#define B200_BAD_MIX(what) fail(B200_EINVAL, what ": unsupported precision combination")

} // namespace b200

extern "C" int b200_spmv(b200_ctx_t ctx, double alpha, b200_csr_t A, b200_vec_t x, double beta,
                         b200_vec_t y) {
    CHECK_CTX(ctx);
    B200_REQUIRE(A && x && y, "null argument");
    touch(ctx, {x, y});
    B200_REQUIRE((int64_t)x->n == A->gl_cols, "spmv: x size != matrix columns");
    B200_REQUIRE((int64_t)y->n == A->gl_rows, "spmv: y size != matrix rows");
    B200_REQUIRE(x != y && (x->ptr != y->ptr || !x->ptr), "spmv: x and y must not alias");
    if (A->kind == B200_CK_GHOST) return B200_OK;          // operator lives on rank 0
    GUARD(ctx);
    if (A->dtype == B200_F32) {
        // FP32 operator (mixed-precision hierarchy): single GPU, persistent ring kernels
        if (all32({x, y})) return spmv_local<PrecFF>(ctx, alpha, A, x, beta, y);
        if (all64({x, y})) return spmv_local<PrecFD>(ctx, alpha, A, x, beta, y);
        if (x->dtype == B200_F32 && y->dtype == B200_F64)
            return spmv_local<PrecFFD>(ctx, alpha, A, x, beta, y);
        return B200_BAD_MIX("spmv");
    }
    if (!all64({x, y})) return B200_BAD_MIX("spmv");
    CsrArgs a = base_args(A);
    a.alpha = alpha; a.beta = beta;
    int rc;
    if (A->kind == B200_CK_PROLONG) {
        B200_REQUIRE(y->kind == B200_VK_DIST, "prolongation: y must be a partitioned vector");
        rc = coarse_to_all(ctx, A, x, &a.x);
        if (rc) return rc;
    } else if (A->kind == B200_CK_RESTRICT) {
        B200_REQUIRE(beta == 0.0, "restriction on a distributed context needs beta == 0");
        B200_REQUIRE(x->kind == B200_VK_DIST, "restriction: x must be a partitioned vector");
        rc = rd(x, &a.x);
        if (rc) return rc;
        a.y = A->cbuf;
        rc = launch_csr<MODE_SPMV>(ctx, A, a);
        if (rc) return rc;
        return partials_to_coarse(ctx, A, y);
    } else {
        rc = rd(x, &a.x);
        if (rc) return rc;
        if (A->kind == B200_CK_SQUARE) {
            B200_REQUIRE(x->kind == B200_VK_DIST && y->kind == B200_VK_DIST,
                         "spmv: vectors must be partitioned like the operator");
            rc = halo_exchange(ctx, A, a.x, a);
            if (rc) return rc;
        }
    }
    if (beta == 0.0 || y->zero_pending) {
        a.y = wr(y);
        return launch_csr<MODE_SPMV>(ctx, A, a);
    }
    a.y = y->ptr;
    return launch_csr<MODE_SPMV_ACC>(ctx, A, a);
}

extern "C" int b200_residual(b200_ctx_t ctx, b200_vec_t f, b200_csr_t A, b200_vec_t x,
                             b200_vec_t r) {
    CHECK_CTX(ctx);
    B200_REQUIRE(f && A && x && r, "null argument");
    touch(ctx, {f, x, r});
    B200_REQUIRE((int64_t)x->n == A->gl_cols, "residual: x size != matrix columns");
    B200_REQUIRE((int64_t)f->n == A->gl_rows && (int64_t)r->n == A->gl_rows,
                 "residual: rhs/r size != matrix rows");
    B200_REQUIRE(x != r && (x->ptr != r->ptr || !x->ptr), "residual: x and r must not alias");
    if (A->kind == B200_CK_GHOST) return B200_OK;
    B200_REQUIRE(A->kind == B200_CK_LOCAL || A->kind == B200_CK_SQUARE,
                 "residual: operator must be square");
    GUARD(ctx);
    if (A->dtype == B200_F32) {
        if (all32({f, x, r})) return residual_local<PrecFF>(ctx, f, A, x, r);
        if (all64({f, x, r})) return residual_local<PrecFD>(ctx, f, A, x, r);
        if (all64({f, x}) && r->dtype == B200_F32) return residual_local<PrecFDF>(ctx, f, A, x, r);
        return B200_BAD_MIX("residual");
    }
    if (!all64({f, x, r})) return B200_BAD_MIX("residual");
    CsrArgs a = base_args(A);
    int rc = rd(x, &a.x);
    if (rc) return rc;
    rc = rd(f, &a.f);
    if (rc) return rc;
    if (A->kind == B200_CK_SQUARE) {
        B200_REQUIRE(x->kind == B200_VK_DIST && f->kind == B200_VK_DIST && r->kind == B200_VK_DIST,
                     "residual: vectors must be partitioned like the operator");
        rc = halo_exchange(ctx, A, a.x, a);
        if (rc) return rc;
    }
    a.y = (f == r) ? r->ptr : wr(r);   // r == f is fine: each row reads f[r] before writing
    return launch_csr<MODE_RESID>(ctx, A, a);
}

extern "C" int b200_clear(b200_ctx_t ctx, b200_vec_t x) {
    CHECK_CTX(ctx);
    B200_REQUIRE(x, "null argument");
    touch(ctx, {x});
    if (x->kind == B200_VK_GHOST) return B200_OK;
    if (ctx->opt_zero_shortcut) {
        x->zero_pending = true;
        return B200_OK;
    }
    GUARD(ctx);
    x->zero_pending = true;
    return materialize(x);
}

extern "C" int b200_copy(b200_ctx_t ctx, b200_vec_t x, b200_vec_t y) {
    CHECK_CTX(ctx);
    B200_REQUIRE(x && y, "null argument");
    touch(ctx, {x, y});
    B200_REQUIRE(same_layout(x, y), "copy: size mismatch");
    if (x->kind == B200_VK_GHOST || x == y || x->ptr == y->ptr) return B200_OK;
    if (x->zero_pending) {
        y->zero_pending = true;
        return B200_OK;
    }
    GUARD(ctx);
    if (all64({x, y}))
        return launch_ew<CopyF<double>, false, false, double>(ctx, x->len, CopyF<double>(), x->ptr, nullptr, nullptr, wr(y));
    if (all32({x, y}))
        return launch_ew<CopyF<float>, false, false, float>(ctx, x->len, CopyF<float>(), tp<float>(x->ptr), nullptr, nullptr, tp<float>(wr(y)));
    if (x->dtype == B200_F64)      // precision-changing copies
        return launch_ew_mixed<CopyF<float>, false, false>(ctx, x->len, CopyF<float>(), x->ptr, (const float *)nullptr, (const float *)nullptr, tp<float>(wr(y)));
    return launch_ew_mixed<CopyF<double>, false, false>(ctx, x->len, CopyF<double>(), tp<float>(x->ptr), (const double *)nullptr, (const double *)nullptr, wr(y));
}

namespace b200 {
template <class T>
static void launch_dot_kernel(b200_ctx_t ctx, b200_vec_t x, b200_vec_t y, double *result_dev) {
    const bool vec_ok = aligned16(x->ptr) && aligned16(y->ptr);
    const int grid = std::min(grid_for(ctx, x->len, 32 / (int)sizeof(T) * 2), kDotMaxBlocks);
    ProfScope prof(ctx, B200_PROF_DOT, (int64_t)x->len, 1, 0);
    const cudaError_t rc = launch_pdl(ctx, dot_kernel<T>, dim3(grid), dim3(kThreads), 0, x->len,
                                      (const T *)tp<T>(x->ptr), (const T *)tp<T>(y->ptr), ctx->dot_partial,
                                      ctx->dot_ticket, result_dev, vec_ok);
    if (rc != cudaSuccess) cuda_fail(rc, "dot_kernel launch", __FILE__, __LINE__);
}
} // namespace b200

extern "C" int b200_dot(b200_ctx_t ctx, b200_vec_t x, b200_vec_t y, double *result) {
    CHECK_CTX(ctx);
    B200_REQUIRE(x && y && result, "null argument");
    NOT_RECORDING(ctx, "dot (host-synchronous)");
    B200_REQUIRE(same_layout(x, y), "dot: size mismatch");
    if (x->dtype != y->dtype) return B200_BAD_MIX("dot");
    GUARD(ctx);
    const bool dist = x->kind == B200_VK_DIST;
    const bool trivial = x->len == 0 || x->zero_pending || y->zero_pending || x->kind == B200_VK_GHOST;
    if (!dist) {
        if (trivial) {
            B200_CUDA(cudaStreamSynchronize(ctx->stream));
            *result = 0.0;
            return B200_OK;
        }
        if (x->dtype == B200_F64) launch_dot_kernel<double>(ctx, x, y, ctx->dot_result_d);
        else launch_dot_kernel<float>(ctx, x, y, ctx->dot_result_d);
        B200_CHECK_LAUNCH();
        ctx->launches++;
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        *result = *reinterpret_cast<volatile double *>(ctx->dot_result_h);
        return B200_OK;
    }
    // partitioned vectors: local partial -> device scalar -> all-reduce -> host
    // (mpi/inner_product.hpp:53-62 does the same with MPI_Allreduce on the host)
    if (trivial) {
        B200_CUDA(cudaMemsetAsync(ctx->dot_dev, 0, sizeof(double), ctx->stream));
    } else {
        launch_dot_kernel<double>(ctx, x, y, ctx->dot_dev);
        B200_CHECK_LAUNCH();
        ctx->launches++;
    }
    if (ctx->p2p) {
        // every rank stores its partial into slot `rank` of every peer; each rank then adds
        // the P partials in rank order (bitwise identical on all ranks) straight into
        // mapped host memory
        const int par = (int)(ctx->dot_seq & 1);
        const unsigned long long seq = ++ctx->dot_seq;
        PeerTargets tgt;
        WaitList w;
        for (int q = 0; q < kMaxRanks; ++q) { tgt.data[q] = nullptr; tgt.flag[q] = nullptr; w.flag[q] = nullptr; }
        for (int q = 0; q < ctx->nranks; ++q) {
            tgt.data[q] = data_at(ctx->dot_pb_peer[q], par, 256) + ctx->rank;
            tgt.flag[q] = flag_at(ctx->dot_pb_peer[q], par, ctx->rank);
            w.flag[q] = flag_at(ctx->dot_pb_local, par, q);
        }
        int rc = launch_push(ctx, 1, ctx->dot_dev, nullptr, tgt, 0, seq);
        if (rc) return rc;
        reduce_sum_kernel<<<1, 32, 0, ctx->stream>>>(1, data_at(ctx->dot_pb_local, par, 256), 1, ctx->nranks,
                                                      w, seq, ctx->dot_dev + 1, ctx->dot_result_d);
        B200_CHECK_LAUNCH();
        ctx->launches++;
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        *result = *reinterpret_cast<volatile double *>(ctx->dot_result_h);
        return B200_OK;
    }
    B200_NCCL(nccl().AllReduce(ctx->dot_dev, ctx->dot_dev, 1, ncclDouble, ncclSum, comm_of(ctx), ctx->stream));
    B200_CUDA(cudaMemcpyAsync(ctx->dot_result_h, ctx->dot_dev, sizeof(double), cudaMemcpyDeviceToHost,
                              ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    *result = *reinterpret_cast<volatile double *>(ctx->dot_result_h);
    return B200_OK;
}

namespace b200 {
template <class T>
static int axpby_t(b200_ctx_t ctx, double a, b200_vec_t x, double b, b200_vec_t y) {
    const double *px;
    int rc = rd(x, &px);
    if (rc) return rc;
    if (b == 0.0 || y->zero_pending) {
        AxF<T> f{(T)a};
        return launch_ew<AxF<T>, false, false, T>(ctx, x->len, f, tp<T>(px), nullptr, nullptr, tp<T>(wr(y)));
    }
    AxpbyF<T> f{(T)a, (T)b};
    return launch_ew<AxpbyF<T>, true, false, T>(ctx, x->len, f, tp<T>(px), tp<T>(y->ptr), nullptr, tp<T>(y->ptr));
}
template <class T>
static int axpbypcz_t(b200_ctx_t ctx, double a, b200_vec_t x, double b, b200_vec_t y, double c, b200_vec_t z) {
    const double *px, *py;
    int rc = rd(x, &px);
    if (rc) return rc;
    rc = rd(y, &py);
    if (rc) return rc;
    if (c == 0.0 || z->zero_pending) {
        AxpbyF<T> f{(T)a, (T)b};
        return launch_ew<AxpbyF<T>, true, false, T>(ctx, x->len, f, tp<T>(px), tp<T>(py), nullptr, tp<T>(wr(z)));
    }
    AxpbypczF<T> f{(T)a, (T)b, (T)c};
    return launch_ew<AxpbypczF<T>, true, true, T>(ctx, x->len, f, tp<T>(px), tp<T>(py), tp<T>(z->ptr), tp<T>(z->ptr));
}
template <class T>
static int vmul_t(b200_ctx_t ctx, double alpha, b200_vec_t x, b200_vec_t y, double beta, b200_vec_t z) {
    const double *px, *py;
    int rc = rd(x, &px);
    if (rc) return rc;
    rc = rd(y, &py);
    if (rc) return rc;
    if (beta == 0.0 || z->zero_pending) {
        VmulF<T> f{(T)alpha};
        return launch_ew<VmulF<T>, true, false, T>(ctx, x->len, f, tp<T>(px), tp<T>(py), nullptr, tp<T>(wr(z)));
    }
    VmulAccF<T> f{(T)alpha, (T)beta};
    return launch_ew<VmulAccF<T>, true, true, T>(ctx, x->len, f, tp<T>(px), tp<T>(py), tp<T>(z->ptr), tp<T>(z->ptr));
}
} // namespace b200

extern "C" int b200_axpby(b200_ctx_t ctx, double a, b200_vec_t x, double b, b200_vec_t y) {
    CHECK_CTX(ctx);
    B200_REQUIRE(x && y, "null argument");
    touch(ctx, {x, y});
    B200_REQUIRE(same_layout(x, y), "axpby: size mismatch");
    if (x->kind == B200_VK_GHOST) return B200_OK;
    GUARD(ctx);
    if (all64({x, y})) return axpby_t<double>(ctx, a, x, b, y);
    if (all32({x, y})) return axpby_t<float>(ctx, a, x, b, y);
    return B200_BAD_MIX("axpby");
}

extern "C" int b200_axpbypcz(b200_ctx_t ctx, double a, b200_vec_t x, double b, b200_vec_t y,
                             double c, b200_vec_t z) {
    CHECK_CTX(ctx);
    B200_REQUIRE(x && y && z, "null argument");
    touch(ctx, {x, y, z});
    B200_REQUIRE(same_layout(x, y) && same_layout(x, z), "axpbypcz: size mismatch");
    if (x->kind == B200_VK_GHOST) return B200_OK;
    GUARD(ctx);
    if (all64({x, y, z})) return axpbypcz_t<double>(ctx, a, x, b, y, c, z);
    if (all32({x, y, z})) return axpbypcz_t<float>(ctx, a, x, b, y, c, z);
    return B200_BAD_MIX("axpbypcz");
}

extern "C" int b200_vmul(b200_ctx_t ctx, double alpha, b200_vec_t x, b200_vec_t y, double beta,
                         b200_vec_t z) {
    CHECK_CTX(ctx);
    B200_REQUIRE(x && y && z, "null argument");
    touch(ctx, {x, y, z});
    B200_REQUIRE(same_layout(x, y) && same_layout(x, z), "vmul: size mismatch");
    if (x->kind == B200_VK_GHOST) return B200_OK;
    GUARD(ctx);
    if (all64({x, y, z})) return vmul_t<double>(ctx, alpha, x, y, beta, z);
    if (all32({x, y, z})) return vmul_t<float>(ctx, alpha, x, y, beta, z);
    if (all32({x, y}) && z->dtype == B200_F64) {
        // FP32 smoother diagonal and residual accumulated into an FP64 iterate
        const double *px, *py;
        int rc = rd(x, &px);
        if (rc) return rc;
        rc = rd(y, &py);
        if (rc) return rc;
        if (beta == 0.0 || z->zero_pending) {
            VmulF<double> f{alpha};
            return launch_ew_mixed<VmulF<double>, true, false>(ctx, x->len, f, tp<float>(px), tp<float>(py),
                                                               (const double *)nullptr, wr(z));
        }
        VmulAccF<double> f{alpha, beta};
        return launch_ew_mixed<VmulAccF<double>, true, true>(ctx, x->len, f, tp<float>(px), tp<float>(py),
                                                            (const double *)z->ptr, z->ptr);
    }
    return B200_BAD_MIX("vmul");
}

// ---------------------------------------------------------------------------
// smoother sweep
// ---------------------------------------------------------------------------
namespace b200 {

template <class TD, class TF, class TX>
static int relax_zero_t(b200_ctx_t ctx, double omega, const double *pd, const double *pf, b200_vec_t x) {
    if (x->len) {
        const int grid = grid_for(ctx, x->len, 2);
        ProfScope prof(ctx, B200_PROF_RELAX_ZERO, (int64_t)x->len, 1, 0);
        B200_CUDA(launch_pdl(ctx, relax_zero_kernel<TD, TF, TX>, dim3(grid), dim3(kThreads), 0, x->len, omega,
                             tp<TD>(pd), tp<TF>(pf), tp<TX>(wr(x))));
        B200_CHECK_LAUNCH();
        ctx->launches++;
    }
    x->zero_pending = false;
    return B200_OK;
}

} // namespace b200

extern "C" int b200_relax(b200_ctx_t ctx, b200_csr_t A, b200_vec_t rhs, b200_vec_t x,
                          b200_vec_t tmp, b200_vec_t diag, double omega) {
    CHECK_CTX(ctx);
    B200_REQUIRE(A && rhs && x && tmp && diag, "null argument");
    touch(ctx, {rhs, x, tmp, diag});
    B200_REQUIRE(A->gl_rows == A->gl_cols, "relax: matrix must be square");
    B200_REQUIRE((int64_t)x->n == A->gl_rows && same_layout(x, rhs) && same_layout(x, diag) &&
                     same_layout(x, tmp),
                 "relax: vector size != matrix rows");
    B200_REQUIRE(x != tmp && x != rhs && tmp != rhs, "relax: x, tmp and rhs must be distinct vectors");
    if (A->kind == B200_CK_GHOST) return B200_OK;
    B200_REQUIRE(x->ptr != tmp->ptr, "relax: x and tmp must not alias");
    GUARD(ctx);

    // precision combination: 0 = FP64 throughout, 1 = FP32 throughout,
    // 2 = FP32 operator + diagonal sweeping an FP64 iterate (finest level of a mixed hierarchy;
    //     tmp is that level's FP32 scratch)
    int mix = -1;
    if (A->dtype == B200_F64 && all64({rhs, x, tmp, diag})) mix = 0;
    else if (A->dtype == B200_F32 && all32({rhs, x, tmp, diag})) mix = 1;
    else if (A->dtype == B200_F32 && all64({rhs, x}) && all32({tmp, diag})) mix = 2;
    if (mix < 0) return B200_BAD_MIX("relax");

    const double *pf, *pd;
    int rc = rd(rhs, &pf);
    if (rc) return rc;
    rc = rd(diag, &pd);
    if (rc) return rc;

    if (x->zero_pending && ctx->opt_zero_shortcut) {
        // residual(rhs, A, 0) == rhs exactly, so the sweep reduces to a scaling
        if (mix == 0) return relax_zero_t<double, double, double>(ctx, omega, pd, pf, x);
        if (mix == 1) return relax_zero_t<float, float, float>(ctx, omega, pd, pf, x);
        return relax_zero_t<float, double, double>(ctx, omega, pd, pf, x);
    }

    if (!ctx->opt_fuse_relax) {
        // the literal reference sequence: tmp = rhs - A x ; x = omega*diag.*tmp + x
        rc = b200_residual(ctx, rhs, A, x, tmp);
        if (rc) return rc;
        return b200_vmul(ctx, omega, diag, tmp, 1.0, x);
    }

    if (mix == 1) {
        CsrArgsT<PrecFF> a = base_args_t<PrecFF>(A);
        const double *px;
        rc = rd(x, &px);
        if (rc) return rc;
        a.x = tp<float>(px); a.f = tp<float>(pf); a.d = tp<float>(pd); a.alpha = omega;
        a.y = tp<float>(wr(tmp));
        rc = launch_csr<MODE_RELAX>(ctx, A, a);
        if (rc) return rc;
        if (x->owned && tmp->owned && x->cap == tmp->cap) std::swap(x->ptr, tmp->ptr);
        else B200_CUDA(cudaMemcpyAsync(x->ptr, tmp->ptr, x->len * x->esz, cudaMemcpyDeviceToDevice, ctx->stream));
        return B200_OK;
    }
    if (mix == 2) {
        // the new FP64 iterate cannot live in the level's FP32 scratch: the operator owns an
        // FP64 buffer that trades places with x exactly like tmp does in the uniform case
        if (!A->scratch64)
            B200_CUDA(cudaMalloc(&A->scratch64, ((size_t)A->nrows + 4) * sizeof(double)));
        if (ctx->recording) touch_slot(ctx, &A->scratch64, nullptr);
        CsrArgsT<PrecFD> a = base_args_t<PrecFD>(A);
        const double *px;
        rc = rd(x, &px);
        if (rc) return rc;
        a.x = px; a.f = pf; a.d = tp<float>(pd); a.alpha = omega;
        a.y = A->scratch64;
        rc = launch_csr<MODE_RELAX>(ctx, A, a);
        if (rc) return rc;
        if (x->owned && x->cap == (size_t)A->nrows) std::swap(x->ptr, A->scratch64);
        else B200_CUDA(cudaMemcpyAsync(x->ptr, A->scratch64, x->len * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        return B200_OK;
    }

    CsrArgs a = base_args(A);
    rc = rd(x, &a.x);
    if (rc) return rc;
    if (A->kind == B200_CK_SQUARE) {
        B200_REQUIRE(x->kind == B200_VK_DIST, "relax: vectors must be partitioned like the operator");
        rc = halo_exchange(ctx, A, a.x, a);
        if (rc) return rc;
    }
    a.f = pf; a.d = pd; a.alpha = omega;
    a.y = wr(tmp);
    rc = launch_csr<MODE_RELAX>(ctx, A, a);
    if (rc) return rc;
    if (x->owned && tmp->owned && x->cap == tmp->cap) {
        std::swap(x->ptr, tmp->ptr);          // x now holds the new iterate
    } else {
        B200_CUDA(cudaMemcpyAsync(x->ptr, tmp->ptr, x->len * sizeof(double),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return B200_OK;
}

// ---------------------------------------------------------------------------
// coarse solve
// ---------------------------------------------------------------------------
namespace b200 {

template <class Ptr, class Col, class Val>
static int coarse_create(b200_ctx_t ctx, int64_t n, const Ptr *ptr, const Col *col,
                         const Val *val, b200_coarse_t *out) {
    CHECK_CTX(ctx);
    NOT_RECORDING(ctx, "coarse solver creation");
    B200_REQUIRE(out != nullptr, "null output pointer");
    *out = nullptr;
    B200_REQUIRE(n > 0 && n <= 16384, "coarse solver: n must be in [1, 16384]");
    B200_REQUIRE(ptr && ptr[0] == 0, "bad row pointer array");
    const int64_t nnz = (int64_t)ptr[n];
    B200_REQUIRE(nnz >= 0 && (nnz == 0 || (col && val)), "bad col/val array");
    GUARD(ctx);
    const bool replicated = ctx->dist && n >= ctx->dist_min_rows;
    if (ctx->dist && ctx->rank != 0 && !replicated) {       // the coarsest level lives on rank 0
        b200_coarse_s *G = new (std::nothrow) b200_coarse_s();
        if (!G) return fail(B200_ENOMEM, "out of host memory");
        G->ctx = ctx; G->n = n; G->ghost = true;
        *out = G;
        return B200_OK;
    }

    std::vector<int32_t> hptr((size_t)n + 1), hcol((size_t)nnz);
    for (int64_t i = 0; i <= n; ++i) hptr[(size_t)i] = (int32_t)ptr[i];
    for (int64_t e = 0; e < nnz; ++e) {
        const int64_t c = (int64_t)col[e];
        if (c < 0 || c >= n) return fail(B200_EINVAL, "coarse solver: column index out of range");
        hcol[(size_t)e] = (int32_t)c;
    }

    // the inverse is always formed and kept in FP64, whatever the hierarchy's precision
    std::vector<double> hval((size_t)nnz);
    for (int64_t e = 0; e < nnz; ++e) hval[(size_t)e] = (double)val[e];
    const int N = (int)n;
    int *dptr = nullptr, *dcol = nullptr, *dpiv = nullptr;
    double *dval = nullptr, *M = nullptr, *colk = nullptr, *pivval = nullptr, *Ainv = nullptr;
    auto cleanup = [&]() {
        cudaFree(dptr); cudaFree(dcol); cudaFree(dval); cudaFree(M);
        cudaFree(colk); cudaFree(dpiv); cudaFree(pivval);
    };
#define CO_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t rc__ = (call);                                             \
        if (rc__ != cudaSuccess) {                                             \
            cleanup();                                                         \
            cudaFree(Ainv);                                                    \
            return cuda_fail(rc__, #call, __FILE__, __LINE__);                 \
        }                                                                      \
    } while (0)
    const size_t Mbytes = (size_t)N * 2 * N * sizeof(double);
    CO_CUDA(cudaMalloc(&dptr, ((size_t)N + 1) * sizeof(int)));
    CO_CUDA(cudaMalloc(&dcol, std::max<size_t>(1, (size_t)nnz) * sizeof(int)));
    CO_CUDA(cudaMalloc(&dval, std::max<size_t>(1, (size_t)nnz) * sizeof(double)));
    CO_CUDA(cudaMalloc(&M, Mbytes));
    CO_CUDA(cudaMalloc(&colk, (size_t)N * sizeof(double)));
    CO_CUDA(cudaMalloc(&dpiv, sizeof(int)));
    CO_CUDA(cudaMalloc(&pivval, ((size_t)N + 1) * sizeof(double)));
    CO_CUDA(cudaMalloc(&Ainv, (size_t)N * N * sizeof(double)));
    cudaStream_t st = ctx->stream;
    CO_CUDA(cudaMemcpyAsync(dptr, hptr.data(), ((size_t)N + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    if (nnz) {
        CO_CUDA(cudaMemcpyAsync(dcol, hcol.data(), (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, st));
        CO_CUDA(cudaMemcpyAsync(dval, hval.data(), (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    CO_CUDA(cudaMemsetAsync(M, 0, Mbytes, st));
    coarse_scatter_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, dptr, dcol, dval, M);
    CO_CUDA(cudaGetLastError());
    ctx->launches++;

    const int gcol2 = (2 * N + kThreads - 1) / kThreads;
    const int gcol1 = (N + kThreads - 1) / kThreads;
    const int ysplit = std::max(1, std::min(N, (ctx->sm_count * 4) / std::max(1, gcol2)));
    for (int k = 0; k < N; ++k) {
        coarse_pivot_kernel<<<1, kThreads, 0, st>>>(N, k, M, dpiv, pivval + k);
        coarse_colk_kernel<<<gcol1, kThreads, 0, st>>>(N, k, M, dpiv, colk);
        coarse_swap_scale_kernel<<<gcol2, kThreads, 0, st>>>(N, k, M, dpiv, pivval + k);
        const int gx = (2 * N - k + kThreads - 1) / kThreads;
        coarse_eliminate_kernel<<<dim3(gx, ysplit), kThreads, 0, st>>>(N, k, M, colk);
        ctx->launches += 4;
    }
    CO_CUDA(cudaGetLastError());
    {
        const size_t tot = (size_t)N * N;
        coarse_extract_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(N, M, Ainv);
        CO_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    std::vector<double> hpiv((size_t)N);
    CO_CUDA(cudaMemcpyAsync(hpiv.data(), pivval, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CO_CUDA(cudaStreamSynchronize(st));
#undef CO_CUDA
    cleanup();
    double pmax = 0.0, pmin = std::numeric_limits<double>::infinity();
    for (double p : hpiv) {
        const double a = std::fabs(p);
        if (!(a == a)) { pmin = 0.0; break; }   // NaN
        pmax = std::max(pmax, a);
        pmin = std::min(pmin, a);
    }
    if (!(pmin > 0.0) || pmin < pmax * 1e-14 || !std::isfinite(pmax)) {
        cudaFree(Ainv);
        return fail(B200_ESINGULAR, "coarse matrix is numerically singular");
    }

    b200_coarse_s *S = new (std::nothrow) b200_coarse_s();
    if (!S) {
        cudaFree(Ainv);
        return fail(B200_ENOMEM, "out of host memory");
    }
    S->ctx = ctx; S->n = n; S->Ainv = Ainv; S->bytes = (size_t)N * N * sizeof(double);
    S->dtype = std::is_same<Val, float>::value ? B200_F32 : B200_F64;
    if (replicated) {
        // the coarsest level is itself partitioned: every rank keeps the inverse and
        // applies its own rows to the all-gathered right-hand side
        S->replicated = true;
        S->block = Partition(n, ctx->nranks).B;
        cudaError_t rc = cudaMalloc(&S->gbuf, ((size_t)S->block * ctx->nranks + 2) * sizeof(double));
        if (rc != cudaSuccess) {
            cudaFree(Ainv);
            delete S;
            return cuda_fail(rc, "cudaMalloc(coarse gather buffer)", __FILE__, __LINE__);
        }
        S->bytes += (size_t)S->block * ctx->nranks * sizeof(double);
    }
    *out = S;
    return B200_OK;
}

} // namespace b200

extern "C" int b200_coarse_create_i64(b200_ctx_t ctx, int64_t n, const int64_t *ptr,
                                      const int64_t *col, const double *val, b200_coarse_t *S) {
    return coarse_create(ctx, n, ptr, col, val, S);
}
extern "C" int b200_coarse_create_i32(b200_ctx_t ctx, int64_t n, const int32_t *ptr,
                                      const int32_t *col, const double *val, b200_coarse_t *S) {
    return coarse_create(ctx, n, ptr, col, val, S);
}

extern "C" int b200_coarse_create_i64_f32(b200_ctx_t ctx, int64_t n, const int64_t *ptr,
                                          const int64_t *col, const float *val, b200_coarse_t *S) {
    CHECK_CTX(ctx);
    B200_REQUIRE_F64_DIST(ctx, "b200_coarse_create_*_f32");
    return coarse_create(ctx, n, ptr, col, val, S);
}
extern "C" int b200_coarse_create_i32_f32(b200_ctx_t ctx, int64_t n, const int32_t *ptr,
                                          const int32_t *col, const float *val, b200_coarse_t *S) {
    CHECK_CTX(ctx);
    B200_REQUIRE_F64_DIST(ctx, "b200_coarse_create_*_f32");
    return coarse_create(ctx, n, ptr, col, val, S);
}

extern "C" int b200_coarse_destroy(b200_coarse_t S) {
    if (!S) return B200_OK;
    NOT_RECORDING(S->ctx, "coarse solver destruction");
    if (S->in_graph) S->ctx->destroy_epoch++;
    GUARD(S->ctx);
    if (S->Ainv) cudaFree(S->Ainv);
    if (S->gbuf) cudaFree(S->gbuf);
    delete S;
    return B200_OK;
}

extern "C" int b200_coarse_bytes(b200_coarse_t S, size_t *bytes) {
    B200_REQUIRE(S && bytes, "null argument");
    *bytes = S->bytes;
    return B200_OK;
}

extern "C" int b200_coarse_solve(b200_ctx_t ctx, b200_coarse_t S, b200_vec_t rhs, b200_vec_t x) {
    CHECK_CTX(ctx);
    B200_REQUIRE(S && rhs && x, "null argument");
    touch(ctx, {rhs, x});
    if (ctx->recording) S->in_graph = true;
    B200_REQUIRE((int64_t)rhs->n == S->n && (int64_t)x->n == S->n, "coarse solve: size mismatch");
    if (S->ghost) return B200_OK;
    GUARD(ctx);
    const int N = (int)S->n;
    const int warps_per_cta = kThreads / 32;
    if (S->replicated) {
        B200_REQUIRE(rhs->kind == B200_VK_DIST && x->kind == B200_VK_DIST &&
                         (int64_t)rhs->cap == S->block && rhs != x,
                     "coarse solve: vectors must be partitioned like the coarsest level");
        int rc = materialize(rhs);
        if (rc) return rc;
        B200_NCCL(nccl().AllGather(rhs->ptr, S->gbuf, (size_t)S->block, ncclDouble, comm_of(ctx), ctx->stream));
        const int nloc = (int)x->len;
        if (nloc) {
            ProfScope prof(ctx, B200_PROF_COARSE, nloc, S->n, (int64_t)nloc * S->n);
            coarse_gemv_kernel<double><<<(nloc + warps_per_cta - 1) / warps_per_cta, kThreads, 0, ctx->stream>>>(
                N, (int)x->off, nloc, S->Ainv, S->gbuf, wr(x));
            B200_CHECK_LAUNCH();
            ctx->launches++;
        }
        x->zero_pending = false;
        return B200_OK;
    }
    B200_REQUIRE(rhs->kind == B200_VK_LOCAL && x->kind == B200_VK_LOCAL,
                 "coarse solve: vectors must live on this rank");
    B200_REQUIRE(rhs != x && rhs->ptr != x->ptr, "coarse solve: rhs and x must not alias");
    if (rhs->dtype != x->dtype) return B200_BAD_MIX("coarse solve");
    const double *pr;
    int rc = rd(rhs, &pr);
    if (rc) return rc;
    ProfScope prof(ctx, B200_PROF_COARSE, S->n, S->n, S->n * S->n);
    if (rhs->dtype == B200_F32)
        B200_CUDA(launch_pdl(ctx, coarse_gemv_kernel<float>, dim3((N + warps_per_cta - 1) / warps_per_cta),
                             dim3(kThreads), 0, N, 0, N, (const double *)S->Ainv, tp<float>(pr),
                             tp<float>(wr(x))));
    else
        B200_CUDA(launch_pdl(ctx, coarse_gemv_kernel<double>, dim3((N + warps_per_cta - 1) / warps_per_cta),
                             dim3(kThreads), 0, N, 0, N, (const double *)S->Ainv, pr, wr(x)));
    B200_CHECK_LAUNCH();
    ctx->launches++;
    return B200_OK;
}

// ---------------------------------------------------------------------------
// CUDA-graph recording of a call sequence (the V-cycle)
// ---------------------------------------------------------------------------
namespace b200 {
static void graph_free(b200_graph_s *g) {
    if (!g) return;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
}
// put every touched object back into the state it had when recording started (nothing that
// was recorded has run)
static void graph_release_deferred(b200_ctx_t ctx) {
    for (void *p : ctx->graph_deferred) cudaFree(p);     // cudaFree waits for the device
    ctx->graph_deferred.clear();
}
static void graph_rollback(b200_graph_s *g) {
    for (const GraphSlot &s : g->slots) {
        *s.slot = s.p0;
        if (s.zp) *s.zp = s.z0;
    }
}
} // namespace b200

extern "C" int b200_graph_begin(b200_ctx_t ctx, int *recording) {
    CHECK_CTX(ctx);
    B200_REQUIRE(recording != nullptr, "null output pointer");
    B200_REQUIRE(!ctx->recording, "graph_begin: already recording");
    *recording = 0;
    // not recordable: per-launch event timing, multi-GPU exchanges (host-side sequence
    // numbers and NCCL calls), the legacy default stream
    if (ctx->profiling || ctx->dist || !ctx->opt_cycle_graph) return B200_OK;
    if (ctx->stream == nullptr || ctx->stream == cudaStreamLegacy) return B200_OK;
    GUARD(ctx);
    b200_graph_s *g = new (std::nothrow) b200_graph_s();
    if (!g) return fail(B200_ENOMEM, "out of host memory");
    g->ctx = ctx;
    g->destroy_epoch = ctx->destroy_epoch;
    g->option_epoch = ctx->option_epoch;
    g->launches0 = ctx->launches;
    const cudaError_t rc = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed);
    if (rc != cudaSuccess) {
        delete g;
        return cuda_fail(rc, "cudaStreamBeginCapture", __FILE__, __LINE__);
    }
    ctx->recording = g;
    *recording = 1;
    return B200_OK;
}

extern "C" int b200_graph_abort(b200_ctx_t ctx) {
    CHECK_CTX(ctx);
    b200_graph_s *g = ctx->recording;
    if (!g) return B200_OK;
    GUARD(ctx);
    cudaGraph_t junk = nullptr;
    cudaStreamEndCapture(ctx->stream, &junk);      // may itself report the capture as invalidated
    if (junk) cudaGraphDestroy(junk);
    cudaGetLastError();
    graph_rollback(g);
    ctx->launches = g->launches0;
    ctx->recording = nullptr;
    graph_free(g);
    graph_release_deferred(ctx);
    return B200_OK;
}

extern "C" int b200_graph_end(b200_ctx_t ctx, b200_graph_t *out) {
    CHECK_CTX(ctx);
    B200_REQUIRE(out != nullptr, "null output pointer");
    *out = nullptr;
    b200_graph_s *g = ctx->recording;
    B200_REQUIRE(g != nullptr, "graph_end: not recording");
    GUARD(ctx);
    ctx->recording = nullptr;
    cudaError_t rc = cudaStreamEndCapture(ctx->stream, &g->graph);
    if (rc == cudaSuccess) rc = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (rc == cudaSuccess) rc = cudaGraphGetNodes(g->graph, nullptr, &g->nodes);
    if (rc == cudaSuccess) rc = cudaGraphLaunch(g->exec, ctx->stream);     // the recorded calls run now
    if (rc != cudaSuccess) {
        cudaGetLastError();
        graph_rollback(g);
        ctx->launches = g->launches0;
        graph_free(g);
        graph_release_deferred(ctx);
        return cuda_fail(rc, "graph_end (capture / instantiate / launch)", __FILE__, __LINE__);
    }
    graph_release_deferred(ctx);
    for (GraphSlot &s : g->slots) {
        s.p1 = *s.slot;
        s.z1 = s.zp ? *s.zp : false;
    }
    g->launches = ctx->launches - g->launches0;
    *out = g;
    return B200_OK;
}

extern "C" int b200_graph_launch(b200_ctx_t ctx, b200_graph_t g, int *launched) {
    CHECK_CTX(ctx);
    B200_REQUIRE(g && launched, "null argument");
    *launched = 0;
    B200_REQUIRE(g->ctx == ctx, "graph belongs to another context");
    B200_REQUIRE(!ctx->recording, "graph_launch: a graph is being recorded");
    if (ctx->profiling || !ctx->opt_cycle_graph) return B200_OK;
    if (g->destroy_epoch != ctx->destroy_epoch || g->option_epoch != ctx->option_epoch)
        return B200_OK;                      // stale: the caller records a new one
    for (const GraphSlot &s : g->slots)
        if (*s.slot != s.p0 || (s.zp && *s.zp != s.z0)) return B200_OK;
    GUARD(ctx);
    B200_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
    for (const GraphSlot &s : g->slots) {
        *s.slot = s.p1;
        if (s.zp) *s.zp = s.z1;
    }
    ctx->launches += g->launches;
    g->replays++;
    *launched = 1;
    return B200_OK;
}

extern "C" int b200_graph_info(b200_graph_t g, int64_t *kernels, int64_t *nodes, int64_t *replays,
                               int *stale) {
    B200_REQUIRE(g != nullptr, "null argument");
    if (kernels) *kernels = (int64_t)g->launches;
    if (nodes) *nodes = (int64_t)g->nodes;
    if (replays) *replays = (int64_t)g->replays;
    if (stale)
        *stale = (g->destroy_epoch != g->ctx->destroy_epoch || g->option_epoch != g->ctx->option_epoch);
    return B200_OK;
}

extern "C" int b200_graph_destroy(b200_graph_t g) {
    if (!g) return B200_OK;
    GUARD(g->ctx);
    graph_free(g);
    return B200_OK;
}
