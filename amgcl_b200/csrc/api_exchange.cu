// api_exchange.cu -- multi-GPU: NCCL / peer-memory set-up, partition views, the exchange steps
//
// Part of the implementation of the C ABI declared in include/amgcl_b200.h (host-side logic
// only: argument checking, bookkeeping, kernel launches; no CPU fallback anywhere).
#include "internal.cuh"
#include "peer.cuh"

using namespace b200;

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
    B200_CUDA(cudaMalloc(&ctx->gather_ticket, sizeof(unsigned int)));
    B200_CUDA(cudaMemset(ctx->gather_ticket, 0, sizeof(unsigned int)));
    B200_CUDA(cudaMalloc(&ctx->ipc_dev, (size_t)kMaxRanks * sizeof(cudaIpcMemHandle_t) + 64));

    // peer-memory exchange: try to map a small buffer of every peer; agree collectively
    ctx->p2p = false;
    if (ctx->opt_p2p && nranks > 1 && nranks <= kMaxRanks) {
        // (peer_alloc agrees on success collectively: either every rank mapped every peer or none did)
        ctx->p2p = peer_alloc(ctx, kFlagBytes + 2 * 256, &ctx->probe_pb_local, ctx->probe_pb_peer) == B200_OK;
        if (!ctx->p2p) cudaGetLastError();
    }
    if (ctx->p2p) {
        // exchange buffers of the device scalar table: in-kernel all-reduce (reduce.cuh)
        void *local = nullptr;
        int rc = peer_alloc(ctx, sizeof(ScalExchange), &local, ctx->scal_x_peer);
        if (rc) return rc;
        ctx->scal_x_local = static_cast<ScalExchange *>(local);
        B200_CUDA(cudaMalloc(&ctx->scal_x_table, kMaxRanks * sizeof(ScalExchange *)));
        B200_CUDA(cudaMemcpy(ctx->scal_x_table, ctx->scal_x_peer, kMaxRanks * sizeof(void *),
                             cudaMemcpyHostToDevice));
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
    B200_REQUIRE(kind >= 1 && kind <= 3, "kind must be 1 (columns partitioned) or 2 (columns replicated)");
    b200_split_s *sp = new (std::nothrow) b200_split_s();
    if (!sp) return fail(B200_ENOMEM, "out of host memory");
    sp->kind = kind;
    // whole rows of `rank`; kind 1 (and the legacy 3): the vector the operator is applied to is
    // partitioned -> local columns + halo slots; kind 2: it is replicated -> columns untouched
    split_rows(Partition(nrows, nranks), Partition(ncols, nranks), kind != 2, rank, ptr, col, sp->m);
    const int64_t nnz = sp->m.ptr.back();
    sp->val.assign(val + sp->m.val_offset, val + sp->m.val_offset + nnz);
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

namespace b200 {

// Collective: every rank allocates `bytes` (zero filled) and maps the allocations of all
// its peers through CUDA IPC.  peers[rank] is the local pointer.
int peer_alloc(b200_ctx_t ctx, size_t bytes, void **local, void **peers) {
    // Collective.  A failure on one rank must not leave the others blocked inside a
    // collective: every rank records its local status, ALWAYS takes part in the all-gather of
    // the handles and in a final all-reduce(min) of the status, and only then cleans up and
    // returns the agreed result.
    bytes = (bytes + 255) & ~size_t(255);
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    char *stage = static_cast<char *>(ctx->ipc_dev);
    *local = nullptr;
    for (int q = 0; q < ctx->nranks; ++q) peers[q] = nullptr;
    cudaError_t lrc = cudaMalloc(local, bytes);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (lrc == cudaSuccess) lrc = cudaMemsetAsync(*local, 0, bytes, ctx->stream);
    if (lrc == cudaSuccess) lrc = cudaIpcGetMemHandle(&mine, *local);
    if (lrc != cudaSuccess) cudaGetLastError();
    // handles travel through a device staging buffer (NCCL works on device memory)
    cudaError_t crc = cudaMemcpyAsync(stage + ctx->rank * hs, &mine, hs, cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t nrc = nccl().AllGather(stage + ctx->rank * hs, stage, hs, ncclChar, comm_of(ctx), ctx->stream);
    std::vector<cudaIpcMemHandle_t> all((size_t)ctx->nranks);
    if (crc == cudaSuccess) crc = cudaMemcpyAsync(all.data(), stage, hs * ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream);
    if (crc == cudaSuccess) crc = cudaStreamSynchronize(ctx->stream);
    int ok = (lrc == cudaSuccess && crc == cudaSuccess && nrc == ncclSuccess) ? 1 : 0;
    // every rank learns whether every allocation exists before anybody maps anything
    int *flag_d = reinterpret_cast<int *>(stage + (size_t)kMaxRanks * hs);
    cudaMemcpyAsync(flag_d, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    nccl().AllReduce(flag_d, flag_d, 1, ncclInt, ncclMin, comm_of(ctx), ctx->stream);
    int all_ok = 0;
    cudaMemcpyAsync(&all_ok, flag_d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (all_ok) {
        for (int q = 0; q < ctx->nranks && all_ok; ++q) {
            if (q == ctx->rank) { peers[q] = *local; continue; }
            if (cudaIpcOpenMemHandle(&peers[q], all[(size_t)q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                peers[q] = nullptr;
                all_ok = 0;
            }
        }
        // (mapping failures are local; agree once more so that all ranks fail together)
        ok = all_ok;
        cudaMemcpyAsync(flag_d, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        nccl().AllReduce(flag_d, flag_d, 1, ncclInt, ncclMin, comm_of(ctx), ctx->stream);
        cudaMemcpyAsync(&all_ok, flag_d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    if (!all_ok) {
        for (int q = 0; q < ctx->nranks; ++q) {
            if (q != ctx->rank && peers[q]) cudaIpcCloseMemHandle(peers[q]);
            peers[q] = nullptr;
        }
        if (*local) cudaFree(*local);
        *local = nullptr;
        cudaGetLastError();
        return fail(B200_ECUDA, "peer-memory exchange buffer: allocation or CUDA-IPC mapping failed on some rank");
    }
    return B200_OK;
}

void peer_release(b200_ctx_t ctx, void *local, void **peers) {
    if (!local) return;
    for (int q = 0; q < ctx->nranks; ++q)
        if (q != ctx->rank && peers[q]) cudaIpcCloseMemHandle(peers[q]);
    // peers may still have this allocation mapped: keep it until the context dies
    ctx->deferred_free.push_back(local);
}


// HALO operators: make the boundary values of x visible to every rank that gathers them.
// Peer transport: nothing is launched here -- the consumer kernel pushes this rank's values
// itself (HaloArgs::push_*) and waits for the peers' flags block by block.  NCCL transport: one
// pack kernel + one in-place ncclAllGather (S doubles per rank) on the stream.
int halo_exchange(b200_ctx_t ctx, b200_csr_t A, const void *x, size_t esz, HaloArgs &a) {
    a = HaloArgs();
    {
        const int trc = tail_flush(ctx);
        if (trc) return trc;
    }
    a.xh = A->halo; a.nloc = (int)A->n_loc;
    if (A->S == 0) return B200_OK;
    if (ctx->p2p) {
        const int par = (int)(A->seq & 1);
        const unsigned long long seq = ++A->seq;
        unsigned int mask = 0;
        for (int q = 0; q < ctx->nranks; ++q) {
            if (q == ctx->rank || !A->xchg[q]) continue;
            a.push_data[q] = data_at(A->pb_peer[q], par, A->pb_half) + (size_t)ctx->rank * A->S * esz;
            a.push_flag[q] = flag_at(A->pb_peer[q], par, ctx->rank);
            mask |= 1u << q;
        }
        A->halo = data_at(A->pb_local, par, A->pb_half);   // what the kernel gathers from
        a.xh = A->halo;
        a.send_idx = A->send_idx;
        a.n_send = (int)A->n_send;
        a.nranks = ctx->nranks;
        a.push_ticket = ctx->push_ticket;
        a.push_seq = mask ? seq : 0;
        a.wait_flags = flag_at(A->pb_local, par, 0);
        a.wait_mask = mask;
        a.wait_seq = seq;
        return B200_OK;
    }
    ProfScope prof(ctx, B200_PROF_COMM, A->n_send, ctx->nranks, 0);
    char *mine = static_cast<char *>(A->halo) + (size_t)ctx->rank * A->S * esz;
    if (A->n_send) {
        const unsigned grid = (unsigned)((A->n_send + kThreads - 1) / kThreads);
        if (esz == sizeof(double))
            halo_pack_kernel<double><<<grid, kThreads, 0, ctx->stream>>>(A->n_send, A->send_idx,
                static_cast<const double *>(x), reinterpret_cast<double *>(mine));
        else
            halo_pack_kernel<float><<<grid, kThreads, 0, ctx->stream>>>(A->n_send, A->send_idx,
                static_cast<const float *>(x), reinterpret_cast<float *>(mine));
        B200_CHECK_LAUNCH();
        ctx->launches++;
    }
    B200_NCCL(nccl().AllGather(mine, A->halo, (size_t)A->S, esz == sizeof(double) ? ncclDouble : ncclFloat,
                               comm_of(ctx), ctx->stream));
    return B200_OK;
}

// Row shares of a replicated result (gather_rows).  begin: where the kernel stores its rows.
// Peer transport: straight into every rank's gather buffer (a.gather_*), local y = own buffer.
// NCCL transport: into this rank's slot of A->ybuf.
int gather_begin(b200_ctx_t ctx, b200_csr_t A, size_t esz, GatherArgs &g) {
    g = GatherArgs();
    if (ctx->p2p) {
        const int par = (int)(A->gseq & 1);
        const unsigned long long seq = ++A->gseq;
        for (int q = 0; q < ctx->nranks; ++q) {
            g.data[q] = data_at(A->gb_peer[q], par, A->gb_half) + (size_t)ctx->rank * A->row_B * esz;
            g.flag[q] = flag_at(A->gb_peer[q], par, ctx->rank);
        }
        g.on = 1;
        g.nranks = ctx->nranks;
        g.ticket = ctx->gather_ticket;
        g.seq = seq;
        g.y_local = g.data[ctx->rank];           // the plain store of a row goes here as well
        g.data[ctx->rank] = nullptr;             // (so it is not stored twice)
        return B200_OK;
    }
    g.y_local = reinterpret_cast<char *>(A->ybuf) + (size_t)ctx->rank * A->row_B * esz;
    return B200_OK;
}

// end: all shares -> the replicated vector y (every rank ends up with the complete result)
int gather_end(b200_ctx_t ctx, b200_csr_t A, const GatherArgs &g, b200_vec_t y) {
    {
        const int trc = tail_flush(ctx);
        if (trc) return trc;
    }
    ProfScope prof(ctx, B200_PROF_COMM, A->gl_rows, ctx->nranks, 2);
    const int64_t count = A->gl_rows;
    const size_t esz = y->esz;
    if (ctx->p2p) {
        const int par = (int)((A->gseq - 1) & 1);
        WaitList w;
        for (int q = 0; q < kMaxRanks; ++q) w.flag[q] = nullptr;
        for (int q = 0; q < ctx->nranks; ++q) w.flag[q] = flag_at(A->gb_local, par, q);
        const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((count + kThreads * 4 - 1) / (kThreads * 4),
                                                                               (int64_t)ctx->sm_count * 4));
        const char *staged = data_at(A->gb_local, par, A->gb_half);
        if (esz == sizeof(double))
            gather_copy_kernel<double><<<grid, kThreads, 0, ctx->stream>>>(
                count, reinterpret_cast<const double *>(staged), w, ctx->nranks, g.seq, wr(y));
        else
            gather_copy_kernel<float><<<grid, kThreads, 0, ctx->stream>>>(
                count, reinterpret_cast<const float *>(staged), w, ctx->nranks, g.seq, tp<float>(wr(y)));
        B200_CHECK_LAUNCH();
        ctx->launches++;
        return B200_OK;
    }
    B200_NCCL(nccl().AllGather(g.y_local, A->ybuf, (size_t)A->row_B, esz == sizeof(double) ? ncclDouble : ncclFloat,
                               comm_of(ctx), ctx->stream));
    B200_CUDA(cudaMemcpyAsync(wr(y), A->ybuf, (size_t)count * esz, cudaMemcpyDeviceToDevice, ctx->stream));
    return B200_OK;
}

// Partitioned inner product: the local partial sum is in ctx->dot_dev; combine the ranks'
// partials and hand the result to the host (mpi/inner_product.hpp:53-62 does the same with
// MPI_Allreduce on the host).
int dist_dot_finish(b200_ctx_t ctx, double *result) {
    // (NCCL transport; with the peer-memory transport the all-reduce happens inside the
    // reducing kernel, reduce.cuh)
    B200_NCCL(nccl().AllReduce(ctx->dot_dev, ctx->dot_dev, 1, ncclDouble, ncclSum, comm_of(ctx), ctx->stream));
    B200_CUDA(cudaMemcpyAsync(ctx->dot_result_h, ctx->dot_dev, sizeof(double), cudaMemcpyDeviceToHost,
                              ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    *result = *reinterpret_cast<volatile double *>(ctx->dot_result_h);
    return B200_OK;
}

} // namespace b200
