// api_krylov.cu -- the device scalar table, in-kernel reductions, and the fused Krylov steps
//
// Part of the implementation of the C ABI declared in include/amgcl_b200.h (host-side logic
// only: argument checking, bookkeeping, kernel launches; no CPU fallback anywhere).
//
// What the reference does per CG iteration (solver/cg.hpp:180-198): P.apply; dot; axpby|copy;
// spmv; dot; axpby; axpby; dot -- seven backend calls, three of them host-synchronous.  Here:
// the V-cycle's last smoother sweep leaves <r,s> in the scalar table, b200_cg_direction is one
// pass, b200_cg_step is the SpMV (leaving <q,p>) plus one pass that forms alpha on the device,
// updates x and r and leaves <r,r>; the host synchronises once per iteration to read <r,r>.
// BiCGStab (solver/bicgstab.hpp:198-236) likewise: two host syncs (its two convergence tests)
// instead of six.
#include "internal.cuh"
#include "krylov_kernels.cuh"

using namespace b200;

// table slots with a fixed role
enum { SLOT_DOT = 0, SLOT_DOT2 = 1, SLOT_PRODUCT0 = 2, SLOT_FIRST_FREE = 6 };

// slots of one Krylov workspace (relative to its base)
enum {
    K_RHO = 0,      // <r,s> (CG) / <r,rh> (BiCGStab): own copy when not taken from a product
    K_RHOP0, K_RHOP1,   // rho of the previous iteration, double-buffered by iteration parity
    K_QP,           // <q,p> (CG) / <rh,v> (BiCGStab)
    K_RR,           // <r,r>
    K_SS,           // <s,s>
    K_TS, K_TT,     // <t,s>, <t,t>
    K_ALPHA, K_OMEGA,
    K_RHO_NEXT,     // <r,rh> left by bicg_update_r for the next iteration
    K_NSLOTS = 16
};

struct b200_krylov_s {
    b200_ctx_t ctx = nullptr;
    size_t     n = 0;
    int        base = 0;        // first table slot
    bool       first = true;    // no search direction computed yet in this solve
    int        parity = 0;      // K_RHOP0 + parity holds the previous rho
    int        rho_slot = -1;   // absolute slot of the current rho (own or a product's)
};

namespace b200 {

int scal_create(b200_ctx_t ctx) {
    B200_CUDA(cudaMalloc(&ctx->scal_d, kScalSlots * sizeof(double)));
    B200_CUDA(cudaMemset(ctx->scal_d, 0, kScalSlots * sizeof(double)));
    B200_CUDA(cudaHostAlloc(&ctx->scal_h, kScalSlots * sizeof(double), cudaHostAllocMapped));
    memset(ctx->scal_h, 0, kScalSlots * sizeof(double));
    B200_CUDA(cudaHostGetDevicePointer(&ctx->scal_hd, ctx->scal_h, 0));
    B200_CUDA(cudaHostAlloc(&ctx->scal_ready_h, kScalSlots * sizeof(unsigned long long), cudaHostAllocMapped));
    memset(ctx->scal_ready_h, 0, kScalSlots * sizeof(unsigned long long));
    B200_CUDA(cudaHostGetDevicePointer(&ctx->scal_ready_hd, ctx->scal_ready_h, 0));
    B200_CUDA(cudaMalloc(&ctx->red_partial, (size_t)kMaxRed * kDotMaxBlocks * sizeof(double)));
    B200_CUDA(cudaMalloc(&ctx->red_ticket, sizeof(unsigned int)));
    B200_CUDA(cudaMemset(ctx->red_ticket, 0, sizeof(unsigned int)));
    for (int i = 0; i < SLOT_FIRST_FREE; ++i) ctx->scal_used[i] = true;
    ctx->product_slot0 = SLOT_PRODUCT0;
    return B200_OK;
}

void scal_destroy(b200_ctx_t ctx) {
    if (ctx->scal_d) cudaFree(ctx->scal_d);
    if (ctx->scal_h) cudaFreeHost(ctx->scal_h);
    if (ctx->scal_ready_h) cudaFreeHost(ctx->scal_ready_h);
    ctx->scal_ready_h = ctx->scal_ready_hd = nullptr;
    if (ctx->red_partial) cudaFree(ctx->red_partial);
    if (ctx->red_ticket) cudaFree(ctx->red_ticket);
    if (ctx->scal_x_table) cudaFree(ctx->scal_x_table);
    ctx->scal_d = ctx->scal_h = ctx->scal_hd = ctx->red_partial = nullptr;
    ctx->red_ticket = nullptr;
    ctx->scal_x_table = nullptr;
}

int scal_alloc(b200_ctx_t ctx, int count) {
    for (int first = SLOT_FIRST_FREE; first + count <= kScalSlots; ++first) {
        bool ok = true;
        for (int i = 0; i < count && ok; ++i) ok = !ctx->scal_used[first + i];
        if (ok) {
            for (int i = 0; i < count; ++i) ctx->scal_used[first + i] = true;
            return first;
        }
    }
    return -1;
}

void scal_free(b200_ctx_t ctx, int first, int count) {
    for (int i = 0; i < count; ++i) ctx->scal_used[first + i] = false;
}

void red_out(b200_ctx_t ctx, int nred, const int *slots, RedOut &o, bool across_ranks, unsigned host_mask) {
    memset(&o, 0, sizeof(o));
    o.partial = ctx->red_partial;
    o.ticket = ctx->red_ticket;
    o.nred = nred;
    o.nranks = (across_ranks && ctx->dist && ctx->scal_x_table) ? ctx->nranks : 1;
    o.rank = ctx->rank;
    o.peers = ctx->scal_x_table;
    for (int k = 0; k < nred; ++k) {
        o.dev[k] = ctx->scal_d + slots[k];
        o.host[k] = (host_mask >> k) & 1u ? ctx->scal_hd + slots[k] : nullptr;
        o.host_seq[k] = ctx->scal_ready_hd + slots[k];
        o.slot[k] = slots[k];
        o.seq[k] = ++ctx->scal_seq[slots[k]];
    }
}

// ---- products left behind by producer kernels ------------------------------------------------
bool product_wanted(b200_ctx_t ctx, size_t n) {
    for (size_t m : ctx->krylov_sizes)
        if (m == n) return true;
    return false;
}

int product_take_slot(b200_ctx_t ctx) {
    const int k = ctx->product_next;
    ctx->product_next = (k + 1) & 3;
    ctx->products[k] = b200_ctx_s::Product();       // whatever it held is about to be overwritten
    return ctx->product_slot0 + k;
}

void product_record(b200_ctx_t ctx, b200_vec_t a, b200_vec_t b, int slot) {
    if (a->escaped || b->escaped) return;
    b200_ctx_s::Product &p = ctx->products[slot - ctx->product_slot0];
    p.a = a; p.b = b; p.gen_a = a->gen; p.gen_b = b->gen; p.slot = slot;
    if (ctx->recording) ctx->recording->products.push_back({a, b, slot});
}

int product_lookup(b200_ctx_t ctx, b200_vec_t a, b200_vec_t b) {
    for (const b200_ctx_s::Product &p : ctx->products) {
        if (p.slot < 0) continue;
        const bool same = (p.a == a && p.b == b && p.gen_a == a->gen && p.gen_b == b->gen) ||
                          (p.a == b && p.b == a && p.gen_a == b->gen && p.gen_b == a->gen);
        if (same) return p.slot;
    }
    return -1;
}

// ---- launch helper for the fused vector passes ---------------------------------------------------
template <class F, int UNR>
static int launch_fused(b200_ctx_t ctx, size_t len, const F &f, const FusedArgs<F::NIN, F::NOUT> &a,
                        const int *slots, int prof_streams, bool across_ranks, unsigned host_mask = 0) {
    RedOut ro;
    memset(&ro, 0, sizeof(ro));
    if (F::NRED > 0) red_out(ctx, F::NRED, slots, ro, across_ranks, host_mask);
    bool vec_ok = true;
    for (int k = 0; k < F::NIN; ++k) vec_ok = vec_ok && aligned16(a.in[k]);
    for (int k = 0; k < F::NOUT; ++k) vec_ok = vec_ok && aligned16(a.out[k]);
    int grid = grid_for(ctx, len, 2 * UNR);
    if (grid > kDotMaxBlocks) grid = kDotMaxBlocks;
    ProfScope prof(ctx, F::NOUT == 0 ? B200_PROF_DOT : B200_PROF_VECTOR + prof_streams - 2, (int64_t)len, 1, 0);
    B200_CUDA(launch_pdl(ctx, fused_vec_kernel<F, UNR>, dim3(grid), dim3(kThreads), 0, len, f, a, ro, vec_ok));
    B200_CHECK_LAUNCH();
    ctx->launches++;
    return B200_OK;
}

// <x,y> (and <x,z>) of FP64 vectors into table slots; one launch, the all-reduce over the
// ranks of a partitioned vector happens inside it
int launch_dot_slots(b200_ctx_t ctx, b200_vec_t x, b200_vec_t y, b200_vec_t z, const int *slots,
                     unsigned host_mask) {
    B200_REQUIRE(x->dtype == B200_F64 && y->dtype == B200_F64 && (!z || z->dtype == B200_F64),
                 "in-kernel reductions need FP64 vectors");
    const double *px, *py, *pz = nullptr;
    int rc = rd(x, &px);
    if (rc) return rc;
    rc = rd(y, &py);
    if (rc) return rc;
    if (z) {
        rc = rd(z, &pz);
        if (rc) return rc;
        FusedArgs<3, 0> a;
        a.in[0] = px; a.in[1] = py; a.in[2] = pz; a.out[0] = nullptr;
        return launch_fused<Dot2F, 2>(ctx, x->len, Dot2F(), a, slots, 0, x->kind == B200_VK_DIST, host_mask);
    }
    if (px == py) {
        FusedArgs<1, 0> a;
        a.in[0] = px; a.out[0] = nullptr;
        return launch_fused<NormF, 4>(ctx, x->len, NormF(), a, slots, 0, x->kind == B200_VK_DIST, host_mask);
    }
    FusedArgs<2, 0> a;
    a.in[0] = px; a.in[1] = py; a.out[0] = nullptr;
    return launch_fused<DotF, 4>(ctx, x->len, DotF(), a, slots, 0, x->kind == B200_VK_DIST, host_mask);
}

static SaveSlot save_slot(b200_ctx_t ctx, int slot) {
    SaveSlot s;
    s.d = ctx->scal_d + slot;
    s.h = ctx->scal_hd + slot;
    return s;
}

int scal_read(b200_ctx_t ctx, int slot, bool mirrored, double *out) {
    if (!mirrored) {
        B200_CUDA(cudaMemcpyAsync(ctx->scal_h + slot, ctx->scal_d + slot, sizeof(double),
                                  cudaMemcpyDeviceToHost, ctx->stream));
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        *out = *reinterpret_cast<volatile double *>(ctx->scal_h + slot);
        return B200_OK;
    }
    // The kernel that finishes the reduction writes the value into mapped host memory and then
    // releases a "ready" word carrying the launch's sequence number: poll that word instead of
    // synchronising the stream (a cudaStreamSynchronize costs ~10 us of driver latency, once per
    // Krylov iteration).  The stream is queried now and then so a failed launch cannot hang us.
    if (ctx->opt_poll_scalars) {
        const unsigned long long want = ctx->scal_seq[slot];
        volatile unsigned long long *ready = ctx->scal_ready_h + slot;
        for (unsigned long long spins = 1;; ++spins) {
            if (*ready >= want) {
                *out = *reinterpret_cast<volatile double *>(ctx->scal_h + slot);
                return B200_OK;
            }
            if ((spins & 0x3ffff) == 0) {
                const cudaError_t q = cudaStreamQuery(ctx->stream);
                if (q == cudaSuccess) break;                 // stream drained: the value must be there
                if (q != cudaErrorNotReady) return cuda_fail(q, "cudaStreamQuery", __FILE__, __LINE__);
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = *reinterpret_cast<volatile double *>(ctx->scal_h + slot);
    return B200_OK;
}
static int sync_read(b200_ctx_t ctx, int slot, double *out) { return scal_read(ctx, slot, true, out); }

} // namespace b200

// ---------------------------------------------------------------------------
// inner product (host-synchronous; interface.hpp:356-371)
// ---------------------------------------------------------------------------
extern "C" int b200_dot(b200_ctx_t ctx, b200_vec_t x, b200_vec_t y, double *result) {
    CHECK_CTX(ctx);
    B200_REQUIRE(x && y && result, "null argument");
    NOT_RECORDING(ctx, "dot (host-synchronous)");
    B200_REQUIRE(same_layout(x, y), "dot: size mismatch");
    if (x->dtype != y->dtype) return B200_BAD_MIX("dot");
    GUARD(ctx);
    const bool dist = x->kind == B200_VK_DIST;
    if (!dist && (x->len == 0 || x->zero_pending || y->zero_pending)) {
        B200_CUDA(cudaStreamSynchronize(ctx->stream));
        *result = 0.0;
        return B200_OK;
    }
    if (x->dtype == B200_F32 || (dist && !ctx->scal_x_table)) return dot_legacy(ctx, x, y, result);
    // a producer kernel may already have left this very product in the table
    const int have = product_lookup(ctx, x, y);
    if (have >= 0) return scal_read(ctx, have, false, result);
    const int slot = SLOT_DOT;
    int rc = launch_dot_slots(ctx, x, y, nullptr, &slot, 1u);
    if (rc) return rc;
    return sync_read(ctx, slot, result);
}

// ---------------------------------------------------------------------------
// Krylov workspace
// ---------------------------------------------------------------------------
extern "C" int b200_krylov_create(b200_ctx_t ctx, size_t n, b200_krylov_t *out) {
    CHECK_CTX(ctx);
    B200_REQUIRE(out != nullptr, "null output pointer");
    *out = nullptr;
    NOT_RECORDING(ctx, "Krylov workspace creation");
    // multi-GPU: the reductions are all-reduced inside the kernels through peer memory; with the
    // NCCL transport (option "p2p" = 0) the solvers issue the reference's sequence instead
    if (ctx->dist && !ctx->scal_x_table)
        return fail(B200_EINVAL, "fused Krylov steps need the peer-memory transport on a multi-GPU context");
    const int base = scal_alloc(ctx, K_NSLOTS);
    if (base < 0) return fail(B200_ENOMEM, "scalar table exhausted (too many live Krylov workspaces)");
    b200_krylov_s *K = new (std::nothrow) b200_krylov_s();
    if (!K) {
        scal_free(ctx, base, K_NSLOTS);
        return fail(B200_ENOMEM, "out of host memory");
    }
    K->ctx = ctx; K->n = n; K->base = base;
    ctx->krylov_sizes.push_back(n);
    *out = K;
    return B200_OK;
}

extern "C" int b200_krylov_destroy(b200_krylov_t K) {
    if (!K) return B200_OK;
    b200_ctx_t ctx = K->ctx;
    scal_free(ctx, K->base, K_NSLOTS);
    for (size_t i = 0; i < ctx->krylov_sizes.size(); ++i)
        if (ctx->krylov_sizes[i] == K->n) {
            ctx->krylov_sizes.erase(ctx->krylov_sizes.begin() + (long)i);
            break;
        }
    delete K;
    return B200_OK;
}

#define CHECK_K(K)                                                             \
    B200_REQUIRE((K) != nullptr, "null Krylov workspace");                     \
    b200_ctx_t ctx = (K)->ctx;                                                 \
    NOT_RECORDING(ctx, "Krylov step")

namespace b200 {
static bool krylov_vec_ok(b200_krylov_t K, std::initializer_list<b200_vec_t> vs) {
    b200_vec_t first = *vs.begin();
    for (b200_vec_t v : vs) {
        if (!v || v->dtype != B200_F64 || v->n != K->n || !same_layout(v, first)) return false;
    }
    return true;
}
} // namespace b200

extern "C" int b200_krylov_scalars(b200_krylov_t K, double *out, int count) {
    CHECK_K(K);
    B200_REQUIRE(out != nullptr && count >= 0 && count <= 9, "bad argument");
    GUARD(ctx);
    // (not every scalar is mirrored to the host by the kernel that forms it: copy the
    // workspace's part of the table)
    double tab[K_NSLOTS];
    B200_CUDA(cudaMemcpyAsync(tab, ctx->scal_d + K->base, sizeof(tab), cudaMemcpyDeviceToHost, ctx->stream));
    B200_CUDA(cudaStreamSynchronize(ctx->stream));
    // logical layout: rho (this iteration's), <q,p> | <rh,v>, alpha, <t,s>, <t,t>, omega, <r,r>,
    // <s,s>, next rho
    const int rel[9] = {K_RHOP0 + K->parity, K_QP, K_ALPHA, K_TS, K_TT, K_OMEGA, K_RR, K_SS, K_RHO_NEXT};
    for (int i = 0; i < count; ++i) out[i] = tab[rel[i]];
    return B200_OK;
}

extern "C" int b200_krylov_residual(b200_krylov_t K, b200_vec_t rhs, b200_csr_t A, b200_vec_t x,
                                    b200_vec_t r, double *rr) {
    CHECK_K(K);
    B200_REQUIRE(rhs && A && x && r && rr, "null argument");
    B200_REQUIRE(krylov_vec_ok(K, {rhs, x, r}), "krylov_residual: FP64 vectors of the workspace's size expected");
    GUARD(ctx);
    K->first = true;
    K->parity = 0;
    K->rho_slot = -1;
    const int slot = K->base + K_RR;
    int rc;
    if (x->zero_pending && ctx->opt_zero_shortcut && r != rhs) {
        // rhs - A*0 == rhs exactly: r = rhs and <r,r> in one element-wise pass
        touch(ctx, {rhs, x, r});
        const double *pf;
        rc = rd(rhs, &pf);
        if (rc) return rc;
        FusedArgs<1, 1> a;
        a.in[0] = pf; a.out[0] = wr(r);
        rc = launch_fused<CopyNormF, 2>(ctx, r->len, CopyNormF(), a, &slot, 3, r->kind == B200_VK_DIST, 1u);
    } else {
        rc = residual_with_norm(ctx, rhs, A, x, r, slot);
    }
    if (rc) return rc;
    return sync_read(ctx, slot, rr);
}

// ---------------------------------------------------------------------------
// CG (solver/cg.hpp:180-198)
// ---------------------------------------------------------------------------
extern "C" int b200_cg_direction(b200_krylov_t K, b200_vec_t r, b200_vec_t s, b200_vec_t p) {
    CHECK_K(K);
    B200_REQUIRE(r && s && p, "null argument");
    B200_REQUIRE(krylov_vec_ok(K, {r, s, p}), "cg_direction: FP64 vectors of the workspace's size expected");
    B200_REQUIRE(s != p && s->ptr != p->ptr, "cg_direction: s and p must be distinct");
    GUARD(ctx);
    // rho = <r,s>: normally left behind by the V-cycle's last smoother sweep
    int rho = product_lookup(ctx, r, s);
    if (rho < 0) {
        rho = K->base + K_RHO;
        int rc = launch_dot_slots(ctx, r, s, nullptr, &rho, 0u);
        if (rc) return rc;
    }
    K->rho_slot = rho;
    const double *ps, *pp = nullptr;
    int rc = rd(s, &ps);
    if (rc) return rc;
    if (!K->first) {
        rc = rd(p, &pp);
        if (rc) return rc;
    }
    CgDirectionF f;
    f.rho = ctx->scal_d + rho;
    f.rho_prev = ctx->scal_d + K->base + K_RHOP0 + K->parity;
    f.rho_save = save_slot(ctx, K->base + K_RHOP0 + (K->parity ^ 1));
    f.first = K->first ? 1 : 0;
    f.beta = 0.0;
    FusedArgs<2, 1> a;
    a.in[0] = ps;
    a.in[1] = K->first ? ps : pp;           // p is not read on the first iteration
    a.out[0] = K->first ? wr(p) : mut(p);
    rc = launch_fused<CgDirectionF, 2>(ctx, p->len, f, a, nullptr, K->first ? 2 : 3, false);
    if (rc) return rc;
    K->parity ^= 1;                          // K_RHOP0 + parity now holds this iteration's rho
    K->first = false;
    return B200_OK;
}

extern "C" int b200_cg_step(b200_krylov_t K, b200_csr_t A, b200_vec_t p, b200_vec_t q, b200_vec_t x,
                            b200_vec_t r, double *rr) {
    CHECK_K(K);
    B200_REQUIRE(A && p && q && x && r && rr, "null argument");
    B200_REQUIRE(krylov_vec_ok(K, {p, q, x, r}), "cg_step: FP64 vectors of the workspace's size expected");
    B200_REQUIRE(!K->first, "cg_step: call b200_cg_direction first");
    GUARD(ctx);
    // q = A p, leaving <q,p>
    const int qp = K->base + K_QP;
    int rc = spmv_with_dots(ctx, A, p, q, p, 1, &qp);
    if (rc) return rc;
    // alpha = rho/<q,p> ; x += alpha p ; r -= alpha q ; <r,r>
    const double *pp, *pq, *px, *pr;
    if ((rc = rd(p, &pp)) || (rc = rd(q, &pq)) || (rc = rd(x, &px)) || (rc = rd(r, &pr))) return rc;
    CgUpdateF f;
    f.rho = ctx->scal_d + K->base + K_RHOP0 + K->parity;     // saved by cg_direction
    f.qp = ctx->scal_d + qp;
    f.alpha_save = save_slot(ctx, K->base + K_ALPHA);
    f.alpha = 0.0;
    FusedArgs<4, 2> a;
    a.in[0] = pp; a.in[1] = pq; a.in[2] = px; a.in[3] = pr;
    a.out[0] = mut(x); a.out[1] = mut(r);
    const int slot = K->base + K_RR;
    rc = launch_fused<CgUpdateF, 2>(ctx, x->len, f, a, &slot, 6, x->kind == B200_VK_DIST, 1u);
    if (rc) return rc;
    return sync_read(ctx, slot, rr);
}

// ---------------------------------------------------------------------------
// BiCGStab, right preconditioning (solver/bicgstab.hpp:176-236)
// ---------------------------------------------------------------------------
extern "C" int b200_bicg_start(b200_krylov_t K, b200_vec_t r, b200_vec_t rh) {
    CHECK_K(K);
    B200_REQUIRE(r && rh, "null argument");
    B200_REQUIRE(krylov_vec_ok(K, {r, rh}), "bicg_start: FP64 vectors of the workspace's size expected");
    B200_REQUIRE(r != rh && r->ptr != rh->ptr, "bicg_start: r and rh must be distinct");
    GUARD(ctx);
    // rh = r (bicgstab.hpp:183) and rho = <r, rh> of the first iteration (bicgstab.hpp:200)
    const double *pr;
    int rc = rd(r, &pr);
    if (rc) return rc;
    FusedArgs<1, 1> a;
    a.in[0] = pr; a.out[0] = wr(rh);
    const int slot = K->base + K_RHO_NEXT;
    rc = launch_fused<CopyNormF, 2>(ctx, r->len, CopyNormF(), a, &slot, 3, r->kind == B200_VK_DIST);
    if (rc) return rc;
    K->first = true;
    K->parity = 0;
    K->rho_slot = slot;
    return B200_OK;
}

extern "C" int b200_bicg_direction(b200_krylov_t K, b200_vec_t r, b200_vec_t v, b200_vec_t p) {
    CHECK_K(K);
    B200_REQUIRE(r && v && p, "null argument");
    B200_REQUIRE(krylov_vec_ok(K, {r, v, p}), "bicg_direction: FP64 vectors of the workspace's size expected");
    B200_REQUIRE(r != p && r->ptr != p->ptr, "bicg_direction: r and p must be distinct");
    GUARD(ctx);
    const double *pr, *pv = nullptr, *pp = nullptr;
    int rc = rd(r, &pr);
    if (rc) return rc;
    if (!K->first) {
        if ((rc = rd(v, &pv)) || (rc = rd(p, &pp))) return rc;
    }
    BicgDirectionF f;
    f.rho = ctx->scal_d + K->base + K_RHO_NEXT;              // <r,rh> left by bicg_start / step_r
    f.rho_prev = ctx->scal_d + K->base + K_RHOP0 + K->parity;
    f.rho_save = save_slot(ctx, K->base + K_RHOP0 + (K->parity ^ 1));
    f.alpha = ctx->scal_d + K->base + K_ALPHA;
    f.omega = ctx->scal_d + K->base + K_OMEGA;
    f.first = K->first ? 1 : 0;
    f.b = f.c = 0.0;
    FusedArgs<3, 1> a;
    a.in[0] = pr;
    a.in[1] = K->first ? pr : pv;
    a.in[2] = K->first ? pr : pp;
    a.out[0] = K->first ? wr(p) : mut(p);
    rc = launch_fused<BicgDirectionF, 2>(ctx, p->len, f, a, nullptr, K->first ? 2 : 4, false);
    if (rc) return rc;
    K->parity ^= 1;
    K->first = false;
    return B200_OK;
}

extern "C" int b200_bicg_step_s(b200_krylov_t K, b200_csr_t A, b200_vec_t rh, b200_vec_t T, b200_vec_t v,
                                b200_vec_t r, b200_vec_t s, b200_vec_t x, double *ss, double *rho) {
    CHECK_K(K);
    B200_REQUIRE(A && rh && T && v && r && s && x && ss, "null argument");
    B200_REQUIRE(krylov_vec_ok(K, {rh, T, v, r, s, x}), "bicg_step_s: FP64 vectors of the workspace's size expected");
    B200_REQUIRE(!K->first, "bicg_step_s: call b200_bicg_direction first");
    GUARD(ctx);
    // v = A T, leaving <v, rh>
    const int rhv = K->base + K_QP;
    int rc = spmv_with_dots(ctx, A, T, v, rh, 1, &rhv);
    if (rc) return rc;
    const double *pT, *px, *pr, *pv;
    if ((rc = rd(T, &pT)) || (rc = rd(x, &px)) || (rc = rd(r, &pr)) || (rc = rd(v, &pv))) return rc;
    BicgUpdateSF f;
    f.rho = ctx->scal_d + K->base + K_RHOP0 + K->parity;
    f.rhv = ctx->scal_d + rhv;
    f.alpha_save = save_slot(ctx, K->base + K_ALPHA);
    f.alpha = 0.0;
    FusedArgs<4, 2> a;
    a.in[0] = pT; a.in[1] = px; a.in[2] = pr; a.in[3] = pv;
    a.out[0] = mut(x); a.out[1] = wr(s);
    const int slot = K->base + K_SS;
    rc = launch_fused<BicgUpdateSF, 2>(ctx, x->len, f, a, &slot, 6, x->kind == B200_VK_DIST, 1u);
    if (rc) return rc;
    rc = sync_read(ctx, slot, ss);
    // rho of this iteration: mirrored by b200_bicg_direction's kernel
    if (rho) *rho = *reinterpret_cast<volatile double *>(ctx->scal_h + K->base + K_RHOP0 + K->parity);
    return rc;
}

extern "C" int b200_bicg_step_r(b200_krylov_t K, b200_csr_t A, b200_vec_t rh, b200_vec_t T, b200_vec_t t,
                                b200_vec_t s, b200_vec_t r, b200_vec_t x, double *rr, double *omega) {
    CHECK_K(K);
    B200_REQUIRE(A && rh && T && t && s && r && x && rr, "null argument");
    B200_REQUIRE(krylov_vec_ok(K, {rh, T, t, s, r, x}), "bicg_step_r: FP64 vectors of the workspace's size expected");
    GUARD(ctx);
    // t = A T, leaving <t,s> and <t,t>
    const int slots_t[2] = {K->base + K_TS, K->base + K_TT};
    int rc = spmv_with_dots(ctx, A, T, t, s, 2, slots_t);
    if (rc) return rc;
    const double *pT, *px, *ps, *pt, *prh;
    if ((rc = rd(T, &pT)) || (rc = rd(x, &px)) || (rc = rd(s, &ps)) || (rc = rd(t, &pt)) || (rc = rd(rh, &prh)))
        return rc;
    BicgUpdateRF f;
    f.ts = ctx->scal_d + slots_t[0];
    f.tt = ctx->scal_d + slots_t[1];
    f.omega_save = save_slot(ctx, K->base + K_OMEGA);
    f.omega = 0.0;
    FusedArgs<5, 2> a;
    a.in[0] = pT; a.in[1] = px; a.in[2] = ps; a.in[3] = pt; a.in[4] = prh;
    a.out[0] = mut(x); a.out[1] = wr(r);
    const int slots_r[2] = {K->base + K_RR, K->base + K_RHO_NEXT};
    rc = launch_fused<BicgUpdateRF, 2>(ctx, x->len, f, a, slots_r, 7, x->kind == B200_VK_DIST, 1u);
    if (rc) return rc;
    rc = sync_read(ctx, slots_r[0], rr);
    if (omega) *omega = *reinterpret_cast<volatile double *>(ctx->scal_h + K->base + K_OMEGA);
    return rc;
}
