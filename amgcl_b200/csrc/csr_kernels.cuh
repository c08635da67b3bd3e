// csr_kernels.cuh -- the streaming CSR kernels of the V-cycle.
//
// One kernel template covers every pass over a CSR operator in the AMGCL solve
// phase:
//
//   MODE_SPMV   y = alpha*A*x (+ beta*y)      backend::spmv      matrix_ops.hpp:47-83
//   MODE_RESID  y = f - A*x                   backend::residual  matrix_ops.hpp:85-115
//   MODE_RELAX  y = x + (alpha*d).*(f - A*x)  damped_jacobi / spai0 sweep
//                                             damped_jacobi.hpp:103-132, spai0.hpp:86-109
//
// Data movement.  The matrix is cut (on the host, once, at upload) into row
// blocks: runs of consecutive rows starting at a multiple of four whose
// non-zeros fit a shared-memory stage.  A block's slice of `val`, `col` and
// `ptr` is contiguous in global memory, so one elected thread fetches it with
// three 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP) that complete on an
// mbarrier; the matrix stream never passes through registers and is tagged
// L2::evict_first so it does not displace the gathered x-vector.  Rows are
// then reduced out of shared memory by groups of L lanes (L = 1..32, chosen
// from the average row length) with a shuffle tree; with L == 1 the summation
// order is exactly the reference's sequential `sum += a.value()*x[a.col()]`.
//
// Two schedulers over the same stage code:
//   variant 0: one row block per CTA, latency hidden by several CTAs per SM;
//   variant 1: persistent CTAs walking the block list with an S-deep ring of
//              stages, so S*CTAs/SM bulk copies are always in flight per SM.
//
// Column formats (template parameter FMT): how the kernel learns the column of an entry.  The
// arithmetic (entry order, lanes, shuffle tree, epilogue) is shared, so every format gives the
// bits of the plain one (tests/test_gpu_formats.py, tests/test_gpu_window.py); measurements in
// DESIGN.md section 3.1b.
//
// Windowed operators (FMT_WINDOW; opt-in: measured slower than the plain path).  For an operator
// whose row blocks gather from few distinct places the upload also stores, per block, the runs
// of x it reads and, per entry, the 16-bit position of its column inside the block's window.
// The kernel fills the window into shared memory with coalesced loads (one 32-byte sector is
// fetched once per block, not once per scattered 8-byte gather that misses the small L1 left
// beside the stages) and reduces the rows entirely out of shared memory; the column stream
// shrinks from 4 to 2 bytes per entry.
//
// Offset-indexed columns (FMT_OFFSET).  In a matrix assembled on a structured grid every entry
// sits on one of a few diagonals: col - row takes a handful of distinct values (7 for the
// Poisson stencil, a few more on a partitioned operator where halo columns are renumbered).  If
// there are at most 256 of them, the upload also stores one BYTE per entry -- the index of its
// offset in a table -- and the kernel streams 1 instead of 4 bytes of column per entry
// (9 instead of 12 bytes per FP64 entry, 5 instead of 8 per FP32 entry), rebuilding the column
// as row + table[index] from shared memory.
//
// Pattern-indexed rows (FMT_PATTERN).  On such a matrix whole ROWS repeat: the tuple of offsets
// (col - row of every entry, in entry order) of a row is one of a few patterns (27 for the
// Poisson problem: interior + boundary combinations).  With at most 256 patterns the upload
// stores one byte per ROW and no column information per entry at all: the kernel streams the
// values (8 or 4 bytes per entry), the row pointers and the pattern ids, and rebuilds
// col = row + pattern[k] for the k-th entry of the row from a table in shared memory -- one
// shared-memory load per entry, as many as the plain path needs for its staged column.  The
// default for every operator that qualifies (the finest level of a structured-grid problem).
//
// Between blocks the CTA synchronises with a barrier (plain, windowed) or not at all
// (offset- / pattern-indexed: the last warp done with a stage refills it) -- see the ring kernel.
//
// Precision.  Every kernel is a template over the element types of the matrix
// values, the gathered vector, the right-hand side, the output and the smoother
// diagonal (struct Prec).  FP64 throughout is the default; the other
// combinations are the ones AMGCL's mixed-precision composition produces (FP32
// hierarchy under an FP64 Krylov solver, tutorial/1.poisson3Db/poisson3Db.cpp:45-51):
// as in the reference (matrix_ops.hpp:57) a row sum is accumulated in the value
// type of the OUTPUT vector.
#pragma once
#include "common.cuh"
#include "reduce.cuh"
#include <type_traits>

namespace b200 {

// storage format of the column indices the kernel streams
enum { FMT_PLAIN = 0,     // int32 column per entry
       FMT_WINDOW = 1,    // 16-bit position in the block's shared-memory window of x
       FMT_OFFSET = 2,    // 8-bit index into a table of (col - row) offsets
       FMT_PATTERN = 3 }; // no per-entry column: 8-bit pattern id per row, col = row + pattern[k]

enum { MODE_SPMV = 0, MODE_SPMV_ACC = 1, MODE_RESID = 2, MODE_RELAX = 3,
       // y = f - A x with x = (alpha*d).*f formed on the fly and written to xw: the smoother's
       // first sweep from x = 0 (relax_zero_kernel) fused into the residual that follows it
       // (amg.hpp:527-534: apply_pre, then residual)
       MODE_RESID_SCALED = 4 };

#ifndef B200_GATHER_BATCH
#define B200_GATHER_BATCH 4
#endif
constexpr int kGatherBatch = B200_GATHER_BATCH;   // x-gathers issued back to back per lane

template <class TV_, class TX_, class TF_, class TY_, class TD_>
struct Prec {
    typedef TV_ TV;   // matrix values
    typedef TX_ TX;   // gathered vector x
    typedef TF_ TF;   // right-hand side f
    typedef TY_ TY;   // output y (and accumulation type of a row sum)
    typedef TD_ TD;   // smoother diagonal
};
typedef Prec<double, double, double, double, double> PrecDD;   // FP64 throughout
typedef Prec<float, float, float, float, float>      PrecFF;   // FP32 level of a mixed hierarchy
typedef Prec<float, double, double, double, float>   PrecFD;   // FP32 operator on FP64 vectors
typedef Prec<float, double, double, float, float>    PrecFDF;  // finest residual into FP32 scratch
typedef Prec<float, float, float, double, float>     PrecFFD;  // prolongation into an FP64 iterate

template <class P>
struct CsrArgsT {
    const int    *ptr;
    const int    *col;
    const typename P::TV *val;
    const int4   *blk;    // [nblocks] in WALK ORDER: {first row (~first row if the block gathers
                          //   halo columns), end row, first non-zero, end non-zero}
    int           nrows;
    int           nblocks;
    int           rows_cap;
    int           nnz_cap;
    const typename P::TX *x;      // gathered vector (local columns)
    const typename P::TX *xh;     // halo values for columns >= nloc (multi-GPU), else nullptr
    int           nloc;   // number of local columns when xh is set
    // multi-GPU, peer-memory transport: the halo is pushed by the peers while this kernel
    // already works on interior rows; a block that gathers remote columns first waits for
    // the flags of the ranks in wait_mask to reach wait_seq (csrc/peer.cuh protocol)
    // (such blocks come last in the walk order, so the transfer overlaps the interior rows)
    const unsigned long long *wait_flags;  // flag row of the current parity (16 slots)
    unsigned int              wait_mask;
    unsigned long long        wait_seq;
    // ... and THIS rank's boundary values are pushed by this very kernel: right after the
    // grid dependency resolves every CTA packs a slice of x[send_idx[.]] into the halo buffers
    // of the ranks that gather them (plain stores over NVLink), the last CTA to finish
    // releases their flags; then all CTAs go on to the interior row blocks
    const int                *send_idx;
    int                       n_send;
    int                       nranks;
    typename P::TX           *push_data[kMaxRanks];   // my segment in peer q's halo buffer (or nullptr)
    unsigned long long       *push_flag[kMaxRanks];   // my flag in peer q's flag row (or nullptr)
    unsigned int             *push_ticket;
    unsigned long long        push_seq;                // 0: nothing to push / NCCL transport
    // row shares of a replicated result (R onto a small level): every row is also stored into
    // every rank's gather buffer; the last CTA releases the flags at the end of the kernel
    int                       gather_on;
    typename P::TY           *gather_data[kMaxRanks]; // my share's place in rank q's buffer
    unsigned long long       *gather_flag[kMaxRanks];
    unsigned int             *gather_ticket;
    unsigned long long        gather_seq;
    // Gather-heavy operators (long rows): the 128-byte lines of x a row block gathers from, as a
    // per-block list (built at upload).  Before reducing a block the CTA touches each of them
    // once with coalesced loads -- one L2->L1 fill per LINE the block uses, instead of one
    // 32-byte sector fill per scattered 8-byte gather that misses (warm_lines below).
    const int    *wl_ptr; // [nblocks+1] (walk order) or nullptr
    const int    *wl;     // line numbers (x index / 16)
    // Windowed operators: window-local column of every entry, the runs of x each block's window
    // is made of ({first column, length | first slot << 16}), and each block's range of runs
    const unsigned short *col16;
    const int2   *wrun;
    const int2   *wblk;   // [nblocks] in walk order
    int           run_cap;// most runs a block has (stage layout); 0: not a windowed launch
    // Offset-indexed operators: index of every entry's (col - row) in off_tab[256]
    const unsigned char *idx8;
    const int    *off_tab;
    // Pattern-indexed operators: pattern id of every row, first table entry of every pattern
    // ([257]), and the table of offsets itself
    const unsigned char  *pid;
    const unsigned short *pat_start;
    const int    *pat_off;
    int           pat_total;   // entries of pat_off in use (<= kPatOffCap)
    typename P::TY       *y;      // output
    typename P::TX       *xw;     // RESID_SCALED: where x = (alpha*d).*f is written
    const typename P::TF *f;      // rhs          (RESID, RELAX)
    const typename P::TD *d;      // diagonal     (RELAX)
    double        alpha;  // SPMV: alpha; RELAX: omega
    double        beta;   // SPMV_ACC
    // scalars produced while the rows are in registers (reduce.cuh; FP64 outputs only):
    //   ndot >= 1: sum_r y_r * w_r  (w == nullptr: RELAX uses the rhs f, otherwise y itself)
    //   ndot == 2: additionally sum_r y_r^2
    int           ndot;
    const double *w;
    RedOut        red;
};
typedef CsrArgsT<PrecDD> CsrArgs;

// ---- shared memory layout of one stage --------------------------------------
struct StageLayout {
    int val_off, col_off, ptr_off, run_off, pid_off, bytes;
};
// fmt: storage format of the columns (FMT_*); run_cap: FMT_WINDOW only, most runs a block has
__host__ __device__ inline StageLayout stage_layout(int rows_cap, int nnz_cap, int val_size,
                                                    int fmt = FMT_PLAIN, int run_cap = 0) {
    StageLayout s;
    s.val_off = 0;
    int val_bytes = nnz_cap * val_size + 16;       // source aligned down to 16 B
    val_bytes = (val_bytes + 15) & ~15;
    s.col_off = s.val_off + val_bytes;
    int col_bytes = fmt == FMT_WINDOW  ? (nnz_cap + 16) * 2   // +7 align down, +7 round up
                  : fmt == FMT_OFFSET  ? (nnz_cap + 32)       // +15 align down, +15 round up
                  : fmt == FMT_PATTERN ? 0
                                       : (nnz_cap + 8) * 4;   // +3 align down, +3 round up
    col_bytes = (col_bytes + 15) & ~15;
    s.ptr_off = s.col_off + col_bytes;
    int ptr_bytes = (rows_cap + 4) * 4;
    ptr_bytes = (ptr_bytes + 15) & ~15;
    s.run_off = s.ptr_off + ptr_bytes;
    int run_bytes = fmt == FMT_WINDOW ? (run_cap + 2) * 8 : 0;   // +1 align down, +1 round up
    run_bytes = (run_bytes + 15) & ~15;
    s.pid_off = s.run_off + run_bytes;
    int pid_bytes = fmt == FMT_PATTERN ? rows_cap + 32 : 0;      // +15 align down, +15 round up
    pid_bytes = (pid_bytes + 15) & ~15;
    s.bytes = s.pid_off + pid_bytes;
    return s;
}
constexpr int kMaxStages   = 8;
constexpr int kHeaderBytes = 448;   // mbarriers [0,64) + reduction scratch [64,128) + descriptors [128,384)
                                    // + per-stage "warps done" counters [384,416)
constexpr int kWinRunLen   = 64;    // longest run of a window (longer ones are cut at upload)
constexpr int kOffTabLen   = 256;   // offset-indexed operators: entries of the (col - row) table
constexpr int kPatCap      = 256;   // pattern-indexed operators: most row patterns ...
constexpr int kPatOffCap   = 1024;  // ... and most offsets in all patterns together
// shared memory behind the stages: the window of x / the offset table / the pattern tables
constexpr int kPatTabBytes = kPatOffCap * 4 + ((kPatCap + 1) * 2 + 15) / 16 * 16;

struct BlockDesc {      // written by the producer thread, read by everyone after the wait
    int r0, r1;         // row range
    int e0, e1;         // non-zero range
    int halo;           // the block gathers columns owned by other ranks
    int q0, q1;         // windowed operators: the block's runs
    int pad_;
};
static_assert(sizeof(BlockDesc) * kMaxStages <= 384 - 128, "descriptors overflow the header");

// ---- issue the bulk copies of one row block ----------------------------------
template <int FMT = FMT_PLAIN, class P>
__device__ __forceinline__ BlockDesc load_desc(const CsrArgsT<P> &a, int b) {
    const int4 q = __ldg(a.blk + b);       // one 16-byte load: nothing else to chase
    BlockDesc d;
    d.halo = q.x < 0;
    d.r0 = q.x < 0 ? ~q.x : q.x;
    d.r1 = q.y; d.e0 = q.z; d.e1 = q.w;
    d.q0 = d.q1 = 0; d.pad_ = 0;
    if (FMT == FMT_WINDOW) {
        const int2 w = __ldg(a.wblk + b);
        d.q0 = w.x; d.q1 = w.y;
    }
    return d;
}

// Returns true if the block was staged (false: too long, use the strided path).
template <int FMT = FMT_PLAIN, class P>
__device__ __forceinline__ bool issue_block(const CsrArgsT<P> &a, const BlockDesc &d, char *stage,
                                            const StageLayout &lay, uint64_t *bar,
                                            uint64_t policy) {
    typedef typename P::TV TV;
    constexpr int VA = 16 / (int)sizeof(TV);        // values per 16 bytes
    const int nnz = d.e1 - d.e0;
    if (FMT == FMT_PATTERN) {
        // (a pattern-indexed operator has no long blocks)
        const int a0 = d.e0 & ~(VA - 1);
        const int nval = ((d.e1 - a0) + VA - 1) & ~(VA - 1);
        const int nptr = ((d.r1 - d.r0 + 1) + 3) & ~3;
        const int p0 = d.r0 & ~15;                  // pattern ids: 16 per 16 bytes
        const int npid = ((d.r1 - p0) + 15) & ~15;
        const uint32_t bytes = nval * (int)sizeof(TV) + nptr * 4 + npid;
        ptx::mbar_expect_tx(bar, bytes);
        if (nval) ptx::bulk_g2s(stage + lay.val_off, a.val + a0, nval * (int)sizeof(TV), bar, policy);
        ptx::bulk_g2s(stage + lay.ptr_off, a.ptr + d.r0, nptr * 4, bar, policy);
        ptx::bulk_g2s(stage + lay.pid_off, a.pid + p0, npid, bar, policy);
        return true;
    }
    if (FMT == FMT_OFFSET) {
        // (an offset-indexed operator has no long blocks)
        const int a0 = d.e0 & ~(VA - 1);
        const int nval = ((d.e1 - a0) + VA - 1) & ~(VA - 1);
        const int c0 = d.e0 & ~15;                  // 8-bit columns: 16 per 16 bytes
        const int ncol = ((d.e1 - c0) + 15) & ~15;
        const int nptr = ((d.r1 - d.r0 + 1) + 3) & ~3;
        const uint32_t bytes = nval * (int)sizeof(TV) + ncol + nptr * 4;
        ptx::mbar_expect_tx(bar, bytes);
        if (nval) ptx::bulk_g2s(stage + lay.val_off, a.val + a0, nval * (int)sizeof(TV), bar, policy);
        if (ncol) ptx::bulk_g2s(stage + lay.col_off, a.idx8 + c0, ncol, bar, policy);
        ptx::bulk_g2s(stage + lay.ptr_off, a.ptr + d.r0, nptr * 4, bar, policy);
        return true;
    }
    if (FMT == FMT_WINDOW) {
        // (a windowed operator has no long blocks)
        const int a0 = d.e0 & ~(VA - 1);
        const int nval = ((d.e1 - a0) + VA - 1) & ~(VA - 1);
        const int c0 = d.e0 & ~7;                   // 16-bit columns: 8 per 16 bytes
        const int ncol = ((d.e1 - c0) + 7) & ~7;
        const int nptr = ((d.r1 - d.r0 + 1) + 3) & ~3;
        const int qa = d.q0 & ~1;                   // runs: 2 per 16 bytes
        const int nrun = ((d.q1 - qa) + 1) & ~1;
        const uint32_t bytes = nval * (int)sizeof(TV) + ncol * 2 + nptr * 4 + nrun * 8;
        ptx::mbar_expect_tx(bar, bytes);
        if (nval) ptx::bulk_g2s(stage + lay.val_off, a.val + a0, nval * (int)sizeof(TV), bar, policy);
        if (ncol) ptx::bulk_g2s(stage + lay.col_off, a.col16 + c0, ncol * 2, bar, policy);
        ptx::bulk_g2s(stage + lay.ptr_off, a.ptr + d.r0, nptr * 4, bar, policy);
        if (nrun) ptx::bulk_g2s(stage + lay.run_off, a.wrun + qa, nrun * 8, bar, policy);
        return true;
    }
    if (nnz > a.nnz_cap) {
        // nothing to stage: complete the phase with a plain arrive
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_addr(bar))
                     : "memory");
        return false;
    }
    const int a0 = d.e0 & ~(VA - 1);                // val source aligned to 16 B
    const int nval = ((d.e1 - a0) + VA - 1) & ~(VA - 1);
    const int c0 = d.e0 & ~3;                       // col source aligned to 16 B
    const int ncol = ((d.e1 - c0) + 3) & ~3;
    const int nptr = ((d.r1 - d.r0 + 1) + 3) & ~3;  // r0 is a multiple of 4
    const uint32_t bytes = nval * (int)sizeof(TV) + ncol * 4 + nptr * 4;
    ptx::mbar_expect_tx(bar, bytes);
    if (nval) ptx::bulk_g2s(stage + lay.val_off, a.val + a0, nval * (int)sizeof(TV), bar, policy);
    if (ncol) ptx::bulk_g2s(stage + lay.col_off, a.col + c0, ncol * 4, bar, policy);
    ptx::bulk_g2s(stage + lay.ptr_off, a.ptr + d.r0, nptr * 4, bar, policy);
    return true;
}

// ---- per-row epilogue ----------------------------------------------------------
// acc: the thread's running contributions to the launch's scalars (reduce.cuh)
struct RowAcc { double s0, s1; };

// what the epilogue of a row reads besides the row sum; loaded BEFORE the row's gathers are
// issued so that it arrives with them (one memory round per row instead of two)
template <class P>
struct RowOps {
    typename P::TF f;     // rhs           (RESID, RESID_SCALED, RELAX)
    typename P::TD d;     // diagonal      (RESID_SCALED, RELAX)
    typename P::TX x;     // old iterate   (RELAX)
    typename P::TY y;     // old output    (SPMV_ACC)
    double         w;     // dot weight    (ndot with w)
};
template <int MODE, class P>
__device__ __forceinline__ RowOps<P> load_row_ops(const CsrArgsT<P> &a, int r) {
    RowOps<P> o;
    o.f = 0; o.d = 0; o.x = 0; o.y = 0; o.w = 0.0;
    if (MODE == MODE_SPMV_ACC) o.y = a.y[r];
    if (MODE == MODE_RESID || MODE == MODE_RESID_SCALED || MODE == MODE_RELAX) o.f = a.f[r];
    if (MODE == MODE_RESID_SCALED || MODE == MODE_RELAX) o.d = a.d[r];
    if (MODE == MODE_RELAX) o.x = a.x[r];
    if (a.ndot && a.w) o.w = a.w[r];
    return o;
}

template <int MODE, class P>
__device__ __forceinline__ void store_row(const CsrArgsT<P> &a, int r, typename P::TY sum, RowAcc &acc,
                                          const RowOps<P> &o) {
    typedef typename P::TY TY;
    TY out;
    double wv = 0.0;
    if (MODE == MODE_SPMV) {
        out = (TY)(a.alpha * sum);
    } else if (MODE == MODE_SPMV_ACC) {
        out = (TY)(a.alpha * sum + a.beta * o.y);
    } else if (MODE == MODE_RESID) {
        out = (TY)(o.f - sum);
    } else if (MODE == MODE_RESID_SCALED) {
        const typename P::TF fr = o.f;
        out = (TY)(fr - sum);
        a.xw[r] = fma((typename P::TX)(a.alpha * o.d), (typename P::TX)fr, (typename P::TX)0);
    } else {
        // x_new = (omega*d)*t + x with t = f - A x; same association as the
        // reference's vmul  z = a*x*y + b*z  (builtin.hpp:1238-1265)
        const typename P::TF fr = o.f;
        const TY t = (TY)(fr - sum);
        const TY w = (TY)(a.alpha * o.d);
        out = fma(w, t, (TY)o.x);
        wv = (double)fr;
    }
    a.y[r] = out;
    if (a.gather_on) {
#pragma unroll 1
        for (int q = 0; q < a.nranks; ++q)
            if (a.gather_data[q]) a.gather_data[q][r] = out;
    }
    if (a.ndot) {
        const double yv = (double)out;
        if (a.w) wv = o.w;
        else if (MODE != MODE_RELAX) wv = yv;
        acc.s0 = fma(yv, wv, acc.s0);
        if (a.ndot > 1) acc.s1 = fma(yv, yv, acc.s1);
    }
}
// (operands loaded on the spot: the paths that do not overlap them with the gathers)
template <int MODE, class P>
__device__ __forceinline__ void store_row(const CsrArgsT<P> &a, int r, typename P::TY sum, RowAcc &acc) {
    store_row<MODE>(a, r, sum, acc, load_row_ops<MODE>(a, r));
}

// ---- x[col]: local columns from the vector, remote ones from the all-gathered halo --
template <bool HALO, class P>
__device__ __forceinline__ typename P::TX gather(const CsrArgsT<P> &a,
                                                 const typename P::TX *__restrict__ x, int c) {
    if (HALO) {
        // halo values may arrive while the kernel runs: read them through L2 (ld.cg)
        if (c < a.nloc) return __ldg(x + c);
        return __ldcg(a.xh + (c - a.nloc));
    }
    return __ldg(x + c);
}

// RESID_SCALED: the gathered vector does not exist yet; its entry c is (alpha*d[c])*f[c],
// evaluated exactly as relax_zero_kernel does
template <int MODE, bool HALO, class P>
__device__ __forceinline__ typename P::TX gather_m(const CsrArgsT<P> &a,
                                                   const typename P::TX *__restrict__ x, int c) {
    typedef typename P::TX TX;
    if (MODE == MODE_RESID_SCALED)
        return fma((TX)(a.alpha * __ldg(a.d + c)), (TX)__ldg(a.f + c), (TX)0);
    return gather<HALO>(a, x, c);
}

// Block until the peers' halo pushes for this exchange have landed (thread 0 polls the
// flags with acquire semantics, the CTA follows through the barrier).
template <bool HALO, class P>
__device__ __forceinline__ void wait_for_halo(const CsrArgsT<P> &a, const BlockDesc &d) {
    if (!HALO) return;
    if (!d.halo || !a.wait_mask) return;                       // uniform per CTA
    if (threadIdx.x == 0) {
        unsigned int m = a.wait_mask;
        while (m) {
            const int q = __ffs(m) - 1;
            m &= m - 1;
            unsigned long long v;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.wait_flags + q) : "memory");
            } while (v < a.wait_seq);
        }
    }
    __syncthreads();
}

// ---- bring the lines of x a block gathers from into L1 with coalesced loads --------------
// Four threads per 128-byte line, one 8-byte load per 32-byte sector; the values are not used.
// The gathers that follow hit (or merge with the outstanding fills).
template <class P>
__device__ __forceinline__ void warm_lines(const CsrArgsT<P> &a, int pos) {
    if (a.wl_ptr == nullptr) return;
    const int w0 = __ldg(a.wl_ptr + pos), w1 = __ldg(a.wl_ptr + pos + 1);
    const char *base = reinterpret_cast<const char *>(a.x);
    for (int t = w0 * 4 + (int)threadIdx.x; t < w1 * 4; t += kThreads) {
        const int line = __ldg(a.wl + (t >> 2));
        const char *p = base + (size_t)line * 128 + (size_t)(t & 3) * 32;
        unsigned long long sink;
        asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(sink) : "l"(p) : "memory");
    }
}

// ---- windowed operators: bring the block's runs of x into shared memory --------------------
// One warp per run (runs are at most kWinRunLen long and start on a 32-byte boundary of x, so a
// warp's loads are whole sectors); ends with a CTA barrier.
template <int MODE, bool HALO, class P>
__device__ __forceinline__ void fill_window(const CsrArgsT<P> &a, const BlockDesc &d, const char *stage,
                                            const StageLayout &lay, typename P::TX *win) {
    const int2 *runs = reinterpret_cast<const int2 *>(stage + lay.run_off) + (d.q0 & 1);
    const int nq   = d.q1 - d.q0;
    const int lane = threadIdx.x & 31;
    const typename P::TX *__restrict__ x = a.x;
    for (int q = threadIdx.x >> 5; q < nq; q += kThreads / 32) {
        const int2 rn   = runs[q];
        const int first = rn.x;
        const int len   = rn.y & 0xffff;
        const int slot  = (int)((unsigned)rn.y >> 16);
#pragma unroll
        for (int i = lane; i < kWinRunLen; i += 32)
            if (i < len) win[slot + i] = gather_m<MODE, HALO>(a, x, first + i);
    }
    __syncthreads();
}

// ---- reduce the rows of a staged block out of shared memory ---------------------
// FMT_WINDOW: x comes from the block's window, `win`, indexed by the 16-bit columns;
// FMT_OFFSET: the column of an entry of row r is r + off[8-bit index];
// FMT_PATTERN: the column of the k-th entry of row r is r + off[pstart[pattern id of r] + k]
template <int MODE, int L, bool HALO, class P, int FMT = FMT_PLAIN>
__device__ __forceinline__ void compute_staged(const CsrArgsT<P> &a, const BlockDesc &d,
                                               const char *stage, const StageLayout &lay, RowAcc &acc,
                                               const typename P::TX *win = nullptr, const int *off = nullptr,
                                               const unsigned short *pstart = nullptr) {
    constexpr bool WIN = FMT == FMT_WINDOW;
    constexpr bool OFF = FMT == FMT_OFFSET;
    constexpr bool PAT = FMT == FMT_PATTERN;
    typedef typename P::TV TV;
    typedef typename P::TX TX;
    typedef typename P::TY TS;                       // row sums live in the output's type
    typedef typename std::conditional<WIN, unsigned short,
                                      typename std::conditional<OFF, unsigned char, int>::type>::type CI;
    static_assert(!((WIN || OFF || PAT) && L >= 16), "compressed column formats use at most 8 lanes per row");
    const unsigned char *pid_s = reinterpret_cast<const unsigned char *>(stage + lay.pid_off) + (d.r0 & 15);
    constexpr int VA = 16 / (int)sizeof(TV);
    const TV *val_s = reinterpret_cast<const TV *>(stage + lay.val_off);
    const CI     *col_s = reinterpret_cast<const CI *>(stage + lay.col_off);
    const int    *ptr_s = reinterpret_cast<const int *>(stage + lay.ptr_off);
    const int vo = d.e0 & ~(VA - 1);
    const int co = WIN ? (d.e0 & ~7) : OFF ? (d.e0 & ~15) : (d.e0 & ~3);
    constexpr int G = kThreads / L;
    const int g    = threadIdx.x / L;
    const int lane = threadIdx.x % L;
    const int nr   = d.r1 - d.r0;
    const TX *__restrict__ x = a.x;

    if (L >= 16) {
        // Wide groups (a half or a full warp per row): consecutive lanes read consecutive
        // entries, so the shared-memory reads are bank-conflict free and the gathers of one
        // row coalesce.  Memory-level parallelism comes from working on RU rows at once
        // instead of batching along one (short) row.
        constexpr int RU = kGatherBatch;
        for (int base = 0; base < nr; base += G * RU) {
            int  beg[RU], end[RU], c[RU];
            TV   v[RU];
            TX   xv[RU];
            TS   sum[RU];
            bool rowok[RU], p[RU];
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const int rr = base + u * G + g;
                rowok[u] = rr < nr;
                beg[u] = rowok[u] ? ptr_s[rr] : 0;
                end[u] = rowok[u] ? ptr_s[rr + 1] : 0;
                sum[u] = 0;
            }
            // first L entries of each of the RU rows: all loads first, then the gathers
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const int e = beg[u] + lane;
                p[u] = e < end[u];
                c[u] = p[u] ? col_s[e - co] : 0;
                v[u] = p[u] ? val_s[e - vo] : (TV)0;
            }
#pragma unroll
            for (int u = 0; u < RU; ++u) xv[u] = p[u] ? gather_m<MODE, HALO>(a, x, c[u]) : (TX)0;
#pragma unroll
            for (int u = 0; u < RU; ++u)
                if (p[u]) sum[u] = (TS)v[u] * (TS)xv[u];
            // rows longer than L: the rest, row by row
#pragma unroll
            for (int u = 0; u < RU; ++u)
                for (int e = beg[u] + lane + L; e < end[u]; e += L)
                    sum[u] = fma((TS)val_s[e - vo], (TS)gather_m<MODE, HALO>(a, x, col_s[e - co]), sum[u]);
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) {
#pragma unroll
                for (int u = 0; u < RU; ++u) sum[u] += __shfl_xor_sync(0xffffffffu, sum[u], o);
            }
#pragma unroll
            for (int u = 0; u < RU; ++u)
                if (rowok[u] && lane == 0) store_row<MODE>(a, d.r0 + base + u * G + g, sum[u], acc);
        }
    } else
    for (int base = 0; base < nr; base += G) {
        const int  rr    = base + g;
        const bool valid = rr < nr;
        TS sum = 0;
        RowOps<P> ops;
        ops.f = 0; ops.d = 0; ops.x = 0; ops.y = 0; ops.w = 0.0;
        if (valid) {
            // the epilogue's operands travel with the row's gathers
            if (lane == 0) ops = load_row_ops<MODE>(a, d.r0 + rr);
            const int beg = ptr_s[rr];
            const int end = ptr_s[rr + 1];
            // PAT: the row's pattern starts at off[pb + beg], so entry e sits at off[pb + e]
            int pb = 0;
            if (PAT) pb = (int)pstart[pid_s[rr]] - beg;
            // U independent gathers in flight per lane, then the FMAs in entry order (one lane per
            // row: short rows, a whole row of up to 8 entries goes out in one round)
            constexpr int U = (L == 1) ? 2 * kGatherBatch : kGatherBatch;
            for (int e = beg + lane; e < end; e += U * L) {
                int  c[U];
                TX   xv[U];
                bool p[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int eu = e + u * L;
                    p[u] = eu < end;
                    if (PAT) c[u] = d.r0 + rr + off[pb + (p[u] ? eu : e)];
                    else c[u] = p[u] ? col_s[eu - co] : col_s[e - co];
                }
                if (OFF) {
#pragma unroll
                    for (int u = 0; u < U; ++u) c[u] = d.r0 + rr + off[c[u]];
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    xv[u] = !p[u] ? (TX)0 : WIN ? win[c[u]] : gather_m<MODE, HALO>(a, x, c[u]);
                // (the values come from shared memory when they are needed: no registers held
                //  across the gathers)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (p[u]) sum = fma((TS)val_s[e + u * L - vo], (TS)xv[u], sum);
            }
        }
        if (L > 1) {
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        }
        if (valid && lane == 0) store_row<MODE>(a, d.r0 + rr, sum, acc, ops);
    }
}

// ---- rows too long to stage: whole CTA strides over each row -----------------------
template <int MODE, bool HALO, class P>
__device__ __forceinline__ void compute_long(const CsrArgsT<P> &a, const BlockDesc &d,
                                             double *red_s /* >= 8 doubles */, RowAcc &acc) {
    typedef typename P::TX TX;
    typedef typename P::TY TS;
    const TX *__restrict__ x = a.x;
    for (int r = d.r0; r < d.r1; ++r) {
        const int beg = __ldg(a.ptr + r), end = __ldg(a.ptr + r + 1);
        TS sum = 0;
        for (int e = beg + threadIdx.x; e < end; e += kThreads)
            sum = fma((TS)__ldg(a.val + e), (TS)gather_m<MODE, HALO>(a, x, __ldg(a.col + e)), sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __syncthreads();                       // red_s free from the previous row
        if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = (double)sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            TS tot = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) tot += (TS)red_s[w];
            store_row<MODE>(a, r, tot, acc);
        }
    }
}

// ---- multi-GPU: this rank's side of the exchanges, done by the consumer kernel itself --------
__device__ __forceinline__ void xchg_st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// all CTAs: pack + push a slice of the boundary values; last CTA: release the consumers' flags
template <class P>
__device__ __forceinline__ void halo_push(const CsrArgsT<P> &a) {
    if (!a.push_seq) return;
    const typename P::TX *__restrict__ x = a.x;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < a.n_send; i += gridDim.x * kThreads) {
        const typename P::TX v = x[a.send_idx[i]];
#pragma unroll 1
        for (int q = 0; q < a.nranks; ++q)
            if (a.push_data[q]) a.push_data[q][i] = v;
    }
    __threadfence_system();                   // my peer stores are visible system-wide ...
    __shared__ bool push_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(a.push_ticket, 1u);    // ... before the ticket moves
        push_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (push_last) {
        // (st.release.sys orders everything that happens-before it -- all CTAs' stores, via
        // their fences and the ticket -- before the flag)
        if (threadIdx.x < a.nranks && a.push_flag[threadIdx.x])
            xchg_st_release_sys(a.push_flag[threadIdx.x], a.push_seq);
        if (threadIdx.x == 0) *a.push_ticket = 0;
    }
}
// end of a kernel that stored row shares into the peers' gather buffers
template <class P>
__device__ __forceinline__ void gather_finish(const CsrArgsT<P> &a) {
    __threadfence_system();
    __shared__ bool gather_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(a.gather_ticket, 1u);
        gather_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (gather_last) {
        if (threadIdx.x < a.nranks && a.gather_flag[threadIdx.x])
            xchg_st_release_sys(a.gather_flag[threadIdx.x], a.gather_seq);
        if (threadIdx.x == 0) *a.gather_ticket = 0;
    }
}

// ---- variant 0: one row block per CTA -----------------------------------------------
template <int MODE, int L, bool HALO, class P>
__global__ void __launch_bounds__(kThreads, 4) csr_block_kernel(const CsrArgsT<P> a) {
    extern __shared__ __align__(128) char smem[];
    uint64_t *bar   = reinterpret_cast<uint64_t *>(smem);
    double   *red_s = reinterpret_cast<double *>(smem + 64);
    char     *stage = smem + kHeaderBytes;
    const StageLayout lay = stage_layout(a.rows_cap, a.nnz_cap, (int)sizeof(typename P::TV));

    const int b = blockIdx.x;
    const BlockDesc d = load_desc(a, b);          // broadcast loads, uniform
    const bool staged = (d.e1 - d.e0) <= a.nnz_cap;
    RowAcc acc = {0.0, 0.0};                      // (this variant produces no scalars)

    if (staged) {
        if (threadIdx.x == 0) {
            ptx::mbar_init(bar, 1);
            ptx::fence_mbar_init();
            issue_block(a, d, stage, lay, bar, ptx::policy_evict_first());
        }
        __syncthreads();
        ptx::mbar_wait(bar, 0);
        wait_for_halo<HALO>(a, d);
        compute_staged<MODE, L, HALO>(a, d, stage, lay, acc);
    } else {
        wait_for_halo<HALO>(a, d);
        compute_long<MODE, HALO>(a, d, red_s, acc);
    }
}

// ---- variant 1: persistent CTAs, S-deep ring of stages ----------------------------------
// FMT: storage format of the columns (FMT_PLAIN / FMT_WINDOW / FMT_OFFSET, see the top of the file)
template <int MODE, int L, bool HALO, class P, int FMT = FMT_PLAIN>
__global__ void __launch_bounds__(kThreads, 4) csr_ring_kernel(const CsrArgsT<P> a, const int nstages) {
    extern __shared__ __align__(128) char smem[];
    uint64_t  *bars  = reinterpret_cast<uint64_t *>(smem);                 // [<=8]
    double    *red_s = reinterpret_cast<double *>(smem + 64);              // [8]
    BlockDesc *descs = reinterpret_cast<BlockDesc *>(smem + 128);          // [<=8]
    unsigned  *done  = reinterpret_cast<unsigned *>(smem + 384);           // [<=8] warps done with a stage
    const StageLayout lay = stage_layout(a.rows_cap, a.nnz_cap, (int)sizeof(typename P::TV), FMT, a.run_cap);
    char *stages = smem + kHeaderBytes;
    // behind the stages: the window of x (FMT_WINDOW), the offset table (FMT_OFFSET), or the
    // patterns' offsets followed by the patterns' first entries (FMT_PATTERN)
    typename P::TX *win = reinterpret_cast<typename P::TX *>(stages + (size_t)nstages * lay.bytes);
    int *off_s = reinterpret_cast<int *>(stages + (size_t)nstages * lay.bytes);
    unsigned short *pstart_s = reinterpret_cast<unsigned short *>(off_s + kPatOffCap);

    const int first = blockIdx.x;
    const int step  = gridDim.x;
    const int mine  = (a.nblocks - first + step - 1) / step;   // blocks this CTA owns
    const uint64_t policy = ptx::policy_evict_first();     // (any warp's first lane may issue a refill)

    // PDL: the prologue below only reads matrix data (never written by a kernel), so it may
    // run before the predecessor's writes are visible
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) { ptx::mbar_init(bars + s, 1); done[s] = 0; }
        ptx::fence_mbar_init();
        const int pre = mine < nstages ? mine : nstages;
        for (int i = 0; i < pre; ++i) {
            const BlockDesc d = load_desc<FMT>(a, first + i * step);
            descs[i] = d;
            issue_block<FMT>(a, d, stages + (size_t)i * lay.bytes, lay, bars + i, policy);
        }
    }
    if constexpr (FMT == FMT_OFFSET) {
        static_assert(kOffTabLen == kThreads, "one table entry per thread");
        off_s[threadIdx.x] = __ldg(a.off_tab + threadIdx.x);
    }
    if constexpr (FMT == FMT_PATTERN) {
        for (int i = threadIdx.x; i < a.pat_total; i += kThreads) off_s[i] = __ldg(a.pat_off + i);
        for (int i = threadIdx.x; i <= kPatCap; i += kThreads) pstart_s[i] = __ldg(a.pat_start + i);
    }
    __syncthreads();
    ptx::pdl_wait();         // vectors (x, f, d, y) come from earlier kernels: from here on
    if (HALO) halo_push(a);  // multi-GPU: my boundary values go out before anything else

    RowAcc acc = {0.0, 0.0};
    int s = 0, parity = 0;
    for (int i = 0; i < mine; ++i) {
        ptx::mbar_wait(bars + s, parity);
        const BlockDesc d = descs[s];
        wait_for_halo<HALO>(a, d);
        const char *stage = stages + (size_t)s * lay.bytes;
        // Compressed formats (the short-row finest operator): no CTA barrier between blocks -- a
        // warp that is done with this block moves on to the next stage at once, and the LAST
        // warp to finish refills the stage.  Measured on the 256^3 solve: -5 % on those passes;
        // on the plain-format operators (long rows, half of the warps without rows in a block)
        // the same scheme costs 10-15 %, so they stay block-synchronous.
        constexpr bool kDecoupled = FMT == FMT_OFFSET || FMT == FMT_PATTERN;
        if constexpr (FMT == FMT_WINDOW) {
            fill_window<MODE, HALO>(a, d, stage, lay, win);
            compute_staged<MODE, (L < 16 ? L : 8), HALO, P, FMT_WINDOW>(a, d, stage, lay, acc, win);
        } else if constexpr (FMT == FMT_OFFSET) {
            compute_staged<MODE, (L < 16 ? L : 8), HALO, P, FMT_OFFSET>(a, d, stage, lay, acc, nullptr, off_s);
        } else if constexpr (FMT == FMT_PATTERN) {
            compute_staged<MODE, (L < 16 ? L : 8), HALO, P, FMT_PATTERN>(a, d, stage, lay, acc, nullptr, off_s,
                                                                            pstart_s);
        } else {
            if (!HALO) warm_lines(a, first + i * step);
            if ((d.e1 - d.e0) <= a.nnz_cap)
                compute_staged<MODE, L, HALO>(a, d, stage, lay, acc);
            else
                compute_long<MODE, HALO>(a, d, red_s, acc);
        }
        if constexpr (kDecoupled) {
            // its arrive.expect_tx (release) / the others' wait (acquire) on the stage's mbarrier
            // publish the new descriptor; the counter orders everyone's reads of the stage before
            // the refill
            __syncwarp();
            if ((threadIdx.x & 31) == 0) {
                __threadfence_block();
                const unsigned arrived = atomicAdd(done + s, 1u);
                if (arrived == kThreads / 32 - 1) {
                    done[s] = 0;
                    __threadfence_block();
                    if (i + nstages < mine) {
                        const BlockDesc n = load_desc<FMT>(a, first + (i + nstages) * step);
                        descs[s] = n;
                        issue_block<FMT>(a, n, stages + (size_t)s * lay.bytes, lay, bars + s, policy);
                    }
                }
            }
        } else {
            __syncthreads();             // every thread is done with stage s, descs[s], the window
            if (threadIdx.x == 0 && i + nstages < mine) {
                const BlockDesc n = load_desc<FMT>(a, first + (i + nstages) * step);
                descs[s] = n;
                issue_block<FMT>(a, n, stages + (size_t)s * lay.bytes, lay, bars + s, policy);
            }
        }
        if (++s == nstages) { s = 0; parity ^= 1; }
    }
    if (HALO && a.gather_on) gather_finish(a);
    if (a.ndot) {
        double v[2] = {acc.s0, acc.s1};
        red_finish<2>(a.red, v);
    }
}

// ---- x == 0 shortcut of the smoother sweep: x = (omega*d).*rhs ---------------------------
template <class TD, class TF, class TX>
__global__ void __launch_bounds__(kThreads) relax_zero_kernel(size_t n, double omega,
                                                              const TD *__restrict__ d,
                                                              const TF *__restrict__ f,
                                                              TX *__restrict__ x) {
    ptx::pdl_wait();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // (omega*d)*(f - 0) + 0, written as the reference evaluates it
        x[i] = fma((TX)(omega * d[i]), (TX)f[i], (TX)0);
    }
}

} // namespace b200
