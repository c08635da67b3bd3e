// csr_launch.cuh -- launching the persistent ring kernel (csr_kernels.cuh).  Shared by
// api_matrices.cu (plain operators), api_window.cu (windowed operators), api_offsets.cu
// (offset-indexed operators) and api_patterns.cu (pattern-indexed operators): the kernel instantiations of each storage format are compiled in
// a translation unit of their own so the build stays parallel.
#pragma once
#include "internal.cuh"
#include "csr_kernels.cuh"

namespace b200 {

// dynamic shared memory a CTA may use when `ctas` of them share an SM: 228 KB per SM, 1 KB
// reserved per CTA, and the kernels' static scratch (reduction + exchange tickets, < 1.5 KB)
inline int ring_budget(int ctas) { return 228 * 1024 / ctas - 1024 - 1536; }

// shared memory of one launch: header + ring of stages (+ the window / the offset table); fewer
// stages if the configured ring does not fit beside opt_ctas_per_sm CTAs.  Returns 0 if a
// windowed launch does not fit even with one stage.
template <class P>
inline int ring_smem(b200_ctx_t ctx, b200_csr_t A, int fmt, int *stages_out) {
    const StageLayout lay = stage_layout(A->rows_cap, A->nnz_cap, (int)sizeof(typename P::TV), fmt, A->win_runs);
    const int extra = fmt == FMT_WINDOW ? (int)(((size_t)A->win_slots * sizeof(typename P::TX) + 15) & ~(size_t)15)
                      : fmt == FMT_OFFSET ? kOffTabLen * (int)sizeof(int)
                      : fmt == FMT_PATTERN ? kPatTabBytes : 0;
    int stages = (int)ctx->opt_stages;
    const int per_cta_budget = ring_budget((int)ctx->opt_ctas_per_sm);
    while (stages > 1 && kHeaderBytes + stages * lay.bytes + extra > per_cta_budget) --stages;
    *stages_out = stages;
    const int smem = kHeaderBytes + stages * lay.bytes + extra;
    return (fmt == FMT_WINDOW && smem > per_cta_budget) ? 0 : smem;
}

template <int MODE, int L, bool HALO, class P, int FMT>
inline int launch_ring_impl(b200_ctx_t ctx, b200_csr_t A, const CsrArgsT<P> &args) {
    int stages = 1;
    const int smem = ring_smem<P>(ctx, A, FMT, &stages);
    static bool attr_set[64] = {};
    if (!attr_set[ctx->device & 63]) {
        // the opt-in limit covers static + dynamic shared memory (red_finish keeps a few
        // hundred bytes of static scratch)
        cudaFuncAttributes fa;
        B200_CUDA(cudaFuncGetAttributes(&fa, csr_ring_kernel<MODE, L, HALO, P, FMT>));
        B200_CUDA(cudaFuncSetAttribute(csr_ring_kernel<MODE, L, HALO, P, FMT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024 - (int)fa.sharedSizeBytes));
        attr_set[ctx->device & 63] = true;
    }
    const int64_t cap = (int64_t)ctx->sm_count * ctx->opt_ctas_per_sm;
    const unsigned grid = (unsigned)std::min<int64_t>(A->nblocks, cap);
    B200_CUDA(launch_pdl(ctx, csr_ring_kernel<MODE, L, HALO, P, FMT>, dim3(grid), dim3(kThreads), (size_t)smem,
                         args, stages));
    return B200_OK;
}

// windowed operators (defined and instantiated in api_window.cu: 1..8 lanes per row)
template <int MODE, int L, bool HALO, class P>
int launch_ring_win(b200_ctx_t ctx, b200_csr_t A, const CsrArgsT<P> &args);
// offset-indexed operators (defined and instantiated in api_offsets.cu: 1..4 lanes per row)
template <int MODE, int L, bool HALO, class P>
int launch_ring_off(b200_ctx_t ctx, b200_csr_t A, const CsrArgsT<P> &args);
// pattern-indexed operators (defined and instantiated in api_patterns.cu: 1..4 lanes per row)
template <int MODE, int L, bool HALO, class P>
int launch_ring_pat(b200_ctx_t ctx, b200_csr_t A, const CsrArgsT<P> &args);

// which storage format does this launch stream?
template <class P>
inline int launch_format(b200_ctx_t ctx, b200_csr_t A) {
    if (ctx->opt_spmv_variant != 1) return FMT_PLAIN;
    if (A->pid && ctx->opt_patterns && A->lanes <= 4) return FMT_PATTERN;
    if (A->idx8 && ctx->opt_offsets && A->lanes <= 4) return FMT_OFFSET;
    if (A->col16 && ctx->opt_window && A->lanes <= 8 && A->win_runs >= 1) {
        int stages;
        if (ring_smem<P>(ctx, A, FMT_WINDOW, &stages) != 0) return FMT_WINDOW;
    }
    return FMT_PLAIN;
}

} // namespace b200
