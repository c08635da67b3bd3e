// offsets.cuh -- host-side construction of the offset-indexed column format (csr_kernels.cuh,
// FMT_OFFSET): if col - row takes at most 256 distinct values over the whole operator, every
// entry is stored as the 8-bit index of its offset in a sorted table.
//
// Pure host logic (exported as b200_offset_plan_i64 for the CPU tests).  Two passes over the
// entries on all host threads; a row usually repeats the pattern of the row before it, which is
// checked first, so the common case costs one comparison per entry.
#pragma once
#include "common.cuh"
#include "csr_kernels.cuh"

#include <algorithm>
#include <cstdint>
#include <vector>
#include <omp.h>

namespace b200 {

struct OffsetPlan {
    std::vector<unsigned char> idx8;     // [nnz] index of (col - row) in tab
    int tab[kOffTabLen] = {};            // sorted distinct offsets, zero-padded
    int count = 0;
};

// Returns false when the operator has more than 256 distinct offsets.
template <class Col>
inline bool build_offsets(int64_t nrows, const int32_t *ptr, const Col *col, OffsetPlan &o) {
    const int64_t nnz = nrows ? ptr[nrows] : 0;
    if (nrows <= 0 || nnz <= 0) return false;
    const int nth = std::max(1, omp_get_max_threads());
    std::vector<std::vector<int>> sets((size_t)nth);
    int bad = 0, used = 1;
#pragma omp parallel num_threads(nth)
    {
        const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#pragma omp single
        used = nt;
        std::vector<int> &set = sets[(size_t)tid];
        std::vector<int> prev, cur;
        bool fail = false;
        const int64_t lo = nrows * tid / nt, hi = nrows * (tid + 1) / nt;
        for (int64_t r = lo; r < hi && !fail; ++r) {
            cur.clear();
            for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) {
                const int d = (int)((int64_t)col[e] - r);
                const size_t j = cur.size();
                cur.push_back(d);
                if (j < prev.size() && prev[j] == d) continue;          // same diagonal as the row above
                std::vector<int>::iterator it = std::lower_bound(set.begin(), set.end(), d);
                if (it != set.end() && *it == d) continue;
                if ((int)set.size() >= kOffTabLen) { fail = true; break; }
                set.insert(it, d);
            }
            prev.swap(cur);
        }
        if (fail) {
#pragma omp atomic write
            bad = 1;
        }
    }
    if (bad) return false;
    std::vector<int> all;
    for (int t = 0; t < used; ++t) all.insert(all.end(), sets[(size_t)t].begin(), sets[(size_t)t].end());
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    if ((int)all.size() > kOffTabLen) return false;
    o.count = (int)all.size();
    std::fill(o.tab, o.tab + kOffTabLen, 0);
    std::copy(all.begin(), all.end(), o.tab);
    o.idx8.assign((size_t)nnz, 0);
    const int *tab = o.tab;
    const int count = o.count;
#pragma omp parallel num_threads(nth)
    {
        const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
        std::vector<int> prev, cur;                  // offsets of the previous / current row
        std::vector<unsigned char> pidx, cidx;       // ... and their table indices
        const int64_t lo = nrows * tid / nt, hi = nrows * (tid + 1) / nt;
        for (int64_t r = lo; r < hi; ++r) {
            cur.clear(); cidx.clear();
            for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) {
                const int d = (int)((int64_t)col[e] - r);
                const size_t j = cur.size();
                unsigned char k;
                if (j < prev.size() && prev[j] == d) k = pidx[j];
                else k = (unsigned char)(std::lower_bound(tab, tab + count, d) - tab);
                cur.push_back(d); cidx.push_back(k);
                o.idx8[(size_t)e] = k;
            }
            prev.swap(cur); pidx.swap(cidx);
        }
    }
    return true;
}

} // namespace b200
