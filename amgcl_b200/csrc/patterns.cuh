// patterns.cuh -- host-side construction of the pattern-indexed row format (csr_kernels.cuh,
// FMT_PATTERN): the tuple (col - row of every entry, in entry order) of a row is its pattern; if
// the operator has at most 256 distinct patterns whose offsets total at most 1024, every row is
// stored as the 8-bit id of its pattern and the entries carry no column information at all.
//
// Pure host logic (exported as b200_pattern_plan_i64 for the CPU tests).  One pass over the
// entries on all host threads: a row usually repeats the pattern of the row before it, which is
// checked first; only a row that differs is looked up in the thread's pattern list.
#pragma once
#include "common.cuh"
#include "csr_kernels.cuh"

#include <algorithm>
#include <cstdint>
#include <map>
#include <vector>
#include <omp.h>

namespace b200 {

struct PatternPlan {
    std::vector<unsigned char>  pid;       // [nrows] pattern of every row
    std::vector<unsigned short> start;     // [kPatCap + 1] first table entry of every pattern
    std::vector<int>            off;       // [kPatOffCap] the patterns' offsets, one after the other
    int count = 0, total = 0;
};

// Returns false when the operator has too many patterns.
template <class Col>
inline bool build_patterns(int64_t nrows, const int32_t *ptr, const Col *col, PatternPlan &o) {
    const int64_t nnz = nrows ? ptr[nrows] : 0;
    if (nrows <= 0 || nnz <= 0) return false;
    const int nth = std::max(1, omp_get_max_threads());
    typedef std::vector<int> Pat;
    std::vector<std::vector<Pat>> lists((size_t)nth);          // each thread's patterns, in order of appearance
    std::vector<int> local((size_t)nrows);                      // row -> index in its thread's list
    int bad = 0, used = 1;
#pragma omp parallel num_threads(nth)
    {
        const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#pragma omp single
        used = nt;
        std::vector<Pat> &list = lists[(size_t)tid];
        std::map<Pat, int> index;
        Pat cur;
        int prev = -1;
        bool fail = false;
        const int64_t lo = nrows * tid / nt, hi = nrows * (tid + 1) / nt;
        for (int64_t r = lo; r < hi && !fail; ++r) {
            const int64_t e0 = ptr[r], e1 = ptr[r + 1];
            const int len = (int)(e1 - e0);
            // the pattern of the row above?
            bool same = prev >= 0 && (int)list[(size_t)prev].size() == len;
            if (same) {
                const Pat &p = list[(size_t)prev];
                for (int k = 0; k < len && same; ++k) same = p[(size_t)k] == (int)((int64_t)col[e0 + k] - r);
            }
            if (!same) {
                if (len > kPatOffCap) { fail = true; break; }
                cur.resize((size_t)len);
                for (int k = 0; k < len; ++k) cur[(size_t)k] = (int)((int64_t)col[e0 + k] - r);
                std::map<Pat, int>::iterator it = index.find(cur);
                if (it == index.end()) {
                    if ((int)list.size() >= kPatCap) { fail = true; break; }
                    prev = (int)list.size();
                    list.push_back(cur);
                    index[cur] = prev;
                } else {
                    prev = it->second;
                }
            }
            local[(size_t)r] = prev;
        }
        if (fail) {
#pragma omp atomic write
            bad = 1;
        }
    }
    if (bad) return false;
    // global pattern list + every thread's local -> global translation
    std::map<Pat, int> gindex;
    std::vector<const Pat *> gl;
    std::vector<std::vector<unsigned char>> xlat((size_t)used);
    int total = 0;
    for (int t = 0; t < used; ++t) {
        xlat[(size_t)t].resize(lists[(size_t)t].size());
        for (size_t i = 0; i < lists[(size_t)t].size(); ++i) {
            const Pat &p = lists[(size_t)t][i];
            std::map<Pat, int>::iterator it = gindex.find(p);
            int g;
            if (it == gindex.end()) {
                if ((int)gl.size() >= kPatCap || total + (int)p.size() > kPatOffCap) return false;
                g = (int)gl.size();
                gl.push_back(&p);
                gindex[p] = g;
                total += (int)p.size();
            } else {
                g = it->second;
            }
            xlat[(size_t)t][i] = (unsigned char)g;
        }
    }
    o.count = (int)gl.size();
    o.total = total;
    o.start.assign((size_t)kPatCap + 1, (unsigned short)total);
    o.off.assign((size_t)kPatOffCap, 0);
    int pos = 0;
    for (int g = 0; g < o.count; ++g) {
        o.start[(size_t)g] = (unsigned short)pos;
        std::copy(gl[(size_t)g]->begin(), gl[(size_t)g]->end(), o.off.begin() + pos);
        pos += (int)gl[(size_t)g]->size();
    }
    o.pid.resize((size_t)nrows);
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < used; ++t) {                   // (the row ranges of the first pass)
        const int64_t lo = nrows * t / used, hi = nrows * (t + 1) / used;
        const std::vector<unsigned char> &x = xlat[(size_t)t];
        for (int64_t r = lo; r < hi; ++r) o.pid[(size_t)r] = x[(size_t)local[(size_t)r]];
    }
    return true;
}

} // namespace b200
