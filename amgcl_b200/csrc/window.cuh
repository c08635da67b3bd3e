// window.cuh -- host-side construction of the windowed operator format (csr_kernels.cuh, WIN).
//
// For every row block (walk order) the distinct 32-byte sectors of x its entries gather from
// are collected, neighbouring sectors are merged into runs (a hole of one sector is fetched
// rather than starting a new run), runs are cut to kWinRunLen elements, and the runs are laid
// one after the other into the block's window.  Every entry gets the 16-bit position of its
// column inside that window.  A block whose window or run list would not fit the shared memory
// set aside for it is cut in two (at a multiple of four rows) until it does.  An operator
// qualifies when the windows are, in total, clearly smaller than the number of entries
// (otherwise gathering straight from global memory moves less data).
//
// Pure host logic (exported as b200_window_plan_i64 for the CPU tests).  Cost: one pass over
// the entries (a bitmap of touched sectors per thread, no sorting of entries), all host threads.
#pragma once
#include "common.cuh"
#include "csr_kernels.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <memory>
#include <vector>
#include <omp.h>

namespace b200 {

constexpr int kWinRunCapMax = 126;   // most runs a block may have (stage layout: 1 KB)

struct WindowPlan {
    std::vector<int4>           blk4;    // [nblocks+1] block descriptors (walk order; blocks may have been cut)
    std::vector<unsigned short> col16;   // [nnz] window-local column of every entry
    std::vector<int2>           runs;    // {first column, length | first slot << 16}, walk order
    std::vector<int2>           wblk;    // [nblocks] {first run, end run}
    int     max_slots = 0, max_runs = 0;
    int64_t total_slots = 0;
    int64_t nsplit = 0;                  // blocks cut because their window did not fit
};

namespace detail {

struct WinScratch {
    uint64_t            *bits = nullptr;     // one bit per sector of x (64 sectors per word), all zero between blocks
    int                 *wmap = nullptr;     // word -> its index among the current block's touched words
    std::vector<int>     tbl;                // [touched words][64] sector -> window slot of the current block
    std::vector<int64_t> touched, sec;
    bool init(int64_t nsec) {
        bits = static_cast<uint64_t *>(calloc((size_t)(nsec / 64 + 2), sizeof(uint64_t)));
        wmap = static_cast<int *>(malloc((size_t)(nsec / 64 + 2) * sizeof(int)));
        return bits && wmap;
    }
    int slot_of(int64_t c) const { return tbl[(size_t)wmap[c >> 8] * 64 + (size_t)((c >> 2) & 63)] + (int)(c & 3); }
    ~WinScratch() { free(bits); free(wmap); }
};

// The window of entries [e0, e1): appends its runs to `out` and fills the sector -> slot table.
// Returns the number of slots, or -1 when a cap is exceeded (out is then unchanged).
template <class Col>
inline int window_of(const Col *col, int64_t e0, int64_t e1, int64_t ncols, int slot_cap, int run_cap,
                     int gap, WinScratch &ws, std::vector<int2> &out) {
    ws.touched.clear();
    ws.sec.clear();
    for (int64_t e = e0; e < e1; ++e) {
        const int64_t s = (int64_t)col[e] >> 2, w = s >> 6;
        if (!ws.bits[w]) ws.touched.push_back(w);
        ws.bits[w] |= (uint64_t)1 << (s & 63);
    }
    std::sort(ws.touched.begin(), ws.touched.end());
    ws.tbl.resize(ws.touched.size() * 64);
    for (size_t k = 0; k < ws.touched.size(); ++k) {
        const int64_t w = ws.touched[k];
        ws.wmap[w] = (int)k;
        uint64_t word = ws.bits[w];
        ws.bits[w] = 0;
        while (word) {
            ws.sec.push_back(w * 64 + __builtin_ctzll(word));
            word &= word - 1;
        }
    }
    const size_t base = out.size();
    const std::vector<int64_t> &sec = ws.sec;
    int slots = 0;
    size_t i = 0;
    while (i < sec.size()) {
        size_t j = i;
        while (j + 1 < sec.size() && sec[j + 1] - sec[j] <= gap) ++j;
        int64_t first = sec[i] * 4;
        int64_t len   = (sec[j] - sec[i] + 1) * 4;
        if (first + len > ncols) len = ncols - first;
        if (slots + len > slot_cap) { out.resize(base); return -1; }
        for (size_t k = i; k <= j; ++k)
            ws.tbl[(size_t)ws.wmap[sec[k] >> 6] * 64 + (size_t)(sec[k] & 63)] = slots + (int)(sec[k] - sec[i]) * 4;
        while (len > 0) {
            const int piece = (int)std::min<int64_t>(len, kWinRunLen);
            if ((int)(out.size() - base) >= run_cap) { out.resize(base); return -1; }
            out.push_back(make_int2((int)first, (int)((unsigned)piece | ((unsigned)slots << 16))));
            slots += piece; first += piece; len -= piece;
        }
        i = j + 1;
    }
    return slots;
}
} // namespace detail

// blk4: the block descriptors in walk order ({r0 or ~r0, r1, e0, e1}); ptr: the (int32) row
// pointers.  Returns false when the operator does not qualify (w is then unspecified).
template <class Col>
inline bool build_windows(const int4 *blk4, int64_t nblocks, const int32_t *ptr, const Col *col, int64_t ncols,
                          int64_t nnz, int slot_cap, int run_cap, double max_ratio, int gap, WindowPlan &w) {
    // gap: sectors s and s + gap still belong to one run (gap 2: a hole of one sector is fetched)
    gap = std::max(1, std::min(gap, 8));
    if (nblocks <= 0 || nnz <= 0 || slot_cap < 4 * kWinRunLen || slot_cap > 65535) return false;
    run_cap = std::min(run_cap, kWinRunCapMax);
    const int64_t nsec = ncols / 4 + 1;
    // a sample first: most operators that do not qualify (restrictions) are rejected here
    {
        detail::WinScratch ws;
        if (!ws.init(nsec)) return false;
        const int64_t stride = std::max<int64_t>(1, nblocks / 256);
        int64_t slots = 0, entries = 0;
        std::vector<int2> runs;
        for (int64_t b = 0; b < nblocks; b += stride) {
            runs.clear();
            const int n = detail::window_of(col, blk4[b].z, blk4[b].w, ncols, 1 << 30, 1 << 30, gap, ws, runs);
            slots += n;
            entries += blk4[b].w - blk4[b].z;
        }
        if ((double)slots > max_ratio * (double)entries) return false;
    }
    const int nth = std::max(1, omp_get_max_threads());
    struct Local {
        std::vector<int4> blk;
        std::vector<int2> wblk, runs;
        int max_slots = 0, max_runs = 0;
        int64_t total = 0, nsplit = 0;
        bool bad = false;
    };
    std::vector<Local> loc((size_t)nth);
    w.col16.assign((size_t)nnz, 0);
    int used = 1;
#pragma omp parallel num_threads(nth)
    {
        const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#pragma omp single
        used = nt;
        Local &L = loc[(size_t)tid];
        detail::WinScratch ws;
        if (!ws.init(nsec)) L.bad = true;
        const int64_t lo = nblocks * tid / nt, hi = nblocks * (tid + 1) / nt;   // contiguous: order is kept
        std::vector<int2> todo;                                                  // row ranges still to place
        for (int64_t b = lo; b < hi && !L.bad; ++b) {
            const bool halo = blk4[b].x < 0;
            todo.clear();
            todo.push_back(make_int2(halo ? ~blk4[b].x : blk4[b].x, blk4[b].y));
            while (!todo.empty() && !L.bad) {
                const int2 rr = todo.back();
                todo.pop_back();
                const int64_t e0 = ptr[rr.x], e1 = ptr[rr.y];
                const size_t q0 = L.runs.size();
                const int n = detail::window_of(col, e0, e1, ncols, slot_cap, run_cap, gap, ws, L.runs);
                if (n < 0) {
                    const int quads = (rr.y - rr.x + 3) / 4;
                    if (quads < 2) { L.bad = true; break; }      // four rows alone do not fit
                    const int mid = rr.x + 4 * (quads / 2);
                    todo.push_back(make_int2(mid, rr.y));        // (second half is placed second)
                    todo.push_back(make_int2(rr.x, mid));
                    L.nsplit++;
                    continue;
                }
                for (int64_t e = e0; e < e1; ++e) {
                    w.col16[(size_t)e] = (unsigned short)ws.slot_of((int64_t)col[e]);
                }
                L.blk.push_back(make_int4(halo ? ~rr.x : rr.x, rr.y, (int)e0, (int)e1));
                L.wblk.push_back(make_int2((int)q0, (int)L.runs.size()));
                L.max_slots = std::max(L.max_slots, n);
                L.max_runs = std::max(L.max_runs, (int)(L.runs.size() - q0));
                L.total += n;
            }
        }
    }
    int64_t nb = 0, nq = 0;
    w.max_slots = w.max_runs = 0; w.total_slots = 0; w.nsplit = 0;
    for (int t = 0; t < used; ++t) {
        if (loc[(size_t)t].bad) return false;
        nb += (int64_t)loc[(size_t)t].blk.size();
        nq += (int64_t)loc[(size_t)t].runs.size();
        w.max_slots = std::max(w.max_slots, loc[(size_t)t].max_slots);
        w.max_runs = std::max(w.max_runs, loc[(size_t)t].max_runs);
        w.total_slots += loc[(size_t)t].total;
        w.nsplit += loc[(size_t)t].nsplit;
    }
    if ((double)w.total_slots > max_ratio * (double)nnz) return false;
    if (nq > (int64_t)std::numeric_limits<int32_t>::max() - 8 || w.max_runs < 1) return false;
    w.blk4.clear(); w.blk4.reserve((size_t)nb + 1);
    w.wblk.clear(); w.wblk.reserve((size_t)nb);
    w.runs.clear(); w.runs.reserve((size_t)nq);
    for (int t = 0; t < used; ++t) {
        const Local &L = loc[(size_t)t];
        const int qbase = (int)w.runs.size();
        w.blk4.insert(w.blk4.end(), L.blk.begin(), L.blk.end());
        for (const int2 &q : L.wblk) w.wblk.push_back(make_int2(q.x + qbase, q.y + qbase));
        w.runs.insert(w.runs.end(), L.runs.begin(), L.runs.end());
    }
    w.blk4.push_back(blk4[nblocks]);      // the sentinel
    return true;
}

} // namespace b200
