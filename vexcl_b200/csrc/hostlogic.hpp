// Host-side structures shared by hostlogic.cu, spmv.cu and dspmat.cu.
#pragma once
#include "common.cuh"
#include <vector>
#include <algorithm>

namespace vexb {

inline int64_t read_index(const void *p, int bytes, size_t i) {
    return bytes == 4 ? (int64_t)((const uint32_t *)p)[i] : (int64_t)((const uint64_t *)p)[i];
}

// SELL-32-sigma layout of a strip of n rows (csrc/spmv.cu, VEXB_FMT_SELL): perm[32 s + l] = the row lane l of slice s
// multiplies (-1: none); rows are taken in order, windows of `sigma` rows sorted by length, longest first, ties in row
// order; sptr[s] = first slot of slice s, a slice being 32 lanes x the length of its longest row.  False when the slot
// count would not fit 32-bit offsets.
inline bool sell_layout(size_t n, const int *rowptr, long sigma, std::vector<int> &perm, std::vector<int> &sptr, size_t *slots_out) {
    sigma = (sigma < 32 ? 32 : sigma > (1l << 20) ? (1l << 20) : sigma) & ~31l;
    const size_t ns = (n + 31) / 32;
    perm.assign(ns * 32, -1);
    for (size_t i = 0; i < n; ++i) perm[i] = (int)i;
    for (size_t w0 = 0; w0 < n; w0 += (size_t)sigma) {
        const size_t w1 = w0 + (size_t)sigma < n ? w0 + (size_t)sigma : n;
        std::stable_sort(perm.begin() + w0, perm.begin() + w1, [&](int a, int b) {
            return rowptr[a + 1] - rowptr[a] > rowptr[b + 1] - rowptr[b]; });
    }
    sptr.assign(ns + 1, 0);
    size_t slots = 0;
    for (size_t sl = 0; sl < ns; ++sl) {
        int wmax = 0;
        for (int l = 0; l < 32; ++l) { const int r = perm[sl * 32 + l]; if (r >= 0 && rowptr[r + 1] - rowptr[r] > wmax) wmax = rowptr[r + 1] - rowptr[r]; }
        slots += (size_t)wmax * 32;
        if (slots >= (size_t)INT32_MAX - 64) return false;
        sptr[sl + 1] = (int)slots;
    }
    *slots_out = slots;
    return true;
}

} // namespace vexb

struct vexb_halo_plan {
    int nparts = 0;
    std::vector<size_t> col_part;
    std::vector<std::vector<int64_t>> ghost;         // per part: sorted global ghost columns
    // reference-equivalent tables (spmat.hpp:319-371)
    std::vector<int64_t> cols_to_send;               // owner-relative
    std::vector<size_t> cidx;
    std::vector<std::vector<int64_t>> cols_to_recv;
    // pairwise form
    std::vector<std::vector<size_t>> send_counts, recv_counts;   // [part][peer]
    std::vector<std::vector<int64_t>> send_cols;                  // [part]: local x indices grouped by peer
};
