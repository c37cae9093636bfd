// Host-only index work of the hot path: slice boundaries and the halo plan.
// Nothing here touches a GPU, so these entry points also work on a CPU-only box
// (the `not gpu` tests check them bit-exactly against oracle/).
//
//   vexb_partition        <- partitioning_scheme<>::get      vexcl/vector.hpp:131-167
//   vexb_strip_ghost_cols <- ghost set construction          vexcl/spmat.hpp:300-316
//   vexb_halo_plan_*      <- the rest of setup_exchange      vexcl/spmat.hpp:319-371
//
// The reference builds these with std::set / unordered_map per device; here
// they are sorted vectors and binary searches (same tables, O(nnz) memory
// traffic instead of O(nnz log nnz) node allocations).
#include "hostlogic.hpp"
#include <algorithm>

using namespace vexb;

extern "C" int vexb_partition(size_t n, int nparts, const double *weights, size_t *part) {
    VEXB_CHECK(nparts >= 1 && part, "bad arguments");
    part[0] = 0;
    if (nparts > 1) {
        std::vector<double> cumsum(nparts + 1, 0.0);
        for (int d = 0; d < nparts; ++d) {
            const double w = weights ? weights[d] : 1.0;
            VEXB_CHECK(w >= 0, "negative weight for part %d", d);
            cumsum[d + 1] = cumsum[d] + w;
        }
        VEXB_CHECK(cumsum[nparts] > 0, "all weights are zero");
        for (int d = 1; d < nparts; ++d) {
            // same expression, evaluated in double, as vector.hpp:157-162
            size_t b = static_cast<size_t>(n * cumsum[d] / cumsum[nparts]);
            b = (b + 15) / 16 * 16;                                   // util.hpp:91-93, m = 16
            part[d] = std::min(n, b);
        }
    }
    part[nparts] = n;
    return VEXB_OK;
}

extern "C" int vexb_strip_ghost_cols(size_t nrows, const void *ptr, int ptr_bytes, const void *col, int col_bytes,
                                     size_t col_begin, size_t col_end, int64_t *out, size_t *count) {
    VEXB_CHECK(count, "count is NULL");
    VEXB_CHECK(ptr_bytes == 4 || ptr_bytes == 8, "ptr_bytes must be 4 or 8");
    VEXB_CHECK(col_bytes == 4 || col_bytes == 8, "col_bytes must be 4 or 8");
    VEXB_CHECK(nrows == 0 || (ptr && col), "NULL matrix arrays");
    std::vector<int64_t> g;
    if (nrows) {
        const int64_t j0 = read_index(ptr, ptr_bytes, 0), j1 = read_index(ptr, ptr_bytes, nrows);
        VEXB_CHECK(j1 >= j0, "row pointers decrease");
        size_t sorted_upto = 0;
        for (int64_t j = j0; j < j1; ++j) {
            const int64_t c = read_index(col, col_bytes, (size_t)(j - j0));
            if ((size_t)c < col_begin || (size_t)c >= col_end) {
                if (!g.empty() && g.back() == c) continue;
                g.push_back(c);
                if (g.size() - sorted_upto > (1u << 22)) {
                    std::sort(g.begin(), g.end());
                    g.erase(std::unique(g.begin(), g.end()), g.end());
                    sorted_upto = g.size();
                }
            }
        }
        std::sort(g.begin(), g.end());
        g.erase(std::unique(g.begin(), g.end()), g.end());
    }
    if (out) {
        VEXB_CHECK(*count >= g.size(), "output buffer too small (%zu < %zu)", *count, g.size());
        std::copy(g.begin(), g.end(), out);
    }
    *count = g.size();
    return VEXB_OK;
}

extern "C" int vexb_halo_plan_create(int nparts, const size_t *col_part, const int64_t *ghost_cols,
                                     const size_t *ghost_off, vexb_halo_plan **plan) {
    VEXB_CHECK(nparts >= 1 && col_part && ghost_off && plan, "bad arguments");
    VEXB_CHECK(ghost_off[nparts] == 0 || ghost_cols, "ghost_cols is NULL");
    auto *p = new vexb_halo_plan();
    p->nparts = nparts;
    p->col_part.assign(col_part, col_part + nparts + 1);
    p->ghost.resize(nparts);
    for (int d = 0; d < nparts; ++d) {
        p->ghost[d].assign(ghost_cols + ghost_off[d], ghost_cols + ghost_off[d + 1]);
        const auto &g = p->ghost[d];
        for (size_t i = 0; i < g.size(); ++i) {
            const bool ok = (i == 0 || g[i - 1] < g[i]) && g[i] >= 0 && (size_t)g[i] < col_part[nparts] &&
                            !((size_t)g[i] >= col_part[d] && (size_t)g[i] < col_part[d + 1]);
            if (!ok) { delete p; VEXB_FAIL(VEXB_ERR_INVALID, "ghost list of part %d is not a sorted set of remote columns (entry %zu)", d, i); }
        }
    }
    // Reference tables: sorted union, owner offsets, receive positions (spmat.hpp:319-358).
    for (int d = 0; d < nparts; ++d) p->cols_to_send.insert(p->cols_to_send.end(), p->ghost[d].begin(), p->ghost[d].end());
    std::sort(p->cols_to_send.begin(), p->cols_to_send.end());
    p->cols_to_send.erase(std::unique(p->cols_to_send.begin(), p->cols_to_send.end()), p->cols_to_send.end());
    p->cidx.resize(nparts + 1);
    for (int d = 0; d <= nparts; ++d)
        p->cidx[d] = std::lower_bound(p->cols_to_send.begin(), p->cols_to_send.end(), (int64_t)col_part[d]) - p->cols_to_send.begin();
    p->cols_to_recv.resize(nparts);
    for (int d = 0; d < nparts; ++d) {
        p->cols_to_recv[d].resize(p->ghost[d].size());
        for (size_t j = 0; j < p->ghost[d].size(); ++j)
            p->cols_to_recv[d][j] = std::lower_bound(p->cols_to_send.begin(), p->cols_to_send.end(), p->ghost[d][j]) - p->cols_to_send.begin();
    }
    // Pairwise segments: part d's sorted ghost list splits into contiguous runs per owner.
    p->recv_counts.assign(nparts, std::vector<size_t>(nparts, 0));
    p->send_counts.assign(nparts, std::vector<size_t>(nparts, 0));
    p->send_cols.resize(nparts);
    for (int d = 0; d < nparts; ++d) {
        const auto &g = p->ghost[d];
        for (int o = 0; o < nparts; ++o) {
            if (o == d) continue;
            const auto lo = std::lower_bound(g.begin(), g.end(), (int64_t)col_part[o]);
            const auto hi = std::lower_bound(g.begin(), g.end(), (int64_t)col_part[o + 1]);
            p->recv_counts[d][o] = hi - lo;
        }
    }
    for (int o = 0; o < nparts; ++o) {
        for (int d = 0; d < nparts; ++d) {
            if (o == d) continue;
            const auto &g = p->ghost[d];
            const auto lo = std::lower_bound(g.begin(), g.end(), (int64_t)col_part[o]);
            const auto hi = std::lower_bound(g.begin(), g.end(), (int64_t)col_part[o + 1]);
            p->send_counts[o][d] = hi - lo;
            for (auto it = lo; it != hi; ++it) p->send_cols[o].push_back(*it - (int64_t)col_part[o]);
        }
    }
    // cols_to_send is stored owner-relative, like spmat.hpp:362-363
    for (int d = 0; d < nparts; ++d)
        for (size_t i = p->cidx[d]; i < p->cidx[d + 1]; ++i) p->cols_to_send[i] -= (int64_t)col_part[d];
    *plan = p;
    return VEXB_OK;
}

extern "C" int vexb_halo_plan_destroy(vexb_halo_plan *plan) { delete plan; return VEXB_OK; }

extern "C" int vexb_halo_plan_ref_sizes(const vexb_halo_plan *plan, size_t *n_send_total) {
    VEXB_CHECK(plan && n_send_total, "NULL argument");
    *n_send_total = plan->cols_to_send.size();
    return VEXB_OK;
}

extern "C" int vexb_halo_plan_ref_tables(const vexb_halo_plan *plan, int64_t *cols_to_send, size_t *cidx) {
    VEXB_CHECK(plan, "plan is NULL");
    if (cols_to_send) std::copy(plan->cols_to_send.begin(), plan->cols_to_send.end(), cols_to_send);
    if (cidx) std::copy(plan->cidx.begin(), plan->cidx.end(), cidx);
    return VEXB_OK;
}

extern "C" int vexb_halo_plan_ref_recv(const vexb_halo_plan *plan, int part, int64_t *cols_to_recv) {
    VEXB_CHECK(plan && part >= 0 && part < plan->nparts && cols_to_recv, "bad arguments");
    std::copy(plan->cols_to_recv[part].begin(), plan->cols_to_recv[part].end(), cols_to_recv);
    return VEXB_OK;
}

extern "C" int vexb_halo_plan_counts(const vexb_halo_plan *plan, int part, size_t *send_counts, size_t *recv_counts) {
    VEXB_CHECK(plan && part >= 0 && part < plan->nparts, "bad arguments");
    if (send_counts) std::copy(plan->send_counts[part].begin(), plan->send_counts[part].end(), send_counts);
    if (recv_counts) std::copy(plan->recv_counts[part].begin(), plan->recv_counts[part].end(), recv_counts);
    return VEXB_OK;
}

extern "C" int vexb_halo_plan_send_cols(const vexb_halo_plan *plan, int part, int64_t *send_cols) {
    VEXB_CHECK(plan && part >= 0 && part < plan->nparts && send_cols, "bad arguments");
    std::copy(plan->send_cols[part].begin(), plan->send_cols[part].end(), send_cols);
    return VEXB_OK;
}
