// Host-side structures shared by hostlogic.cu, spmv.cu and dspmat.cu.
#pragma once
#include "common.cuh"
#include <vector>

namespace vexb {

inline int64_t read_index(const void *p, int bytes, size_t i) {
    return bytes == 4 ? (int64_t)((const uint32_t *)p)[i] : (int64_t)((const uint64_t *)p)[i];
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
