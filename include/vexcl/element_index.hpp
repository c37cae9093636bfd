#ifndef VEXCL_ELEMENT_INDEX_HPP
#define VEXCL_ELEMENT_INDEX_HPP
// vex::element_index(offset, length): the global position of the element being computed
// (vexcl/element_index.hpp:40-111).  The per-device part_start is supplied by the launch
// (index_offset argument of vexb_eval / vexb_reduce).
#include "operations.hpp"

namespace vex {

struct elem_index : vector_expr_tag {
    VEXCL_NODE_COMMON
    typedef size_t value_type;
    size_t offset, length;
    elem_index(size_t offset = 0, size_t length = 0) : offset(offset), length(length) {}
    int lower(detail::ir_builder &b) const { b.push_index(static_cast<long long>(offset)); return VEXB_U64; }
    void props(detail::expr_props &p) const { if (length) p.see_size(length); }
};

inline elem_index element_index(size_t offset = 0, size_t length = 0) { return elem_index(offset, length); }

} // namespace vex
#endif
