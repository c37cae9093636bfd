#ifndef VEXCL_SPMAT_CCSR_HPP
#define VEXCL_SPMAT_CCSR_HPP
/*
 * vex::SpMatCCSR<val_t, col_t, idx_t> (vexcl/spmat/ccsr.hpp:54-86): "compressed CSR".  Only the unique rows of the
 * matrix are stored, column positions are relative to the diagonal, and idx[i] names the unique row of matrix row i.
 * Single device, like the reference.  The product is libvexb200's ccsr_kernel (csrc/ccsr.cu); `A * x` takes part in
 * `y = A * x`, `y += A * x`, `y = expr + A * x`, ... through the same additive-operator rules as vex::SpMat, for
 * vectors and multivectors.
 */
#include <memory>
#include "../vector.hpp"
#include "../multivector.hpp"

namespace vex {

template <typename val_t, typename col_t = ptrdiff_t, typename idx_t = size_t>
struct SpMatCCSR {
    static_assert(std::is_signed<col_t>::value, "Column type for CCSR format has to be signed.");
    static_assert(sizeof(col_t) == 4 || sizeof(col_t) == 8, "column type must be 32 or 64 bit");
    static_assert(sizeof(idx_t) == 4 || sizeof(idx_t) == 8, "index type must be 32 or 64 bit");
    typedef val_t value_type;

    /// n rows, m unique rows; idx: n entries, row: m+1 offsets into col/val (ccsr.hpp:70-78).
    SpMatCCSR(const backend::command_queue &queue, size_t n, size_t m,
              const idx_t *idx, const idx_t *row, const col_t *col, const val_t *val)
        : queue(queue), n(n)
    {
        vexb_ccsr *h = nullptr;
        VEXB_CHECKED(vexb_ccsr_create(queue.ordinal(), queue.raw(), n, m, idx, sizeof(idx_t), row, sizeof(idx_t),
                                      col, sizeof(col_t), val, dtype_of<val_t>::value, &h));
        mtx.reset(h, [](vexb_ccsr *p) { vexb_ccsr_destroy(p); });
    }

    void apply(const vex::vector<val_t> &x, vex::vector<val_t> &y, val_t alpha = 1, bool append = false) const {
        precondition(x.nparts() == 1 && y.nparts() == 1, "SpMatCCSR works with single-device vectors only");
        precondition(x.size() == n && y.size() == n, "SpMatCCSR::apply: vector sizes do not match the matrix");
        precondition(x.queue_list()[0].ordinal() == queue.ordinal(), "SpMatCCSR and its vectors must live on the same device");
        VEXB_CHECKED(vexb_ccsr_spmv(queue.ordinal(), y.queue_list()[0].raw(), mtx.get(), x(0).raw(), y(0).raw(),
                                    static_cast<double>(alpha), append));
    }
    size_t rows() const { return n; }
    size_t cols() const { return n; }
    vexb_ccsr_info info() const { vexb_ccsr_info i; VEXB_CHECKED(vexb_ccsr_get_info(mtx.get(), &i)); return i; }

    backend::command_queue queue;
    size_t n;
    std::shared_ptr<vexb_ccsr> mtx;
};

template <typename val_t, typename col_t, typename idx_t>
additive_operator<SpMatCCSR<val_t, col_t, idx_t>, vector<val_t>>
operator*(const SpMatCCSR<val_t, col_t, idx_t> &A, const vector<val_t> &x) {
    return additive_operator<SpMatCCSR<val_t, col_t, idx_t>, vector<val_t>>(A, x);
}
template <typename val_t, typename col_t, typename idx_t, size_t N>
additive_operator<SpMatCCSR<val_t, col_t, idx_t>, multivector<val_t, N>>
operator*(const SpMatCCSR<val_t, col_t, idx_t> &A, const multivector<val_t, N> &x) {
    return additive_operator<SpMatCCSR<val_t, col_t, idx_t>, multivector<val_t, N>>(A, x);
}

} // namespace vex
#endif
