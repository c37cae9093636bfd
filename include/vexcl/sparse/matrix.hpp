#ifndef VEXCL_SPARSE_MATRIX_HPP
#define VEXCL_SPARSE_MATRIX_HPP
/*
 * vex::sparse::csr / ell / matrix: single-device sparse matrices (vexcl/sparse/csr.hpp:47-196,
 * vexcl/sparse/ell.hpp:61-508, vexcl/sparse/matrix.hpp:10-150).  Default index types are int.
 * csr -> the TMA-staged CSR row-block kernel, ell -> hybrid ELL (the reference's layout and width
 * rule; its device-side csr2ell conversion, ell.hpp:348-506, happens on the host at upload here),
 * matrix -> libvexb200's own choice (the reference picks csr on CPUs and ell on GPUs).
 */
#include <iterator>
#include <memory>
#include "../vector.hpp"
#include "product.hpp"

namespace vex {
namespace sparse {

namespace detail_sparse {
template <class R> inline size_t range_size(const R &r) { return static_cast<size_t>(std::distance(std::begin(r), std::end(r))); }
template <class R> inline auto range_data(const R &r) -> decltype(&*std::begin(r)) { return range_size(r) ? &*std::begin(r) : nullptr; }
}

template <int Format, typename Val, typename Col, typename Ptr>
class single_device_matrix {
    public:
        typedef Val value_type; typedef Val val_type; typedef Col col_type; typedef Ptr ptr_type;

        template <class PtrRange, class ColRange, class ValRange>
        single_device_matrix(const std::vector<backend::command_queue> &q, size_t nrows, size_t ncols,
                             const PtrRange &ptr, const ColRange &col, const ValRange &val, bool /*fast_setup*/ = true)
            : q(q), n(nrows), m(ncols), nnz(detail_sparse::range_size(val))
        {
            precondition(q.size() == 1, "sparse matrices of this kind are only supported for single-device contexts");
            static_assert(sizeof(Col) == 4 || sizeof(Col) == 8, "column type must be 32 or 64 bit");
            static_assert(sizeof(Ptr) == 4 || sizeof(Ptr) == 8, "pointer type must be 32 or 64 bit");
            vexb_spmat *h = nullptr;
            VEXB_CHECKED(vexb_csr_create(q[0].ordinal(), q[0].raw(), nrows, ncols, detail_sparse::range_data(ptr), sizeof(Ptr),
                                         detail_sparse::range_data(col), sizeof(Col), detail_sparse::range_data(val),
                                         dtype_of<Val>::value, Format, &h));
            A.reset(h, [](vexb_spmat *p) { vexb_spmat_destroy(p); });
        }
        single_device_matrix() : n(0), m(0), nnz(0) {}

        size_t rows() const { return n; }
        size_t cols() const { return m; }
        size_t nonzeros() const { return nnz; }
        const std::vector<backend::command_queue>& queue_list() const { return q; }

        /// y = A * x
        void mul(const vex::vector<Val> &x, vex::vector<Val> &y, Val alpha = 1, bool append = false) const {
            precondition(x.size() == m && y.size() == n, "sparse product: vector sizes do not match the matrix");
            VEXB_CHECKED(vexb_spmv(q[0].ordinal(), q[0].raw(), A.get(), x(0).raw(), y(0).raw(), static_cast<double>(alpha), append));
        }

        /// The strip when a generated kernel can walk its rows (CSR / hybrid ELL), else NULL.
        const vexb_spmat* inline_strip(unsigned = 0) const {
            vexb_spmat_info i;
            if (!A || vexb_spmat_get_info(A.get(), &i) != VEXB_OK) return nullptr;
            return (i.fmt == VEXB_FMT_CSR || i.fmt == VEXB_FMT_HELL) ? A.get() : nullptr;
        }

        template <class Expr>
        friend typename std::enable_if<is_vector_expr<Expr>::value, matrix_vector_product<single_device_matrix, Expr> >::type
        operator*(const single_device_matrix &A, const Expr &x) { return matrix_vector_product<single_device_matrix, Expr>(A, x); }
    private:
        std::vector<backend::command_queue> q;
        size_t n, m, nnz;
        std::shared_ptr<vexb_spmat> A;
};

template <typename Val, typename Col = int, typename Ptr = Col> using csr    = single_device_matrix<VEXB_FMT_CSR,  Val, Col, Ptr>;
template <typename Val, typename Col = int, typename Ptr = Col> using ell    = single_device_matrix<VEXB_FMT_HELL, Val, Col, Ptr>;
template <typename Val, typename Col = int, typename Ptr = Col> using matrix = single_device_matrix<VEXB_FMT_AUTO, Val, Col, Ptr>;

} // namespace sparse
} // namespace vex
#endif
