#ifndef VEXCL_SPARSE_PRODUCT_HPP
#define VEXCL_SPARSE_PRODUCT_HPP
/*
 * `A * x` for the vex::sparse classes as a *terminal* of vector expressions
 * (vexcl/sparse/product.hpp:45-130): usable anywhere a vector is, e.g.
 *     y = x + A * x;      s = sum(f - A * x);      y = A * (2 * x + z);
 * In the reference the row loop is emitted into the consumer's kernel.  Here the product
 * is evaluated by the SpMV kernels into a temporary when the enclosing expression is
 * launched, and the temporary is the terminal the IR sees (fusing the epilogue into the
 * SpMV kernel is listed as next work in DESIGN.md).
 */
#include <memory>
#include "../operations.hpp"
#include "../vector.hpp"

namespace vex {
namespace sparse {

template <class Matrix, class Vector>
struct matrix_vector_product : vector_expr_tag {
    static const bool hold_by_reference = false;
    typedef typename Matrix::value_type value_type;

    const Matrix &A;
    typename detail::operand<Vector>::type x;
    mutable std::shared_ptr<vex::vector<value_type>> y, xt;

    matrix_vector_product(const Matrix &A, const Vector &x) : A(A), x(detail::operand<Vector>::wrap(x)) {}

    // Evaluated once per launch of the enclosing expression, before lowering.
    void props(detail::expr_props &p) const {
        const vex::vector<value_type> &xv = materialize(x);
        if (!y || y->size() != A.rows()) y = std::make_shared<vex::vector<value_type>>(A.queue_list(), A.rows());
        A.mul(xv, *y);
        y->props(p);
    }
    int lower(detail::ir_builder &b) const { return y->lower(b); }
    private:
        const vex::vector<value_type>& materialize(const vex::vector<value_type> &v) const { return v; }
        template <class E>
        const vex::vector<value_type>& materialize(const E &e) const {
            if (!xt || xt->size() != A.cols()) xt = std::make_shared<vex::vector<value_type>>(A.queue_list(), A.cols());
            *xt = e;
            return *xt;
        }
};

} // namespace sparse
} // namespace vex
#endif
