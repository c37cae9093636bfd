// The reference's tests/multivector_create.cpp, tests/multivector_arithmetics.cpp and the multivector cases of
// tests/spmv.cpp (:262-343) against include/vexcl, plus the read-before-write rule of tied assignments
// (vexcl/operations.hpp:2238).
#include <array>
#include <numeric>
#include "testing.hpp"

BOOST_AUTO_TEST_CASE(empty_constructor)                   // multivector_create.cpp:7-17
{
    vex::multivector<double, 3> m;
    BOOST_CHECK(0U == m.size());
    BOOST_CHECK(m.end() - m.begin() == 0);
    for (int i = 0; i < 3; ++i) BOOST_CHECK(0U == m(i).size());
}

BOOST_AUTO_TEST_CASE(copy_constructor)                    // :19-35
{
    typedef std::array<double, 3> elem_t;
    const size_t n = 1024;
    vex::multivector<double, 3> m(ctx, n);
    m(0) = 1; m(1) = 2; m(2) = 3;
    vex::multivector<double, 3> c(m);
    BOOST_CHECK(c.size() == m.size());
    check_sample(c, m, [](size_t, elem_t a, elem_t b) { BOOST_CHECK(a == b); });
}

BOOST_AUTO_TEST_CASE(element_access)                      // :38-65
{
    const size_t n = 256, m = 4;
    typedef std::array<double, m> elem_t;
    std::vector<double> host = random_vector<double>(n * m);
    vex::multivector<double, m> x(ctx, n);
    copy(host, x);
    for (size_t i = 0; i < n; ++i) {
        elem_t val = x[i];
        for (size_t j = 0; j < m; ++j) { BOOST_CHECK(val[j] == host[j * n + i]); val[j] = 0; }
        x[i] = val;
    }
    copy(x, host);
    BOOST_CHECK(0 == *std::min_element(host.begin(), host.end()));
    BOOST_CHECK(0 == *std::max_element(host.begin(), host.end()));
}

BOOST_AUTO_TEST_CASE(stl_container_of_multivector)        // :67-91
{
    const size_t N = 1024, D = 2, M = 16 + generator<size_t>::get();
    std::vector<vex::multivector<unsigned, D>> x;
    std::vector<std::array<const void*, D>> bufs;
    for (size_t i = 0; i < M; ++i) {
        x.push_back(vex::multivector<unsigned, D>(ctx, N));
        x.back() = i;
        std::array<const void*, D> b = {{x.back()(0)(0).raw(), x.back()(1)(0).raw()}};
        bufs.push_back(b);
    }
    for (size_t i = 0; i < M; ++i) {
        BOOST_CHECK_EQUAL(bufs[i][0], x[i](0)(0).raw());
        BOOST_CHECK_EQUAL(bufs[i][1], x[i](1)(0).raw());
        BOOST_CHECK_EQUAL(x[i](1)[7], i);
    }
}

BOOST_AUTO_TEST_CASE(arithmetics)                         // multivector_arithmetics.cpp:10-35
{
    typedef std::array<double, 4> elem_t;
    const size_t n = 1024;
    vex::multivector<double, 4> x(ctx, n);
    vex::multivector<double, 4> y(ctx, random_vector<double>(n * 4));
    vex::multivector<double, 4> z(ctx, random_vector<double>(n * 4));
    vex::Reductor<double, vex::MIN> min(ctx);
    vex::Reductor<double, vex::MAX> max(ctx);
    elem_t v = {{0, 1, 2, 3}};
    x = v;
    BOOST_CHECK(min(x) == v);
    BOOST_CHECK(max(x) == v);
    x = std::make_tuple(1, 2, 3, 4) * y + z;
    check_sample(x, y, z, [](size_t, elem_t a, elem_t b, elem_t c) {
        for (size_t i = 0; i < 4; ++i) BOOST_CHECK(a[i] == (i + 1) * b[i] + c[i]);           // no contraction: exact
    });
}

BOOST_AUTO_TEST_CASE(multivector_multiexpressions)        // :37-55
{
    typedef std::array<double, 2> elem_t;
    const size_t n = 1024;
    vex::multivector<double, 2> x(ctx, n);
    vex::multivector<double, 2> y(ctx, random_vector<double>(n * 2));
    x = std::tie(sin(y(0)) + cos(y(1)), cos(y(0)) + sin(y(1)));
    check_sample(x, y, [](size_t, elem_t a, elem_t b) {
        BOOST_CHECK_CLOSE(a[0], sin(b[0]) + cos(b[1]), 1e-8);
        BOOST_CHECK_CLOSE(a[1], cos(b[0]) + sin(b[1]), 1e-8);
    });
}

BOOST_AUTO_TEST_CASE(tied_vectors)                        // :57-76
{
    const size_t n = 1024;
    vex::vector<double> X(ctx, random_vector<double>(n)), Y(ctx, random_vector<double>(n));
    vex::vector<double> A(ctx, n), B(ctx, n);
    vex::tie(A, B) = std::tie(X + Y, X - Y);
    check_sample(A, X, Y, [](size_t, double a, double x, double y) { BOOST_CHECK(a == x + y); });
    check_sample(B, X, Y, [](size_t, double b, double x, double y) { BOOST_CHECK(b == x - y); });
}

BOOST_AUTO_TEST_CASE(tied_vectors_read_before_write)      // operations.hpp:2238: vex::tie(x,y) = std::make_tuple(x + y, y - x)
{
    const size_t n = 1024;
    std::vector<double> x = random_vector<double>(n), y = random_vector<double>(n);
    vex::vector<double> X(ctx, x), Y(ctx, y);
    vex::tie(X, Y) = std::make_tuple(X + Y, Y - X);
    std::vector<double> rx(n), ry(n);
    copy(X, rx); copy(Y, ry);
    for (size_t i = 0; i < n; ++i) { BOOST_CHECK(rx[i] == x[i] + y[i]); BOOST_CHECK(ry[i] == y[i] - x[i]); }
    // the same through a multivector whose components feed each other
    vex::multivector<double, 3> m(ctx, n);
    m(0) = X; m(1) = Y; m(2) = 1;
    m = std::tie(m(1), m(2) + m(0), m(0) * m(1));
    std::vector<double> h(3 * n);
    copy(m, h);
    for (size_t i = 0; i < n; ++i) {
        BOOST_CHECK(h[i] == ry[i]); BOOST_CHECK(h[n + i] == 1 + rx[i]); BOOST_CHECK(h[2 * n + i] == rx[i] * ry[i]);
    }
}

BOOST_AUTO_TEST_CASE(multiexpression_in_one_kernel)       // assign_multiexpression, operations.hpp:2081-2185: one launch per device
{
    const size_t n = 4096;
    std::vector<double> x = random_vector<double>(n), y = random_vector<double>(n);
    vex::vector<double> X(ctx, x), Y(ctx, y);
    vexb_set_param("eval.jit", 1);                        // compile at once, so that the first assignment is already fused
    uint64_t l0 = 0, l1 = 0;
    vexb_launch_count(&l0);
    vex::tie(X, Y) = std::make_tuple(X + Y, Y - X);       // reads of element i before its writes: no temporaries
    vexb_launch_count(&l1);
    vexb_set_param("eval.jit", 2);
    BOOST_CHECK(l1 - l0 == ctx.size());
    std::vector<double> rx(n), ry(n);
    copy(X, rx); copy(Y, ry);
    for (size_t i = 0; i < n; ++i) { BOOST_CHECK(rx[i] == x[i] + y[i]); BOOST_CHECK(ry[i] == y[i] - x[i]); }
    // compound assignment on three components, mixed functions
    vex::multivector<double, 3> m(ctx, n);
    m(0) = X; m(1) = Y; m(2) = 0.5;
    vexb_set_param("eval.jit", 1);
    vexb_launch_count(&l0);
    m += std::tie(sin(m(1)) * 2, m(0) * m(2), m(0) - m(1));
    vexb_launch_count(&l1);
    vexb_set_param("eval.jit", 2);
    BOOST_CHECK(l1 - l0 == ctx.size());
    std::vector<double> h(3 * n);
    copy(m, h);
    for (size_t i = 0; i < n; ++i) {
        BOOST_CHECK_CLOSE(h[i], rx[i] + sin(ry[i]) * 2, 1e-12);
        BOOST_CHECK(h[n + i] == ry[i] + rx[i] * 0.5);
        BOOST_CHECK(h[2 * n + i] == 0.5 + (rx[i] - ry[i]));
    }
}

BOOST_AUTO_TEST_CASE(builtin_functions)                   // :78-92
{
    typedef std::array<double, 2> elem_t;
    const size_t n = 1024;
    vex::multivector<double, 2> x(ctx, random_vector<double>(n * 2));
    vex::multivector<double, 2> y(ctx, n);
    y = pow(sin(x), 2.0) + pow(cos(x), 2.0);
    check_sample(y, [](size_t, elem_t a) { BOOST_CHECK_CLOSE(a[0], 1, 1e-8); BOOST_CHECK_CLOSE(a[1], 1, 1e-8); });
}

BOOST_AUTO_TEST_CASE(user_defined_functions)              // :94-118
{
    typedef std::array<double, 2> elem_t;
    const size_t n = 1024, m = 2;
    vex::multivector<double, m> x(ctx, n), y(ctx, n);
    elem_t v1 = {{1, 2}}, v2 = {{2, 1}};
    VEX_FUNCTION(size_t, greater, (double, x)(double, y), return x > y;);
    x = v1; y = v2;
    x = greater(x, y);
    check_sample(x, [](size_t, elem_t a) { BOOST_CHECK(a[0] == 0); BOOST_CHECK(a[1] == 1); });
}

BOOST_AUTO_TEST_CASE(reduction)                           // :120-149
{
    typedef std::array<double, 2> elem_t;
    const size_t n = 1024;
    std::vector<double> x = random_vector<double>(n), y = random_vector<double>(n);
    vex::multivector<double, 2> m(ctx, n);
    copy(x, m(0)); copy(y, m(1));
    vex::Reductor<double, vex::SUM> sum(ctx);
    vex::Reductor<double, vex::MIN> min(ctx);
    vex::Reductor<double, vex::MAX> max(ctx);
    elem_t summ = sum(m), minm = min(m), maxm = max(m);
    BOOST_CHECK_CLOSE(summ[0], std::accumulate(x.begin(), x.end(), 0.0), 1e-6);
    BOOST_CHECK_CLOSE(summ[1], std::accumulate(y.begin(), y.end(), 0.0), 1e-6);
    BOOST_CHECK(minm[0] == *std::min_element(x.begin(), x.end()));
    BOOST_CHECK(minm[1] == *std::min_element(y.begin(), y.end()));
    BOOST_CHECK(maxm[0] == *std::max_element(x.begin(), x.end()));
    BOOST_CHECK(maxm[1] == *std::max_element(y.begin(), y.end()));
    elem_t dot = sum(m * std::make_tuple(2, 3));
    BOOST_CHECK_CLOSE(dot[0], 2 * summ[0], 1e-10);
    BOOST_CHECK_CLOSE(dot[1], 3 * summ[1], 1e-10);
}

BOOST_AUTO_TEST_CASE(element_index)                       // :151-175
{
    typedef std::array<double, 2> elem_t;
    const size_t N = 1024;
    vex::multivector<double, 2> x(ctx, N);
    x = 0.5 * vex::element_index();
    check_sample(x, [](size_t idx, elem_t a) { BOOST_CHECK(a[0] == 0.5 * idx); BOOST_CHECK(a[1] == 0.5 * idx); });
    x = std::tie(sin(0.5 * vex::element_index()), cos(0.5 * vex::element_index()));
    check_sample(x, [](size_t idx, elem_t a) {
        BOOST_CHECK_CLOSE(a[0], sin(0.5 * idx), 1e-6);
        BOOST_CHECK_CLOSE(a[1], cos(0.5 * idx), 1e-6);
    });
}

BOOST_AUTO_TEST_CASE(compound_assignment)                 // :177-211
{
    const size_t n = 1024, m = 2;
    typedef std::array<double, m> elem_t;
    vex::multivector<double, m> x(ctx, n);
    vex::multivector<double, m> y(ctx, random_vector<double>(n * m));
    x = 0;
    x += sin(2 * y);
    check_sample(x, y, [&](size_t, elem_t a, elem_t b) { for (size_t i = 0; i < m; ++i) BOOST_CHECK_CLOSE(a[i], sin(2 * b[i]), 1e-8); });
    x = 0;
    x -= sin(2 * y);
    check_sample(x, y, [&](size_t, elem_t a, elem_t b) { for (size_t i = 0; i < m; ++i) BOOST_CHECK_CLOSE(a[i], -sin(2 * b[i]), 1e-8); });
    x = 1;
    x *= std::tie(y(1), sin(y(0)));
    check_sample(x, y, [](size_t, elem_t a, elem_t b) {
        BOOST_CHECK_CLOSE(a[0], b[1], 1e-8);
        BOOST_CHECK_CLOSE(a[1], sin(b[0]), 1e-8);
    });
}

BOOST_AUTO_TEST_CASE(expression_size_check)               // :234-241
{
    vex::multivector<int, 2> x(ctx, 16), y(ctx, 32);
    BOOST_CHECK_THROW(x = y + 1, std::runtime_error);
}

template <class RT, class CT>
static double row_sum(const std::vector<RT> &row, const std::vector<CT> &col, const std::vector<double> &val,
                      const double *x, size_t idx) {
    double sum = 0;
    for (size_t j = row[idx]; j < row[idx + 1]; j++) sum += val[j] * x[col[j]];
    return sum;
}

BOOST_AUTO_TEST_CASE(multivector_product)                 // spmv.cpp:262-307
{
    const size_t n = 1024, m = 2;
    typedef std::array<double, m> elem_t;
    std::vector<size_t> row, col; std::vector<double> val;
    random_matrix(n, n, 16, row, col, val);
    std::vector<double> x = random_vector<double>(n * m);
    vex::SpMat<double> A(ctx, n, n, row.data(), col.data(), val.data());
    vex::multivector<double, m> X(ctx, x), Y(ctx, n);
    Y = A * X;
    check_sample(Y, [&](size_t idx, elem_t a) {
        BOOST_CHECK_CLOSE(a[0], row_sum(row, col, val, x.data(), idx), 1e-8);
        BOOST_CHECK_CLOSE(a[1], row_sum(row, col, val, x.data() + n, idx), 1e-8);
    });
    Y = X + A * X;
    check_sample(Y, [&](size_t idx, elem_t a) {
        BOOST_CHECK_CLOSE(a[0], x[idx] + row_sum(row, col, val, x.data(), idx), 1e-8);
        BOOST_CHECK_CLOSE(a[1], x[n + idx] + row_sum(row, col, val, x.data() + n, idx), 1e-8);
    });
    Y -= 2 * (A * X);
    check_sample(Y, [&](size_t idx, elem_t a) {
        BOOST_CHECK_CLOSE(a[0], x[idx] - row_sum(row, col, val, x.data(), idx), 1e-6);
        BOOST_CHECK_CLOSE(a[1], x[n + idx] - row_sum(row, col, val, x.data() + n, idx), 1e-6);
    });
}

// SpMat * multivector<double,4> on one slice: all four products come out of ONE pass over the matrix (hell_multi_kernel),
// with the same bits as four single products.
BOOST_AUTO_TEST_CASE(multivector_product_reads_the_matrix_once)
{
    const size_t n = 4096, m = 4;
    std::vector<vex::command_queue> queue(1, ctx.queue(0));
    std::vector<size_t> row, col; std::vector<double> val;
    random_matrix(n, n, 12, row, col, val);
    std::vector<double> x = random_vector<double>(n * m);
    vex::SpMat<double> A(queue, n, n, row.data(), col.data(), val.data());
    vex::multivector<double, m> X(queue, x), Y(queue, n), Z(queue, n);
    uint64_t l0 = 0, l1 = 0;
    vexb_launch_count(&l0);
    Y = A * X;
    vexb_launch_count(&l1);
    if (A.info().loc.fmt == VEXB_FMT_HELL) BOOST_CHECK_EQUAL(l1 - l0, 1u);
    for (size_t i = 0; i < m; ++i) Z(i) = A * X(i);
    std::vector<double> y(n * m), z(n * m);
    vex::copy(Y, y); vex::copy(Z, z);
    size_t diff = 0;
    for (size_t k = 0; k < n * m; ++k) diff += y[k] != z[k];
    BOOST_CHECK_EQUAL(diff, 0u);
    Y += 0.5 * (A * X);
    for (size_t i = 0; i < m; ++i) Z(i) += 0.5 * (A * X(i));
    vex::copy(Y, y); vex::copy(Z, z);
    for (size_t k = 0; k < n * m; ++k) diff += y[k] != z[k];
    BOOST_CHECK_EQUAL(diff, 0u);
}

BOOST_AUTO_TEST_CASE(inline_multivector_product)          // spmv.cpp:309-343
{
    const size_t n = 1024, m = 2;
    typedef std::array<double, m> elem_t;
    std::vector<vex::command_queue> queue(1, ctx.queue(0));
    std::vector<size_t> row, col; std::vector<double> val;
    random_matrix(n, n, 16, row, col, val);
    std::vector<double> x = random_vector<double>(n * m);
    vex::SpMat<double> A(queue, n, n, row.data(), col.data(), val.data());
    vex::multivector<double, m> X(queue, x), Y(queue, n);
    Y = cos(vex::make_inline(A * X));
    check_sample(Y, [&](size_t idx, elem_t a) {
        BOOST_CHECK_CLOSE(a[0], cos(row_sum(row, col, val, x.data(), idx)), 1e-8);
        BOOST_CHECK_CLOSE(a[1], cos(row_sum(row, col, val, x.data() + n, idx)), 1e-8);
    });
}

// tests/spmv.cpp:148-231 (ccsr_vector_product) and :345-437 (ccsr_multivector_product): 3-D Poisson, two unique rows.
static void poisson_ccsr(size_t n, std::vector<size_t> &idx, std::vector<size_t> &row, std::vector<int> &col, std::vector<double> &val) {
    const double h2i = (n - 1) * (n - 1);
    row = {0, 1, 8};
    col = {0, -static_cast<int>(n * n), -static_cast<int>(n), -1, 0, 1, static_cast<int>(n), static_cast<int>(n * n)};
    val = {1, -h2i, -h2i, -h2i, 6 * h2i, -h2i, -h2i, -h2i};
    idx.clear();
    for (size_t k = 0; k < n; k++) for (size_t j = 0; j < n; j++) for (size_t i = 0; i < n; i++)
        idx.push_back((i == 0 || i + 1 == n || j == 0 || j + 1 == n || k == 0 || k + 1 == n) ? 0 : 1);
}
static double ccsr_row(const std::vector<size_t> &idx, const std::vector<size_t> &row, const std::vector<int> &col,
                       const std::vector<double> &val, const double *x, size_t ii) {
    double sum = 0;
    for (size_t j = row[idx[ii]]; j < row[idx[ii] + 1]; j++) sum += val[j] * x[ii + col[j]];
    return sum;
}

BOOST_AUTO_TEST_CASE(ccsr_vector_product)
{
    const size_t n = 32, N = n * n * n;
    std::vector<size_t> idx, row; std::vector<int> col; std::vector<double> val;
    poisson_ccsr(n, idx, row, col, val);
    std::vector<double> x = random_vector<double>(N);
    std::vector<vex::command_queue> queue(1, ctx.queue(0));
    vex::SpMatCCSR<double, int> A(queue[0], N, row.size() - 1, idx.data(), row.data(), col.data(), val.data());
    vex::vector<double> X(queue, x), Y(queue, N);
    Y = A * X;
    check_sample(Y, [&](size_t ii, double a) { BOOST_CHECK(a == ccsr_row(idx, row, col, val, x.data(), ii)); });
    Y = X + A * X;
    check_sample(Y, [&](size_t ii, double a) { BOOST_CHECK_CLOSE(a, x[ii] + ccsr_row(idx, row, col, val, x.data(), ii), 1e-8); });
}

BOOST_AUTO_TEST_CASE(ccsr_multivector_product)
{
    const size_t n = 32, N = n * n * n;
    typedef std::array<double, 2> elem_t;
    std::vector<size_t> idx, row; std::vector<int> col; std::vector<double> val;
    poisson_ccsr(n, idx, row, col, val);
    std::vector<double> x = random_vector<double>(N * 2);
    std::vector<vex::command_queue> queue(1, ctx.queue(0));
    vex::SpMatCCSR<double, int> A(queue[0], N, row.size() - 1, idx.data(), row.data(), col.data(), val.data());
    vex::multivector<double, 2> X(queue, x), Y(queue, N);
    Y = A * X;
    check_sample(Y, [&](size_t ii, elem_t a) {
        BOOST_CHECK(a[0] == ccsr_row(idx, row, col, val, x.data(), ii));
        BOOST_CHECK(a[1] == ccsr_row(idx, row, col, val, x.data() + N, ii));
    });
    Y = X + A * X;
    check_sample(Y, [&](size_t ii, elem_t a) {
        BOOST_CHECK_CLOSE(a[0], x[ii] + ccsr_row(idx, row, col, val, x.data(), ii), 1e-8);
        BOOST_CHECK_CLOSE(a[1], x[N + ii] + ccsr_row(idx, row, col, val, x.data() + N, ii), 1e-8);
    });
}
