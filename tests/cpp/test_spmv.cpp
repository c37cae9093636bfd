// Hot-path cases of the reference's tests/spmv.cpp (vector_product :10-59, non_square_matrix :61-87,
// non_default_types :89-114, empty_rows :116-146) against include/vexcl.
#include "testing.hpp"

template <class RT, class CT>
static double row_sum(const std::vector<RT> &row, const std::vector<CT> &col, const std::vector<double> &val,
                      const std::vector<double> &x, size_t idx) {
    double sum = 0;
    for (size_t j = row[idx]; j < row[idx + 1]; j++) sum += val[j] * x[col[j]];
    return sum;
}

BOOST_AUTO_TEST_CASE(vector_product)
{
    const size_t n = 1024;
    std::vector<size_t> row, col; std::vector<double> val;
    random_matrix(n, n, 16, row, col, val);
    std::vector<double> x = random_vector<double>(n);
    vex::SpMat<double> A(ctx, n, n, row.data(), col.data(), val.data());
    vex::vector<double> X(ctx, x), Y(ctx, n);
    Y = A * X;
    check_sample(Y, [&](size_t idx, double a) { BOOST_CHECK_CLOSE(a, row_sum(row, col, val, x, idx), 1e-8); });
    Y -= A * X;
    check_sample(Y, [&](size_t, double a) { BOOST_CHECK_SMALL(a, 1e-8); });
    Y += 42 * (A * X);
    check_sample(Y, [&](size_t idx, double a) { BOOST_CHECK_CLOSE(a, 42 * row_sum(row, col, val, x, idx), 1e-8); });
    Y = X + A * X;
    check_sample(Y, [&](size_t idx, double a) { BOOST_CHECK_CLOSE(a, x[idx] + row_sum(row, col, val, x, idx), 1e-8); });
    Y = X - 0.5 * (A * X) + A * X;
    check_sample(Y, [&](size_t idx, double a) { BOOST_CHECK_CLOSE(a, x[idx] + 0.5 * row_sum(row, col, val, x, idx), 1e-8); });
    BOOST_CHECK_EQUAL(A.rows(), n); BOOST_CHECK_EQUAL(A.cols(), n); BOOST_CHECK_EQUAL(A.nonzeros(), row.back());
}

BOOST_AUTO_TEST_CASE(non_square_matrix)
{
    const size_t n = 1024, m = 2 * n;
    std::vector<size_t> row, col; std::vector<double> val;
    random_matrix(n, m, 16, row, col, val);
    std::vector<double> x = random_vector<double>(m);
    vex::SpMat<double> A(ctx, n, m, row.data(), col.data(), val.data());
    vex::vector<double> X(ctx, x), Y(ctx, n);
    Y = A * X;
    check_sample(Y, [&](size_t idx, double a) { BOOST_CHECK_CLOSE(a, row_sum(row, col, val, x, idx), 1e-8); });
}

BOOST_AUTO_TEST_CASE(non_default_types)
{
    const size_t n = 1024;
    std::vector<unsigned> row; std::vector<int> col; std::vector<double> val;
    random_matrix(n, n, 16, row, col, val);
    std::vector<double> x = random_vector<double>(n);
    vex::SpMat<double, int, unsigned> A(ctx, n, n, row.data(), col.data(), val.data());
    vex::vector<double> X(ctx, x), Y(ctx, n);
    Y = A * X;
    check_sample(Y, [&](size_t idx, double a) { BOOST_CHECK_CLOSE(a, row_sum(row, col, val, x, idx), 1e-8); });
}

BOOST_AUTO_TEST_CASE(empty_rows)
{
    const size_t n = 1024, non_empty_part = 256;
    std::vector<size_t> row, col; std::vector<double> val;
    random_matrix(non_empty_part, n, 16, row, col, val);
    while (row.size() < n + 1) row.push_back(col.size());
    std::vector<double> x = random_vector<double>(n);
    vex::SpMat<double> A(ctx, n, n, row.data(), col.data(), val.data());
    vex::vector<double> X(ctx, x), Y(ctx, n);
    Y = 1;
    Y = A * X;
    check_sample(Y, [&](size_t idx, double a) { BOOST_CHECK_CLOSE(a, row_sum(row, col, val, x, idx), 1e-8); });
    BOOST_CHECK_EQUAL(Y[n - 1], 0.0);
}

BOOST_AUTO_TEST_CASE(inline_spmv)                         // tests/spmv.cpp:233-260
{
    const size_t n = 1024;
    std::vector<vex::command_queue> queue(1, ctx.queue(0));
    std::vector<size_t> row, col; std::vector<double> val;
    random_matrix(n, n, 16, row, col, val);
    std::vector<double> x = random_vector<double>(n);
    vex::SpMat<double> A(queue, n, n, row.data(), col.data(), val.data());
    vex::vector<double> X(queue, x), Y(queue, n);
    Y = sin(vex::make_inline(A * X));
    check_sample(Y, [&](size_t idx, double a) { BOOST_CHECK_CLOSE(a, sin(row_sum(row, col, val, x, idx)), 1e-8); });
    vex::Reductor<double, vex::MAX> max(queue);
    std::vector<double> y(n);
    for (size_t i = 0; i < n; ++i) y[i] = row_sum(row, col, val, x, i);
    vex::vector<double> F(queue, y);
    BOOST_CHECK_SMALL(max(fabs(F - vex::make_inline(A * X))), 1e-10);
}

BOOST_AUTO_TEST_CASE(poisson_benchmark_matrix)          // examples/benchmark.cpp:357-473 at n = 32
{
    const size_t n = 32, N = n * n * n;
    const double h2i = (n - 1) * (n - 1);
    std::vector<size_t> row; std::vector<unsigned> col; std::vector<double> val;
    std::vector<double> X(N, 1e-2), Y(N, 0);
    row.push_back(0);
    for (size_t k = 0, idx = 0; k < n; k++) for (size_t j = 0; j < n; j++) for (size_t i = 0; i < n; i++, idx++) {
        if (i == 0 || i == n - 1 || j == 0 || j == n - 1 || k == 0 || k == n - 1) {
            col.push_back(idx); val.push_back(1);
        } else {
            const long off[] = {-(long)(n * n), -(long)n, -1, 0, 1, (long)n, (long)(n * n)};
            for (int t = 0; t < 7; ++t) { col.push_back(idx + off[t]); val.push_back(t == 3 ? 6 * h2i : -h2i); }
        }
        row.push_back(col.size());
    }
    vex::SpMat<double, unsigned> A(ctx, N, N, row.data(), col.data(), val.data());
    vex::vector<double> x(ctx, X), y(ctx, Y);
    const size_t M = 8;
    for (size_t i = 0; i < M; i++) y += A * x;
    for (size_t k = 0; k < M; k++) for (size_t i = 0; i < N; i++) {
        double s = 0;
        for (size_t j = row[i]; j < row[i + 1]; j++) s += val[j] * X[col[j]];
        Y[i] += s;
    }
    vex::copy(Y, x);
    y -= x;
    vex::Reductor<double, vex::SUM> sum(ctx);
    BOOST_CHECK_SMALL(sum(y * y), 1e-12);               // "res" of the reference benchmark
}
