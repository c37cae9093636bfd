// The cases of the reference's tests/sparse_matrices.cpp that sit on the hot path (csr :66, ell :95,
// matrix :124, distributed :153-193, distributed_single :195) against include/vexcl/sparse.
#include "testing.hpp"
#include <vexcl/sparse/matrix.hpp>
#include <vexcl/sparse/distributed.hpp>

template <class M>
static void single_device_case()
{
    const size_t n = 1024;
    std::vector<vex::command_queue> q(1, ctx.queue(0));
    std::vector<int> row, col; std::vector<double> val;
    random_matrix(n, n, 16, row, col, val);
    std::vector<double> x = random_vector<double>(n);
    M A(q, n, n, row, col, val);
    vex::vector<double> X(q, x), Y(q, n);
    Y = A * X;
    check_sample(Y, [&](size_t idx, double a) {
        double sum = 0;
        for (int j = row[idx]; j < row[idx + 1]; j++) sum += val[j] * x[col[j]];
        BOOST_CHECK_CLOSE(a, sum, 1e-8);
    });
    // the product is a terminal: it composes with any vector expression and with reductions
    Y = X + 2 * (A * X);
    check_sample(Y, [&](size_t idx, double a) {
        double sum = 0;
        for (int j = row[idx]; j < row[idx + 1]; j++) sum += val[j] * x[col[j]];
        BOOST_CHECK_CLOSE(a, x[idx] + 2 * sum, 1e-8);
    });
    Y = A * (2 * X + 1);
    check_sample(Y, [&](size_t idx, double a) {
        double sum = 0;
        for (int j = row[idx]; j < row[idx + 1]; j++) sum += val[j] * (2 * x[col[j]] + 1);
        BOOST_CHECK_CLOSE(a, sum, 1e-8);
    });
    vex::Reductor<double, vex::SUM> sum(q);
    Y = A * X;
    std::vector<double> y(n); vex::copy(Y, y);
    double ref = 0; for (size_t i = 0; i < n; ++i) ref += x[i] - y[i];
    BOOST_CHECK_CLOSE(sum(X - A * X), ref, 1e-6);
}

BOOST_AUTO_TEST_CASE(csr)    { single_device_case<vex::sparse::csr<double>>(); }
BOOST_AUTO_TEST_CASE(ell)    { single_device_case<vex::sparse::ell<double>>(); }
BOOST_AUTO_TEST_CASE(matrix) { single_device_case<vex::sparse::matrix<double>>(); }

static void tridiagonal(int n, std::vector<int> &ptr, std::vector<int> &col, std::vector<double> &val) {
    ptr.push_back(0);
    for (int i = 0; i < n; ++i) {
        if (i > 0) { col.push_back(i - 1); val.push_back(-1); }
        col.push_back(i); val.push_back(2);
        if (i + 1 < n) { col.push_back(i + 1); val.push_back(-1); }
        ptr.push_back(static_cast<int>(col.size()));
    }
}

template <class Q>
static void distributed_case(const Q &queues)
{
    const int n = 1024;
    std::vector<int> ptr, col; std::vector<double> val;
    tridiagonal(n, ptr, col, val);
    vex::sparse::distributed<vex::sparse::ell<double>> A(queues, n, n, ptr, col, val);
    std::vector<double> x = random_vector<double>(n);
    vex::vector<double> X(queues, x), Y(queues, n);
    Y = A * X;
    for (int i = 0; i < n; ++i) {                          // all rows, as the reference does
        double y = Y[i], sum = 0;
        for (int j = ptr[i]; j < ptr[i + 1]; j++) sum += val[j] * x[col[j]];
        BOOST_CHECK_CLOSE(y, sum, 1e-8);
    }
    vex::sparse::distributed<vex::sparse::csr<double>> B(queues, n, n, ptr, col, val);
    Y = X - B * X;
    for (int i = 0; i < n; i += 37) {
        double y = Y[i], sum = 0;
        for (int j = ptr[i]; j < ptr[i + 1]; j++) sum += val[j] * x[col[j]];
        BOOST_CHECK_CLOSE(y, x[i] - sum, 1e-8);
    }
}

BOOST_AUTO_TEST_CASE(distributed) { distributed_case(ctx.queue()); }
BOOST_AUTO_TEST_CASE(distributed_single) { distributed_case(std::vector<vex::command_queue>(1, ctx.queue(0))); }
