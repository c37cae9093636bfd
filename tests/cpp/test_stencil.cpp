// The reference's tests/stencil.cpp (stencil_convolution :18-56, two_stencils :60-75, small_vector :78-107,
// multivector :109-154, big_stencil :156-181) against include/vexcl.  user_defined_stencil (:183-217) needs
// VEX_STENCIL_OPERATOR, which is not provided.
#include <array>
#include "testing.hpp"

struct clamp_index {
    size_t n;
    clamp_index(size_t n) : n(n) {}
    size_t operator()(size_t i, long shift) const {
        return std::min<size_t>(n - 1, std::max<long>(0, static_cast<long>(i) + shift));
    }
};

static double conv_at(const std::vector<double> &s, int center, const double *x, size_t n, size_t i, double sum) {
    clamp_index idx(n);
    int k = -center;
    for (size_t j = 0; j < s.size(); k++, j++) sum += s[j] * x[idx(i, k)];
    return sum;
}

BOOST_AUTO_TEST_CASE(stencil_convolution)
{
    const size_t n = 1024;
    std::vector<double> s = random_vector<double>(rand() % 64 + 1);
    int center = rand() % s.size();
    vex::stencil<double> S(ctx, s, center);
    std::vector<double> x = random_vector<double>(n);
    vex::vector<double> X(ctx, x), Y(ctx, n);
    Y = 1;
    Y += X * S;
    check_sample(Y, [&](size_t i, double a) { BOOST_CHECK_CLOSE(a, conv_at(s, center, x.data(), n, i, 1), 1e-8); });
    Y = 42 * (X * S);
    check_sample(Y, [&](size_t i, double a) { BOOST_CHECK_CLOSE(a, 42 * conv_at(s, center, x.data(), n, i, 0), 1e-8); });
    Y = S * X;                                            // every element, and exactly: tap-order sums without contraction
    std::vector<double> y(n);
    copy(Y, y);
    for (size_t i = 0; i < n; ++i) BOOST_CHECK(y[i] == conv_at(s, center, x.data(), n, i, 0));
}

BOOST_AUTO_TEST_CASE(two_stencils)
{
    const size_t n = 32;
    std::vector<double> s(5, 1);
    vex::stencil<double> S(ctx, s, 3);
    vex::vector<double> X(ctx, n), Y(ctx, n);
    X = 0;
    Y = X * S + X * S;
    BOOST_CHECK(Y[ 0] == 0);
    BOOST_CHECK(Y[16] == 0);
    BOOST_CHECK(Y[31] == 0);
    X = 1;
    Y = X * S + X * S;
    BOOST_CHECK(Y[ 0] == 10);
    BOOST_CHECK(Y[16] == 10);
    BOOST_CHECK(Y[31] == 10);
}

BOOST_AUTO_TEST_CASE(small_vector)
{
    const size_t n = 128;
    std::vector<double> s = random_vector<double>(rand() % 64 + 1);
    int center = rand() % s.size();
    vex::stencil<double> S(ctx, s, center);
    std::vector<double> x = random_vector<double>(n);
    vex::vector<double> X(ctx, x), Y(ctx, n);
    Y = 1;
    Y += X * S;
    check_sample(Y, [&](size_t i, double a) { BOOST_CHECK_CLOSE(a, conv_at(s, center, x.data(), n, i, 1), 1e-8); });
}

BOOST_AUTO_TEST_CASE(tiny_vectors_and_wide_stencils)      // halos longer than the neighbouring slices
{
    for (size_t n : {1u, 2u, 17u, 33u}) {
        std::vector<double> s = random_vector<double>(41);
        for (int center : {0, 20, 40}) {
            vex::stencil<double> S(ctx, s, center);
            std::vector<double> x = random_vector<double>(n), y(n);
            vex::vector<double> X(ctx, x), Y(ctx, n);
            Y = X * S;
            copy(Y, y);
            for (size_t i = 0; i < n; ++i) BOOST_CHECK(y[i] == conv_at(s, center, x.data(), n, i, 0));
        }
    }
}

BOOST_AUTO_TEST_CASE(multivector)
{
    typedef std::array<double, 2> elem_t;
    const size_t n = 1024;
    std::vector<double> s = random_vector<double>(rand() % 64 + 1);
    int center = rand() % s.size();
    vex::stencil<double> S(ctx, s.begin(), s.end(), center);
    std::vector<double> x = random_vector<double>(2 * n);
    vex::multivector<double, 2> X(ctx, x), Y(ctx, n);
    Y = 1;
    Y += X * S;
    check_sample(Y, [&](size_t i, elem_t a) {
        BOOST_CHECK_CLOSE(a[0], conv_at(s, center, x.data(), n, i, 1), 1e-8);
        BOOST_CHECK_CLOSE(a[1], conv_at(s, center, x.data() + n, n, i, 1), 1e-8);
    });
    Y = 42 * (X * S);
    check_sample(Y, [&](size_t i, elem_t a) {
        BOOST_CHECK_CLOSE(a[0], 42 * conv_at(s, center, x.data(), n, i, 0), 1e-8);
        BOOST_CHECK_CLOSE(a[1], 42 * conv_at(s, center, x.data() + n, n, i, 0), 1e-8);
    });
}

BOOST_AUTO_TEST_CASE(big_stencil)
{
    const size_t n = 1 << 16;
    std::vector<double> s = random_vector<double>(2048);
    int center = rand() % s.size();
    vex::stencil<double> S(ctx, s, center);
    std::vector<double> x = random_vector<double>(n);
    vex::vector<double> X(ctx, x), Y(ctx, n);
    Y = X * S;
    check_sample(Y, [&](size_t i, double a) { BOOST_CHECK_CLOSE(a, conv_at(s, center, x.data(), n, i, 0), 1e-8); });
}

BOOST_AUTO_TEST_CASE(single_precision)
{
    const size_t n = 5000;
    std::vector<float> s = random_vector<float>(21), x = random_vector<float>(n), y(n);
    vex::stencil<float> S(ctx, s, 10);
    vex::vector<float> X(ctx, x), Y(ctx, n);
    Y = X * S;
    copy(Y, y);
    clamp_index idx(n);
    for (size_t i = 0; i < n; ++i) {
        float sum = 0;
        for (int k = 0; k < 21; ++k) sum += s[k] * x[idx(i, k - 10)];
        BOOST_CHECK(y[i] == sum);
    }
}
