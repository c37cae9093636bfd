// Stencil convolution through include/vexcl.  The cases are those of the reference's tests/stencil.cpp
// (stencil_convolution :18-56, two_stencils :60-75, small_vector :78-107, multivector :109-154, big_stencil :156-181)
// written table-driven, plus exact (bit-for-bit) checks, halos longer than a neighbouring slice and single precision.
// user_defined_stencil (:183-217) is the last case; its NVRTC path is opt-in until it has been seen on a GPU.
#include <array>
#include "testing.hpp"

namespace {

// x[clamp(i + shift)] as the reference's test helper does (tests/stencil.cpp:8-16)
template <class T>
T clamped(const T *x, size_t n, size_t i, long shift) {
    const long j = static_cast<long>(i) + shift;
    return x[j < 0 ? 0 : (static_cast<size_t>(j) >= n ? n - 1 : static_cast<size_t>(j))];
}

// init + sum_k s[k] * x[clamp(i + k - center)], taps in order
template <class T>
T convolve_at(const std::vector<T> &s, int center, const T *x, size_t n, size_t i, T init) {
    T acc = init;
    for (size_t k = 0; k < s.size(); ++k) acc += s[k] * clamped(x, n, i, static_cast<long>(k) - center);
    return acc;
}

struct shape { size_t n; size_t width; int center; };          // width 0: random in [1, 64]; center < 0: random

void run_vector_case(const shape &sh, bool exact_everywhere) {
    const size_t width = sh.width ? sh.width : static_cast<size_t>(rand() % 64 + 1);
    const int center = sh.center >= 0 ? sh.center : static_cast<int>(rand() % width);
    const std::vector<double> taps = random_vector<double>(width), host = random_vector<double>(sh.n);
    vex::stencil<double> S(ctx, taps, center);
    vex::vector<double> X(ctx, host), Y(ctx, sh.n);

    Y = 1;
    Y += X * S;                                                                      // stencil.cpp:32-33
    check_sample(Y, [&](size_t i, double got) { BOOST_CHECK_CLOSE(got, convolve_at(taps, center, host.data(), sh.n, i, 1.0), 1e-8); });
    Y = 42 * (X * S);                                                                // stencil.cpp:47
    check_sample(Y, [&](size_t i, double got) { BOOST_CHECK_CLOSE(got, 42 * convolve_at(taps, center, host.data(), sh.n, i, 0.0), 1e-8); });
    if (exact_everywhere) {
        Y = S * X;
        std::vector<double> back(sh.n);
        copy(Y, back);
        for (size_t i = 0; i < sh.n; ++i) BOOST_CHECK(back[i] == convolve_at(taps, center, host.data(), sh.n, i, 0.0));
    }
}

} // namespace

BOOST_AUTO_TEST_CASE(stencil_convolution) { run_vector_case({1024, 0, -1}, true); }

BOOST_AUTO_TEST_CASE(two_stencils)
{
    vex::stencil<double> S(ctx, std::vector<double>(5, 1.0), 3);
    vex::vector<double> X(ctx, 32), Y(ctx, 32);
    for (double fill : {0.0, 1.0}) {
        X = fill;
        Y = X * S + X * S;
        for (size_t i : {0u, 16u, 31u}) BOOST_CHECK(Y[i] == 10 * fill);
    }
}

BOOST_AUTO_TEST_CASE(small_vector) { run_vector_case({128, 0, -1}, true); }

BOOST_AUTO_TEST_CASE(tiny_vectors_and_wide_stencils)      // halos longer than the neighbouring slices
{
    for (size_t n : {1u, 2u, 17u, 33u})
        for (int center : {0, 20, 40}) run_vector_case({n, 41, center}, true);
}

BOOST_AUTO_TEST_CASE(multivector)
{
    typedef std::array<double, 2> pair_t;
    const size_t n = 1024;
    const std::vector<double> taps = random_vector<double>(rand() % 64 + 1);
    const int center = rand() % taps.size();
    vex::stencil<double> S(ctx, taps.begin(), taps.end(), center);
    const std::vector<double> host = random_vector<double>(2 * n);
    vex::multivector<double, 2> X(ctx, host), Y(ctx, n);
    const double scale[] = {1, 42}, init[] = {1, 0};
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 0) { Y = 1; Y += X * S; } else Y = 42 * (X * S);
        check_sample(Y, [&](size_t i, pair_t got) {
            for (size_t c = 0; c < 2; ++c)
                BOOST_CHECK_CLOSE(got[c], scale[pass] * convolve_at(taps, center, host.data() + c * n, n, i, init[pass]), 1e-8);
        });
    }
}

BOOST_AUTO_TEST_CASE(big_stencil) { run_vector_case({size_t(1) << 16, 2048, -1}, false); }

BOOST_AUTO_TEST_CASE(single_precision)
{
    const size_t n = 5000;
    const std::vector<float> taps = random_vector<float>(21), host = random_vector<float>(n);
    std::vector<float> back(n);
    vex::stencil<float> S(ctx, taps, 10);
    vex::vector<float> X(ctx, host), Y(ctx, n);
    Y = X * S;
    copy(Y, back);
    for (size_t i = 0; i < n; ++i) BOOST_CHECK(back[i] == convolve_at(taps, 10, host.data(), n, i, 0.0f));
}

// tests/stencil.cpp:183-217 (user_defined_stencil).  The operator is compiled by NVRTC when first applied.
BOOST_AUTO_TEST_CASE(user_defined_stencil)
{
    const size_t n = 1024;
    VEX_STENCIL_OPERATOR(oscillate, double, 3, 1, "return sin(X[1] - X[0]) + sin(X[0] - X[-1]);", ctx);
    const std::vector<double> host = random_vector<double>(n);
    vex::vector<double> X(ctx, host), Y(ctx, n);
    auto expected = [&](size_t i) {
        return sin(clamped(host.data(), n, i, +1) - host[i]) + sin(host[i] - clamped(host.data(), n, i, -1));
    };
    Y = oscillate(X);
    check_sample(Y, [&](size_t i, double got) { BOOST_CHECK_CLOSE(got, expected(i), 1e-8); });
    Y = 41 * oscillate(X) + oscillate(X);
    check_sample(Y, [&](size_t i, double got) { BOOST_CHECK_CLOSE(got, 42 * expected(i), 1e-8); });
}
