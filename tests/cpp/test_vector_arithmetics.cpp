// Hot-path cases of the reference's tests/vector_arithmetics.cpp, tests/vector_create.cpp and
// tests/vector_copy.cpp, compiled against include/vexcl (no Boost) and run on the GPU box.
#include "testing.hpp"
#include <numeric>

BOOST_AUTO_TEST_CASE(access_element)                    // vector_arithmetics.cpp:10-31
{
    const size_t N = 5;
    vex::vector<double> x(ctx, N);
    for (size_t i = 0; i < N; ++i) x[i] = static_cast<double>(i);
    BOOST_CHECK_EQUAL(x.at(2), 2.0);
    BOOST_CHECK_EQUAL(x[3], 3.0);
    BOOST_CHECK_THROW(x.at(5), std::out_of_range);
}

BOOST_AUTO_TEST_CASE(assign_expression)                 // :33-48
{
    const size_t N = 1024;
    vex::vector<double> x(ctx, N), y(ctx, N), z(ctx, N);
    y = 42;
    z = 67;
    x = 5 * sin(y) + z;
    check_sample(x, [](size_t, double a) { BOOST_CHECK_CLOSE(a, 5 * sin(42.0) + 67, 1e-12); });
}

BOOST_AUTO_TEST_CASE(compound_assignment)               // :50-64
{
    const size_t n = 1024;
    vex::vector<double> x(ctx, n);
    x = 0;
    x += 1;
    check_sample(x, [](size_t, double a) { BOOST_CHECK(a == 1); });
    x -= 2;
    check_sample(x, [](size_t, double a) { BOOST_CHECK(a == -1); });
    x *= 4;  x /= 8;
    check_sample(x, [](size_t, double a) { BOOST_CHECK(a == -0.5); });
    vex::vector<int> k(ctx, n);
    k = 13;  k %= 5;  k <<= 3;  k |= 1;  k ^= 3;  k &= 0xfe;  k >>= 1;
    check_sample(k, [](size_t, int a) { BOOST_CHECK(a == ((((((13 % 5) << 3) | 1) ^ 3) & 0xfe) >> 1)); });
}

BOOST_AUTO_TEST_CASE(reduce_expression)                 // :66-99
{
    const size_t N = 1024;
    std::vector<double> x = random_vector<double>(N);
    for (auto &v : x) v = (v - 0.5) * 1e8;
    vex::vector<double> X(ctx, x);
    vex::Reductor<double, vex::SUM> sum(ctx);
    vex::Reductor<double, vex::MIN> min(ctx);
    vex::Reductor<double, vex::MAX> max(ctx);
    vex::Reductor<double, vex::SUM_Kahan> csum(ctx);
    double ks = 0, c = 0;                                                  // Kahan accumulator
    for (double v : x) { double y = v - c, t = ks + y; c = (t - ks) - y; ks = t; }
    BOOST_CHECK_CLOSE(sum(X), ks, 1e-8);
    BOOST_CHECK_CLOSE(csum(X), ks, 1e-8);
    BOOST_CHECK_EQUAL(min(X), *std::min_element(x.begin(), x.end()));
    BOOST_CHECK_EQUAL(max(X), *std::max_element(x.begin(), x.end()));
    vex::Reductor<double, vex::MIN_MAX> minmax(ctx);
    auto mm = minmax(X);
    BOOST_CHECK_EQUAL(mm.s[0], *std::min_element(x.begin(), x.end()));
    BOOST_CHECK_EQUAL(mm.s[1], *std::max_element(x.begin(), x.end()));
    BOOST_CHECK_EQUAL(max(fabs(X - X)), 0.0);
}

// vex::CombineReductors (vexcl/reductor.hpp:132-280): several reductions of one expression in a single pass
BOOST_AUTO_TEST_CASE(combined_reductors)
{
    const size_t N = 100003;
    std::vector<double> x = random_vector<double>(N);
    for (auto &v : x) v = (v - 0.25) * 1e3;
    vex::vector<double> X(ctx, x), Y(ctx, N);
    Y = 2 * X;
    const double lo = *std::min_element(x.begin(), x.end()), hi = *std::max_element(x.begin(), x.end());
    double ks = 0, c = 0;
    for (double v : x) { double y = v - c, t = ks + y; c = (t - ks) - y; ks = t; }
    vex::Reductor<double, vex::CombineReductors<vex::MIN, vex::MAX, vex::SUM>> mms(ctx);
    auto r = mms(X);                                                       // three components -> a 4-wide CL-style vector
    static_assert(sizeof(r.s) == 4 * sizeof(double), "three reductors fit cl_double4");
    BOOST_CHECK_EQUAL(r.s[0], lo);
    BOOST_CHECK_EQUAL(r.s[1], hi);
    BOOST_CHECK_CLOSE(r.s[2], ks, 1e-8);
    vex::Reductor<double, vex::CombineReductors<vex::SUM, vex::SUM_Kahan, vex::MAX, vex::MIN, vex::MAX>> five(ctx);
    auto q = five(X + Y);                                                  // an expression, five components -> 8-wide
    static_assert(sizeof(q.s) == 8 * sizeof(double), "five reductors fit cl_double8");
    BOOST_CHECK_CLOSE(q.s[0], 3 * ks, 1e-8);
    BOOST_CHECK_CLOSE(q.s[1], 3 * ks, 1e-8);
    BOOST_CHECK_EQUAL(q.s[2], 3 * hi);
    BOOST_CHECK_EQUAL(q.s[3], 3 * lo);
    BOOST_CHECK_EQUAL(q.s[4], 3 * hi);
    vex::Reductor<int, vex::CombineReductors<vex::SUM, vex::MAX>> cnt(ctx);
    auto k = cnt(X > 0.0);
    size_t pos = 0; for (double v : x) pos += v > 0.0;
    BOOST_CHECK_EQUAL(k.s[0], static_cast<int>(pos));
    BOOST_CHECK_EQUAL(k.s[1], 1);
    vex::vector<double> E(ctx, 0);                                         // empty: the initial values (reductor.hpp:318-321)
    auto e = mms(E);
    BOOST_CHECK_EQUAL(e.s[0], std::numeric_limits<double>::max());
    BOOST_CHECK_EQUAL(e.s[1], std::numeric_limits<double>::lowest());
    BOOST_CHECK_EQUAL(e.s[2], 0.0);
}

BOOST_AUTO_TEST_CASE(builtin_functions)                 // :101-111
{
    const size_t N = 1024;
    std::vector<double> x = random_vector<double>(N);
    vex::vector<double> X(ctx, x), Y(ctx, N);
    Y = pow(sin(X), 2.0) + pow(cos(X), 2.0);
    check_sample(Y, [](size_t, double a) { BOOST_CHECK_CLOSE(a, 1, 1e-8); });
    Y = sqrt(fabs(X)) + exp(-X) * log(X + 1.0);
    check_sample(Y, [&](size_t i, double a) { BOOST_CHECK_CLOSE(a, sqrt(fabs(x[i])) + exp(-x[i]) * log(x[i] + 1.0), 1e-8); });
}

BOOST_AUTO_TEST_CASE(counting_reductions)               // :113-145 (the checks that do not need user functions)
{
    const size_t N = 1024;
    vex::vector<double> x(ctx, N), y(ctx, N);
    x = 1;  y = 2;
    vex::Reductor<size_t, vex::SUM> count(ctx);
    BOOST_CHECK_EQUAL(count(x > y), 0u);
    BOOST_CHECK_EQUAL(count(x < y), N);
    vex::Reductor<double, vex::SUM> sum(ctx);
    BOOST_CHECK_EQUAL(sum(x * 2), 2.0 * N);
    vex::Reductor<size_t, vex::SUM> isum(ctx);
    BOOST_CHECK_EQUAL(isum(vex::element_index(0, N)), N * (N - 1) / 2);
}

VEX_FUNCTION(size_t, greater_fn, (double, x)(double, y), return x > y;);
VEX_FUNCTION(double, times2, (double, x), return x * 2;);
VEX_FUNCTION_V1(sqr_v1, double(double), "return prm1 * prm1;");

BOOST_AUTO_TEST_CASE(user_defined_functions)            // :113-145
{
    const size_t N = 1024;
    vex::vector<double> x(ctx, N), y(ctx, N);
    x = 1;  y = 2;
    vex::Reductor<size_t, vex::SUM> sum(ctx);
    BOOST_CHECK_EQUAL(sum(greater_fn(x, y)), 0u);
    BOOST_CHECK_EQUAL(sum(greater_fn(y, x)), N);
    vex::Reductor<double, vex::SUM> dsum(ctx);
    BOOST_CHECK_EQUAL(dsum(times2(x)), 2.0 * N);
    vex::vector<double> z(ctx, N);
    z = times2(x) + sqr_v1(y) * 0.5;
    check_sample(z, [](size_t, double a) { BOOST_CHECK_EQUAL(a, 2.0 + 4.0 * 0.5); });
    z += times2(z) * greater_fn(y, x);
    check_sample(z, [](size_t, double a) { BOOST_CHECK_EQUAL(a, 12.0); });
}

BOOST_AUTO_TEST_CASE(ternary_operator)                  // :238-252
{
    const size_t n = 1024;
    vex::vector<double> x(ctx, random_vector<double>(n)), y(ctx, n);
    y = if_else(x > 0.5, sin(x), cos(x));
    check_sample(x, y, [](size_t, double X, double Y) { BOOST_CHECK_CLOSE(Y, X > 0.5 ? sin(X) : cos(X), 1e-8); });
}

BOOST_AUTO_TEST_CASE(combine_expressions)               // :271-297
{
    const size_t n = 1024;
    vex::vector<double> x(ctx, n);
    auto alpha  = vex::element_index() * (2 * M_PI / n);
    auto sine   = sin(alpha);
    auto cosine = cos(alpha);
    x = pow(sine, 2.0) + pow(cosine, 2.0);
    check_sample(x, [](size_t, double v) { BOOST_CHECK_CLOSE(v, 1.0, 1e-8); });
    x = vex::element_index() * 2.0 + 42;
    check_sample(x, [](size_t i, double v) { BOOST_CHECK_EQUAL(v, 2.0 * i + 42); });
}

BOOST_AUTO_TEST_CASE(expression_size_check)             // :319-327
{
    vex::vector<int> x(ctx, 16), y(ctx, 32);
    BOOST_CHECK_THROW(x = y, std::runtime_error);
    BOOST_CHECK_THROW(x = 2 * y + 1, std::runtime_error);
}

BOOST_AUTO_TEST_CASE(mixed_types_and_tags)
{
    const size_t n = 4096;
    std::vector<float> f = random_vector<float>(n);
    std::vector<int> k = random_vector<int>(n);
    vex::vector<float> F(ctx, f);
    vex::vector<int> K(ctx, k);
    vex::vector<double> D(ctx, n);
    D = K * F + 0.5;
    check_sample(D, [&](size_t i, double a) { BOOST_CHECK_EQUAL(a, static_cast<double>(static_cast<float>(k[i]) * f[i]) + 0.5); });
    // tagged terminals: benchmark_saxpy's spelling (examples/benchmark.cpp:100-107)
    vex::vector<double> a(ctx, n), b(ctx, random_vector<double>(n));
    a = 1.0;
    auto ta = vex::tag<1>(a);
    ta = 0.5 * ta + b;
    check_sample(a, b, [](size_t, double A, double B) { BOOST_CHECK_EQUAL(A, 0.5 * 1.0 + B); });
}

BOOST_AUTO_TEST_CASE(some_devices_are_empty)            // vector_create.cpp:189-194
{
    vex::vector<double> x(ctx, 1);
    x = 0;
    BOOST_CHECK(x[0] == 0);
}

BOOST_AUTO_TEST_CASE(create_copy_and_iterate)           // vector_create.cpp / vector_copy.cpp basics
{
    const size_t N = 1000;
    std::vector<double> h = random_vector<double>(N);
    vex::vector<double> x(ctx, h);
    vex::vector<double> y = x;                           // copy constructor: allocate + assign
    vex::vector<double> z(x * 2);                        // from an expression
    std::vector<double> back(N);
    vex::copy(y, back);
    BOOST_CHECK(back == h);
    vex::copy(z, back);
    for (size_t i = 0; i < N; i += 97) BOOST_CHECK_EQUAL(back[i], 2 * h[i]);
    vex::copy(x.begin() + 10, x.begin() + 20, back.data());
    for (size_t i = 0; i < 10; ++i) BOOST_CHECK_EQUAL(back[i], h[10 + i]);
    std::vector<double> piece(5, 7.0);
    vex::copy(piece.data(), piece.data() + 5, x.begin() + 3);
    BOOST_CHECK_EQUAL(x[3], 7.0); BOOST_CHECK_EQUAL(x[7], 7.0); BOOST_CHECK_EQUAL(x[8], h[8]);
    BOOST_CHECK_EQUAL(x.size(), N);
    BOOST_CHECK_EQUAL(x.nparts(), ctx.size());
    size_t tot = 0;
    for (unsigned d = 0; d < x.nparts(); ++d) { BOOST_CHECK_EQUAL(x.part_start(d), tot); tot += x.part_size(d); }
    BOOST_CHECK_EQUAL(tot, N);
    BOOST_CHECK(x.partition() == vex::partition(N, ctx.queue()));
    { auto m = x.map(0); m[0] = -1; }
    BOOST_CHECK_EQUAL(x[0], -1.0);
    vex::vector<double> w;
    w.resize(ctx, 10);  w = 3;
    x.swap(w);
    BOOST_CHECK_EQUAL(x.size(), 10u);
    BOOST_CHECK_EQUAL(x[9], 3.0);
}
