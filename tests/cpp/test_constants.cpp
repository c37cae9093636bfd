// vex::constants, std::integral_constant and VEX_CONSTANT in expressions (vexcl/constants.hpp; the reference uses them
// in tests/vector_arithmetics.cpp:270-297 and tests/multivector_arithmetics.cpp:213-232).
#include <array>
#include "testing.hpp"

VEX_CONSTANT(answer, 42);
VEX_CONSTANT(golden, 1.61803398874989484820);

BOOST_AUTO_TEST_CASE(constants_as_whole_right_hand_sides)
{
    const size_t n = 1000;
    vex::vector<double> x(ctx, n);
    x = vex::constants::pi();                                                    // vector_arithmetics.cpp:295
    check_sample(x, [](size_t, double v) { BOOST_CHECK(v == 3.14159265358979323846); });
    x = std::integral_constant<int, 42>();                                       // multivector_arithmetics.cpp:221
    check_sample(x, [](size_t, double v) { BOOST_CHECK(v == 42); });
    x = answer();
    check_sample(x, [](size_t, double v) { BOOST_CHECK(v == 42); });
    x = golden();
    check_sample(x, [](size_t, double v) { BOOST_CHECK(v == 1.61803398874989484820); });
    BOOST_CHECK(static_cast<double>(vex::constants::two_pi) == 2 * 3.14159265358979323846);
    BOOST_CHECK(static_cast<int>(answer) == 42);
    vex::multivector<double, 3> m(ctx, n);
    m = std::integral_constant<int, 7>();
    check_sample(m, [](size_t, std::array<double, 3> v) { for (double c : v) BOOST_CHECK(c == 7); });
}

BOOST_AUTO_TEST_CASE(constants_inside_expressions)
{
    const size_t n = 1000;
    vex::vector<double> x(ctx, n), y(ctx, n);
    y = 0.001 * vex::element_index();
    x = sin(vex::constants::two_pi() * y) + vex::constants::half();              // vector_arithmetics.cpp:277
    check_sample(x, [](size_t i, double v) { BOOST_CHECK_CLOSE(v, sin(2 * 3.14159265358979323846 * (0.001 * i)) + 0.5, 1e-8); });
    x = y * std::integral_constant<int, 3>() + golden() * answer();
    check_sample(x, [](size_t i, double v) { BOOST_CHECK_CLOSE(v, 0.001 * i * 3 + 1.61803398874989484820 * 42, 1e-10); });
    vex::multivector<double, 2> m(ctx, n);
    m = sin(vex::constants::e() * vex::element_index());                         // multivector_arithmetics.cpp:226
    check_sample(m, [](size_t i, std::array<double, 2> v) {
        for (double c : v) BOOST_CHECK_CLOSE(c, sin(2.71828182845904523536 * i), 1e-8);
    });
}
