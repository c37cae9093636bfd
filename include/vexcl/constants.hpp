#ifndef VEXCL_CONSTANTS_HPP
#define VEXCL_CONSTANTS_HPP
/*
 * Constants in vector expressions (vexcl/constants.hpp): `x = sin(vex::constants::two_pi() * vex::element_index());`,
 * `x = std::integral_constant<int, 42>();`, user constants through VEX_CONSTANT(name, value).
 *
 * The reference pastes the value into the generated kernel source; here a constant is a by-value scalar terminal of
 * the expression IR, so it costs nothing at run time either.  Values are written out (no Boost.Math): correctly
 * rounded doubles.
 */
#include <type_traits>
#include "operations.hpp"

namespace vex {

/// Terminal produced by a VEX_CONSTANT functor (user_constant<Impl>, constants.hpp:93-136).
template <class Impl>
struct user_constant : vector_expr_tag {
    static const bool hold_by_reference = false;
    typedef typename detail::promoted<typename Impl::value_type>::type value_type;
    int lower(detail::ir_builder &b) const { b.push_scalar(Impl::value()); return dtype_of<value_type>::value; }
    void props(detail::expr_props&) const {}
};

namespace detail {
// std::integral_constant<T, v> as an operand (constants.hpp:52-88)
template <class T, T v>
struct operand<std::integral_constant<T, v>, void> {
    typedef scalar_term<T> type;
    static type wrap(const std::integral_constant<T, v>&) { return type(v); }
};
template <class T, T v> struct is_operand<std::integral_constant<T, v>> : std::true_type {};
} // namespace detail

} // namespace vex

/// VEX_CONSTANT(name, value): `name()` is usable in vector expressions, `name` converts to the value (constants.hpp:142-161).
#define VEX_CONSTANT(name, val)                                                        \
    struct constant_##name {                                                           \
        typedef decltype(val) value_type;                                              \
        static value_type value() { static const value_type v = val; return v; }       \
        const vex::user_constant<constant_##name> operator()() const { return vex::user_constant<constant_##name>(); } \
        operator value_type() const { return value(); }                                \
    };                                                                                 \
    const constant_##name name = {}

namespace vex {
/// Mathematical constants (the list of constants.hpp:167-212).
namespace constants {
VEX_CONSTANT(pi,                     3.14159265358979323846264338327950288);
VEX_CONSTANT(root_pi,                1.77245385090551602729816748334114518);
VEX_CONSTANT(root_half_pi,           1.25331413731550025120788264240552263);
VEX_CONSTANT(root_two_pi,            2.50662827463100050241576528481104525);
VEX_CONSTANT(root_ln_four,           1.17741002251547469101156932645969963);
VEX_CONSTANT(e,                      2.71828182845904523536028747135266250);
VEX_CONSTANT(half,                   0.5);
VEX_CONSTANT(euler,                  0.57721566490153286060651209008240243);
VEX_CONSTANT(root_two,               1.41421356237309504880168872420969808);
VEX_CONSTANT(ln_two,                 0.69314718055994530941723212145817657);
VEX_CONSTANT(ln_ln_two,             -0.36651292058166432701243915823266947);
VEX_CONSTANT(third,                  0.33333333333333333333333333333333333);
VEX_CONSTANT(twothirds,              0.66666666666666666666666666666666667);
VEX_CONSTANT(pi_minus_three,         0.14159265358979323846264338327950288);
VEX_CONSTANT(four_minus_pi,          0.85840734641020676153735661672049712);
VEX_CONSTANT(two_pi,                 6.28318530717958647692528676655900577);
VEX_CONSTANT(half_root_two,          0.70710678118654752440084436210484904);
VEX_CONSTANT(exp_minus_half,         0.60653065971263342360379953499118045);
VEX_CONSTANT(one_div_two_pi,         0.15915494309189533576888376337251436);
VEX_CONSTANT(catalan,                0.91596559417721901505460351493238411);
VEX_CONSTANT(cbrt_pi,                1.46459188756152326302014252726379039);
VEX_CONSTANT(cosh_one,               1.54308063481524377847790562075706168);
VEX_CONSTANT(cos_one,                0.54030230586813971740093660744297660);
VEX_CONSTANT(degree,                 0.01745329251994329576923690768488613);
VEX_CONSTANT(e_pow_pi,              23.14069263277926900572908636794854738);
VEX_CONSTANT(euler_sqr,              0.33317792380771866431337145307588440);
VEX_CONSTANT(four_thirds_pi,         4.18879020478639098461685784437267051);
VEX_CONSTANT(glaisher,               1.28242712910062263687534256886979172);
VEX_CONSTANT(half_pi,                1.57079632679489661923132169163975144);
VEX_CONSTANT(khinchin,               2.68545200106530644530971483548179569);
VEX_CONSTANT(ln_phi,                 0.48121182505960344749775891342436842);
VEX_CONSTANT(ln_ten,                 2.30258509299404568401799145468436421);
VEX_CONSTANT(log10_e,                0.43429448190325182765112891891660508);
VEX_CONSTANT(one_div_cbrt_pi,        0.68278406325529568146702083315816455);
VEX_CONSTANT(one_div_euler,          1.73245471460063347358302531586082969);
VEX_CONSTANT(one_div_root_two,       0.70710678118654752440084436210484904);
VEX_CONSTANT(one_div_root_two_pi,    0.39894228040143267793994605993438187);
VEX_CONSTANT(phi,                    1.61803398874989484820458683436563811);
VEX_CONSTANT(pi_cubed,              31.00627668029982017547631506710139520);
VEX_CONSTANT(pi_pow_e,              22.45915771836104547342715220454373502);
VEX_CONSTANT(pi_sqr,                 9.86960440108935861883449099987615114);
VEX_CONSTANT(pi_sqr_div_six,         1.64493406684822643647241516664602519);
VEX_CONSTANT(rad,                   57.29577951308232087679815481410517033);
VEX_CONSTANT(root_e,                 1.64872127070012814684865078781416357);
VEX_CONSTANT(root_one_div_pi,        0.56418958354775628694807945156077259);
VEX_CONSTANT(root_three,             1.73205080756887729352744634150587237);
VEX_CONSTANT(sinh_one,               1.17520119364380145688238185059560082);
VEX_CONSTANT(sin_one,                0.84147098480789650665250232163029900);
VEX_CONSTANT(sixth_pi,               0.52359877559829887307710723054658381);
VEX_CONSTANT(three_quarters,         0.75);
VEX_CONSTANT(three_quarters_pi,      2.35619449019234492884698253745962716);
VEX_CONSTANT(two_div_pi,             0.63661977236758134307553505349005745);
VEX_CONSTANT(two_thirds,             0.66666666666666666666666666666666667);
VEX_CONSTANT(two_thirds_pi,          2.09439510239319549230842892218633526);
VEX_CONSTANT(zeta_three,             1.20205690315959428539973816151144999);
VEX_CONSTANT(zeta_two,               1.64493406684822643647241516664602519);
} // namespace constants
} // namespace vex
#endif
