#ifndef VEXCL_FUNCTION_HPP
#define VEXCL_FUNCTION_HPP
/*
 * Builtin device functions usable inside vector expressions (vexcl/function.hpp:255-268 and the
 * list that follows :287).  Each maps to one IR opcode evaluated with CUDA's math library.
 * The result type is the common type of the arguments (as the reference assumes,
 * operations.hpp:1750-1764); integer arguments of floating-only functions are promoted to double.
 *
 * User-defined functions (VEX_FUNCTION, function.hpp:225) carry a C source body and therefore
 * need run-time compilation; they are not part of this build ("out of scope").
 */
#include "operations.hpp"

namespace vex {

template <int OP, bool FloatOnly, class... Args>
struct function_node : vector_expr_tag {
    VEXCL_NODE_COMMON
    typedef typename std::common_type<typename detail::value_of<Args>::type...>::type common;
    typedef typename std::conditional<FloatOnly && std::is_integral<common>::value, double,
                                      typename detail::promoted<common>::type>::type value_type;
    std::tuple<Args...> args;
    explicit function_node(Args... a) : args(a...) {}

    int lower(detail::ir_builder &b) const {
        lower_args(b, std::index_sequence_for<Args...>());
        b.emit(OP, dtype_of<value_type>::value);
        return dtype_of<value_type>::value;
    }
    void props(detail::expr_props &p) const { props_args(p, std::index_sequence_for<Args...>()); }
    private:
        template <size_t... I> void lower_args(detail::ir_builder &b, std::index_sequence<I...>) const {
            int dummy[] = {0, (b.cvt(std::get<I>(args).lower(b), dtype_of<value_type>::value), 0)...}; (void)dummy;
        }
        template <size_t... I> void props_args(detail::expr_props &p, std::index_sequence<I...>) const {
            int dummy[] = {0, (std::get<I>(args).props(p), 0)...}; (void)dummy;
        }
};

#define VEXCL_BUILTIN_1(name, OP, FLOATONLY) \
    template <class A> \
    typename std::enable_if<is_vector_expr<A>::value, function_node<OP, FLOATONLY, typename detail::operand<A>::type> >::type \
    name(const A &a) { return function_node<OP, FLOATONLY, typename detail::operand<A>::type>(detail::operand<A>::wrap(a)); }

#define VEXCL_BUILTIN_2(name, OP, FLOATONLY) \
    template <class A, class B> \
    typename std::enable_if<detail::is_operand<A>::value && detail::is_operand<B>::value && \
                            (is_vector_expr<A>::value || is_vector_expr<B>::value), \
        function_node<OP, FLOATONLY, typename detail::operand<A>::type, typename detail::operand<B>::type> >::type \
    name(const A &a, const B &b) { \
        return function_node<OP, FLOATONLY, typename detail::operand<A>::type, typename detail::operand<B>::type>( \
                detail::operand<A>::wrap(a), detail::operand<B>::wrap(b)); }

#define VEXCL_BUILTIN_3(name, OP, FLOATONLY) \
    template <class A, class B, class C> \
    typename std::enable_if<detail::is_operand<A>::value && detail::is_operand<B>::value && detail::is_operand<C>::value && \
                            (is_vector_expr<A>::value || is_vector_expr<B>::value || is_vector_expr<C>::value), \
        function_node<OP, FLOATONLY, typename detail::operand<A>::type, typename detail::operand<B>::type, typename detail::operand<C>::type> >::type \
    name(const A &a, const B &b, const C &c) { \
        return function_node<OP, FLOATONLY, typename detail::operand<A>::type, typename detail::operand<B>::type, typename detail::operand<C>::type>( \
                detail::operand<A>::wrap(a), detail::operand<B>::wrap(b), detail::operand<C>::wrap(c)); }

VEXCL_BUILTIN_1(sin, VEXB_OP_SIN, true)     VEXCL_BUILTIN_1(cos, VEXB_OP_COS, true)     VEXCL_BUILTIN_1(tan, VEXB_OP_TAN, true)
VEXCL_BUILTIN_1(asin, VEXB_OP_ASIN, true)   VEXCL_BUILTIN_1(acos, VEXB_OP_ACOS, true)   VEXCL_BUILTIN_1(atan, VEXB_OP_ATAN, true)
VEXCL_BUILTIN_1(sinh, VEXB_OP_SINH, true)   VEXCL_BUILTIN_1(cosh, VEXB_OP_COSH, true)   VEXCL_BUILTIN_1(tanh, VEXB_OP_TANH, true)
VEXCL_BUILTIN_1(exp, VEXB_OP_EXP, true)     VEXCL_BUILTIN_1(exp2, VEXB_OP_EXP2, true)   VEXCL_BUILTIN_1(log, VEXB_OP_LOG, true)
VEXCL_BUILTIN_1(log2, VEXB_OP_LOG2, true)   VEXCL_BUILTIN_1(log10, VEXB_OP_LOG10, true) VEXCL_BUILTIN_1(sqrt, VEXB_OP_SQRT, true)
VEXCL_BUILTIN_1(rsqrt, VEXB_OP_RSQRT, true) VEXCL_BUILTIN_1(cbrt, VEXB_OP_CBRT, true)   VEXCL_BUILTIN_1(fabs, VEXB_OP_FABS, false)
VEXCL_BUILTIN_1(abs, VEXB_OP_FABS, false)   VEXCL_BUILTIN_1(floor, VEXB_OP_FLOOR, true) VEXCL_BUILTIN_1(ceil, VEXB_OP_CEIL, true)
VEXCL_BUILTIN_1(round, VEXB_OP_ROUND, true) VEXCL_BUILTIN_1(trunc, VEXB_OP_TRUNC, true)
VEXCL_BUILTIN_2(pow, VEXB_OP_POW, true)     VEXCL_BUILTIN_2(atan2, VEXB_OP_ATAN2, true) VEXCL_BUILTIN_2(fmod, VEXB_OP_FMOD, true)
VEXCL_BUILTIN_2(hypot, VEXB_OP_HYPOT, true) VEXCL_BUILTIN_2(fmin, VEXB_OP_FMIN, false)  VEXCL_BUILTIN_2(fmax, VEXB_OP_FMAX, false)
VEXCL_BUILTIN_2(min, VEXB_OP_FMIN, false)   VEXCL_BUILTIN_2(max, VEXB_OP_FMAX, false)
VEXCL_BUILTIN_3(fma, VEXB_OP_FMA, true)     VEXCL_BUILTIN_3(mad, VEXB_OP_FMA, true)

#undef VEXCL_BUILTIN_1
#undef VEXCL_BUILTIN_2
#undef VEXCL_BUILTIN_3

} // namespace vex
#endif
