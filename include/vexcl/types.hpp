#ifndef VEXCL_TYPES_HPP
#define VEXCL_TYPES_HPP
// Element types understood by the back end and their names (subset of vexcl/types.hpp:202-260:
// scalars only -- OpenCL vector types are outside the hot path).
#include <cstdint>
#include <string>
#include <type_traits>
#include "../vexb200.h"

typedef unsigned int  uint;
typedef double        cl_double;
typedef float         cl_float;
typedef int           cl_int;
typedef unsigned int  cl_uint;
typedef long long     cl_long;
typedef unsigned long long cl_ulong;

namespace vex {

/// N-component result of combined reductions (stands in for cl_double2 / cl_double4 / ...; fields as in CL: s[0], s[1], ...).
template <class T, unsigned N> struct vecn { T s[N]; };
template <class T> using vec2 = vecn<T, 2>;
/// Smallest CL vector width that holds n components (vexcl/types.hpp cl_fit_vec_size): 1, 2, 4, 8 or 16.
template <unsigned n> struct cl_fit_vec_size { static const unsigned value = n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : n <= 8 ? 8 : 16; };

template <class T, class Enable = void> struct dtype_of;   // no definition: unsupported element type
#define VEXB_DTYPE(T, code, nm) \
    template <> struct dtype_of<T> { static const int value = code; static const char *name() { return nm; } };
VEXB_DTYPE(double, VEXB_F64, "double")
VEXB_DTYPE(float, VEXB_F32, "float")
VEXB_DTYPE(int, VEXB_I32, "int")
VEXB_DTYPE(unsigned int, VEXB_U32, "uint")
VEXB_DTYPE(long, VEXB_I64, "long")
VEXB_DTYPE(unsigned long, VEXB_U64, "ulong")
VEXB_DTYPE(long long, VEXB_I64, "long")
VEXB_DTYPE(unsigned long long, VEXB_U64, "ulong")
#undef VEXB_DTYPE
// small integers and bool appear only as scalar terminals: promoted to int, as C does
template <> struct dtype_of<bool>  { static const int value = VEXB_I32; static const char *name() { return "bool"; } };
template <> struct dtype_of<char>  { static const int value = VEXB_I32; static const char *name() { return "char"; } };
template <> struct dtype_of<short> { static const int value = VEXB_I32; static const char *name() { return "short"; } };
template <> struct dtype_of<signed char>    { static const int value = VEXB_I32; static const char *name() { return "char"; } };
template <> struct dtype_of<unsigned char>  { static const int value = VEXB_I32; static const char *name() { return "uchar"; } };
template <> struct dtype_of<unsigned short> { static const int value = VEXB_I32; static const char *name() { return "ushort"; } };

template <class T> inline std::string type_name() { return dtype_of<typename std::decay<T>::type>::name(); }

template <class T> struct is_cl_native : std::is_arithmetic<T> {};
template <class T> struct cl_scalar_of { typedef T type; };
template <class T> struct cl_vector_length { static const unsigned value = 1; };

} // namespace vex
#endif
