#ifndef VEXCL_UTIL_HPP
#define VEXCL_UTIL_HPP
// Small helpers with the reference's names (vexcl/util.hpp:67-93).
#include <cstddef>
#include <stdexcept>
#include <string>

namespace vex {

/// Throws std::runtime_error(message) when the condition does not hold.
template <class Condition, class Message>
inline void precondition(const Condition &condition, const Message &message) {
    if (!condition) throw std::runtime_error(message);
}

inline size_t nextpow2(size_t x) {
    size_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

/// Round n up to a multiple of m (slice boundaries use m = 16).
inline size_t alignup(size_t n, size_t m = 16U) { return (n + m - 1) / m * m; }

} // namespace vex
#endif
