// intp_b200/util.hpp -- the few helpers of the reference's util.hpp (src/include/util.hpp) that its
// public API and test programs name: index sequences (:18-37), a compile-time power (:43-48),
// remove_cvref_t (:195-197) and get_range (:266-270).  The reference carries C++11 polyfills; this
// header requires C++17 and forwards to the standard library.
#ifndef INTP_B200_UTIL_HPP
#define INTP_B200_UTIL_HPP

#include <cstddef>
#include <type_traits>
#include <utility>

namespace intp {
namespace util {

template <std::size_t... I>
using index_sequence = std::index_sequence<I...>;
template <std::size_t N>
using make_index_sequence = std::make_index_sequence<N>;
template <typename... T>
using make_index_sequence_for = std::index_sequence_for<T...>;

// base^exp for an unsigned exponent, usable in constant expressions
template <typename T1, typename T2>
constexpr std::enable_if_t<std::is_unsigned_v<T2>, T1> pow(T1 base, T2 exp) {
    T1 r{1};
    for (T2 i = 0; i < exp; ++i) r *= base;
    return r;
}

template <typename X>
using remove_cvref_t = std::remove_cv_t<std::remove_reference_t<X>>;

template <typename C>
auto get_range(C& c) -> std::pair<decltype(c.begin()), decltype(c.end())> {
    return std::make_pair(c.begin(), c.end());
}

}  // namespace util
}  // namespace intp

#endif  // INTP_B200_UTIL_HPP
