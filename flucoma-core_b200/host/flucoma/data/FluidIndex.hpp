// Host-side mirror of flucoma-core's index type (reference: include/flucoma/data/FluidIndex.hpp:9-21).
#pragma once
#include <cstddef>
#include <type_traits>

namespace fluid {
using index = std::ptrdiff_t;

template <typename T>
constexpr std::make_unsigned_t<T> asUnsigned(T x) { return static_cast<std::make_unsigned_t<T>>(x); }
template <typename T>
constexpr std::make_signed_t<T> asSigned(T x) { return static_cast<std::make_signed_t<T>>(x); }
} // namespace fluid
