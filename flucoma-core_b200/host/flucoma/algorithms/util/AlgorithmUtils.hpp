// reference: include/flucoma/algorithms/util/AlgorithmUtils.hpp:19
#pragma once
#include <limits>
namespace fluid {
namespace algorithm {
constexpr double epsilon = std::numeric_limits<double>::epsilon();
}
} // namespace fluid
