// Drop-in for algorithm::RatioMask (reference: include/flucoma/algorithms/public/RatioMask.hpp:33-57).
// As a stand-alone object this is a trivially cheap elementwise host loop; inside BufNMF / NMFFilter the same
// arithmetic runs fused on the device for all components at once (kernel k_mask, csrc/kernels_stft.cu).
#pragma once
#include "../util/AlgorithmUtils.hpp"
#include "../../data/TensorTypes.hpp"
#include <cassert>
#include <cmath>
#include <vector>

namespace fluid {
namespace algorithm {

class RatioMask
{
public:
  RatioMask(index maxRows, index maxCols, Allocator& = FluidDefaultAllocator()) : mMultiplier(asUnsigned(maxRows * maxCols)) {}

  void init(RealMatrixView denominator)
  { // :33-42
    mRows = denominator.rows();
    mCols = denominator.cols();
    assert(asUnsigned(mRows * mCols) <= mMultiplier.size());
    auto m = mMultiplier.begin();
    for (auto it = denominator.begin(); it != denominator.end(); ++it, ++m) *m = 1.0 / std::max(*it, epsilon);
    mInitialized = true;
  }

  void process(const ComplexMatrixView& mixture, RealMatrixView targetMag, index exponent, ComplexMatrixView out)
  { // :44-57
    assert(mInitialized);
    assert(mixture.cols() == targetMag.cols());
    assert(mixture.rows() == targetMag.rows());
    auto m = mMultiplier.begin();
    auto t = targetMag.begin();
    auto o = out.begin();
    for (auto x = mixture.begin(); x != mixture.end(); ++x, ++t, ++m, ++o)
      *o = *x * std::min(1.0, std::pow(*t, double(exponent)) * std::pow(*m, double(exponent)));
  }

private:
  std::vector<double> mMultiplier;
  bool                mInitialized{false};
  index               mRows{0}, mCols{0};
};
} // namespace algorithm
} // namespace fluid
