// Drop-in for algorithm::MelBands (reference: include/flucoma/algorithms/public/MelBands.hpp:27-108) over the C ABI
// (fb200_melbands -> kernel k_melbands, csrc/kernels_spectral.cu).  init() keeps the reference's arguments; the filter
// bank itself is built on the device from the same LinSpaced / triangle rule.  processFrame() is the reference's
// per-frame call (one small device round trip); processFrames() takes a whole magnitude spectrogram in one call.
// There is no CPU fallback: without the library or a CUDA device construction of the plan throws.
#pragma once
#include "../util/AlgorithmUtils.hpp"
#include "../util/B200Backend.hpp"
#include "../../data/FluidMemory.hpp"
#include "../../data/TensorTypes.hpp"
#include <cassert>
#include <cmath>
#include <memory>
#include <vector>

namespace fluid {
namespace algorithm {

class MelBands
{
public:
  MelBands(index maxBands, index maxFFT, Allocator& = FluidDefaultAllocator()) : mMaxBands(maxBands), mMaxBins(maxFFT / 2 + 1) {}

  static inline double hz2mel(double x) { return 1127.01048 * std::log(x / 700.0 + 1.0); } // :37-40

  void init(double lo, double hi, index nBands, index nBins, double sampleRate, index windowSize, Allocator& = FluidDefaultAllocator())
  { // :42-78
    assert(hi > lo);
    assert(nBands > 1);
    assert(nBins <= mMaxBins && nBands <= mMaxBands);
    mLo = lo; mHi = hi; mSampleRate = sampleRate;
    mNBands = nBands; mNBins = nBins;
    const index fftSize = 2 * (nBins - 1);
    mScale1 = 1.0 / (windowSize / 4.0);                      // :50
    mScale2 = 1.0 / (2.0 * double(fftSize) / windowSize);    // :53
    if (!mPlan || mWin != windowSize || mFFT != fftSize) {
      mPlan = std::make_unique<b200::Plan>(windowSize, fftSize, std::max<index>(1, windowSize / 2));
      mWin = windowSize; mFFT = fftSize;
    }
    mInitialized = true;
  }

  void processFrame(const RealVectorView in, RealVectorView out, bool magNorm, bool usePower, bool logOutput, Allocator& = FluidDefaultAllocator())
  { // :80-101
    assert(mInitialized && in.size() == mNBins && out.size() == mNBands);
    std::vector<float> x(asUnsigned(mNBins)), y(asUnsigned(mNBands));
    for (index i = 0; i < mNBins; ++i) x[asUnsigned(i)] = static_cast<float>(in(i));
    run(x.data(), 1, y.data(), magNorm, usePower, logOutput);
    for (index i = 0; i < mNBands; ++i) out(i) = y[asUnsigned(i)];
  }

  // additive: all frames of a magnitude spectrogram [frames][nBins] -> [frames][nBands] in one device pass
  void processFrames(const RealMatrixView in, RealMatrixView out, bool magNorm, bool usePower, bool logOutput)
  {
    assert(mInitialized && in.cols() == mNBins && out.cols() == mNBands && in.rows() == out.rows());
    const index F = in.rows();
    std::vector<float> x(asUnsigned(F * mNBins)), y(asUnsigned(F * mNBands));
    for (index f = 0; f < F; ++f)
      for (index i = 0; i < mNBins; ++i) x[asUnsigned(f * mNBins + i)] = static_cast<float>(in(f, i));
    run(x.data(), F, y.data(), magNorm, usePower, logOutput);
    for (index f = 0; f < F; ++f)
      for (index i = 0; i < mNBands; ++i) out(f, i) = y[asUnsigned(f * mNBands + i)];
  }

  double mScale1{1.0};
  double mScale2{1.0};

private:
  void run(const float* mags, index frames, float* bands, bool magNorm, bool usePower, bool logOutput)
  {
    fb200_melbands_args a{};
    a.struct_size = sizeof a;
    a.mem = FB200_HOST;
    a.batch = 1; a.frames = frames; a.n_samples = 0;
    a.n_bands = static_cast<int32_t>(mNBands);
    a.flags = (magNorm ? 1 : 0) | (usePower ? 2 : 0) | (logOutput ? 4 : 0);
    a.lo = mLo; a.hi = mHi; a.sample_rate = mSampleRate;
    a.mags = mags; a.audio = nullptr; a.bands = bands;
    mPlan->check(b200::B200Backend::get().melbands(mPlan->get(), &a));
  }

  index  mMaxBands, mMaxBins;
  index  mNBands{0}, mNBins{0}, mWin{0}, mFFT{0};
  double mLo{20}, mHi{20000}, mSampleRate{44100};
  bool   mInitialized{false};
  std::unique_ptr<b200::Plan> mPlan;
};
} // namespace algorithm
} // namespace fluid
