// FFT size rules only (reference: include/flucoma/clients/common/ParameterTypes.hpp:260-348).  The compile-time
// ParameterSet machinery of the reference is a host-framework concern and is not rebuilt (out of scope).
#pragma once
#include "../../data/FluidIndex.hpp"
#include <cstdint>

namespace fluid {
namespace client {

class FFTParams
{
public:
  constexpr FFTParams(intptr_t win, intptr_t hop, intptr_t fft, intptr_t max = -1)
      : mWindowSize{win}, mHopSize{hop}, mFFTSize{fft}, mMaxFFTSize{max}
  {}
  index    fftSize() const noexcept { return mFFTSize < 0 ? nextPow2(static_cast<uint32_t>(mWindowSize), true) : mFFTSize; }
  intptr_t winSize() const noexcept { return mWindowSize; }
  intptr_t hopSize() const noexcept { return mHopSize > 0 ? mHopSize : mWindowSize >> 1; }
  intptr_t frameSize() const { return (fftSize() >> 1) + 1; }
  index    max() const noexcept { return mMaxFFTSize < 0 ? fftSize() : mMaxFFTSize; }
  intptr_t maxFrameSize() const { return (max() >> 1) + 1; }
  intptr_t fftRaw() const noexcept { return mFFTSize; }
  intptr_t hopRaw() const noexcept { return mHopSize; }

  // zero padding either side of the analysed segment (reference :315-323): 0 none, 1 half a window, 2 window - hop
  static index padding(const FFTParams& settings, index option)
  {
    switch (option)
    {
    case 1: return settings.winSize() >> 1;
    case 2: return settings.winSize() - settings.hopSize();
    default: return 0;
    }
  }

  static index nextPow2(uint32_t x, bool up)
  {
    if (!x) return static_cast<index>(x);
    uint32_t p = 1;
    while (p < x) p <<= 1; // smallest power of two >= x
    return static_cast<index>(up ? p : p >> 1); // the reference's bit trick halves exact powers too
  }

private:
  intptr_t mWindowSize, mHopSize, mFFTSize, mMaxFFTSize;
};
} // namespace client
} // namespace fluid
