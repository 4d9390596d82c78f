// Drop-in for flucoma-core's algorithm::STFT / algorithm::ISTFT (reference: include/flucoma/algorithms/public/STFT.hpp).
// Same class names, constructor arguments and method signatures; the arithmetic runs on the B200 through the C ABI
// (fb200_stft / fb200_istft).  Only the Hann window (windowType 0) exists on the device, which is all the NMF clients
// ever ask for (clients/nrt/NMFClient.hpp:213-214).
#pragma once
#include "../util/B200Backend.hpp"
#include "../../data/TensorTypes.hpp"
#include <cassert>
#include <cmath>
#include <complex>
#include <memory>

namespace fluid {
namespace algorithm {

class STFT
{
public:
  // STFT.hpp:36-47
  STFT(index windowSize, index fftSize, index hopSize, index windowType = 0, Allocator& = FluidDefaultAllocator())
      : mWindowSize(windowSize), mFFTSize(fftSize), mHopSize(hopSize), mFrameSize(fftSize / 2 + 1),
        mMaxWindowSize(windowSize)
  {
    if (windowType != 0) throw std::runtime_error("flucoma-b200 STFT: only the Hann window (type 0) is implemented");
    makePlan();
  }

  void resize(index windowSize, index fftSize, index hopSize)
  { // STFT.hpp:49-59
    assert(windowSize <= mMaxWindowSize && "STFT: Window Size greater than Max");
    mWindowSize = windowSize;
    mFFTSize = fftSize;
    mHopSize = hopSize;
    mFrameSize = fftSize / 2 + 1;
    makePlan();
  }

  // STFT.hpp:61-73 -- elementwise |X|; stays on the host (the fused device path is process() + wantMagnitude)
  static void magnitude(const FluidTensorView<std::complex<double>, 2> in, FluidTensorView<double, 2> out)
  {
    assert(in.rows() == out.rows() && in.cols() == out.cols());
    auto o = out.begin();
    for (auto i = in.begin(); i != in.end(); ++i, ++o) *o = std::abs(*i);
  }
  static void magnitude(const FluidTensorView<std::complex<double>, 1> in, FluidTensorView<double, 1> out)
  {
    assert(in.size() == out.size());
    auto o = out.begin();
    for (auto i = in.begin(); i != in.end(); ++i, ++o) *o = std::abs(*i);
  }
  static void phase(const FluidTensorView<std::complex<double>, 2> in, FluidTensorView<double, 2> out)
  { // STFT.hpp:75-80
    auto o = out.begin();
    for (auto i = in.begin(); i != in.end(); ++i, ++o) *o = std::arg(*i);
  }

  // STFT.hpp:90-108: audio[n] -> spectrogram[(n+hop)/hop][fft/2+1]
  void process(const RealVectorView audio, ComplexMatrixView spectrogram)
  {
    const auto& api = b200::B200Backend::get();
    index       nFrames = static_cast<index>(api.num_frames(audio.size(), int32_t(mWindowSize), int32_t(mHopSize)));
    assert(spectrogram.rows() == nFrames && spectrogram.cols() == mFrameSize);
    auto                              in = b200::pack(audio);
    std::vector<std::complex<double>> out(asUnsigned(nFrames * mFrameSize));
    mPlan->check(api.stft(mPlan->get(), in.data(), 1, audio.size(), out.data(), nullptr, FB200_F64, FB200_HOST));
    b200::unpack(out, spectrogram);
  }

  // additive batched entry point: `batch` equal-length buffers at once, spectrum and/or magnitudes (either may be null)
  void processBatch(const double* audio, index batch, index nSamples, std::complex<double>* spectrogram, double* mags)
  {
    const auto& api = b200::B200Backend::get();
    mPlan->check(api.stft(mPlan->get(), audio, batch, nSamples, spectrogram, mags, FB200_F64, FB200_HOST));
  }

  // STFT.hpp:110-118: one already-cut frame of `windowSize` samples (window * frame -> rFFT)
  void processFrame(const RealVectorView frame, ComplexVectorView out)
  {
    assert(frame.size() == mWindowSize);
    // A lone frame equals the STFT of [frame | zeros] read at the frame whose start is sample 0: feed it as a signal
    // of length win + win/2 preceded by nothing: frame index (win/2)/hop is not integral in general, so use hop = win.
    const auto& api = b200::B200Backend::get();
    if (!mFramePlan || mFramePlanWin != mWindowSize || mFramePlanFFT != mFFTSize)
    {
      mFramePlan = std::make_unique<b200::Plan>(mWindowSize, mFFTSize, mWindowSize - mWindowSize / 2 > 0 ? mWindowSize / 2 : 1);
      mFramePlanWin = mWindowSize;
      mFramePlanFFT = mFFTSize;
    }
    // with hop = win/2 the centred STFT's frame 1 covers exactly samples [0, win)
    index               hop = mWindowSize / 2 > 0 ? mWindowSize / 2 : 1;
    auto                in = b200::pack(frame);
    index               nFrames = static_cast<index>(api.num_frames(frame.size(), int32_t(mWindowSize), int32_t(hop)));
    std::vector<std::complex<double>> spec(asUnsigned(nFrames * mFrameSize));
    mFramePlan->check(api.stft(mFramePlan->get(), in.data(), 1, frame.size(), spec.data(), nullptr, FB200_F64, FB200_HOST));
    auto o = out.begin();
    for (index b = 0; b < mFrameSize; ++b, ++o) *o = spec[asUnsigned(1 * mFrameSize + b)];
  }

  index windowSize() const { return mWindowSize; }
  index hopSize() const { return mHopSize; }
  index frameSize() const { return mFrameSize; }

  b200::Plan& plan() { return *mPlan; }

private:
  void makePlan() { mPlan = std::make_unique<b200::Plan>(mWindowSize, mFFTSize, mHopSize); }

  index                       mWindowSize, mFFTSize, mHopSize, mFrameSize, mMaxWindowSize;
  std::unique_ptr<b200::Plan> mPlan;
  std::unique_ptr<b200::Plan> mFramePlan;
  index                       mFramePlanWin{0}, mFramePlanFFT{0};
};

class ISTFT
{
public:
  // STFT.hpp:154-164
  ISTFT(index windowSize, index fftSize, index hopSize, index windowType = 0, Allocator& = FluidDefaultAllocator())
      : mWindowSize(windowSize), mFFTSize(fftSize), mHopSize(hopSize), mMaxWindowSize(windowSize)
  {
    if (windowType != 0) throw std::runtime_error("flucoma-b200 ISTFT: only the Hann window (type 0) is implemented");
    mPlan = std::make_unique<b200::Plan>(mWindowSize, mFFTSize, mHopSize);
  }

  void resize(index windowSize, index fftSize, index hopSize)
  { // STFT.hpp:166-176
    assert(windowSize <= mMaxWindowSize && "STFT: Window Size greater than Max");
    mWindowSize = windowSize;
    mFFTSize = fftSize;
    mHopSize = hopSize;
    mPlan = std::make_unique<b200::Plan>(mWindowSize, mFFTSize, mHopSize);
  }

  // STFT.hpp:178-199
  void process(const ComplexMatrixView spectrogram, RealVectorView audio)
  {
    const auto&         api = b200::B200Backend::get();
    auto                in = b200::pack(spectrogram);
    std::vector<double> out(asUnsigned(audio.size()));
    mPlan->check(api.istft(mPlan->get(), in.data(), 1, spectrogram.rows(), out.data(), audio.size(), FB200_F64, FB200_HOST));
    b200::unpack(out, audio);
  }

  void processBatch(const std::complex<double>* spectrogram, index batch, index nFrames, double* audio, index nSamples)
  {
    const auto& api = b200::B200Backend::get();
    mPlan->check(api.istft(mPlan->get(), spectrogram, batch, nFrames, audio, nSamples, FB200_F64, FB200_HOST));
  }

private:
  index                       mWindowSize, mFFTSize, mHopSize, mMaxWindowSize;
  std::unique_ptr<b200::Plan> mPlan;
};

} // namespace algorithm
} // namespace fluid
