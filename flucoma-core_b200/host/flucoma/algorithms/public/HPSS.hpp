// Drop-in for algorithm::HPSS (reference: include/flucoma/algorithms/public/HPSS.hpp:30-186) over the C ABI
// (fb200_hpss -> kernel k_hpss, csrc/kernels_spectral.cu: exact medians, the streaming delay lines in closed form).
// processFrame() keeps the reference's streaming contract: every call takes one input spectrum and emits the harmonic /
// percussive / residual spectra of the frame the delay line has reached ((hSize - 1) frames behind, zeros while it fills).
// The reference object's state (frame / vertical delay lines of hSize columns, and the per-bin median filters whose output
// enters the horizontal delay line (hSize - 1) / 2 + 1 columns before it is used, HPSS.hpp:85-105) is a function of its
// last hSize + (hSize - 1) / 2 + 1 input frames, so the mirror keeps those on the host and asks the device for the window's
// last output frame.  processFrames() handles a whole spectrogram in one call from init() state.
// There is no CPU fallback: without the library or a CUDA device construction of the plan throws.
#pragma once
#include "../util/AlgorithmUtils.hpp"
#include "../util/B200Backend.hpp"
#include "../../data/FluidMemory.hpp"
#include "../../data/TensorTypes.hpp"
#include <cassert>
#include <complex>
#include <deque>
#include <memory>
#include <vector>

namespace fluid {
namespace algorithm {

class HPSS
{
public:
  enum HPSSMode { kClassic, kCoupled, kAdvanced };

  HPSS(index maxFFTSize, index maxHSize, Allocator& = FluidDefaultAllocator()) : mMaxBins(maxFFTSize / 2 + 1), mMaxHSize(maxHSize) {}

  void init(index nBins, index hSize)
  { // :48-64
    assert(hSize % 2);
    assert(nBins <= mMaxBins);
    assert(hSize <= mMaxHSize);
    mNBins = nBins; mHSize = hSize;
    mHistory.clear();
    const index fft = 2 * (nBins - 1);
    if (!mPlan || mFFT != fft) {
      mPlan = std::make_unique<b200::Plan>(fft, fft, std::max<index>(1, fft / 2));
      mFFT = fft;
    }
    mInitialized = true;
  }

  void processFrame(const ComplexVectorView in, ComplexMatrixView out, index vSize, index hSize, index mode, double hThresholdX1,
                    double hThresholdY1, double hThresholdX2, double hThresholdY2, double pThresholdX1, double pThresholdY1,
                    double pThresholdX2, double pThresholdY2)
  { // :66-162; out is [nBins][3]
    assert(mInitialized);
    assert(in.size() == mNBins && out.rows() == mNBins && out.cols() == 3);
    if (hSize != mHSize) { // the reference resets its horizontal filters when the size changes (:83-90)
      assert(hSize % 2 && hSize <= mMaxHSize);
      mHSize = hSize;
      mHistory.clear();
    }
    std::vector<float> fr(asUnsigned(2 * mNBins));
    for (index i = 0; i < mNBins; ++i) { fr[asUnsigned(2 * i)] = float(in(i).real()); fr[asUnsigned(2 * i + 1)] = float(in(i).imag()); }
    mHistory.push_back(std::move(fr));
    if (static_cast<index>(mHistory.size()) > mHSize + (mHSize - 1) / 2 + 1) mHistory.pop_front();
    const index F = static_cast<index>(mHistory.size());
    std::vector<float> spec(asUnsigned(F * 2 * mNBins)), res(asUnsigned(3 * F * 2 * mNBins));
    for (index f = 0; f < F; ++f) std::copy(mHistory[asUnsigned(f)].begin(), mHistory[asUnsigned(f)].end(), spec.begin() + f * 2 * mNBins);
    run(spec.data(), F, res.data(), vSize, mode, hThresholdX1, hThresholdY1, hThresholdX2, hThresholdY2, pThresholdX1, pThresholdY1,
        pThresholdX2, pThresholdY2);
    for (index c = 0; c < 3; ++c)
      for (index i = 0; i < mNBins; ++i) {
        const float* v = res.data() + ((c * F + (F - 1)) * mNBins + i) * 2;
        out(i, c) = std::complex<double>(v[0], v[1]);
      }
  }

  // additive: all frames of a spectrogram [frames][nBins] from init() state -> out[3][frames][nBins], frame t of the output
  // belonging to input frame t - (hSize - 1) exactly as consecutive processFrame calls would emit them
  void processFrames(const ComplexMatrixView in, FluidTensorView<std::complex<double>, 3> out, index vSize, index hSize, index mode,
                     double hThresholdX1, double hThresholdY1, double hThresholdX2, double hThresholdY2, double pThresholdX1,
                     double pThresholdY1, double pThresholdX2, double pThresholdY2)
  {
    assert(mInitialized && in.cols() == mNBins);
    mHSize = hSize;
    const index F = in.rows();
    std::vector<float> spec(asUnsigned(F * 2 * mNBins)), res(asUnsigned(3 * F * 2 * mNBins));
    for (index f = 0; f < F; ++f)
      for (index i = 0; i < mNBins; ++i) {
        spec[asUnsigned((f * mNBins + i) * 2)] = float(in(f, i).real());
        spec[asUnsigned((f * mNBins + i) * 2 + 1)] = float(in(f, i).imag());
      }
    run(spec.data(), F, res.data(), vSize, mode, hThresholdX1, hThresholdY1, hThresholdX2, hThresholdY2, pThresholdX1, pThresholdY1,
        pThresholdX2, pThresholdY2);
    for (index c = 0; c < 3; ++c)
      for (index f = 0; f < F; ++f)
        for (index i = 0; i < mNBins; ++i) {
          const float* v = res.data() + ((c * F + f) * mNBins + i) * 2;
          out(c, f, i) = std::complex<double>(v[0], v[1]);
        }
  }

private:
  void run(const float* spec, index frames, float* res, index vSize, index mode, double hx1, double hy1, double hx2, double hy2,
           double px1, double py1, double px2, double py2)
  {
    fb200_hpss_args a{};
    a.struct_size = sizeof a;
    a.mem = FB200_HOST;
    a.batch = 1; a.frames = frames;
    a.v_size = static_cast<int32_t>(vSize); a.h_size = static_cast<int32_t>(mHSize); a.mode = static_cast<int32_t>(mode);
    const double th[8] = {hx1, hy1, hx2, hy2, px1, py1, px2, py2};
    for (int i = 0; i < 8; ++i) a.thresholds[i] = th[i];
    a.spectrum = spec; a.out = res;
    mPlan->check(b200::B200Backend::get().hpss(mPlan->get(), &a));
  }

  index mMaxBins, mMaxHSize;
  index mNBins{0}, mHSize{0}, mFFT{0};
  bool  mInitialized{false};
  std::deque<std::vector<float>> mHistory; // the last hSize + (hSize - 1) / 2 + 1 input frames (interleaved complex)
  std::unique_ptr<b200::Plan>    mPlan;
};
} // namespace algorithm
} // namespace fluid
