// NMFFilter streaming client (reference: include/flucoma/clients/rt/NMFFilterClient.hpp:25-141), rerouted to the B200.
// process() keeps the reference's streaming behaviour sample for sample: one audio channel in, `rank` channels out, each
// the ratio-masked resynthesis of one basis, delayed by the latency of one window (:64), overlap-added and divided by the
// accumulated window^2 (BufferedProcess.hpp:219-237).  Per frame the device runs STFT::processFrame -> magnitude ->
// NMF::processFrame -> estimate + RatioMask per component -> ISTFT::processFrame (:98-117) for all frames that fall due
// in the host block.  The compile-time ParameterSet is the plain NMFFilterParams struct (same names and defaults, :27-32).
#pragma once
#include "../common/BufferAdaptor.hpp"
#include "../common/BufferedProcess.hpp"
#include "../../algorithms/util/B200Backend.hpp"
#include <memory>
#include <vector>

namespace fluid {
namespace client {
namespace nmffilter {

struct NMFFilterParams
{
  std::shared_ptr<const BufferAdaptor> bases;
  index                                maxComponents{20};
  index                                iterations{10};
  index                                seed{-1};
  FFTParams                            fftSettings{1024, -1, -1};
};

class NMFFilterClient
{
public:
  using ParamSetViewType = NMFFilterParams;
  NMFFilterClient(ParamSetViewType& p, FluidContext& c)
      : mParams(&p), mSTFTProcessor(p.fftSettings, 1, p.maxComponents, c.hostVectorSize(), c.allocator())
  {}
  void  setParams(ParamSetViewType& p) { mParams = &p; }
  index latency() const { return mParams->fftSettings.winSize(); } // :64
  void  reset(FluidContext&) { mSTFTProcessor.reset(); }
  index audioChannelsOut() const { return mParams->maxComponents; }

  template <typename T>
  void process(std::vector<HostVector<T>>& input, std::vector<HostVector<T>>& output, FluidContext& c)
  {
    if (!input[0].data()) return;
    auto& P = *mParams;
    assert(output.size() >= asUnsigned(P.maxComponents) && "Too few output channels");
    if (!P.bases) return;
    BufferAdaptor::ReadAccess filterBuffer(P.bases.get());
    if (!filterBuffer.valid()) return;
    const FFTParams& fft = P.fftSettings;
    const index      rank = std::min<index>(filterBuffer.numChans(), P.maxComponents);
    const index      frameSize = fft.frameSize(), win = fft.winSize(), maxRank = P.maxComponents;
    if (filterBuffer.numFrames() != frameSize) return; // :83
    mFilter.resize(asUnsigned(rank * frameSize));
    for (index i = 0; i < rank; ++i)
    {
      auto ch = filterBuffer.samps(i);
      for (index b = 0; b < frameSize; ++b) mFilter[asUnsigned(i * frameSize + b)] = ch(b);
    }
    mSTFTProcessor.process(fft, input, output, c, [&](const float* frames, index nFrames, float* out) {
      if (!mPlan || mWin != win || mFFT != fft.fftSize() || mHop != fft.hopSize())
      {
        mPlan = std::make_unique<b200::Plan>(win, fft.fftSize(), fft.hopSize(), maxRank);
        mWin = win; mFFT = fft.fftSize(); mHop = fft.hopSize();
      }
      mFrames.resize(asUnsigned(nFrames * rank * win));
      fb200_filter_frames_args a{};
      a.struct_size = sizeof(a);
      a.mem = FB200_HOST;
      a.frames = nFrames;
      a.rank = int32_t(rank);
      a.iterations = int32_t(P.iterations); // :104
      a.seed = P.seed;
      a.in = frames;
      a.bases = mFilter.data();
      a.out = mFrames.data();
      mPlan->check(b200::B200Backend::get().nmf_filter_frames(mPlan->get(), &a));
      // the sink has maxComponents channels; channels >= rank stay silent
      for (index f = 0; f < nFrames; ++f)
        for (index ch = 0; ch < maxRank; ++ch)
          for (index j = 0; j < win; ++j)
            out[(f * maxRank + ch) * win + j] = ch < rank ? mFrames[asUnsigned((f * rank + ch) * win + j)] : 0.f;
    });
  }

private:
  NMFFilterParams*            mParams;
  std::vector<float>          mFilter, mFrames;
  STFTBufferedProcess<true>   mSTFTProcessor;
  std::unique_ptr<b200::Plan> mPlan;
  index                       mWin{0}, mFFT{0}, mHop{0};
};
} // namespace nmffilter
} // namespace client
} // namespace fluid
