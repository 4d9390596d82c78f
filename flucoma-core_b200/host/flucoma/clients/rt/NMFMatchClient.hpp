// NMFMatch streaming client (reference: include/flucoma/clients/rt/NMFMatchClient.hpp:25-136), rerouted to the B200.
// Same block semantics as the reference's process(): the activations written to the control output are those of the LAST
// frame of the PREVIOUS host block (they are copied out before this block's frames are analysed, :110-113), every frame
// is solved from h0 = U(seed) with a fresh copy of the bases (:106-107), and the iteration count is the reference's
// hard-coded 10 (:115), not the `iterations` parameter.
// Differences, both deliberate: the compile-time ParameterSet is the plain NMFMatchParams struct (same names and defaults,
// :32-38), and all frames that fall due in one host block go to the device in one call.
#pragma once
#include "../common/BufferAdaptor.hpp"
#include "../common/BufferedProcess.hpp"
#include "../../algorithms/util/B200Backend.hpp"
#include <memory>
#include <vector>

namespace fluid {
namespace client {
namespace nmfmatch {

struct NMFMatchParams
{
  std::shared_ptr<const BufferAdaptor> bases;
  index                                maxComponents{20};
  index                                iterations{10};
  index                                seed{-1};
  FFTParams                            fftSettings{1024, -1, -1};
};

class NMFMatchClient
{
public:
  using ParamSetViewType = NMFMatchParams;
  NMFMatchClient(ParamSetViewType& p, FluidContext& c)
      : mParams(&p), mActivations(asUnsigned(p.maxComponents), 0.0),
        mSTFTProcessor(p.fftSettings, 1, 0, c.hostVectorSize(), c.allocator())
  {}
  void  setParams(ParamSetViewType& p) { mParams = &p; }
  index latency() const { return mParams->fftSettings.winSize(); } // :71
  void  reset(FluidContext&) { mSTFTProcessor.reset(); }
  index controlChannelsOut() const { return mRank; }

  template <typename T>
  void process(std::vector<HostVector<T>>& input, std::vector<HostVector<T>>& output, FluidContext& c)
  {
    if (!input[0].data()) return;
    auto& P = *mParams;
    if (!P.bases) return;
    BufferAdaptor::ReadAccess filterBuffer(P.bases.get());
    if (!filterBuffer.valid()) return;
    const FFTParams& fft = P.fftSettings;
    const index      rank = std::min<index>(filterBuffer.numChans(), P.maxComponents);
    const index      frameSize = fft.frameSize();
    if (filterBuffer.numFrames() != frameSize) return; // :96
    mRank = rank;
    mFilter.resize(asUnsigned(rank * frameSize));
    for (index i = 0; i < rank; ++i)
    {
      auto ch = filterBuffer.samps(i);
      for (index b = 0; b < frameSize; ++b) mFilter[asUnsigned(i * frameSize + b)] = ch(b);
    }
    // :110-111 -- last block's activations go out first
    for (index i = 0; i < rank; ++i) output[0](i) = static_cast<T>(mActivations[asUnsigned(i)]);
    for (index i = rank; i < P.maxComponents && i < output[0].size(); ++i) output[0](i) = 0;

    mSTFTProcessor.processInput(fft, input, c, [&](const float* frames, index nFrames) {
      if (!mPlan || mWin != fft.winSize() || mFFT != fft.fftSize() || mHop != fft.hopSize())
      {
        mPlan = std::make_unique<b200::Plan>(fft.winSize(), fft.fftSize(), fft.hopSize(), P.maxComponents);
        mWin = fft.winSize(); mFFT = fft.fftSize(); mHop = fft.hopSize();
      }
      mActs.resize(asUnsigned(nFrames * rank));
      fb200_filter_frames_args a{};
      a.struct_size = sizeof(a);
      a.mem = FB200_HOST;
      a.frames = nFrames;
      a.rank = int32_t(rank);
      a.iterations = 10; // :115
      a.seed = P.seed;
      a.in = frames;
      a.bases = mFilter.data();
      a.acts_out = mActs.data();
      mPlan->check(b200::B200Backend::get().nmf_filter_frames(mPlan->get(), &a));
      // every frame overwrites `activations` (:114-117): the last one survives the block
      for (index i = 0; i < rank; ++i) mActivations[asUnsigned(i)] = mActs[asUnsigned((nFrames - 1) * rank + i)];
    });
  }

private:
  NMFMatchParams*             mParams;
  std::vector<double>         mActivations;
  std::vector<float>          mFilter, mActs;
  STFTBufferedProcess<false>  mSTFTProcessor;
  std::unique_ptr<b200::Plan> mPlan;
  index                       mWin{0}, mFFT{0}, mHop{0}, mRank{0};
};
} // namespace nmfmatch
} // namespace client
} // namespace fluid
