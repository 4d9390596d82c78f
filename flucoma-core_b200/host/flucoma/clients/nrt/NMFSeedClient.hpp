// BufNMFSeed offline client (reference: include/flucoma/clients/nrt/NMFSeedClient.hpp:25-140), rerouted to the B200.
// Same validation, messages and buffer shapes as NMFSeedClient::process<T>(): mono source only, bases resized to
// (bins x rank) at sr / fft, activations to (frames / hop + 1 x rank) at sr / hop and scaled by 1 / max (:119-128).
// STFT, magnitude and NNDSVD (algorithms/public/NNDSVD.hpp:30-131) run in one device call (fb200_nmfseed).
// The compile-time ParameterSet is the plain NMFSeedParams struct (same names and defaults, :39-51).
#pragma once
#include "../common/BufferAdaptor.hpp"
#include "../common/FluidTask.hpp"
#include "../common/ParameterTypes.hpp"
#include "../common/Result.hpp"
#include "../../algorithms/util/B200Backend.hpp"
#include <cmath>
#include <memory>
#include <vector>

namespace fluid {
namespace client {
namespace nndsvd {

struct NMFSeedParams
{
  std::shared_ptr<const BufferAdaptor> source;
  std::shared_ptr<BufferAdaptor>       bases;
  std::shared_ptr<BufferAdaptor>       activations;
  index                                minComponents{1};
  index                                maxComponents{200};
  double                               coverage{0.5};
  index                                method{0}; // 0 NMF-SVD, 1 NNDSVDar, 2 NNDSVDa, 3 NNDSVD
  index                                seed{-1};
  FFTParams                            fftSettings{1024, -1, -1};
};

class NMFSeedClient
{
public:
  using ParamSetViewType = NMFSeedParams;
  NMFSeedClient(ParamSetViewType& p, FluidContext&) : mParams(&p) {}
  void setParams(ParamSetViewType& p) { mParams = &p; }

  template <typename T>
  Result process(FluidContext&)
  {
    auto& P = *mParams;
    BufferAdaptor::ReadAccess source(P.source.get());
    if (!source.exists()) return {Result::Status::kError, "Source Buffer Supplied But Invalid"}; // :79-80
    const double sampleRate = source.sampleRate();
    const index  nFrames = source.numFrames();
    const auto&  fft = P.fftSettings;
    const index  nWindows = static_cast<index>(std::floor((nFrames + fft.hopSize()) / fft.hopSize())); // :85-86
    const index  nBins = fft.frameSize();
    if (source.numChans() > 1) return {Result::Status::kError, "Only one channel supported"};   // :89-90

    std::vector<float> audio(asUnsigned(nFrames));
    FluidTensorView<float, 1>(audio.data(), 0, nFrames) <<= source.samps(0, nFrames, 0);        // :95
    const index        maxRank = P.maxComponents;
    std::vector<float> W(asUnsigned(maxRank * nBins)), H(asUnsigned(nWindows * maxRank));
    int32_t            rank = 0;
    fb200_nmfseed_args a{};
    a.struct_size = sizeof(a);
    a.mem = FB200_HOST;
    a.n_samples = nFrames;
    a.min_rank = int32_t(P.minComponents); a.max_rank = int32_t(maxRank);
    a.coverage = P.coverage;
    a.method = int32_t(P.method);
    a.scale_acts = 1;                                                                            // :119-128
    a.seed = P.seed;
    a.audio = audio.data();
    a.bases = W.data(); a.acts = H.data();
    a.rank_out = &rank;
    try
    {
      b200::Plan plan(fft.winSize(), fft.fftSize(), fft.hopSize(), maxRank);
      int32_t    st = b200::B200Backend::get().nmfseed(plan.get(), &a);
      if (st < 0) return {Result::Status::kError, b200::B200Backend::get().last_error(plan.get())};
    }
    catch (const std::exception& e)
    {
      return {Result::Status::kError, e.what()};
    }
    BufferAdaptor::Access filters(P.bases.get());
    Result                r = filters.resize(nBins, rank, sampleRate / double(fft.fftSize()));  // :107-109
    if (!r.ok()) return r;
    for (index j = 0; j < rank; ++j) filters.samps(j) <<= FluidTensorView<float, 1>(W.data(), j * nBins, nBins); // :111-112
    BufferAdaptor::Access envelopes(P.activations.get());
    r = envelopes.resize((nFrames / fft.hopSize()) + 1, rank, sampleRate / double(fft.hopSize())); // :115-117
    if (!r.ok()) return r;
    for (index j = 0; j < rank; ++j)
      envelopes.samps(j) <<= FluidTensorView<float, 2>(H.data(), 0, nWindows, maxRank).col(j);   // :122-127
    mRank = rank;
    return {Result::Status::kOk, ""};
  }
  index lastRank() const { return mRank; } // additive: the rank NNDSVD chose

private:
  NMFSeedParams* mParams;
  index          mRank{0};
};
} // namespace nndsvd
} // namespace client
} // namespace fluid
