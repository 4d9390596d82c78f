// BufNMF offline client (reference: include/flucoma/clients/nrt/NMFClient.hpp:54-337), rerouted to the B200.
//
// Same validation, messages, buffer shapes and per-channel semantics as the reference's NMFClient::process<T>():
//   bases (B frames x C*K chans, sr/fft), activations (N/hop+1 x C*K, sr/hop, scaled by 1/max(H) per channel),
//   resynth (N x C*K, sr); basesMode / actMode 0 none, 1 seed, 2 fixed; every channel uses the same `seed`.
// Differences, both deliberate: (1) the reference's compile-time ParameterSet is replaced by the plain BufNMFParams
// struct below (same names and defaults, NMFClient.hpp:54-71); (2) all channels go to the device in ONE batched call
// (the reference loops over channels sequentially, :233) -- progress is reported per NMF iteration for the batch.
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
namespace bufnmf {

struct BufNMFParams
{ // NMFClient.hpp:54-71
  std::shared_ptr<const BufferAdaptor> source;
  index                                startFrame{0};
  index                                numFrames{-1};
  index                                startChan{0};
  index                                numChans{-1};
  std::shared_ptr<BufferAdaptor>       resynth;
  index                                resynthMode{0};
  std::shared_ptr<BufferAdaptor>       bases;
  index                                basesMode{0};
  std::shared_ptr<BufferAdaptor>       activations;
  index                                actMode{0};
  index                                components{1};
  index                                iterations{100};
  index                                seed{-1};
  FFTParams                            fftSettings{1024, -1, -1};
};

class NMFClient
{
public:
  using ParamSetViewType = BufNMFParams;
  NMFClient(ParamSetViewType& p, FluidContext&) : mParams(&p) {}
  void setParams(ParamSetViewType& p) { mParams = &p; }

  template <typename T>
  Result process(FluidContext& c)
  {
    auto& P = *mParams;
    index nFrames = P.numFrames, nChannels = P.numChans;
    auto  rangeCheck = bufferRangeCheck(P.source.get(), P.startFrame, nFrames, P.startChan, nChannels); // :102-105
    if (!rangeCheck.ok()) return rangeCheck;

    auto        source = BufferAdaptor::ReadAccess(P.source.get());
    double      sampleRate = source.sampleRate();
    const auto& fft = P.fftSettings;
    const index hop = fft.hopSize(), win = fft.winSize(), fftSize = fft.fftSize();
    const index nWindows = static_cast<index>(std::floor((nFrames + hop) / hop)); // :111-112
    const index nBins = fft.frameSize();
    const index rank = P.components;

    bool       hasFilters{false};
    const bool seedFilters{P.basesMode > 0}, fixFilters{P.basesMode == 2}, shouldResynth{P.resynthMode != 0};
    if (P.bases)
    { // :120-133
      BufferAdaptor::Access buf(P.bases.get());
      if (!buf.exists()) return {Result::Status::kError, "Bases Buffer Supplied But Invalid"};
      if (seedFilters && (!buf.valid() || buf.numFrames() != nBins || buf.numChans() != rank * nChannels))
        return {Result::Status::kError, "Supplied bases buffer for seeding must be [(FFTSize / 2) + 1] frames long, and "
                                        "have [rank] * [channels] channels"};
      hasFilters = true;
    }
    else if (seedFilters)
      return {Result::Status::kError, "Bases Mode set to Seed or Fix , but no Bases Buffer supplied"};

    bool       hasEnvelopes{false};
    const bool seedEnvelopes{P.actMode > 0}, fixEnvelopes{P.actMode == 2};
    const bool needsAnalysis = !(fixEnvelopes && fixFilters); // :141
    if (!needsAnalysis && !shouldResynth)
      return {Result::Status::kWarning, "Bases and Activations buffers both fixed, but resynthesis disabled: no work to do"};
    if (P.activations)
    { // :147-163
      BufferAdaptor::Access buf(P.activations.get());
      if (!buf.exists()) return {Result::Status::kError, "Activations Buffer Supplied But Invalid"};
      if (seedEnvelopes && (!buf.valid() || buf.numFrames() != (nFrames / hop) + 1 || buf.numChans() != rank * nChannels))
        return {Result::Status::kError, "Supplied activations buffer for seeding must be [(num samples / hop size)  + 1] "
                                        "frames long, and have [rank] * [channels] channels"};
      hasEnvelopes = true;
    }
    else if (seedEnvelopes)
      return {Result::Status::kError, "Activations Mode set to Seed or Fix , but no Activations Buffer supplied"};

    bool hasResynth{false};
    if (shouldResynth)
    { // :172-195
      if (!P.resynth) return {Result::Status::kError, "Resynthesis requested but no buffer supplied"};
      BufferAdaptor::Access buf(P.resynth.get());
      if (!buf.exists()) return {Result::Status::kError, "Resynthesis Buffer Supplied But Invalid"};
      hasResynth = true;
      Result r = buf.resize(nFrames, nChannels * rank, sampleRate);
      if (!r.ok()) return r;
    }
    if (hasFilters && !seedFilters)
    { // :197-203
      Result r = BufferAdaptor::Access(P.bases.get()).resize(nBins, nChannels * rank, sampleRate / double(fftSize));
      if (!r.ok()) return r;
    }
    if (hasEnvelopes && !seedEnvelopes)
    { // :204-211
      Result r = BufferAdaptor::Access(P.activations.get()).resize((nFrames / hop) + 1, nChannels * rank, sampleRate / double(hop));
      if (!r.ok()) return r;
    }

    // ---- gather: channels -> planar float [C][nFrames] (:240), seeds -> [C][K][B] / [C][F][K] (:246-258)
    std::vector<float> audio(asUnsigned(nChannels * nFrames));
    for (index i = 0; i < nChannels; ++i)
    {
      auto ch = source.samps(P.startFrame, nFrames, P.startChan + i);
      FluidTensorView<float, 1>(audio.data(), i * nFrames, nFrames) <<= ch;
    }
    std::vector<float> basesIn, actsIn;
    if (seedFilters)
    {
      basesIn.resize(asUnsigned(nChannels * rank * nBins));
      BufferAdaptor::Access filters(P.bases.get());
      for (index i = 0; i < nChannels; ++i)
        for (index j = 0; j < rank; ++j)
          FluidTensorView<float, 1>(basesIn.data(), (i * rank + j) * nBins, nBins) <<= filters.samps(i * rank + j);
    }
    if (seedEnvelopes)
    {
      actsIn.resize(asUnsigned(nChannels * nWindows * rank));
      BufferAdaptor::Access envelopes(P.activations.get());
      for (index i = 0; i < nChannels; ++i)
        for (index j = 0; j < rank; ++j)
        {
          auto                      e = envelopes.samps(i * rank + j);
          FluidTensorView<float, 2> dst(actsIn.data(), i * nWindows * rank, nWindows, rank);
          dst.col(j) <<= e;
        }
    }

    // ---- device: STFT -> |X| -> NMF -> (masks -> ISTFT) for every channel in one call (:241-333)
    std::vector<float>   basesOut(asUnsigned(nChannels * rank * nBins)), actsOut(asUnsigned(nChannels * nWindows * rank));
    std::vector<float>   resynthOut(hasResynth ? asUnsigned(nChannels * rank * nFrames) : 0);
    std::vector<int64_t> seeds(asUnsigned(nChannels), static_cast<int64_t>(P.seed));
    const double         progressTotal = static_cast<double>(needsAnalysis * P.iterations + (hasResynth ? 3 * rank : 0)); // :230-231
    struct Ctx { FluidContext* c; double total; } ctx{&c, progressTotal};
    fb200_bufnmf_args a{};
    a.struct_size = sizeof(a);
    a.mem = FB200_HOST;
    a.batch = nChannels; a.n_samples = nFrames;
    a.rank = int32_t(rank); a.iterations = int32_t(P.iterations);
    a.bases_mode = int32_t(P.basesMode); a.acts_mode = int32_t(P.actMode);
    a.audio = audio.data(); a.seeds = seeds.data();
    a.bases_in = seedFilters ? basesIn.data() : nullptr;
    a.acts_in = seedEnvelopes ? actsIn.data() : nullptr;
    a.bases_out = (hasFilters && !fixFilters) ? basesOut.data() : nullptr;
    a.acts_out = (hasEnvelopes && !fixEnvelopes) ? actsOut.data() : nullptr;
    a.resynth_out = hasResynth ? resynthOut.data() : nullptr;
    if (c.task())
    {
      a.progress = [](void* u, int64_t it) -> int {
        auto* x = static_cast<Ctx*>(u);
        return x->c->task()->processUpdate(static_cast<double>(it), x->total) ? 1 : 0; // :261-267
      };
      a.progress_user = &ctx;
      // The reference polls the task once per iteration and throws every output of a cancelled job away (:273-274), so
      // the update loop need not stop at an exact iteration: the persistent device loop keeps running, the callback is
      // fed from its pass counters and a cancel is honoured at the next pass boundary (fb200_progress_fn).
      a.progress_stride = FB200_PROGRESS_ASYNC;
    }
    int32_t st;
    try
    {
      b200::Plan plan(win, fftSize, hop, rank);
      st = b200::B200Backend::get().bufnmf(plan.get(), &a);
      if (st < 0) return {Result::Status::kError, b200::B200Backend::get().last_error(plan.get())};
      b200::B200Backend::get().get_stats(plan.get(), &mStats);
    }
    catch (const std::exception& e)
    {
      return {Result::Status::kError, e.what()};
    }
    if (st == FB200_CANCELLED || (c.task() && c.task()->cancelled())) return {Result::Status::kCancelled, ""}; // :273-274

    // ---- scatter back into the host buffers (:277-300, :329)
    if (hasFilters && !fixFilters)
    {
      BufferAdaptor::Access filters(P.bases.get());
      for (index i = 0; i < nChannels; ++i)
        for (index j = 0; j < rank; ++j)
          filters.samps(i * rank + j) <<= FluidTensorView<float, 1>(basesOut.data(), (i * rank + j) * nBins, nBins);
    }
    if (hasEnvelopes && !fixEnvelopes)
    {
      BufferAdaptor::Access envelopes(P.activations.get());
      for (index i = 0; i < nChannels; ++i)
        for (index j = 0; j < rank; ++j)
          envelopes.samps(i * rank + j) <<= FluidTensorView<float, 2>(actsOut.data(), i * nWindows * rank, nWindows, rank).col(j);
    }
    if (hasResynth)
    {
      BufferAdaptor::Access resynth(P.resynth.get());
      for (index i = 0; i < nChannels; ++i)
        for (index j = 0; j < rank; ++j)
          resynth.samps(i * rank + j) <<= FluidTensorView<float, 1>(resynthOut.data(), (i * rank + j) * nFrames, nFrames);
    }
    return {Result::Status::kOk, ""};
  }

  // instrumentation of the last process() call (engine used, stage times): additive, not in the reference
  const fb200_stats& lastStats() const { return mStats; }

private:
  BufNMFParams* mParams;
  fb200_stats   mStats{};
};
} // namespace bufnmf
} // namespace client
} // namespace fluid
