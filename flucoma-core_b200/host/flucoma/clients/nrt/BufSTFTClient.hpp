// BufSTFT offline client (reference: include/flucoma/clients/nrt/BufSTFTClient.hpp:31-287), rerouted to the B200.
//
// Same validation, messages, buffer shapes and padding rules as the reference's BufferSTFTClient::process<T>():
//   forward  (invert == 0): source channel -> magnitude and/or phase buffers, numHops frames x numBins channels,
//                           sample rate sr / hop (:133-147); padding mode 0 none / 1 win/2 / 2 win - hop (:121-131)
//   inverse  (invert != 0): magnitude + phase -> resynth, (numHops - 1) * hop + win - padding samples, sr * hop (:241-246)
// The reference's compile-time ParameterSet is replaced by the plain BufSTFTParams struct below (same names and
// defaults, :31-52); the per-frame STFT/ISTFT loops (:162-164, :262-268) become ONE fb200_bufstft call.
#pragma once
#include "../common/BufferAdaptor.hpp"
#include "../common/FluidTask.hpp"
#include "../common/ParameterTypes.hpp"
#include "../common/Result.hpp"
#include "../../algorithms/util/B200Backend.hpp"
#include <memory>
#include <vector>

namespace fluid {
namespace client {
namespace bufstft {

struct BufSTFTParams
{ // BufSTFTClient.hpp:31-52
  std::shared_ptr<const BufferAdaptor> source;
  index                                startFrame{0};
  index                                numFrames{-1};
  index                                startChan{0};
  std::shared_ptr<BufferAdaptor>       magnitude;
  std::shared_ptr<BufferAdaptor>       phase;
  std::shared_ptr<BufferAdaptor>       resynth;
  index                                inverse{0};
  index                                padding{1};
  FFTParams                            fftSettings{1024, -1, -1};
};

class BufferSTFTClient
{
public:
  using ParamSetViewType = BufSTFTParams;
  BufferSTFTClient(ParamSetViewType& p, FluidContext&) : mParams(&p) {}
  void setParams(ParamSetViewType& p) { mParams = &p; }

  template <typename T>
  Result process(FluidContext& c)
  { // :73-79
    return mParams->inverse == 0 ? processFwd<T>(c) : processInverse<T>(c);
  }

private:
  template <typename T>
  Result processFwd(FluidContext&)
  {
    auto& P = *mParams;
    if (!P.source) return {Result::Status::kError, "No input buffer supplied"}; // :86
    const bool haveMag = P.magnitude != nullptr, havePhase = P.phase != nullptr;
    if (!haveMag && !havePhase) return {Result::Status::kError, "Neither magnitude nor phase buffer supplied"}; // :94-96
    index  numFrames = P.numFrames, numChans = 1;
    Result rangeOK = bufferRangeCheck(P.source.get(), P.startFrame, numFrames, P.startChan, numChans); // :103-106
    if (!rangeOK.ok()) return rangeOK;
    auto source = BufferAdaptor::ReadAccess(P.source.get());
    if (haveMag && !BufferAdaptor::Access(P.magnitude.get()).exists())
      return {Result::Status::kError, "Magnitude buffer not found"}; // :110-111
    if (havePhase && !BufferAdaptor::Access(P.phase.get()).exists())
      return {Result::Status::kError, "Phase buffer not found"}; // :113-114

    const auto& fft = P.fftSettings;
    const index fftSize = fft.fftSize(), winSize = fft.winSize(), hopSize = fft.hopSize();
    const index numBins = (fftSize >> 1) + 1;
    int64_t     padding = 0, numHops = 0;
    if (b200::B200Backend::get().bufstft_sizes(int32_t(winSize), int32_t(hopSize), int32_t(P.padding), 0, numFrames, &padding,
                                               &numHops) != FB200_OK)
      return {Result::Status::kError, "Input shorter than one analysis window"};
    if (numChans * numBins >= 65536) // :137-141
      return {Result::Status::kError, "Can produce up to 65536 channels. Split your data up and try again"};
    const double frameRate = source.sampleRate() / double(hopSize);
    if (haveMag)
    { // :143-148
      Result r = BufferAdaptor::Access(P.magnitude.get()).resize(numHops, numBins * numChans, frameRate);
      if (!r.ok()) return r;
    }
    if (havePhase)
    { // :150-155
      Result r = BufferAdaptor::Access(P.phase.get()).resize(numHops, numBins * numChans, frameRate);
      if (!r.ok()) return r;
    }

    std::vector<float> audio(asUnsigned(numFrames));
    FluidTensorView<float, 1>(audio.data(), 0, numFrames) <<= source.samps(P.startFrame, numFrames, P.startChan); // :157
    std::vector<float> mags(haveMag ? asUnsigned(numHops * numBins) : 0), phases(havePhase ? asUnsigned(numHops * numBins) : 0);
    fb200_bufstft_args a{};
    a.struct_size = sizeof(a);
    a.mem = FB200_HOST;
    a.invert = 0;
    a.padding_mode = int32_t(P.padding);
    a.batch = 1; a.n_samples = numFrames; a.frames = numHops;
    a.audio = audio.data();
    a.mag = haveMag ? mags.data() : nullptr;
    a.phase = havePhase ? phases.data() : nullptr;
    try
    {
      b200::Plan plan(winSize, fftSize, hopSize);
      if (b200::B200Backend::get().bufstft(plan.get(), &a) < 0)
        return {Result::Status::kError, b200::B200Backend::get().last_error(plan.get())};
    }
    catch (const std::exception& e)
    {
      return {Result::Status::kError, e.what()};
    }
    // buffers hold numHops frames x numBins channels: the transposed copies of the reference (:171-181)
    if (haveMag)
      BufferAdaptor::Access(P.magnitude.get()).allFrames().transpose() <<= FluidTensorView<float, 2>(mags.data(), 0, numHops, numBins);
    if (havePhase)
      BufferAdaptor::Access(P.phase.get()).allFrames().transpose() <<= FluidTensorView<float, 2>(phases.data(), 0, numHops, numBins);
    return {};
  }

  template <typename T>
  Result processInverse(FluidContext&)
  {
    auto& P = *mParams;
    if (!P.magnitude || !P.phase)
      return {Result::Status::kError, "Need both magnutude and phase buffers for inverse transform"}; // :201-203
    if (!P.resynth) return {Result::Status::kError, "No resynthesis buffer supplied"};                  // :207
    auto mags = BufferAdaptor::ReadAccess(P.magnitude.get());
    auto phases = BufferAdaptor::ReadAccess(P.phase.get());
    if (mags.numFrames() != phases.numFrames() || mags.numChans() != phases.numChans())
      return {Result::Status::kError, "Magnitude and Phase buffer sizes don't match"}; // :212-215
    const auto& fft = P.fftSettings;
    const index fftSize = fft.fftSize(), winSize = fft.winSize(), hopSize = fft.hopSize();
    const index numBins = (fftSize >> 1) + 1;
    if (mags.numChans() != numBins) return {Result::Status::kError, "Wrong number of channels for FFT size"}; // :221-228
    const index numFrames = mags.numFrames();
    int64_t     padding = 0, outSize = 0;
    if (b200::B200Backend::get().bufstft_sizes(int32_t(winSize), int32_t(hopSize), int32_t(P.padding), 1, numFrames, &padding,
                                               &outSize) != FB200_OK)
      return {Result::Status::kError, "Magnitude buffer is empty"};
    auto   resynth = BufferAdaptor::Access(P.resynth.get());
    Result resizeResult = resynth.resize(outSize, 1, mags.sampleRate() * double(hopSize)); // :243-246
    if (!resizeResult.ok()) return resizeResult;

    std::vector<float> m(asUnsigned(numFrames * numBins)), ph(asUnsigned(numFrames * numBins)), out(asUnsigned(outSize));
    FluidTensorView<float, 2>(m.data(), 0, numFrames, numBins) <<= mags.allFrames().transpose();   // :256-257
    FluidTensorView<float, 2>(ph.data(), 0, numFrames, numBins) <<= phases.allFrames().transpose();
    fb200_bufstft_args a{};
    a.struct_size = sizeof(a);
    a.mem = FB200_HOST;
    a.invert = 1;
    a.padding_mode = int32_t(P.padding);
    a.batch = 1; a.frames = numFrames;
    a.mag = m.data(); a.phase = ph.data(); a.resynth = out.data();
    try
    {
      b200::Plan plan(winSize, fftSize, hopSize);
      if (b200::B200Backend::get().bufstft(plan.get(), &a) < 0)
        return {Result::Status::kError, b200::B200Backend::get().last_error(plan.get())};
    }
    catch (const std::exception& e)
    {
      return {Result::Status::kError, e.what()};
    }
    resynth.samps(0) <<= FluidTensorView<float, 1>(out.data(), 0, outSize); // :279
    return {};
  }

  BufSTFTParams* mParams;
};
} // namespace bufstft
} // namespace client
} // namespace fluid
