// BufNMFCross offline client (reference: include/flucoma/clients/nrt/NMFCrossClient.hpp:30-193), rerouted to the B200.
// Same validation, messages, buffer shapes and progress accounting as the reference's NMFCrossClient::process<T>():
// channel 0 of `source` and `target`, output resized to (target frames x 1) at the source's sample rate, progress total =
// iterations + 3 (:152), polyphony clamped to the number of source windows (:160), 50 Griffin-Lim iterations (:172).
// The compile-time ParameterSet is the plain NMFCrossParams struct (same names and defaults, :39-49); everything between
// the two STFTs and the final ISTFT runs in ONE device call (fb200_bufnmfcross).
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
namespace nmfcross {

struct NMFCrossParams
{
  std::shared_ptr<const BufferAdaptor> source;
  std::shared_ptr<const BufferAdaptor> target;
  std::shared_ptr<BufferAdaptor>       output;
  index                                timeSparsity{7};
  index                                polyphony{11};
  index                                continuity{7};
  index                                iterations{50};
  index                                seed{-1};
  FFTParams                            fftSettings{1024, -1, -1};
};

class NMFCrossClient
{
public:
  using ParamSetViewType = NMFCrossParams;
  NMFCrossClient(ParamSetViewType& p, FluidContext&) : mParams(&p) {}
  void setParams(ParamSetViewType& p) { mParams = &p; }

  template <typename T>
  Result process(FluidContext& c)
  {
    auto& P = *mParams;
    BufferAdaptor::ReadAccess source(P.source.get());
    BufferAdaptor::ReadAccess target(P.target.get());
    BufferAdaptor::Access     output(P.output.get());
    if (!source.exists()) return {Result::Status::kError, "Source Buffer Supplied But Invalid"}; // :92-97
    if (!target.exists()) return {Result::Status::kError, "Target Buffer Supplied But Invalid"};
    if (!output.exists()) return {Result::Status::kError, "Output Buffer Supplied But Invalid"};
    const double sampleRate = source.sampleRate();
    const auto&  fft = P.fftSettings;
    const index  srcFrames = source.numFrames(), tgtFrames = target.numFrames();
    const index  tgtWindows = static_cast<index>(std::floor((tgtFrames + fft.hopSize()) / fft.hopSize())); // :107-108
    if (srcFrames <= 0) return {Result::Status::kError, "Empty source buffer"};                  // :110-112
    if (tgtFrames <= 0) return {Result::Status::kError, "Empty target buffer"};
    if (P.timeSparsity > tgtWindows) return {Result::Status::kError, "Time Sparsity is larger than target frames"}; // :114-119
    if (P.continuity > tgtWindows) return {Result::Status::kError, "Continuity is larger than target frames"};
    Result resizeResult = output.resize(tgtFrames, 1, sampleRate);                               // :131-132
    if (!resizeResult.ok()) return resizeResult;

    std::vector<float> src(asUnsigned(srcFrames)), tgt(asUnsigned(tgtFrames)), out(asUnsigned(tgtFrames));
    FluidTensorView<float, 1>(src.data(), 0, srcFrames) <<= source.samps(0, srcFrames, 0);      // :134, :137
    FluidTensorView<float, 1>(tgt.data(), 0, tgtFrames) <<= target.samps(0, tgtFrames, 0);

    const double progressTotal = static_cast<double>(P.iterations + 3);                         // :152
    struct Ctx { FluidContext* c; double total; } ctx{&c, progressTotal};
    fb200_nmfcross_args a{};
    a.struct_size = sizeof(a);
    a.mem = FB200_HOST;
    a.n_source = srcFrames; a.n_target = tgtFrames;
    a.time_sparsity = int32_t(P.timeSparsity); a.polyphony = int32_t(P.polyphony); a.continuity = int32_t(P.continuity);
    a.iterations = int32_t(P.iterations);
    a.seed = P.seed;
    a.griffinlim_iterations = 50;                                                                // :172
    a.source = src.data(); a.target = tgt.data(); a.out = out.data();
    if (c.task())
    {
      a.progress = [](void* u, int64_t it) -> int {
        auto* x = static_cast<Ctx*>(u);
        return x->c->task()->processUpdate(static_cast<double>(it), x->total) ? 1 : 0;          // :154-158, checkTask :69-77
      };
      a.progress_user = &ctx;
    }
    if (c.task() && c.task()->cancelled()) return {Result::Status::kCancelled, ""};
    int32_t st;
    try
    {
      b200::Plan plan(fft.winSize(), fft.fftSize(), fft.hopSize());
      st = b200::B200Backend::get().bufnmfcross(plan.get(), &a);
      if (st < 0) return {Result::Status::kError, b200::B200Backend::get().last_error(plan.get())};
    }
    catch (const std::exception& e)
    {
      return {Result::Status::kError, e.what()};
    }
    if (st == FB200_CANCELLED || (c.task() && c.task()->cancelled())) return {Result::Status::kCancelled, ""};
    output.samps(0) <<= FluidTensorView<float, 1>(out.data(), 0, tgtFrames);                    // :183
    return {Result::Status::kOk, ""};
  }

private:
  NMFCrossParams* mParams;
};
} // namespace nmfcross
} // namespace client
} // namespace fluid
