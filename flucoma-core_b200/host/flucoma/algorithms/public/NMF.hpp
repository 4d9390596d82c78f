// Drop-in for flucoma-core's algorithm::NMF (reference: include/flucoma/algorithms/public/NMF.hpp).
// process() / processFrame() keep the reference signatures and semantics (layouts X[F][B], W[K][B], H[F][K];
// random init from mt19937_64(seed) for W and H independently; eps clamps; W0 mutated by processFrame; progress
// callbacks polled once per iteration, a false return cancels and leaves V1 = X) and run on the B200 through
// fb200_nmf_process / fb200_nmf_process_frames.  processBatch() is the additive batched entry point.
#pragma once
#include "../util/AlgorithmUtils.hpp"
#include "../util/B200Backend.hpp"
#include "../../data/TensorTypes.hpp"
#include <cassert>
#include <functional>
#include <memory>
#include <vector>

namespace fluid {
namespace algorithm {

class NMF
{
public:
  using ProgressCallback = std::function<bool(index)>; // NMF.hpp:31

  // NMF.hpp:33-42 -- rank-1 estimate of component idx: V[f][b] = H[f][idx] * W[idx][b]
  static void estimate(const RealMatrixView W, const RealMatrixView H, index idx, RealMatrixView V)
  {
    for (index f = 0; f < H.rows(); ++f)
      for (index b = 0; b < W.cols(); ++b) V(f, b) = H(f, idx) * W(idx, b);
  }

  // NMF.hpp:45-89
  void processFrame(const RealVectorView x, const RealMatrixView W0, RealVectorView out, index nIterations,
                    RealVectorView v, index randomSeed, Allocator& = FluidDefaultAllocator())
  {
    const auto& api = b200::B200Backend::get();
    index       rank = W0.extent(0), nBins = x.extent(0);
    auto        xs = b200::pack(x);
    auto        w = b200::pack(W0);
    std::vector<double> h(asUnsigned(rank)), vv(asUnsigned(nBins)), wn(asUnsigned(rank * nBins));
    fb200_frames_args a{};
    a.struct_size = sizeof(a);
    a.dtype = FB200_F64; a.mem = FB200_HOST;
    a.frames = 1; a.bins = nBins; a.rank = int32_t(rank); a.iterations = int32_t(nIterations);
    a.seed = randomSeed;
    a.X = xs.data(); a.W0 = w.data(); a.W_norm = wn.data(); a.H = h.data();
    a.V = v.data() ? vv.data() : nullptr;
    plan().check(api.nmf_process_frames(plan().get(), &a));
    // the reference clamps and row-normalises the caller's W0 in place (:58, :63-64)
    b200::unpack(wn, FluidTensorView<double, 2>(W0.descriptor(), const_cast<double*>(W0.baseData())));
    if (out.data()) b200::unpack(h, out); // :85-86
    if (v.data()) b200::unpack(vv, v);    // :88
  }

  // NMF.hpp:91-134
  void process(const RealMatrixView X, RealMatrixView W1, RealMatrixView H1, RealMatrixView V1, index rank,
               index nIterations, bool updateW, bool updateH = false, index randomSeed = -1,
               RealMatrixView W0 = RealMatrixView(nullptr, 0, 0, 0), RealMatrixView H0 = RealMatrixView(nullptr, 0, 0, 0))
  {
    index nFrames = X.extent(0), nBins = X.extent(1);
    bool  hasW0 = !(W0.extent(0) == 0 && W0.extent(1) == 0); // :102
    bool  hasH0 = !(H0.extent(0) == 0 && H0.extent(1) == 0); // :114
    if (hasW0) { assert(W0.extent(0) == rank); assert(W0.extent(1) == nBins); } // :109-110
    if (hasH0) { assert(H0.extent(0) == nFrames); assert(H0.extent(1) == rank); } // :121-122
    auto                x = b200::pack(X);
    std::vector<double> w0, h0;
    if (hasW0) w0 = b200::pack(W0);
    if (hasH0) h0 = b200::pack(H0);
    std::vector<double> w1(asUnsigned(rank * nBins)), h1(asUnsigned(nFrames * rank)), v1(asUnsigned(nFrames * nBins));
    int64_t seed = randomSeed;
    processBatch(x.data(), 1, nFrames, nBins, rank, nIterations, updateW, updateH, &seed, hasW0 ? w0.data() : nullptr,
                 hasH0 ? h0.data() : nullptr, w1.data(), h1.data(), v1.data());
    b200::unpack(v1, V1); // :131-133
    b200::unpack(w1, W1);
    b200::unpack(h1, H1);
  }

  // Additive: `batch` independent factorisations in one device pass.  Dense arrays X[batch][F][B] etc., seeds[batch].
  // Returns false when a progress callback cancelled (V1 then equals X, NMF.hpp:175-176).
  bool processBatch(const double* X, index batch, index nFrames, index nBins, index rank, index nIterations, bool updateW,
                    bool updateH, const int64_t* seeds, const double* W0, const double* H0, double* W1, double* H1,
                    double* V1)
  {
    const auto&    api = b200::B200Backend::get();
    fb200_nmf_args a{};
    a.struct_size = sizeof(a);
    a.dtype = FB200_F64; a.mem = FB200_HOST;
    a.batch = batch; a.frames = nFrames; a.bins = nBins;
    a.rank = int32_t(rank); a.iterations = int32_t(nIterations);
    a.update_w = updateW; a.update_h = updateH;
    a.X = X; a.seeds = seeds; a.W0 = W0; a.H0 = H0; a.W1 = W1; a.H1 = H1; a.V1 = V1;
    if (!mCallbacks.empty())
    {
      a.progress = &NMF::trampoline;
      a.progress_user = this;
      a.progress_stride = 1;
    }
    int32_t st = api.nmf_process(plan().get(), &a);
    plan().check(st);
    return st != FB200_CANCELLED;
  }

  void addProgressCallback(ProgressCallback&& callback) { mCallbacks.emplace_back(std::move(callback)); } // :136-139

private:
  static int trampoline(void* user, int64_t iteration)
  {
    auto* self = static_cast<NMF*>(user);
    for (auto& cb : self->mCallbacks)
      if (!cb(static_cast<index>(iteration))) return 0;
    return 1;
  }
  b200::Plan& plan()
  {
    if (!mPlan) mPlan = std::make_unique<b200::Plan>(64, 64, 32); // NMF itself has no FFT state; any valid sizes do
    return *mPlan;
  }

  std::vector<ProgressCallback> mCallbacks;
  std::unique_ptr<b200::Plan>   mPlan;
};
} // namespace algorithm
} // namespace fluid
