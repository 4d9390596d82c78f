// Progress + cooperative cancel token (reference: include/flucoma/clients/common/FluidTask.hpp:22-39,
// FluidContext.hpp:20-52).  processUpdate()/iterationUpdate() return false once cancel() was requested; NMF polls it
// once per iteration through its progress callback (NMFClient.hpp:261-267).
#pragma once
#include "../../data/FluidIndex.hpp"
#include "../../data/FluidMemory.hpp"
#include <atomic>

namespace fluid {
namespace client {

class FluidTask
{
public:
  bool processUpdate(double samplesDone, double taskLength)
  {
    mProgress = (mIteration + samplesDone / taskLength) / mTotalIterations;
    return !mCancel;
  }
  bool iterationUpdate(double iterationsDone, double totalIterations)
  {
    mIteration = iterationsDone;
    mTotalIterations = totalIterations;
    mProgress = mIteration / mTotalIterations;
    return !mCancel;
  }
  double progress() const { return mProgress; }
  void   cancel() { mCancel = true; }
  bool   cancelled() const { return mCancel; }

private:
  std::atomic<double> mProgress{0.0};
  double              mIteration{0.0}, mTotalIterations{1.0};
  std::atomic<bool>   mCancel{false};
};

class FluidContext
{
public:
  FluidContext() = default;
  explicit FluidContext(FluidTask& t) : mTask(&t) {}
  FluidTask* task() { return mTask; }
  void       task(FluidTask* t) { mTask = t; }
  index      hostVectorSize() const { return mVectorSize; }
  void       hostVectorSize(index s) { mVectorSize = s; }
  Allocator& allocator() { return FluidDefaultAllocator(); }

private:
  FluidTask* mTask{nullptr};
  index      mVectorSize{64};
};
} // namespace client
} // namespace fluid
