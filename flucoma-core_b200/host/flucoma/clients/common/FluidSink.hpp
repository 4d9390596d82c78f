// Output ring buffer with overlap-add (reference: include/flucoma/clients/common/FluidSink.hpp:22-176).
// push(x, frameTime) ADDS a frame at `frameTime` samples after the read head; pull() copies one host block from the read
// head, zeroes it and advances.
#pragma once
#include "../../data/FluidMemory.hpp"
#include "../../data/FluidTensor.hpp"
#include <algorithm>
#include <cassert>
#include <vector>

namespace fluid {
namespace client {

template <typename T>
class FluidSink
{
public:
  using View = FluidTensorView<T, 2>;
  FluidSink(index size, index channels, index maxHostVectorSize, Allocator& = FluidDefaultAllocator())
      : mSize(size), mChannels(channels), mHostBufferSize(maxHostVectorSize), mMaxHostBufferSize(maxHostVectorSize),
        mData(asUnsigned(channels * (size + maxHostVectorSize)), T(0))
  {}
  FluidSink() : FluidSink(0, 1, 0) {}
  FluidSink(const FluidSink&) = delete;
  FluidSink& operator=(const FluidSink&) = delete;
  FluidSink(FluidSink&&) noexcept = default;
  FluidSink& operator=(FluidSink&&) noexcept = default;

  // reference :49-69
  void push(View x, index frameTime)
  {
    assert(x.rows() == mChannels);
    const index block = x.cols(), L = bufferSize();
    assert(block <= L);
    if (frameTime + block > L) return;
    const index start = (frameTime + mCounter) % L;
    for (index c = 0; c < mChannels; ++c)
      for (index i = 0; i < block; ++i) at(c, (start + i) % L) += x(c, i);
  }

  // reference :72-89
  template <typename U>
  void pull(FluidTensorView<U, 2> out)
  {
    const index block = out.cols(), L = bufferSize();
    if (block > L) return;
    for (index c = 0; c < mChannels; ++c)
      for (index i = 0; i < block; ++i)
      {
        T& v = at(c, (mCounter + i) % L);
        out(c, i) = static_cast<U>(v);
        v = T(0);
      }
    mCounter = (mCounter + block) % L;
  }

  void reset()
  {
    std::fill(mData.begin(), mData.end(), T(0));
    mCounter = 0;
  }
  void setHostBufferSize(index size)
  {
    assert(size <= mMaxHostBufferSize);
    mHostBufferSize = size;
  }
  index channels() const noexcept { return mChannels; }
  index size() const noexcept { return mSize; }
  index hostBufferSize() const noexcept { return mHostBufferSize; }

private:
  index bufferSize() const { return mSize + mHostBufferSize; } // reference: ring length follows the CURRENT host block size
  T&    at(index c, index i) { return mData[asUnsigned(c * (mSize + mMaxHostBufferSize) + i)]; }

  index          mCounter = 0;
  index          mSize, mChannels, mHostBufferSize, mMaxHostBufferSize;
  std::vector<T> mData;
};
} // namespace client
} // namespace fluid
