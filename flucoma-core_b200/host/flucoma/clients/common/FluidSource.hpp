// Input ring buffer with overlapped reads (reference: include/flucoma/clients/common/FluidSource.hpp:22-165).
// push() appends one host block per channel at the write head; pull(out, frameTime) returns the out.cols() samples that
// END at host time `blockStart + frameTime`, zeros before the stream started -- so frame f of a stream covers
// [f*hop - win, f*hop) and a client's latency is one window (tests/clients/common/TestFluidSource.cpp:39-55).
#pragma once
#include "../../data/FluidMemory.hpp"
#include "../../data/FluidTensor.hpp"
#include <algorithm>
#include <cassert>
#include <vector>

namespace fluid {
namespace client {

template <typename T>
class FluidSource
{
public:
  using View = FluidTensorView<T, 2>;
  FluidSource(index size, index channels, index maxHostBufferSize, Allocator& = FluidDefaultAllocator())
      : mSize(size), mChannels(channels), mHostBufferSize(maxHostBufferSize), mMaxHostBufferSize(maxHostBufferSize),
        mData(asUnsigned(channels * (size + maxHostBufferSize)), T(0))
  {}
  FluidSource() : FluidSource(0, 1, 0) {}
  FluidSource(const FluidSource&) = delete;
  FluidSource& operator=(const FluidSource&) = delete;
  FluidSource(FluidSource&&) noexcept = default;
  FluidSource& operator=(FluidSource&&) noexcept = default;

  template <typename U>
  void push(const std::vector<FluidTensorView<U, 1>>& in)
  {
    assert(in.size() == asUnsigned(mChannels));
    const index block = in[0].size(), L = bufferSize();
    assert(block <= L);
    for (index c = 0; c < mChannels; ++c)
      for (index i = 0; i < block; ++i) at(c, (mCounter + i) % L) = static_cast<T>(in[asUnsigned(c)](i));
    mCounter = (mCounter + block) % L;
  }
  template <typename U>
  void push(FluidTensorView<U, 2> in)
  {
    assert(in.rows() == mChannels);
    const index block = in.cols(), L = bufferSize();
    assert(block <= L);
    for (index c = 0; c < mChannels; ++c)
      for (index i = 0; i < block; ++i) at(c, (mCounter + i) % L) = static_cast<T>(in(c, i));
    mCounter = (mCounter + block) % L;
  }

  // reference :68-89
  void pull(View out, index frameTime)
  {
    const index block = out.cols(), L = bufferSize();
    index       back = mHostBufferSize - frameTime; // distance of the frame's END behind the write head
    if (back > L)
    {
      out.fill(0);
      return;
    }
    back += block;
    const index start = ((mCounter - back) % L + L) % L;
    for (index c = 0; c < mChannels; ++c)
      for (index i = 0; i < block; ++i) out(c, i) = at(c, (start + i) % L);
  }

  void setHostBufferSize(index size)
  {
    assert(size <= mMaxHostBufferSize);
    mHostBufferSize = size;
  }
  void reset()
  {
    std::fill(mData.begin(), mData.end(), T(0));
    mCounter = 0;
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
