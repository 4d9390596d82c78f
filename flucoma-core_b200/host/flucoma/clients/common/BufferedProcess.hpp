// Streaming hop scheduler (reference: include/flucoma/clients/common/BufferedProcess.hpp:34-336).
// BufferedProcess keeps the reference's interface and timing: frames start every `hop` samples of host time, a frame is
// the `win` samples that end at its start time, outputs are overlap-added `frameTime` samples after the read head.
// STFTBufferedProcess differs from the reference in ONE deliberate way: the per-frame spectral work (window, FFT, the
// client's processing, inverse FFT, window) runs on the B200, so instead of a lambda over one spectrum the client
// passes a BATCH function over all frames that fall due in the current host block (usually one, several when the host
// block is longer than the hop).  Pulls do not depend on pushes, so batching them is exact.
#pragma once
#include "FluidSink.hpp"
#include "FluidSource.hpp"
#include "FluidTask.hpp"
#include "ParameterTypes.hpp"
#include "../../data/TensorTypes.hpp"
#include <algorithm>
#include <cmath>
#include <vector>

namespace fluid {
namespace client {

template <typename T>
using HostVector = FluidTensorView<T, 1>;
template <typename T>
using HostMatrix = FluidTensorView<T, 2>;

class BufferedProcess
{
public:
  BufferedProcess(index maxFramesIn, index maxFramesOut, index maxChannelsIn, index maxChannelsOut, index hostSize,
                  Allocator& alloc = FluidDefaultAllocator())
      : mHostSize(hostSize), mMaxHostSize(hostSize), mSource(maxFramesIn, maxChannelsIn, hostSize, alloc),
        mSink(maxFramesOut, maxChannelsOut, hostSize, alloc), mFrameIn(asUnsigned(maxChannelsIn * maxFramesIn)),
        mFrameOut(asUnsigned(maxChannelsOut * maxFramesOut))
  {}

  // reference :49-72
  template <typename F>
  void process(index windowSizeIn, index windowSizeOut, index hopSize, FluidContext& c, F processFunc)
  {
    assert(windowSizeIn <= maxWindowSizeIn() && "Window in bigger than maximum");
    assert(windowSizeOut <= maxWindowSizeOut() && "Window out bigger than maximum");
    for (; mFrameTime < mHostSize; mFrameTime += hopSize)
    {
      RealMatrixView windowIn{mFrameIn.data(), 0, channelsIn(), windowSizeIn};
      RealMatrixView windowOut{mFrameOut.data(), 0, channelsOut(), windowSizeOut};
      mSource.pull(windowIn, mFrameTime);
      processFunc(windowIn, windowOut);
      mSink.push(windowOut, mFrameTime);
      if (FluidTask* t = c.task())
        if (!t->processUpdate(static_cast<double>(std::min(mFrameTime + hopSize, mHostSize)), static_cast<double>(mHostSize))) break;
    }
    mFrameTime = mFrameTime < mHostSize ? mFrameTime : mFrameTime - mHostSize;
  }

  // reference :74-93
  template <typename F>
  void processInput(index windowSize, index hopSize, FluidContext& c, F processFunc)
  {
    assert(windowSize <= maxWindowSizeIn() && "Window bigger than maximum");
    for (; mFrameTime < mHostSize; mFrameTime += hopSize)
    {
      RealMatrixView windowIn{mFrameIn.data(), 0, channelsIn(), windowSize};
      mSource.pull(windowIn, mFrameTime);
      processFunc(windowIn);
      if (FluidTask* t = c.task())
        if (!t->processUpdate(static_cast<double>(std::min(mFrameTime + hopSize, mHostSize)), static_cast<double>(mHostSize))) break;
    }
    mFrameTime = mFrameTime < mHostSize ? mFrameTime : mFrameTime - mHostSize;
  }

  // Additive (device batching): the start times of the frames that fall due in this host block, and the time carried into
  // the next block -- exactly the values `mFrameTime` takes in the loops above.
  std::vector<index> dueFrameTimes(index hopSize) const
  {
    std::vector<index> t;
    for (index ft = mFrameTime; ft < mHostSize; ft += hopSize) t.push_back(ft);
    return t;
  }
  void advance(index hopSize)
  {
    for (; mFrameTime < mHostSize; mFrameTime += hopSize) {}
    mFrameTime = mFrameTime < mHostSize ? mFrameTime : mFrameTime - mHostSize;
  }
  void pullFrame(RealMatrixView windowIn, index frameTime) { mSource.pull(windowIn, frameTime); }
  void pushFrame(RealMatrixView windowOut, index frameTime) { mSink.push(windowOut, frameTime); }

  index hostSize() const noexcept { return mHostSize; }
  void  hostSize(index size) noexcept
  {
    assert(size <= mMaxHostSize);
    mHostSize = size;
    mSource.setHostBufferSize(size);
    mSink.setHostBufferSize(size);
    reset();
  }
  index maxWindowSizeIn() const noexcept { return mSource.size(); }
  index maxWindowSizeOut() const noexcept { return mSink.size(); }
  index channelsIn() const noexcept { return mSource.channels(); }
  index channelsOut() const noexcept { return mSink.channels(); }

  template <typename T>
  void push(const std::vector<HostVector<T>>& in) { mSource.push(in); }
  template <typename T>
  void push(HostMatrix<T> in) { mSource.push(in); }
  template <typename T>
  void pull(HostMatrix<T> out) { mSink.pull(out); }

  void reset()
  {
    mSource.reset();
    mSink.reset();
    mFrameTime = 0;
  }

private:
  index               mFrameTime = 0;
  index               mHostSize;
  index               mMaxHostSize;
  FluidSource<double> mSource;
  FluidSink<double>   mSink;
  std::vector<double> mFrameIn, mFrameOut;
};

// One audio channel in, `channelsOut` channels out (all the NMF clients need).  batchFunc(in, nFrames, out):
//   in  float [nFrames][win]   raw frames as pulled from the source (the device applies the analysis window)
//   out float [nFrames][channelsOut][win]   frames ready to overlap-add (inverse FFT, synthesis window, 1/fft applied)
template <bool Normalise = true>
class STFTBufferedProcess
{
public:
  STFTBufferedProcess(FFTParams fftParams, index channelsIn, index channelsOut, index hostVectorSize,
                      Allocator& alloc = FluidDefaultAllocator())
      : mBufferedProcess(fftParams.max(), fftParams.max(), channelsIn, channelsOut + Normalise, hostVectorSize, alloc),
        mFrameAndWindow(asUnsigned((Normalise + channelsOut) * std::max<index>(fftParams.max(), hostVectorSize)))
  {
    assert(channelsIn == 1 && "the B200 mirror batches one input channel");
  }

  // reference :187-241
  template <typename T, typename F>
  void process(FFTParams p, const std::vector<HostVector<T>>& input, std::vector<HostVector<T>>& output, FluidContext& c,
               F&& batchFunc)
  {
    if (!input[0].data()) return;
    assert(mBufferedProcess.channelsIn() == asSigned(input.size()));
    assert(mBufferedProcess.channelsOut() == asSigned(output.size() + Normalise));
    const index win = p.winSize(), hop = p.hopSize();
    const index chansOut = mBufferedProcess.channelsOut() - Normalise;
    setup(p);
    mBufferedProcess.push(input);

    const std::vector<index> times = mBufferedProcess.dueFrameTimes(hop);
    const index              nf = asSigned(times.size());
    if (nf > 0)
    {
      std::vector<double> frame(asUnsigned(win));
      mIn.resize(asUnsigned(nf * win));
      mOut.resize(asUnsigned(nf * chansOut * win));
      for (index i = 0; i < nf; ++i)
      {
        mBufferedProcess.pullFrame(RealMatrixView{frame.data(), 0, 1, win}, times[asUnsigned(i)]);
        for (index j = 0; j < win; ++j) mIn[asUnsigned(i * win + j)] = static_cast<float>(frame[asUnsigned(j)]);
      }
      batchFunc(mIn.data(), nf, mOut.data());
      std::vector<double> out(asUnsigned((chansOut + Normalise) * win));
      for (index i = 0; i < nf; ++i)
      {
        for (index ch = 0; ch < chansOut; ++ch)
          for (index j = 0; j < win; ++j)
            out[asUnsigned(ch * win + j)] = static_cast<double>(mOut[asUnsigned((i * chansOut + ch) * win + j)]);
        if (Normalise) // :219-224: the analysis window times the synthesis window rides along as an extra channel
          for (index j = 0; j < win; ++j) out[asUnsigned(chansOut * win + j)] = mWindow[asUnsigned(j)] * mWindow[asUnsigned(j)];
        mBufferedProcess.pushFrame(RealMatrixView{out.data(), 0, chansOut + Normalise, win}, times[asUnsigned(i)]);
        if (FluidTask* t = c.task()) // :65-69
          t->processUpdate(static_cast<double>(std::min(times[asUnsigned(i)] + hop, mBufferedProcess.hostSize())),
                           static_cast<double>(mBufferedProcess.hostSize()));
      }
    }
    mBufferedProcess.advance(hop);

    const index    block = input[0].size();
    RealMatrixView unnormalisedFrame{mFrameAndWindow.data(), 0, Normalise + chansOut, block};
    mBufferedProcess.pull(unnormalisedFrame);
    for (index i = 0; i < chansOut; ++i)
    {
      if (Normalise) // :231-237
        for (index j = 0; j < block; ++j)
        {
          double&      x = unnormalisedFrame(i, j);
          const double g = unnormalisedFrame(chansOut, j);
          if (x != 0) x /= (g > 0) ? g : 1;
        }
      if (output[asUnsigned(i)].data())
        for (index j = 0; j < block; ++j) output[asUnsigned(i)](j) = static_cast<T>(unnormalisedFrame(i, j));
    }
  }

  // reference :243-263
  template <typename T, typename F>
  void processInput(FFTParams p, const std::vector<HostVector<T>>& input, FluidContext& c, F&& batchFunc)
  {
    if (!input[0].data()) return;
    assert(mBufferedProcess.channelsIn() == asSigned(input.size()));
    const index win = p.winSize(), hop = p.hopSize();
    setup(p);
    mBufferedProcess.push(input);
    const std::vector<index> times = mBufferedProcess.dueFrameTimes(hop);
    const index              nf = asSigned(times.size());
    if (nf > 0)
    {
      std::vector<double> frame(asUnsigned(win));
      mIn.resize(asUnsigned(nf * win));
      for (index i = 0; i < nf; ++i)
      {
        mBufferedProcess.pullFrame(RealMatrixView{frame.data(), 0, 1, win}, times[asUnsigned(i)]);
        for (index j = 0; j < win; ++j) mIn[asUnsigned(i * win + j)] = static_cast<float>(frame[asUnsigned(j)]);
      }
      batchFunc(mIn.data(), nf);
      if (FluidTask* t = c.task())
        t->processUpdate(static_cast<double>(mBufferedProcess.hostSize()), static_cast<double>(mBufferedProcess.hostSize()));
    }
    mBufferedProcess.advance(hop);
  }

  void reset() { mBufferedProcess.reset(); }

private:
  void setup(FFTParams p)
  {
    const index win = p.winSize();
    if (asSigned(mWindow.size()) != win)
    { // WindowFuncs.hpp:41-45, periodic Hann (the NMF clients never choose another window)
      mWindow.resize(asUnsigned(win));
      for (index i = 0; i < win; ++i)
        mWindow[asUnsigned(i)] = 0.5 - 0.5 * std::cos((3.14159265358979323846 * 2.0 * static_cast<double>(i)) / static_cast<double>(win));
    }
  }

  BufferedProcess     mBufferedProcess;
  std::vector<double> mFrameAndWindow;
  std::vector<double> mWindow;
  std::vector<float>  mIn, mOut;
};
} // namespace client
} // namespace fluid
