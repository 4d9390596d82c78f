// Host audio-buffer interface (reference: include/flucoma/clients/common/BufferAdaptor.hpp:18-208): float32 samples,
// channel views, RAII ReadAccess/Access (acquire/release, refresh on destruction), bufferRangeCheck.
// MemoryBufferAdaptor mirrors clients/common/MemoryBufferAdaptor.hpp:92-143 (frames x channels, interleaved).
#pragma once
#include "Result.hpp"
#include "../../data/FluidTensor.hpp"
#include <memory>
#include <string>

namespace fluid {
namespace client {

class BufferAdaptor
{
public:
  class ReadAccess
  {
  public:
    ReadAccess(const BufferAdaptor* adaptor) : mAdaptor(nullptr)
    {
      if (adaptor && adaptor->acquire()) mAdaptor = adaptor;
    }
    ~ReadAccess()
    {
      if (mAdaptor) mAdaptor->release();
    }
    ReadAccess(const ReadAccess&) = delete;
    ReadAccess& operator=(const ReadAccess&) = delete;

    bool   valid() const { return mAdaptor ? mAdaptor->valid() : false; }
    bool   exists() const { return mAdaptor ? mAdaptor->exists() : false; }
    index  numFrames() const { return mAdaptor ? mAdaptor->numFrames() : 0; }
    index  numChans() const { return mAdaptor ? mAdaptor->numChans() : 0; }
    double sampleRate() const { return mAdaptor ? mAdaptor->sampleRate() : 0; }
    FluidTensorView<const float, 2> allFrames() const { return mAdaptor->allFrames(); }
    FluidTensorView<const float, 1> samps(index channel) const { return mAdaptor->samps(channel); }
    FluidTensorView<const float, 1> samps(index offset, index nframes, index chanoffset) const
    {
      return mAdaptor->samps(offset, nframes, chanoffset);
    }

  private:
    const BufferAdaptor* mAdaptor;
  };

  class Access : public ReadAccess
  {
  public:
    Access(BufferAdaptor* adaptor) : ReadAccess(adaptor), mMutable(adaptor) {}
    ~Access()
    {
      if (mMutable) mMutable->refresh();
    }
    FluidTensorView<float, 2> allFrames() { return mMutable->allFrames(); }
    FluidTensorView<float, 1> samps(index channel) { return mMutable->samps(channel); }
    FluidTensorView<float, 1> samps(index offset, index nframes, index chanoffset)
    {
      return mMutable->samps(offset, nframes, chanoffset);
    }
    const Result resize(index frames, index channels, double sampleRate)
    {
      return mMutable ? mMutable->resize(frames, channels, sampleRate)
                      : Result{Result::Status::kError, "Trying to resize null buffer"};
    }

  private:
    BufferAdaptor* mMutable;
  };

  virtual ~BufferAdaptor() = default;

private:
  virtual bool         acquire() const = 0;
  virtual void         release() const = 0;
  virtual bool         valid() const = 0;
  virtual bool         exists() const = 0;
  virtual const Result resize(index frames, index channels, double sampleRate) = 0;
  virtual std::string  asString() const = 0;
  virtual FluidTensorView<float, 1>       samps(index channel) = 0;
  virtual FluidTensorView<float, 1>       samps(index offset, index nframes, index chanoffset) = 0;
  virtual FluidTensorView<const float, 1> samps(index channel) const = 0;
  virtual FluidTensorView<const float, 1> samps(index offset, index nframes, index chanoffset) const = 0;
  virtual FluidTensorView<float, 2>       allFrames() = 0;
  virtual FluidTensorView<const float, 2> allFrames() const = 0;
  virtual index                           numFrames() const = 0;
  virtual index                           numChans() const = 0;
  virtual double                          sampleRate() const = 0;
  virtual void                            refresh() {}
};

// BufferAdaptor.hpp:175-208
inline Result bufferRangeCheck(const BufferAdaptor* b, index startFrame, index& nFrames, index startChan, index& nChans)
{
  if (!b) return {Result::Status::kError, "Input buffer not set"};
  BufferAdaptor::ReadAccess in(b);
  if (!in.exists()) return {Result::Status::kError, "Input buffer not found."};
  if (!in.valid()) return {Result::Status::kError, "Input buffer invalid (possibly zero-size?)"};
  if (startFrame >= in.numFrames() || startFrame < 0)
    return {Result::Status::kError, "Input buffer invalid start frame ", startFrame};
  if (startChan >= in.numChans() || startChan < 0)
    return {Result::Status::kError, "Input buffer invalid start channel ", startChan};
  nFrames = nFrames < 0 ? in.numFrames() - startFrame : nFrames;
  if (nFrames <= 0 || nFrames > in.numFrames() - startFrame)
    return {Result::Status::kError, "Input buffer: not enough frames"};
  nChans = nChans < 0 ? in.numChans() - startChan : nChans;
  if (nChans <= 0 || nChans > in.numChans() - startChan)
    return {Result::Status::kError, "Input buffer: not enough channels"};
  return {Result::Status::kOk, ""};
}

class MemoryBufferAdaptor : public BufferAdaptor
{
public:
  MemoryBufferAdaptor(index chans, index frames, double sampleRate) : mData(frames, chans), mSampleRate(sampleRate), mExists(true) {}
  FluidTensor<float, 2>& data() { return mData; }

private:
  bool         acquire() const override { return true; }
  void         release() const override {}
  bool         valid() const override { return mData.rows() > 0 && mData.cols() > 0; }
  bool         exists() const override { return mExists; }
  const Result resize(index frames, index channels, double sampleRate) override
  {
    mData.resize(frames, channels);
    mData.fill(0);
    mSampleRate = sampleRate;
    return {};
  }
  std::string               asString() const override { return "MemoryBufferAdaptor"; }
  FluidTensorView<float, 1> samps(index channel) override { return FluidTensorView<float, 2>(mData).col(channel); }
  FluidTensorView<float, 1> samps(index offset, index nframes, index chanoffset) override
  {
    return FluidTensorView<float, 2>(mData)(Slice(offset, nframes), Slice(chanoffset, 1)).col(0);
  }
  FluidTensorView<const float, 1> samps(index channel) const override
  {
    return FluidTensorView<const float, 2>(mData).col(channel);
  }
  FluidTensorView<const float, 1> samps(index offset, index nframes, index chanoffset) const override
  {
    return FluidTensorView<const float, 2>(mData)(Slice(offset, nframes), Slice(chanoffset, 1)).col(0);
  }
  FluidTensorView<float, 2>       allFrames() override { return FluidTensorView<float, 2>(mData).transpose(); }
  FluidTensorView<const float, 2> allFrames() const override { return FluidTensorView<const float, 2>(mData).transpose(); }
  index                           numFrames() const override { return mData.rows(); }
  index                           numChans() const override { return mData.cols(); }
  double                          sampleRate() const override { return mSampleRate; }

  FluidTensor<float, 2> mData; // frames x channels
  double                mSampleRate;
  bool                  mExists;
};

} // namespace client
} // namespace fluid
