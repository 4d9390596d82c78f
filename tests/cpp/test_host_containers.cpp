// CPU test of the host-side containers: the behaviours the reference pins in tests/data/TestFluidTensor.cpp,
// TestFluidTensorView.cpp, TestFluidTensorSupport.cpp (construction, row-major layout, views, slices, transposes, deep
// copies, iteration order) restated as plain asserts.  Built and run by tests/test_host_cpp.py.
#include <flucoma/data/TensorTypes.hpp>
#include <flucoma/clients/common/BufferAdaptor.hpp>
#include <flucoma/clients/common/ParameterTypes.hpp>
#include <cstdio>
#include <numeric>
#include <vector>

#define CHECK(x)                                                                    \
  do {                                                                              \
    if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); return 1; } \
  } while (0)

using namespace fluid;

int main()
{
  { // construction + row-major layout
    FluidTensor<double, 2> t(3, 4);
    CHECK(t.rows() == 3 && t.cols() == 4 && t.size() == 12);
    std::iota(t.begin(), t.end(), 0.0);
    CHECK(t(1, 2) == 6.0 && t(2, 3) == 11.0);
    CHECK(t.descriptor().strides[0] == 4 && t.descriptor().strides[1] == 1);
    FluidTensor<double, 2> u{{1, 2, 3}, {4, 5, 6}};
    CHECK(u.rows() == 2 && u.cols() == 3 && u(1, 0) == 4);
    FluidTensor<double, 1> v{{1, 0, 1, 0}};
    CHECK(v.size() == 4 && v(2) == 1);
    FluidTensor<int, 3> w(2, 3, 4);
    CHECK(w.size() == 24 && w.descriptor().strides[0] == 12);
  }
  { // row / col / transpose views alias the storage
    FluidTensor<double, 2> t(3, 4);
    std::iota(t.begin(), t.end(), 0.0);
    auto r = t.row(1);
    CHECK(r.size() == 4 && r(0) == 4 && r(3) == 7);
    auto c = t.col(2);
    CHECK(c.size() == 3 && c(0) == 2 && c(2) == 10);
    c(1) = 99;
    CHECK(t(1, 2) == 99);
    auto tt = t.transpose();
    CHECK(tt.rows() == 4 && tt.cols() == 3 && tt(2, 1) == 99 && tt(3, 2) == 11);
    std::vector<double> order;
    for (auto x : tt) order.push_back(x); // logical row-major order of the transposed view
    CHECK(order[0] == 0 && order[1] == 4 && order[2] == 8 && order[3] == 1);
  }
  { // slices, with the -1 = "all" convention and strides
    FluidTensor<double, 2> t(4, 6);
    std::iota(t.begin(), t.end(), 0.0);
    auto s = t(Slice(1, 2), Slice(0));
    CHECK(s.rows() == 2 && s.cols() == 6 && s(0, 0) == 6 && s(1, 5) == 17);
    auto e = t(Slice(0), Slice(0, 3, 2));
    CHECK(e.cols() == 3 && e(0, 1) == 2 && e(3, 2) == 22);
    auto m = t(2, Slice(1, 3));
    CHECK(m.rows() == 1 && m.cols() == 3 && m(0, 0) == 13);
    FluidTensor<double, 1> v(10);
    std::iota(v.begin(), v.end(), 0.0);
    auto vs = v(Slice(2, 4));
    CHECK(vs.size() == 4 && vs(0) == 2 && vs(3) == 5);
  }
  { // deep copy <<= between views/tensors, with conversion (float <-> double as BufferAdaptor needs)
    FluidTensor<double, 2> a(2, 3);
    std::iota(a.begin(), a.end(), 1.0);
    FluidTensor<float, 2> b(2, 3);
    FluidTensorView<float, 2>(b) <<= FluidTensorView<double, 2>(a);
    CHECK(b(1, 2) == 6.0f);
    FluidTensor<double, 2> c(3, 2);
    FluidTensorView<double, 2>(c) <<= a.transpose();
    CHECK(c(2, 1) == 6 && c(0, 1) == 4);
    FluidTensor<double, 1> col(2);
    col <<= a.col(1);
    CHECK(col(0) == 2 && col(1) == 5);
    FluidTensor<double, 2> d{FluidTensorView<double, 2>(a)};
    CHECK(d == a);
    d(0, 0) = 42;
    CHECK(d != a);
  }
  { // pointer/start/dims view ctor, new-axis ctor, null views (NMF.hpp:94-95 default arguments)
    std::vector<double> buf(12);
    std::iota(buf.begin(), buf.end(), 0.0);
    FluidTensorView<double, 2> v(buf.data(), 2, 2, 5);
    CHECK(v(0, 0) == 2 && v(1, 4) == 11 && v.data() == buf.data() + 2);
    FluidTensorView<double, 1> r(buf.data(), 0, 4);
    FluidTensorView<double, 2> up(r);
    CHECK(up.rows() == 1 && up.cols() == 4 && up(0, 3) == 3);
    FluidTensorView<double, 2> nul(nullptr, 0, 0, 0);
    CHECK(nul.data() == nullptr && nul.extent(0) == 0 && nul.extent(1) == 0);
    FluidTensorView<double, 1> nv{nullptr, 0, 0};
    CHECK(nv.data() == nullptr);
  }
  { // apply / fill / resize
    FluidTensor<double, 2> t(2, 2);
    t.fill(3);
    t.apply([](double& x) { x *= 2; });
    CHECK(t(1, 1) == 6);
    t.resize(3, 5);
    CHECK(t.rows() == 3 && t.cols() == 5 && t.size() == 15);
    FluidTensorView<double, 2>(t).row(0).fill(7);
    CHECK(t(0, 4) == 7);
  }
  { // MemoryBufferAdaptor: frames x chans interleaved, channel views (MemoryBufferAdaptor.hpp:92-112)
    using namespace fluid::client;
    MemoryBufferAdaptor buf(2, 5, 44100.0);
    {
      BufferAdaptor::Access a(&buf);
      CHECK(a.valid() && a.exists() && a.numFrames() == 5 && a.numChans() == 2);
      auto ch1 = a.samps(1);
      for (index i = 0; i < 5; ++i) ch1(i) = float(i);
      CHECK(buf.data()(3, 1) == 3.0f && buf.data()(3, 0) == 0.0f);
      auto part = a.samps(1, 3, 1);
      CHECK(part.size() == 3 && part(0) == 1.0f && part(2) == 3.0f);
      CHECK(a.allFrames().rows() == 2 && a.allFrames().cols() == 5);
      CHECK(a.resize(7, 3, 48000.0).ok());
      CHECK(a.numFrames() == 7 && a.numChans() == 3 && a.sampleRate() == 48000.0);
    }
    index nf = -1, nc = -1;
    CHECK(bufferRangeCheck(&buf, 0, nf, 0, nc).ok() && nf == 7 && nc == 3);
    nf = 10;
    CHECK(!bufferRangeCheck(&buf, 0, nf, 0, nc).ok());
    CHECK(!bufferRangeCheck(nullptr, 0, nf, 0, nc).ok());
    FFTParams p(1024, -1, -1);
    CHECK(p.fftSize() == 1024 && p.hopSize() == 512 && p.frameSize() == 513);
    FFTParams q(1000, 256, -1);
    CHECK(q.fftSize() == 1024 && q.hopSize() == 256);
  }
  std::printf("host containers ok\n");
  return 0;
}
