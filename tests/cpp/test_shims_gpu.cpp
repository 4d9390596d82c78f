// GPU test of the C++ host mirror: the reference's own NMF tests (tests/algorithms/public/TestNMF.cpp:11-73) compiled
// against OUR headers, plus STFT/ISTFT and the BufNMF client.  Writes the BufNMF outputs to argv[1] so that
// tests/test_host_cpp.py can compare them with the oracle.  Needs libflucoma_b200.so (FLUCOMA_B200_LIB) and a GPU.
#include <flucoma/algorithms/public/NMF.hpp>
#include <flucoma/algorithms/public/RatioMask.hpp>
#include <flucoma/algorithms/public/STFT.hpp>
#include <flucoma/clients/nrt/NMFClient.hpp>
#include <flucoma/data/FluidTensor.hpp>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#define CHECK(x)                                                                    \
  do {                                                                              \
    if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); return 1; } \
  } while (0)

using namespace fluid;

template <class A, class B>
bool rangeEquals(const A& a, const B& b)
{
  return std::equal(a.begin(), a.end(), b.begin(), b.end());
}

int main(int argc, char** argv)
{
  { // TestNMF.cpp:11-46
    using algorithm::NMF;
    using Tensor = FluidTensor<double, 2>;
    NMF    algo;
    Tensor input{{1, 2, 3}, {4, 5, 6}, {7, 8, 9}};
    std::vector<Tensor> Vs(4, Tensor(3, 3));
    std::vector<Tensor> Ws(4, Tensor(2, 3));
    std::vector<Tensor> Hs(4, Tensor(3, 2));
    algo.process(input, Ws[0], Hs[0], Vs[0], 2, 1, true, true, 42);
    algo.process(input, Ws[1], Hs[1], Vs[1], 2, 1, true, true, 42);
    algo.process(input, Ws[2], Hs[2], Vs[2], 2, 1, true, true, 5063);
    algo.process(input, Ws[3], Hs[3], Vs[3], 2, 1, true, true, 5063);
    CHECK(rangeEquals(Ws[1], Ws[0]) && rangeEquals(Hs[1], Hs[0]) && rangeEquals(Vs[1], Vs[0]));
    CHECK(rangeEquals(Ws[3], Ws[2]) && rangeEquals(Hs[3], Hs[2]) && rangeEquals(Vs[3], Vs[2]));
    CHECK(!rangeEquals(Ws[1], Ws[2]) && !rangeEquals(Hs[1], Hs[2]) && !rangeEquals(Vs[1], Vs[2]));
    // W rows unit norm after a W-update (NMF.hpp:162)
    for (index k = 0; k < 2; ++k)
    {
      double s = 0;
      for (index b = 0; b < 3; ++b) s += Ws[0](k, b) * Ws[0](k, b);
      CHECK(std::abs(std::sqrt(s) - 1.0) < 1e-5);
    }
    std::printf("W42 %.9g %.9g %.9g\n", Ws[0](0, 0), Ws[0](0, 1), Ws[0](1, 2));
  }
  { // TestNMF.cpp:48-73
    using algorithm::NMF;
    using Tensor = FluidTensor<double, 2>;
    using Vector = FluidTensor<double, 1>;
    NMF    algo;
    Vector input{{1, 0, 1, 0}};
    Tensor bases{{0, 0, 1, 0}, {1, 0, 0, 0}};
    Vector v(4);
    std::vector<Vector> outputs(3, Vector(2));
    algo.processFrame(input, bases, outputs[0], 0, v, 42, FluidDefaultAllocator());
    algo.processFrame(input, bases, outputs[1], 0, v, 42, FluidDefaultAllocator());
    algo.processFrame(input, bases, outputs[2], 0, v, 7863, FluidDefaultAllocator());
    CHECK(rangeEquals(outputs[1], outputs[0]));
    CHECK(!rangeEquals(outputs[1], outputs[2]));
    CHECK(std::abs(bases(0, 2) - 1.0) < 1e-6 && bases(0, 0) > 0); // W0 clamped + row-normalised in place (NMF.hpp:58-64)
  }
  { // progress callback + cancel (NMF.hpp:136-139,175-176)
    using algorithm::NMF;
    FluidTensor<double, 2> X(20, 12), W(3, 12), H(20, 3), V(20, 12);
    for (index i = 0; i < 20; ++i)
      for (index j = 0; j < 12; ++j) X(i, j) = 1.0 + ((i * 7 + j * 3) % 5);
    NMF   algo;
    index last = 0;
    algo.addProgressCallback([&last](index it) { last = it; return it < 3; });
    algo.process(X, W, H, V, 3, 10, true, true, 1);
    CHECK(last == 3);
    CHECK(rangeEquals(V, X)); // cancelled: V1 is X
  }
  { // STFT -> magnitude -> ISTFT identity (STFT.hpp)
    const index n = 8192, win = 512, hop = 128;
    FluidTensor<double, 1> audio(n), back(n);
    for (index i = 0; i < n; ++i) audio(i) = std::sin(0.01 * double(i)) + 0.3 * std::sin(0.37 * double(i));
    algorithm::STFT  stft(win, win, hop);
    algorithm::ISTFT istft(win, win, hop);
    index            nWin = (n + hop) / hop;
    FluidTensor<std::complex<double>, 2> spec(nWin, win / 2 + 1);
    FluidTensor<double, 2>               mag(nWin, win / 2 + 1);
    stft.process(audio, spec);
    algorithm::STFT::magnitude(spec, mag);
    CHECK(mag(10, 1) >= 0);
    istft.process(spec, back);
    double err = 0;
    for (index i = 0; i < n; ++i) err = std::max(err, std::abs(back(i) - audio(i)));
    CHECK(err < 1e-4);
    FluidTensor<std::complex<double>, 1> fr(win / 2 + 1);
    stft.processFrame(audio(Slice(hop * 8 - win / 2, win)), fr); // == frame 8 of the centred STFT
    double ferr = 0;
    for (index b = 0; b <= win / 2; ++b) ferr = std::max(ferr, std::abs(fr(b) - spec(8, b)));
    CHECK(ferr < 1e-3);
  }
  { // BufNMF client, 2 channels, resynthesis
    using namespace fluid::client;
    const index n = 6000, chans = 2, rank = 3;
    auto        src = std::make_shared<MemoryBufferAdaptor>(chans, n, 44100.0);
    for (index c = 0; c < chans; ++c)
      for (index i = 0; i < n; ++i)
      {
        double t = double(i) / 44100.0;
        double env1 = (i / 1500) % 2 ? 1.0 : 0.0, env2 = (i / 1000) % 2 ? 0.0 : 1.0;
        src->data()(i, c) = float(0.4 * env1 * std::sin(2 * M_PI * (440.0 + 200 * c) * t) + 0.3 * env2 * std::sin(2 * M_PI * 1500.0 * t) +
                                  0.001 * std::sin(12345.678 * t * (c + 1)));
      }
    auto bases = std::make_shared<MemoryBufferAdaptor>(1, 1, 44100.0);
    auto acts = std::make_shared<MemoryBufferAdaptor>(1, 1, 44100.0);
    auto res = std::make_shared<MemoryBufferAdaptor>(1, 1, 44100.0);
    bufnmf::BufNMFParams p;
    p.source = src; p.bases = bases; p.activations = acts; p.resynth = res; p.resynthMode = 1;
    p.components = rank; p.iterations = 30; p.seed = 7; p.fftSettings = FFTParams(256, 64, -1);
    FluidContext      ctx;
    bufnmf::NMFClient client(p, ctx);
    Result            r = client.process<float>(ctx);
    if (!r.ok()) std::printf("BufNMF: %s\n", r.message().c_str());
    CHECK(r.ok());
    CHECK(bases->data().rows() == 129 && bases->data().cols() == chans * rank);
    CHECK(acts->data().rows() == n / 64 + 1 && acts->data().cols() == chans * rank);
    CHECK(res->data().rows() == n && res->data().cols() == chans * rank);
    for (index c = 0; c < chans; ++c)
    { // 1/max(H) scaling per channel (NMFClient.hpp:289-298) and masks summing to one
      float mx = 0;
      for (index f = 0; f < acts->data().rows(); ++f)
        for (index k = 0; k < rank; ++k) mx = std::max(mx, acts->data()(f, c * rank + k));
      CHECK(std::abs(mx - 1.0f) < 1e-5f);
      double err = 0;
      for (index i = 0; i < n; ++i)
      {
        double s = 0;
        for (index k = 0; k < rank; ++k) s += res->data()(i, c * rank + k);
        err = std::max(err, std::abs(s - double(src->data()(i, c))));
      }
      CHECK(err < 1e-4);
    }
    if (argc > 1)
    {
      FILE* f = std::fopen(argv[1], "wb");
      CHECK(f);
      auto dump = [f](FluidTensor<float, 2>& t) {
        int64_t hdr[2] = {t.rows(), t.cols()};
        std::fwrite(hdr, sizeof(int64_t), 2, f);
        std::fwrite(t.data(), sizeof(float), size_t(t.size()), f);
      };
      dump(src->data()); dump(bases->data()); dump(acts->data()); dump(res->data());
      std::fclose(f);
    }
    // error paths keep the reference's messages
    bufnmf::BufNMFParams q = p;
    q.basesMode = 1; q.bases = nullptr;
    bufnmf::NMFClient c2(q, ctx);
    CHECK(c2.process<float>(ctx).status() == Result::Status::kError);
    // cancellation through the task (NMFClient.hpp:235-238, 273-274)
    FluidTask    task;
    FluidContext ctx2(task);
    task.cancel();
    bufnmf::NMFClient c3(p, ctx2);
    CHECK(c3.process<float>(ctx2).status() == Result::Status::kCancelled);
  }
  { // BufNMF at a shape the tcgen05 engine takes (fft 256, rank 16, >= 128 frames) WITH a FluidTask: the reference
    // always installs a progress callback (NMFClient.hpp:261-267); it must not push the job off the tensor-core engine
    using namespace fluid::client;
    const index n = 128 * 64, chans = 3, rank = 16;
    auto        src = std::make_shared<MemoryBufferAdaptor>(chans, n, 44100.0);
    for (index c = 0; c < chans; ++c)
      for (index i = 0; i < n; ++i)
      {
        double t = double(i) / 44100.0, s = 0;
        for (int h = 1; h <= 5; ++h) s += ((i / (700 * h)) % 2 ? 0.2 : 0.02) * std::sin(2 * M_PI * (300.0 * h + 50 * c) * t);
        src->data()(i, c) = float(s);
      }
    auto bases = std::make_shared<MemoryBufferAdaptor>(1, 1, 44100.0);
    auto acts = std::make_shared<MemoryBufferAdaptor>(1, 1, 44100.0);
    bufnmf::BufNMFParams p;
    p.source = src; p.bases = bases; p.activations = acts;
    p.components = rank; p.iterations = 40; p.seed = 3; p.fftSettings = FFTParams(256, 64, -1);
    FluidTask         task;
    FluidContext      ctx(task);
    bufnmf::NMFClient client(p, ctx);
    Result            r = client.process<float>(ctx);
    CHECK(r.ok());
    CHECK(client.lastStats().backend_used == FB200_BACKEND_TCGEN05);
    CHECK(std::abs(task.progress() - 1.0) < 1e-9); // every iteration was reported (FluidTask.hpp:27-33)
    // the same job without a task gives bit-identical buffers: the callback path runs the same single launch
    auto bases2 = std::make_shared<MemoryBufferAdaptor>(1, 1, 44100.0);
    auto acts2 = std::make_shared<MemoryBufferAdaptor>(1, 1, 44100.0);
    bufnmf::BufNMFParams q = p;
    q.bases = bases2; q.activations = acts2;
    FluidContext      ctxn;
    bufnmf::NMFClient c2(q, ctxn);
    CHECK(c2.process<float>(ctxn).ok());
    CHECK(rangeEquals(bases->data(), bases2->data()) && rangeEquals(acts->data(), acts2->data()));
  }
  std::printf("host shims ok\n");
  return 0;
}
