// GPU test of the BufSTFT host client mirror (flucoma/clients/nrt/BufSTFTClient.hpp): forward then inverse for the three
// padding modes, shapes and sample rates as the reference sets them, error messages; dumps source / magnitude / phase /
// resynth of mode 1 to argv[1] for tests/test_host_cpp.py to compare with the oracle.
#include <flucoma/clients/nrt/BufSTFTClient.hpp>
#include <cmath>
#include <cstdio>

#define CHECK(x)                                                                    \
  do {                                                                              \
    if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); return 1; } \
  } while (0)

using namespace fluid;
using namespace fluid::client;

int main(int argc, char** argv)
{
  const index n = 5000, win = 256, hop = 64, bins = 129;
  auto        src = std::make_shared<MemoryBufferAdaptor>(2, n, 44100.0);
  for (index i = 0; i < n; ++i)
  {
    double t = double(i) / 44100.0;
    src->data()(i, 0) = float(0.5 * std::sin(2 * M_PI * 440.0 * t) + 0.2 * std::sin(2 * M_PI * 3000.0 * t + 1.0));
    src->data()(i, 1) = float(0.3 * std::sin(2 * M_PI * 1234.5 * t) * ((i / 700) % 2));
  }
  FluidContext ctx;
  for (index mode = 0; mode < 3; ++mode)
  {
    auto mag = std::make_shared<MemoryBufferAdaptor>(1, 1, 1.0);
    auto phase = std::make_shared<MemoryBufferAdaptor>(1, 1, 1.0);
    auto res = std::make_shared<MemoryBufferAdaptor>(1, 1, 1.0);
    bufstft::BufSTFTParams p;
    p.source = src; p.startChan = 1; p.magnitude = mag; p.phase = phase; p.padding = mode;
    p.fftSettings = FFTParams(win, hop, -1);
    bufstft::BufferSTFTClient fwd(p, ctx);
    Result                    r = fwd.process<float>(ctx);
    if (!r.ok()) std::printf("BufSTFT fwd: %s\n", r.message().c_str());
    CHECK(r.ok());
    const index pad = mode == 0 ? 0 : (mode == 1 ? win / 2 : win - hop);
    index       padded = n + 2 * pad;
    if (mode == 2) padded = (padded + hop - 1) / hop * hop;
    const index hops = 1 + (padded - win) / hop; // BufSTFTClient.hpp:121-131
    CHECK(mag->data().rows() == hops && mag->data().cols() == bins);
    CHECK(phase->data().rows() == hops && phase->data().cols() == bins);
    bufstft::BufSTFTParams q;
    q.magnitude = mag; q.phase = phase; q.resynth = res; q.inverse = 1; q.padding = mode;
    q.fftSettings = FFTParams(win, hop, -1);
    bufstft::BufferSTFTClient inv(q, ctx);
    r = inv.process<float>(ctx);
    if (!r.ok()) std::printf("BufSTFT inv: %s\n", r.message().c_str());
    CHECK(r.ok());
    CHECK(res->data().rows() == (hops - 1) * hop + win - pad && res->data().cols() == 1); // :241-246
    double err = 0; // Hann at 75 % overlap: interior samples come back
    for (index i = win; i < n - win; ++i) err = std::max(err, std::abs(double(res->data()(i, 0)) - double(src->data()(i, 1))));
    CHECK(err < 1e-4);
    if (mode == 1 && argc > 1)
    {
      FILE* f = std::fopen(argv[1], "wb");
      CHECK(f);
      auto dump = [f](FluidTensor<float, 2>& t) {
        int64_t hdr[2] = {t.rows(), t.cols()};
        std::fwrite(hdr, sizeof(int64_t), 2, f);
        std::fwrite(t.data(), sizeof(float), size_t(t.size()), f);
      };
      dump(src->data()); dump(mag->data()); dump(phase->data()); dump(res->data());
      std::fclose(f);
    }
  }
  { // error paths keep the reference's messages (:86, :94-96, :201-203)
    bufstft::BufSTFTParams p;
    bufstft::BufferSTFTClient c0(p, ctx);
    CHECK(c0.process<float>(ctx).status() == Result::Status::kError);
    p.source = src;
    bufstft::BufferSTFTClient c1(p, ctx);
    Result r = c1.process<float>(ctx);
    CHECK(r.status() == Result::Status::kError && r.message().find("Neither magnitude nor phase") != std::string::npos);
    p.inverse = 1;
    bufstft::BufferSTFTClient c2(p, ctx);
    CHECK(c2.process<float>(ctx).status() == Result::Status::kError);
  }
  std::printf("bufstft client ok\n");
  return 0;
}
