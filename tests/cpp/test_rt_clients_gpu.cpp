// GPU test of the streaming client mirrors (clients/rt/NMFFilterClient.hpp, NMFMatchClient.hpp, common/BufferedProcess.hpp):
// a mono stream is pushed through the clients in host blocks of argv[2] samples, the way a host's audio callback would;
// outputs go to argv[1] for tests/test_host_cpp.py to compare with the oracle's stream simulation.
// Ring-buffer semantics are checked here on the CPU side first (tests/clients/common/TestFluidSource.cpp:39-55,
// TestBufferedProcess.cpp:20-70 restated).
#include <flucoma/clients/rt/NMFFilterClient.hpp>
#include <flucoma/clients/rt/NMFMatchClient.hpp>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                    \
  do {                                                                              \
    if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); return 1; } \
  } while (0)

using namespace fluid;
using namespace fluid::client;

int main(int argc, char** argv)
{
  { // FluidSource: a window is the input delayed by its own length (TestFluidSource.cpp:39-55)
    const index hostSize = 64, frameSize = 100;
    FluidSource<double> src(frameSize, 1, hostSize);
    src.reset();
    std::vector<double> data(1000);
    for (index i = 0; i < 1000; ++i) data[asUnsigned(i)] = double(i + 1);
    FluidTensor<double, 2> out(1, frameSize);
    for (index blk = 0; blk + hostSize <= 1000; blk += hostSize)
    {
      FluidTensorView<double, 2> in(data.data(), blk, 1, hostSize);
      src.push(in);
      src.pull(out, 0); // the frame that ends at the START of this block
      for (index j = 0; j < frameSize; ++j)
      {
        const index t = blk - frameSize + j;
        CHECK(out(0, j) == (t < 0 ? 0.0 : double(t + 1)));
      }
    }
  }
  { // BufferedProcess + Hann^2 normalisation: identity processing reconstructs the input delayed by one window
    // (TestBufferedProcess.cpp:20-70: COLA with window = hann, hop = win / 4)
    const index win = 128, hop = 32, host = 50, n = 2000;
    BufferedProcess bp(win, win, 1, 2, host);
    bp.hostSize(host);
    std::vector<double> w(asUnsigned(win));
    for (index i = 0; i < win; ++i) w[asUnsigned(i)] = 0.5 - 0.5 * std::cos(2.0 * M_PI * double(i) / double(win));
    std::vector<double> in(asUnsigned(n)), out(asUnsigned(n));
    for (index i = 0; i < n; ++i) in[asUnsigned(i)] = std::sin(0.05 * double(i)) + 0.2;
    FluidContext c;
    std::vector<double> blockOut(asUnsigned(2 * host));
    for (index blk = 0; blk + host <= n; blk += host)
    {
      FluidTensorView<double, 2> bin(in.data(), blk, 1, host);
      bp.push(bin);
      bp.process(win, win, hop, c, [&](RealMatrixView fin, RealMatrixView fout) {
        for (index j = 0; j < win; ++j)
        {
          fout(0, j) = fin(0, j) * w[asUnsigned(j)] * w[asUnsigned(j)];
          fout(1, j) = w[asUnsigned(j)] * w[asUnsigned(j)];
        }
      });
      FluidTensorView<double, 2> bout(blockOut.data(), 0, 2, host);
      bp.pull(bout);
      for (index j = 0; j < host; ++j) out[asUnsigned(blk + j)] = bout(1, j) > 0 ? bout(0, j) / bout(1, j) : bout(0, j);
    }
    double err = 0;
    for (index i = 2 * win; i + host < n; ++i) err = std::max(err, std::abs(out[asUnsigned(i)] - in[asUnsigned(i - win)]));
    CHECK(err < 1e-12);
  }
  if (argc < 3) { std::printf("rt clients cpu ok\n"); return 0; }

  // ---- device part: stream through NMFFilter and NMFMatch in host blocks of argv[2] samples
  const index host = std::atoi(argv[2]);
  const index n = 6000, win = 256, hop = 64, rank = 5, bins = win / 2 + 1;
  std::vector<float> audio(asUnsigned(n));
  for (index i = 0; i < n; ++i)
  {
    const double t = double(i) / 44100.0;
    audio[asUnsigned(i)] = float(0.4 * ((i / 1500) % 2) * std::sin(2 * M_PI * 440.0 * t) + 0.3 * (((i / 1000) + 1) % 2) * std::sin(2 * M_PI * 1500.0 * t) +
                                 0.05 * std::sin(2 * M_PI * 3300.0 * t));
  }
  auto bases = std::make_shared<MemoryBufferAdaptor>(rank, bins, 44100.0);
  for (index k = 0; k < rank; ++k)
    for (index b = 0; b < bins; ++b)
      bases->data()(b, k) = float(0.01 + std::exp(-0.5 * std::pow((double(b) - 3.0 - 17.0 * double(k)) / (2.0 + double(k)), 2)));

  FluidContext ctx;
  ctx.hostVectorSize(host);
  nmffilter::NMFFilterParams fp;
  fp.bases = bases; fp.maxComponents = rank + 2; fp.iterations = 10; fp.seed = 42; fp.fftSettings = FFTParams(win, hop, -1);
  nmffilter::NMFFilterClient filter(fp, ctx);
  nmfmatch::NMFMatchParams mp;
  mp.bases = bases; mp.maxComponents = rank + 2; mp.seed = 42; mp.fftSettings = FFTParams(win, hop, -1);
  nmfmatch::NMFMatchClient match(mp, ctx);
  CHECK(filter.latency() == win && match.latency() == win);

  const index nblocks = n / host;
  std::vector<float> fout(asUnsigned((rank + 2) * nblocks * host)), mout(asUnsigned(nblocks * (rank + 2)));
  std::vector<float> block(asUnsigned(host)), ctl(asUnsigned(rank + 2));
  std::vector<std::vector<float>> outs(asUnsigned(rank + 2), std::vector<float>(asUnsigned(host)));
  for (index blk = 0; blk < nblocks; ++blk)
  {
    for (index j = 0; j < host; ++j) block[asUnsigned(j)] = audio[asUnsigned(blk * host + j)];
    std::vector<HostVector<float>> in{HostVector<float>(block.data(), 0, host)};
    std::vector<HostVector<float>> out;
    for (auto& o : outs) out.emplace_back(o.data(), 0, host);
    filter.process(in, out, ctx);
    for (index ch = 0; ch < rank + 2; ++ch)
      for (index j = 0; j < host; ++j) fout[asUnsigned((ch * nblocks + blk) * host + j)] = outs[asUnsigned(ch)][asUnsigned(j)];
    std::vector<HostVector<float>> cout{HostVector<float>(ctl.data(), 0, rank + 2)};
    match.process(in, cout, ctx);
    for (index k = 0; k < rank + 2; ++k) mout[asUnsigned(blk * (rank + 2) + k)] = ctl[asUnsigned(k)];
  }
  FILE* f = std::fopen(argv[1], "wb");
  CHECK(f);
  auto dump = [f](const float* p, int64_t r, int64_t c) {
    int64_t hdr[2] = {r, c};
    std::fwrite(hdr, sizeof(int64_t), 2, f);
    std::fwrite(p, sizeof(float), size_t(r * c), f);
  };
  dump(audio.data(), 1, n);
  std::vector<float> W(asUnsigned(rank * bins));
  for (index k = 0; k < rank; ++k)
    for (index b = 0; b < bins; ++b) W[asUnsigned(k * bins + b)] = bases->data()(b, k);
  dump(W.data(), rank, bins);
  dump(fout.data(), rank + 2, nblocks * host);
  dump(mout.data(), nblocks, rank + 2);
  std::fclose(f);
  std::printf("rt clients ok\n");
  return 0;
}
