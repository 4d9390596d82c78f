// GPU test of the MelBands and HPSS host mirrors: per-frame (streaming) calls against the batched calls, and dumps of the
// batched outputs to argv[1] for tests/test_host_cpp.py to compare with the oracle.
#include <flucoma/algorithms/public/HPSS.hpp>
#include <flucoma/algorithms/public/MelBands.hpp>
#include <cmath>
#include <cstdio>
#include <vector>

#define CHECK(x)                                                                    \
  do {                                                                              \
    if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); return 1; } \
  } while (0)

using namespace fluid;
using namespace fluid::algorithm;

int main(int argc, char** argv)
{
  const index F = 60, B = 129; // fft 256
  FluidTensor<std::complex<double>, 2> S(F, B);
  FluidTensor<double, 2>               M(F, B);
  unsigned                             s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return double(s >> 8) / double(1u << 24); };
  for (index f = 0; f < F; ++f)
    for (index b = 0; b < B; ++b) {
      const double mag = (0.05 + std::pow(rnd(), 3.0)) * (1.0 + (b % 16 == 3 ? 4.0 : 0.0)) * (f % 11 == 5 ? 5.0 : 1.0);
      const double ph = 6.283185307179586 * rnd();
      S(f, b) = std::polar(mag, ph);
      M(f, b) = mag;
    }
  // ---- MelBands: per-frame == batched
  MelBands mel(40, 256);
  mel.init(20.0, 20000.0, 13, B, 44100.0, 256);
  FluidTensor<double, 2> bands(F, 13), bands1(F, 13);
  mel.processFrames(M, bands, true, false, false);
  for (index f = 0; f < F; ++f) mel.processFrame(M.row(f), bands1.row(f), true, false, false);
  for (index f = 0; f < F; ++f)
    for (index k = 0; k < 13; ++k) CHECK(std::abs(bands(f, k) - bands1(f, k)) <= 1e-6 * std::abs(bands(f, k)) + 1e-12);
  // ---- HPSS: streaming processFrame == batched from init state (hSize 9: state reaches back 14 frames)
  const index vSize = 7, hSize = 9;
  HPSS                                 hp(256, 17);
  FluidTensor<std::complex<double>, 3> all(3, F, B);
  hp.init(B, hSize);
  hp.processFrames(S, all, vSize, hSize, 0, 0, 1, 1, 1, 0, 1, 1, 1);
  hp.init(B, hSize);
  FluidTensor<std::complex<double>, 2> one(B, 3);
  double                               worst = 0, scale = 0;
  for (index f = 0; f < F; ++f) {
    hp.processFrame(S.row(f), one, vSize, hSize, 0, 0, 1, 1, 1, 0, 1, 1, 1);
    for (index c = 0; c < 3; ++c)
      for (index b = 0; b < B; ++b) {
        worst = std::max(worst, std::abs(one(b, c) - all(c, f, b)));
        scale = std::max(scale, std::abs(all(c, f, b)));
      }
  }
  CHECK(scale > 0.1 && worst <= 1e-6 * scale);
  if (argc > 1) {
    FILE* fp = std::fopen(argv[1], "wb");
    CHECK(fp);
    auto dump = [fp](const std::vector<float>& v, long long r, long long c) {
      long long hdr[2] = {r, c};
      std::fwrite(hdr, sizeof(long long), 2, fp);
      std::fwrite(v.data(), sizeof(float), v.size(), fp);
    };
    std::vector<float> m, bn, sp, hh;
    for (index f = 0; f < F; ++f) for (index b = 0; b < B; ++b) m.push_back(float(M(f, b)));
    for (index f = 0; f < F; ++f) for (index k = 0; k < 13; ++k) bn.push_back(float(bands(f, k)));
    for (index f = 0; f < F; ++f) for (index b = 0; b < B; ++b) { sp.push_back(float(S(f, b).real())); sp.push_back(float(S(f, b).imag())); }
    for (index c = 0; c < 3; ++c) for (index f = 0; f < F; ++f) for (index b = 0; b < B; ++b) { hh.push_back(float(all(c, f, b).real())); hh.push_back(float(all(c, f, b).imag())); }
    dump(m, F, B); dump(bn, F, 13); dump(sp, F, 2 * B); dump(hh, 3 * F, 2 * B);
    std::fclose(fp);
  }
  std::printf("spectral mirrors ok\n");
  return 0;
}
