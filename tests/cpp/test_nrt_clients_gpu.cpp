// GPU test of the offline client mirrors added in round 2: NMFSeedClient -> (seeded) NMFClient, and NMFCrossClient.
// Writes the outputs to argv[1] for tests/test_host_cpp.py to compare with the oracle.
#include <flucoma/clients/nrt/NMFClient.hpp>
#include <flucoma/clients/nrt/NMFCrossClient.hpp>
#include <flucoma/clients/nrt/NMFSeedClient.hpp>
#include <cmath>
#include <cstdio>
#include <vector>

#define CHECK(x)                                                                    \
  do {                                                                              \
    if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); return 1; } \
  } while (0)

using namespace fluid;
using namespace fluid::client;

static float tone(index i, double f0, double f1, index gate)
{
  const double t = double(i) / 44100.0;
  return float(0.4 * ((i / gate) % 2) * std::sin(2 * M_PI * f0 * t) + 0.3 * (((i / (gate / 2 + 300)) + 1) % 2) * std::sin(2 * M_PI * f1 * t) +
               0.02 * std::sin(2 * M_PI * 3300.0 * t));
}

int main(int argc, char** argv)
{
  const index ns = 9000, nt = 7000;
  auto        src = std::make_shared<MemoryBufferAdaptor>(1, ns, 44100.0);
  auto        tgt = std::make_shared<MemoryBufferAdaptor>(1, nt, 44100.0);
  for (index i = 0; i < ns; ++i) src->data()(i, 0) = tone(i, 440.0, 1500.0, 1500);
  for (index i = 0; i < nt; ++i) tgt->data()(i, 0) = tone(i, 660.0, 990.0, 1100);
  FluidContext ctx;

  // ---- NMFSeed: bases / activations buffers sized by the chosen rank, activations scaled to max 1
  auto                  bases = std::make_shared<MemoryBufferAdaptor>(1, 1, 44100.0);
  auto                  acts = std::make_shared<MemoryBufferAdaptor>(1, 1, 44100.0);
  nndsvd::NMFSeedParams sp;
  sp.source = src; sp.bases = bases; sp.activations = acts;
  sp.minComponents = 2; sp.maxComponents = 10; sp.coverage = 0.9; sp.method = 0; sp.fftSettings = FFTParams(256, 64, -1);
  nndsvd::NMFSeedClient seed(sp, ctx);
  Result                r = seed.process<float>(ctx);
  if (!r.ok()) std::printf("NMFSeed: %s\n", r.message().c_str());
  CHECK(r.ok());
  const index rank = seed.lastRank();
  CHECK(rank >= 2 && rank <= 10);
  CHECK(bases->data().rows() == 129 && bases->data().cols() == rank);
  CHECK(acts->data().rows() == ns / 64 + 1 && acts->data().cols() == rank);
  float mx = 0;
  for (index f = 0; f < acts->data().rows(); ++f)
    for (index k = 0; k < rank; ++k) mx = std::max(mx, acts->data()(f, k));
  CHECK(std::abs(mx - 1.0f) < 1e-5f);
  { // a stereo source is refused with the reference's message (:89-90)
    auto                  st2 = std::make_shared<MemoryBufferAdaptor>(2, 100, 44100.0);
    nndsvd::NMFSeedParams q = sp;
    q.source = st2;
    nndsvd::NMFSeedClient c2(q, ctx);
    Result                e = c2.process<float>(ctx);
    CHECK(e.status() == Result::Status::kError && e.message() == "Only one channel supported");
  }
  // ---- the seeds drive BufNMF (basesMode = actMode = 1)
  bufnmf::BufNMFParams np;
  np.source = src; np.bases = bases; np.activations = acts; np.basesMode = 1; np.actMode = 1;
  np.components = rank; np.iterations = 15; np.seed = 3; np.fftSettings = FFTParams(256, 64, -1);
  bufnmf::NMFClient nmf(np, ctx);
  CHECK(nmf.process<float>(ctx).ok());

  // ---- NMFCross
  auto                     out = std::make_shared<MemoryBufferAdaptor>(1, 1, 44100.0);
  nmfcross::NMFCrossParams cp;
  cp.source = src; cp.target = tgt; cp.output = out;
  cp.timeSparsity = 7; cp.polyphony = 11; cp.continuity = 7; cp.iterations = 20; cp.seed = 5; cp.fftSettings = FFTParams(256, 64, -1);
  FluidTask                task;
  FluidContext             tctx(task);
  nmfcross::NMFCrossClient cross(cp, tctx);
  r = cross.process<float>(tctx);
  if (!r.ok()) std::printf("NMFCross: %s\n", r.message().c_str());
  CHECK(r.ok());
  CHECK(out->data().rows() == nt && out->data().cols() == 1);
  CHECK(std::abs(task.progress() - 1.0) < 1e-9); // iterations + 3 progress steps, all reported (:152)
  { // parameter checks keep the reference's messages (:114-119)
    nmfcross::NMFCrossParams q = cp;
    q.timeSparsity = 1001;
    nmfcross::NMFCrossClient c2(q, ctx);
    Result                   e = c2.process<float>(ctx);
    CHECK(e.status() == Result::Status::kError && e.message() == "Time Sparsity is larger than target frames");
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
    dump(src->data()); dump(tgt->data()); dump(out->data());
    std::fclose(f);
  }
  std::printf("nrt clients ok\n");
  return 0;
}
