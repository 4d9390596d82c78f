// C ABI of libflucoma_b200.so (include/flucoma_b200.h): plan management and the batched pipelines.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <random>
#include <thread>

using namespace fb200;

struct fb200_plan : public fb200::Plan {};

static thread_local std::string g_create_error;

namespace {

// FB200_DEVICE arrays are used in place on the plan's own stream, which is not ordered behind the streams the caller
// produced them on: wait for the device to go idle first (a few microseconds when it already is).  Outputs are complete
// on return, so the caller needs no further synchronisation.
int32_t enter(Plan* p, int mem)
{
  FB_CUDA(p, cudaSetDevice(p->cfg.device));
  if (mem == FB200_DEVICE) FB_CUDA(p, cudaDeviceSynchronize());
  return FB200_OK;
}

int64_t next_pow2_i64(int64_t x)
{ // clients/common/ParameterTypes.hpp:323-335 (up = true)
  if (x <= 0) return 0;
  uint32_t v = (uint32_t) x;
  --v;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return (int64_t) (v + 1);
}

size_t dsize(int dtype) { return dtype == FB200_F64 ? 8 : 4; }

// cuFFT plans are cached per (type, batch); the cache is a small LRU so that a long-lived plan fed buffers of many
// different lengths does not accumulate handles (each owns device workspace)
int32_t get_fft_plan(Plan* p, cufftType type, int64_t batch, cufftHandle* out)
{
  constexpr size_t kMaxPlans = 8;
  auto key = std::make_pair((int) type, batch);
  auto it = p->fft_plans.find(key);
  if (it != p->fft_plans.end()) {
    p->fft_lru[key] = ++p->fft_tick;
    *out = it->second;
    return FB200_OK;
  }
  if (batch > 0x7fffffffLL) { p->err = "cuFFT batch too large"; return FB200_ERR_INVALID; }
  if (p->fft_plans.size() >= kMaxPlans) {
    auto victim = p->fft_lru.begin();
    for (auto j = p->fft_lru.begin(); j != p->fft_lru.end(); ++j)
      if (j->second < victim->second) victim = j;
    FB_CUDA(p, cudaStreamSynchronize(p->stream)); // the evicted plan may still be executing
    cufftDestroy(p->fft_plans[victim->first]);
    p->fft_plans.erase(victim->first);
    p->fft_lru.erase(victim);
  }
  cufftHandle h;
  int n[1] = {p->fft};
  FB_CUFFT(p, cufftPlanMany(&h, 1, n, nullptr, 1, 0, nullptr, 1, 0, type, (int) batch));
  FB_CUFFT(p, cufftSetStream(h, p->stream));
  p->fft_plans[key] = h;
  p->fft_lru[key] = ++p->fft_tick;
  *out = h;
  return FB200_OK;
}

// buffers per wave so that the cuFFT real-side scratch stays <= ~512 MB
int64_t wave_size(const Plan* p, int64_t F, int64_t batch, int64_t mult)
{
  int64_t per_buf = std::max<int64_t>(1, F * p->fft * mult);
  int64_t w = std::max<int64_t>(1, ((int64_t) 1 << 27) / per_buf);
  return std::min(w, batch);
}

struct StageTimer {
  Plan* p;
  explicit StageTimer(Plan* pl) : p(pl)
  {
    p->launches = 0; p->launches_nmf = 0; p->kev_used = 0; p->backend_used = FB200_BACKEND_SIMT;
    std::memset(&p->stats, 0, sizeof(p->stats));
  }
  void mark(int i)
  {
    cudaEventRecord(p->ev[i], p->stream);
    static const bool dbg_sync = getenv("FB200_DEBUG_SYNC") != nullptr; // developer aid: localise an asynchronous fault
    if (dbg_sync) {
      cudaError_t e = cudaStreamSynchronize(p->stream);
      if (e != cudaSuccess) fprintf(stderr, "[fb200] stage mark %d: %s\n", i, cudaGetErrorString(e));
    }
  }
  float ms(int a, int b)
  {
    float t = 0.f;
    cudaEventElapsedTime(&t, p->ev[a], p->ev[b]);
    return t;
  }
};

// STFT of `batch` buffers already on the device (float [batch][n]) -> V (padded magnitudes, may be null) and/or
// spec_all (float2 [batch][F][B], may be null).  STFT.hpp:90-108 + :61-66.
// h_audio != nullptr: the audio still sits in (pinned) host memory; every wave is uploaded into d_audio on the plan's
// copy stream while the previous waves run their kernels, so that only the first upload is exposed.
int32_t run_stft(Plan* p, const float* d_audio, int64_t batch, int64_t n, int64_t F, float* V, int64_t Fp, int64_t Bp,
                 float2* spec_all, int64_t half, const float* h_audio = nullptr, int hop_override = 0, bool guard_staging = true)
{
  const int B = p->bins;
  int64_t wave = wave_size(p, F, batch, 1);
  if (h_audio) {
    wave = std::min<int64_t>(wave, std::max<int64_t>(1, (batch + 7) / 8)); // at least ~8 waves to overlap
    size_t nw = (size_t) ((batch + wave - 1) / wave);
    while (p->cev.size() < nw + 1) { cudaEvent_t e; FB_CUDA(p, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); p->cev.push_back(e); }
    // the copy stream must not overwrite the staging buffer before everything queued so far has finished with it
    // (not for the second part of a split call: its region of the buffer is untouched, and the wait would put the
    // upload behind the first part's update kernel)
    if (guard_staging) {
      FB_CUDA(p, cudaEventRecord(p->cev[nw], p->stream));
      FB_CUDA(p, cudaStreamWaitEvent(p->copy_stream, p->cev[nw], 0));
    }
    size_t i = 0;
    for (int64_t b0 = 0; b0 < batch; b0 += wave, i++) {
      int64_t nb = std::min(wave, batch - b0);
      FB_CUDA(p, cudaMemcpyAsync(const_cast<float*>(d_audio) + b0 * n, h_audio + b0 * n, sizeof(float) * (size_t) (nb * n),
                                 cudaMemcpyHostToDevice, p->copy_stream));
      FB_CUDA(p, cudaEventRecord(p->cev[i], p->copy_stream));
    }
  }
  FB_CUDA(p, p->frames.ensure(sizeof(float) * (size_t) (wave * F * p->fft)));
  float2* spec_wave = nullptr;
  if (!spec_all) {
    FB_CUDA(p, p->spec.ensure(sizeof(float2) * (size_t) (wave * F * B)));
    spec_wave = p->spec.as<float2>();
  }
  const int hop_eff = hop_override > 0 ? hop_override : p->hop;
  const bool fused = stft_fused_eligible(p, d_audio, n, std::min(wave, batch), hop_eff);
  size_t wi = 0;
  for (int64_t b0 = 0; b0 < batch; b0 += wave, wi++) {
    int64_t nb = std::min(wave, batch - b0);
    if (h_audio) FB_CUDA(p, cudaStreamWaitEvent(p->stream, p->cev[wi], 0));
    if (fused) { // one kernel from samples to |X| (and the complex spectrum when it is kept); TMA in, no intermediates
      FB_TRY(launch_stft_fused(p, d_audio + b0 * n, nb, n, F, V ? V + b0 * Fp * Bp : nullptr, Fp, Bp,
                               spec_all ? spec_all + b0 * F * B : nullptr, half, hop_eff));
      continue;
    }
    launch_frame_window(p, d_audio + b0 * n, n, nb, F, p->frames.as<float>(), half, hop_override);
    cufftHandle h;
    FB_TRY(get_fft_plan(p, CUFFT_R2C, nb * F, &h));
    float2* sp = spec_all ? spec_all + b0 * F * B : spec_wave;
    FB_CUFFT(p, cufftExecR2C(h, p->frames.as<float>(), reinterpret_cast<cufftComplex*>(sp)));
    p->launches++;
    launch_magnitude(p, sp, nb, F, V ? V + b0 * Fp * Bp : nullptr, Fp, Bp);
  }
  return FB200_OK;
}

// ISTFT of nsig spectra (float2 [nsig][F][B], destroyed) -> out float [nsig][out_stride], n samples each.
// STFT.hpp:178-199; stream_norm: the streaming clients' normalisation (BufferedProcess.hpp:231-237).
int32_t run_istft(Plan* p, float2* spec, int64_t nsig, int64_t F, int64_t n, float* out, int64_t half,
                  int64_t out_stride = -1, int stream_norm = 0)
{
  if (out_stride < 0) out_stride = n;
  if (istft_fused_eligible(p, nsig, F, n, half)) { // one kernel from spectra to samples; cuFFT for the sizes it does not take
    const int32_t st = launch_istft_fused(p, spec, nsig, F, n, out, half, out_stride, stream_norm);
    if (st != FB200_ERR_UNSUPPORTED) return st;
  }
  int64_t wave = wave_size(p, F, nsig, 1);
  FB_CUDA(p, p->frames.ensure(sizeof(float) * (size_t) (wave * F * p->fft)));
  for (int64_t s0 = 0; s0 < nsig; s0 += wave) {
    int64_t ns = std::min(wave, nsig - s0);
    cufftHandle h;
    FB_TRY(get_fft_plan(p, CUFFT_C2R, ns * F, &h));
    FB_CUFFT(p, cufftExecC2R(h, reinterpret_cast<cufftComplex*>(spec + s0 * F * p->bins), p->frames.as<float>()));
    p->launches++;
    launch_ola(p, p->frames.as<float>(), ns, F, n, out + s0 * out_stride, half, out_stride, stream_norm);
  }
  return FB200_OK;
}

int32_t alloc_nmf(Plan* p, NmfDev& d)
{
  FB_TRY(simt_configure(p, d));
  FB_CUDA(p, p->V.ensure(sizeof(float) * (size_t) d.batch * d.Fp * d.Bp));
  FB_CUDA(p, p->W.ensure(sizeof(float) * (size_t) d.batch * d.KP * d.Bp));
  FB_CUDA(p, p->H.ensure(sizeof(float) * (size_t) d.batch * d.Fp * d.KP));
  FB_CUDA(p, p->hden.ensure(sizeof(float) * (size_t) d.batch * d.KP));
  d.V = p->V.as<float>(); d.W = p->W.as<float>(); d.H = p->H.as<float>(); d.hden = p->hden.as<float>();
  d.wnum_part = nullptr; d.wden_part = nullptr; d.ticket = nullptr;
  return FB200_OK;
}

int32_t alloc_partials(Plan* p, NmfDev& d)
{
  FB_CUDA(p, p->wnum_part.ensure(sizeof(float) * (size_t) d.batch * d.ctas_per_buf * d.KP * d.Bp));
  FB_CUDA(p, p->wden_part.ensure(sizeof(float) * (size_t) d.batch * d.ctas_per_buf * d.KP));
  d.wnum_part = p->wnum_part.as<float>(); d.wden_part = p->wden_part.as<float>();
  // the kernels leave the tickets at zero; they are cleared once per call anyway (a launch that failed half-way must not
  // poison the next call)
  FB_CUDA(p, p->ticket.ensure(sizeof(int) * (size_t) d.batch));
  FB_CUDA(p, cudaMemsetAsync(p->ticket.p, 0, sizeof(int) * (size_t) d.batch, p->stream));
  d.ticket = p->ticket.as<int>();
  return FB200_OK;
}

// seeds (host) -> device, two per buffer: [0, batch) seed the W draws, [batch, 2 batch) the H draws.  For seed >= 0 both
// are the seed itself (the reference restarts the same stream for W and H, NMF.hpp:104-105,116-117); a negative seed
// means std::random_device, and the reference then builds two generators with independent draws (EigenRandom.hpp:80) --
// so do we.  Unseeded runs are therefore statistically, not bitwise, equivalent to the reference.
int32_t upload_seeds(Plan* p, const int64_t* seeds, int64_t batch, bool* any_random)
{
  FB_CUDA(p, p->seeds.ensure(sizeof(int64_t) * 2 * (size_t) batch));
  FB_CUDA(p, p->pin_a.ensure(sizeof(int64_t) * 2 * (size_t) batch));
  int64_t* s = reinterpret_cast<int64_t*>(p->pin_a.p);
  std::random_device rd;
  bool neg = false;
  // pin_a may still feed the previous call's copy only if that call failed mid-way; every successful call ends synchronised
  for (int64_t i = 0; i < batch; i++) {
    int64_t v = seeds ? seeds[i] : -1;
    if (v < 0) { neg = true; s[i] = (int64_t) rd(); s[batch + i] = (int64_t) rd(); }
    else { s[i] = v; s[batch + i] = v; }
  }
  FB_CUDA(p, cudaMemcpyAsync(p->seeds.p, s, sizeof(int64_t) * 2 * (size_t) batch, cudaMemcpyHostToDevice, p->stream));
  if (any_random) *any_random = neg;
  return FB200_OK;
}

// uniform draws for the random initialisation: U_w (and U_h when the two streams differ) in p->rnd
int32_t draw_uniforms(Plan* p, int64_t batch, int64_t u_stride, bool any_random, const float** U_w, const float** U_h)
{
  FB_CUDA(p, p->rnd.ensure(sizeof(float) * (size_t) (batch * u_stride) * (any_random ? 2 : 1)));
  launch_mt_uniform(p, p->seeds.as<int64_t>(), batch, u_stride, p->rnd.as<float>());
  *U_w = p->rnd.as<float>();
  *U_h = *U_w;
  if (any_random) {
    launch_mt_uniform(p, p->seeds.as<int64_t>() + batch, batch, u_stride, p->rnd.as<float>() + batch * u_stride);
    *U_h = p->rnd.as<float>() + batch * u_stride;
  }
  return FB200_OK;
}

// h0 of a frame stream (NMF.hpp:55): seed >= 0 -> the same K draws for every frame (*per_frame = 0); seed < 0 -> the
// reference redraws from random_device for every processFrame call: here every frame gets its own K draws out of
// ceil(frames / 4096) randomly seeded streams (*per_frame = 1) -- statistically equivalent
int32_t draw_frame_h0(Plan* p, int64_t seed, int64_t frames, int64_t K, const float** U, int* per_frame)
{
  bool neg = false;
  if (seed >= 0) {
    FB_TRY(upload_seeds(p, &seed, 1, &neg));
    FB_CUDA(p, p->rnd.ensure(sizeof(float) * (size_t) K));
    launch_mt_uniform(p, p->seeds.as<int64_t>(), 1, K, p->rnd.as<float>());
    *per_frame = 0;
  } else {
    const int64_t ns = (frames + 4095) / 4096;
    std::vector<int64_t> none((size_t) ns, -1);
    FB_TRY(upload_seeds(p, none.data(), ns, &neg));
    FB_CUDA(p, p->rnd.ensure(sizeof(float) * (size_t) (ns * 4096 * K)));
    launch_mt_uniform(p, p->seeds.as<int64_t>(), ns, 4096 * K, p->rnd.as<float>());
    *per_frame = 1;
  }
  *U = p->rnd.as<float>();
  return FB200_OK;
}

// n fused iterations on the SIMT engine; leaves W,H exactly as after n reference iterations (NMF.hpp:154-172)
void simt_run_iters(Plan* p, NmfDev& d, int n, bool upd_w, bool upd_h)
{
  if (upd_w && upd_h) {
    // (the W finalisation runs inside the tile launch, in the last CTA of every buffer, when the plan has a ticket array)
    simt_launch_tile(p, d, 0, 1, 1); // W-numerator of the first iteration
    if (!d.ticket) simt_launch_w_finalize(p, d);
    for (int it = 1; it < n; it++) {
      simt_launch_tile(p, d, 1, 1, 1); // H-update of iteration `it` fused with the W-numerator of `it+1`
      if (!d.ticket) simt_launch_w_finalize(p, d);
    }
    simt_launch_tile(p, d, 1, 0, 1); // H-update of the last iteration
  } else if (upd_w) {
    for (int it = 0; it < n; it++) { simt_launch_tile(p, d, 0, 1, 1); if (!d.ticket) simt_launch_w_finalize(p, d); }
  } else {
    for (int it = 0; it < n; it++) simt_launch_tile(p, d, 1, 0, 1);
  }
}

// Asynchronous progress on the persistent tensor-core engine: ONE launch runs all iterations.  The control words live
// in DEVICE memory (a first version kept them in host-mapped memory: 148 CTAs sampling a sysmem word once per pass were
// serialised on the PCIe read path and doubled the step time); the calling thread reads the per-CTA pass counters with
// small copies on the plan's copy stream while the kernel runs, reports iterations 1..n (each exactly once, in order) as
// the batch advances, and raises a cancel word in host-mapped memory when a callback returns 0 (CTA 0 relays it into the
// device word all CTAs sample).  See `Ctl`.
// `parts`: consecutive slices of the batch, each its own persistent launch (a call whose audio sits in host memory runs a
// first slice of one buffer per SM while the rest is still being uploaded: see fb200_bufnmf); `before_launch(i)` queues
// whatever slice i needs on the stream ahead of its launch.  Without a callback the launches are simply queued.
int32_t run_tc_parts(Plan* p, NmfDev* parts, int nparts, const std::function<int32_t(int)>& before_launch, int iters, bool upd_w, bool upd_h,
                     fb200_progress_fn progress, void* user)
{
  constexpr int SLOT = 256; // control words per part: [0] cancel, [1 + cta] pass counters
  if (!progress) {
    for (int i = 0; i < nparts; i++) {
      if (before_launch) FB_TRY(before_launch(i));
      FB_TRY(tc_run(p, parts[i], iters, upd_w, upd_h, nullptr, nullptr));
    }
    return FB200_OK;
  }
  FB_CUDA(p, p->ctrl_dev.ensure(sizeof(unsigned int) * 1025));
  FB_CUDA(p, p->ctrl.ensure(sizeof(unsigned int) * 1026));
  unsigned int* host = reinterpret_cast<unsigned int*>(p->ctrl.p); // [0, 1024): counters read back; [1024]: the word "1"
  unsigned int* dev = p->ctrl_dev.as<unsigned int>();
  FB_CUDA(p, cudaMemsetAsync(dev, 0, sizeof(unsigned int) * 1025, p->stream));
  if (!p->ev_async) FB_CUDA(p, cudaEventCreateWithFlags(&p->ev_async, cudaEventDisableTiming));
  host[1024] = 0u; // the cancel request: written by this thread, read by CTA 0 through the mapping below
  unsigned int* host_dev = nullptr;
  FB_CUDA(p, cudaHostGetDevicePointer(reinterpret_cast<void**>(&host_dev), host, 0));
  const int64_t npass = (upd_w && upd_h) ? iters + 1 : iters;
  int64_t total = 0;
  int grid = 0;
  for (int i = 0; i < nparts; i++) {
    if (before_launch) FB_TRY(before_launch(i));
    FB_TRY(tc_run(p, parts[i], iters, upd_w, upd_h, dev + i * SLOT, host_dev + 1024));
    total += (int64_t) parts[i].batch * npass;
    grid = std::max(grid, tc_grid(p, parts[i]));
  }
  FB_CUDA(p, cudaEventRecord(p->ev_async, p->stream));
  int64_t reported = 0;
  bool cancelled = false;
  auto report_up_to = [&](int64_t it) -> int32_t {
    while (!cancelled && reported < it) {
      if (!progress(user, ++reported)) {
        cancelled = true;
        // a plain store into pinned, device-mapped memory.  (A 4-byte cudaMemcpyAsync looked equivalent but is executed
        // by a helper kernel, which cannot be scheduled while the persistent CTAs own every SM: the request arrived
        // after the launch had finished.)
        *reinterpret_cast<volatile unsigned int*>(host + 1024) = 1u;
      }
    }
    return FB200_OK;
  };
  for (;;) {
    cudaError_t q = cudaEventQuery(p->ev_async);
    if (q == cudaSuccess) break;
    if (q != cudaErrorNotReady) { p->err = std::string("CUDA error: ") + cudaGetErrorString(q); return FB200_ERR_CUDA; }
    if (!cancelled) {
      FB_CUDA(p, cudaMemcpyAsync(host, dev, sizeof(unsigned int) * (size_t) (nparts * SLOT), cudaMemcpyDeviceToHost, p->poll_stream));
      FB_CUDA(p, cudaStreamSynchronize(p->poll_stream));
      int64_t done = 0;
      for (int i = 0; i < nparts; i++)
        for (int c = 0; c < grid; c++) done += host[i * SLOT + 1 + c];
      // iteration `it` is reported once the batch as a whole has done the work of `it` iterations; the last one at the end
      FB_TRY(report_up_to(std::min<int64_t>(iters - 1, done * iters / std::max<int64_t>(1, total))));
    }
    std::this_thread::sleep_for(std::chrono::microseconds(250));
  }
  FB_CUDA(p, cudaStreamSynchronize(p->poll_stream));
  FB_TRY(report_up_to(iters));
  return cancelled ? FB200_CANCELLED : FB200_OK;
}

// The multiplicative-update loop (NMF.hpp:154-181) on the device state in `d`.  Returns FB200_CANCELLED when the
// progress callback asked to stop.
//   stride >= 1 (0 -> 1): exact mode.  The engine runs `stride` iterations per launch group, then the callbacks of those
//     iterations are replayed in order; a cancel leaves W,H as after the LAST iteration of the group (== the reference
//     at stride 1; cancellation granularity is `stride`).
//   stride == FB200_PROGRESS_ASYNC: the loop is never interrupted (see run_tc_async; on the SIMT engine: groups of 8).
int32_t run_nmf_loop(Plan* p, NmfDev& d, int iters, bool upd_w, bool upd_h, fb200_progress_fn progress, void* user,
                     int stride)
{
  if (iters <= 0 || (!upd_w && !upd_h)) {
    // nothing changes W/H; the reference still calls the callbacks every iteration (:175-176)
    for (int it = 1; progress && it <= iters; it++)
      if (!progress(user, it)) return FB200_CANCELLED;
    return FB200_OK;
  }
  // engine choice: the resident tensor-core engine for its shapes (rank 16, <= 513 bins, <= 512 frames, W updated or not),
  // the streamed one for everything else it covers (rank 9..64, any bins = 128 m + 1, any frame count, fixed-W frame
  // streams) when there are enough independent work units to fill the SMs, the SIMT engine for the rest
  const int be = p->cfg.backend;
  const bool tc1 = (be == FB200_BACKEND_AUTO || be == FB200_BACKEND_TCGEN05) && tc_eligible(d);
  const int64_t units = upd_w ? d.batch : (int64_t) d.batch * ((d.Fp / 128 + 1) / 2);
  const bool tc2 = !tc1 && be != FB200_BACKEND_SIMT && tcs_eligible(d) && (be != FB200_BACKEND_AUTO || units >= 32);
  if (!tc1 && !tc2 && (be == FB200_BACKEND_TCGEN05 || be == FB200_BACKEND_TCGEN05_STREAMED)) {
    p->err = "tensor-core backend requested but the shape does not qualify (rank 9..64, bins = 128 m + 1, frames padded to 128)";
    return FB200_ERR_UNSUPPORTED;
  }
  const bool use_tc = tc1;
  p->backend_used = tc1 ? FB200_BACKEND_TCGEN05 : (tc2 ? FB200_BACKEND_TCGEN05_STREAMED : FB200_BACKEND_SIMT);
  if (!tc1 && !tc2 && upd_w) FB_TRY(alloc_partials(p, d));
  // one launch group = n complete iterations: the tensor-core engines run them in one persistent launch (their state
  // round-trips exactly through the fp32 W/H arrays), the SIMT engine as 2n+1 fused launches
  auto run_group = [&](int n) -> int32_t {
    if (tc1) return tc_run(p, d, n, upd_w, upd_h);
    if (tc2) return tcs_run(p, d, n, upd_w, upd_h);
    simt_run_iters(p, d, n, upd_w, upd_h);
    return FB200_OK;
  };
  if (!progress) return run_group(iters);
  if (stride == FB200_PROGRESS_ASYNC && use_tc) return run_tc_parts(p, &d, 1, nullptr, iters, upd_w, upd_h, progress, user);
  const int s = stride == FB200_PROGRESS_ASYNC ? 8 : std::max(1, stride);
  for (int it0 = 0; it0 < iters; it0 += s) {
    const int n = std::min(s, iters - it0);
    FB_TRY(run_group(n));
    FB_CUDA(p, cudaStreamSynchronize(p->stream));
    for (int j = it0 + 1; j <= it0 + n; j++)
      if (!progress(user, j)) return FB200_CANCELLED;
  }
  return FB200_OK;
}

// activation-only solve with a fixed dictionary (NMF.hpp:72-83): the streamed tensor-core engine when the shape allows,
// else one SIMT launch that runs all iterations per frame tile
int32_t run_h_only(Plan* p, NmfDev& d, int iters)
{
  const int be = p->cfg.backend;
  const int64_t units = (int64_t) d.batch * ((d.Fp / 128 + 1) / 2);
  if (be != FB200_BACKEND_SIMT && tcs_eligible(d) && (be != FB200_BACKEND_AUTO || units >= 32)) {
    p->backend_used = FB200_BACKEND_TCGEN05_STREAMED;
    return tcs_run(p, d, iters, false, true);
  }
  if (be == FB200_BACKEND_TCGEN05 || be == FB200_BACKEND_TCGEN05_STREAMED) {
    p->err = "tensor-core backend requested but the shape does not qualify (rank 9..64, bins = 128 m + 1)";
    return FB200_ERR_UNSUPPORTED;
  }
  simt_launch_tile(p, d, 1, 0, iters);
  return FB200_OK;
}

// host<->device helpers for caller arrays --------------------------------------------------------------------------
// Brings a caller array (dtype/mem) of `count` elements to a device pointer of the same dtype.
int32_t to_device_raw(Plan* p, const void* src, int mem, size_t bytes, DevBuf& stage, const void** out)
{
  if (mem == FB200_DEVICE) { *out = src; return FB200_OK; }
  FB_CUDA(p, stage.ensure(bytes));
  FB_CUDA(p, cudaMemcpyAsync(stage.p, src, bytes, cudaMemcpyHostToDevice, p->stream));
  *out = stage.p;
  return FB200_OK;
}

int32_t finish(Plan* p, StageTimer& t, int last_ev)
{
  FB_CUDA(p, cudaStreamSynchronize(p->stream));
  FB_CUDA(p, cudaGetLastError());
  p->stats.ms_total = t.ms(0, last_ev);
  p->stats.launches_total = p->launches;
  p->stats.launches_nmf = p->launches_nmf;
  p->stats.backend_used = p->backend_used;
  float sum = 0.f;
  for (size_t i = 0; i + 1 < p->kev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p->kev[i], p->kev[i + 1]) == cudaSuccess) sum += ms;
  }
  p->stats.ms_update_kernel = sum;
  p->stats.update_kernel_launches = (int32_t) (p->kev_used / 2);
  return FB200_OK;
}

} // namespace

// =====================================================================================================================
extern "C" {

uint32_t fb200_abi_version(void) { return FB200_ABI_VERSION; }

int32_t fb200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int64_t fb200_num_frames(int64_t n_samples, int32_t win, int32_t hop)
{ // STFT.hpp:94-99, NMFClient.hpp:111-113
  if (hop <= 0) hop = win >> 1;
  if (hop <= 0) return 0;
  return (n_samples + win + hop - win) / hop;
}

int32_t fb200_resolve_fft(int32_t win, int32_t hop, int32_t fft, int32_t* out_hop, int32_t* out_fft, int32_t* out_bins)
{ // ParameterTypes.hpp:295-313
  if (win <= 0) return FB200_ERR_INVALID;
  int32_t f = fft < 0 ? (int32_t) next_pow2_i64(win) : fft;
  int32_t h = hop > 0 ? hop : (win >> 1);
  if (f < win || (f & (f - 1)) != 0 || f < 4 || h <= 0) return FB200_ERR_INVALID;
  if (out_hop) *out_hop = h;
  if (out_fft) *out_fft = f;
  if (out_bins) *out_bins = (f >> 1) + 1;
  return FB200_OK;
}

void fb200_shard_range(int64_t total, int32_t world, int32_t rank, int64_t* begin, int64_t* count)
{ // contiguous, balanced: the first (total % world) ranks take one extra buffer
  if (world <= 0) world = 1;
  int64_t base = total / world, rem = total % world;
  int64_t b = rank * base + std::min<int64_t>(rank, rem);
  int64_t c = base + (rank < rem ? 1 : 0);
  if (begin) *begin = b;
  if (count) *count = c;
}

int32_t fb200_plan_create(const fb200_config* cfg, fb200_plan** out)
{
  if (!cfg || !out || cfg->struct_size != sizeof(fb200_config)) { g_create_error = "bad config"; return FB200_ERR_INVALID; }
  *out = nullptr;
  int32_t hop, fft, bins;
  if (fb200_resolve_fft(cfg->win, cfg->hop, cfg->fft, &hop, &fft, &bins) != FB200_OK) {
    g_create_error = "invalid window/hop/fft: fft must be a power of two >= win";
    return FB200_ERR_INVALID;
  }
  if (fb200_device_count() <= cfg->device || cfg->device < 0) { g_create_error = "no such CUDA device"; return FB200_ERR_NO_DEVICE; }
  if (cudaSetDevice(cfg->device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return FB200_ERR_CUDA; }
  fb200_plan* p = new fb200_plan();
  p->cfg = *cfg;
  p->win = cfg->win; p->hop = hop; p->fft = fft; p->bins = bins;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) p->sm_count = prop.multiProcessorCount;
  bool ok = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&p->poll_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (auto& e : p->ev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
  ok = ok && p->window.ensure(sizeof(float) * (size_t) p->win) == cudaSuccess;
  if (!ok) {
    g_create_error = std::string("CUDA setup failed: ") + cudaGetErrorString(cudaGetLastError());
    fb200_plan_destroy(p);
    return FB200_ERR_CUDA;
  }
  launch_hann(p);
  if (cudaStreamSynchronize(p->stream) != cudaSuccess) {
    g_create_error = std::string("kernel launch failed (is this an sm_100a device?): ") + cudaGetErrorString(cudaGetLastError());
    fb200_plan_destroy(p);
    return FB200_ERR_CUDA;
  }
  *out = p;
  return FB200_OK;
}

void fb200_plan_destroy(fb200_plan* p)
{
  if (!p) return;
  cudaSetDevice(p->cfg.device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  for (auto& kv : p->fft_plans) cufftDestroy(kv.second);
  DevBuf* bufs[] = {&p->window, &p->audio, &p->stage, &p->frames, &p->spec, &p->cspec, &p->V, &p->W, &p->H, &p->hden,
                    &p->wnum_part, &p->wden_part, &p->ticket, &p->rnd, &p->seeds, &p->scale, &p->out_a, &p->out_b, &p->ctrl_dev, &p->wop_buf, &p->hop_buf, &p->twiddle, &p->x0, &p->x1, &p->x2, &p->x3, &p->x4, &p->x5};
  for (auto* b : bufs) b->release();
  p->pin_a.release(); p->pin_b.release(); p->ctrl.release();
  if (p->ev_async) cudaEventDestroy(p->ev_async);
  for (auto& e : p->ev) if (e) cudaEventDestroy(e);
  for (auto& e : p->kev) cudaEventDestroy(e);
  for (auto& e : p->cev) cudaEventDestroy(e);
  if (p->copy_stream) { cudaStreamSynchronize(p->copy_stream); cudaStreamDestroy(p->copy_stream); }
  if (p->poll_stream) { cudaStreamSynchronize(p->poll_stream); cudaStreamDestroy(p->poll_stream); }
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

const char* fb200_last_error(const fb200_plan* p) { return p ? p->err.c_str() : g_create_error.c_str(); }

int32_t fb200_get_stats(const fb200_plan* p, fb200_stats* out)
{
  if (!p || !out) return FB200_ERR_INVALID;
  *out = p->stats;
  return FB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int32_t fb200_stft(fb200_plan* p, const void* audio, int64_t batch, int64_t n, void* spectrum, void* magnitude,
                   int32_t dtype, int32_t mem)
{
  if (!p) return FB200_ERR_INVALID;
  if (!audio || batch <= 0 || n < 0 || (!spectrum && !magnitude)) { p->err = "fb200_stft: bad arguments"; return FB200_ERR_INVALID; }
  FB_TRY(enter(p, mem));
  StageTimer t(p);
  t.mark(0);
  const int B = p->bins;
  const int64_t F = fb200_num_frames(n, p->win, p->hop);
  const void* d_raw;
  FB_TRY(to_device_raw(p, audio, mem, dsize(dtype) * (size_t) (batch * n), p->stage, &d_raw));
  FB_CUDA(p, p->audio.ensure(sizeof(float) * (size_t) std::max<int64_t>(1, batch * n)));
  launch_copy3d(p, d_raw, dtype, n, n, p->audio.p, FB200_F32, n, n, batch, 1, n, nullptr, 0);
  t.mark(1);
  FB_CUDA(p, p->cspec.ensure(sizeof(float2) * (size_t) (batch * F * B)));
  float* mag = nullptr;
  if (magnitude) {
    FB_CUDA(p, p->V.ensure(sizeof(float) * (size_t) (batch * F * B)));
    mag = p->V.as<float>();
  }
  FB_TRY(run_stft(p, p->audio.as<float>(), batch, n, F, mag, F, B, p->cspec.as<float2>(), p->win / 2));
  t.mark(2);
  // export
  if (spectrum) {
    size_t bytes = dsize(dtype) * 2 * (size_t) (batch * F * B);
    void* dst = spectrum;
    if (mem == FB200_HOST) { FB_CUDA(p, p->out_a.ensure(bytes)); dst = p->out_a.p; }
    launch_copy3d(p, p->cspec.p, FB200_F32, F * B * 2, 2 * B, dst, dtype, F * B * 2, 2 * B, batch, F, 2 * B, nullptr, 0);
    if (mem == FB200_HOST) FB_CUDA(p, cudaMemcpyAsync(spectrum, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
  }
  if (magnitude) {
    size_t bytes = dsize(dtype) * (size_t) (batch * F * B);
    void* dst = magnitude;
    if (mem == FB200_HOST) { FB_CUDA(p, p->out_b.ensure(bytes)); dst = p->out_b.p; }
    launch_copy3d(p, mag, FB200_F32, F * B, B, dst, dtype, F * B, B, batch, F, B, nullptr, 0);
    if (mem == FB200_HOST) FB_CUDA(p, cudaMemcpyAsync(magnitude, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
  }
  t.mark(3);
  FB_TRY(finish(p, t, 3));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_stft = t.ms(1, 2); p->stats.ms_d2h = t.ms(2, 3);
  return FB200_OK;
}

int32_t fb200_istft(fb200_plan* p, const void* spectrum, int64_t batch, int64_t F, void* audio, int64_t n,
                    int32_t dtype, int32_t mem)
{
  if (!p) return FB200_ERR_INVALID;
  if (!spectrum || !audio || batch <= 0 || F <= 0 || n <= 0) { p->err = "fb200_istft: bad arguments"; return FB200_ERR_INVALID; }
  if (p->win / 2 + n > p->win + (F - 1) * p->hop + p->win + p->hop) { p->err = "fb200_istft: n_samples exceeds the overlap-add length"; return FB200_ERR_INVALID; }
  FB_TRY(enter(p, mem));
  StageTimer t(p);
  t.mark(0);
  const int B = p->bins;
  const void* d_raw;
  FB_TRY(to_device_raw(p, spectrum, mem, dsize(dtype) * 2 * (size_t) (batch * F * B), p->stage, &d_raw));
  FB_CUDA(p, p->cspec.ensure(sizeof(float2) * (size_t) (batch * F * B)));
  launch_copy3d(p, d_raw, dtype, F * B * 2, 2 * B, p->cspec.p, FB200_F32, F * B * 2, 2 * B, batch, F, 2 * B, nullptr, 0);
  t.mark(1);
  FB_CUDA(p, p->audio.ensure(sizeof(float) * (size_t) (batch * n)));
  FB_TRY(run_istft(p, p->cspec.as<float2>(), batch, F, n, p->audio.as<float>(), p->win / 2));
  t.mark(2);
  size_t bytes = dsize(dtype) * (size_t) (batch * n);
  void* dst = audio;
  if (mem == FB200_HOST) { FB_CUDA(p, p->out_a.ensure(bytes)); dst = p->out_a.p; }
  launch_copy3d(p, p->audio.p, FB200_F32, n, n, dst, dtype, n, n, batch, 1, n, nullptr, 0);
  if (mem == FB200_HOST) FB_CUDA(p, cudaMemcpyAsync(audio, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
  t.mark(3);
  FB_TRY(finish(p, t, 3));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_resynth = t.ms(1, 2); p->stats.ms_d2h = t.ms(2, 3);
  return FB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int32_t fb200_nmf_process(fb200_plan* p, const fb200_nmf_args* a)
{
  if (!p) return FB200_ERR_INVALID;
  if (!a || a->struct_size != sizeof(fb200_nmf_args) || !a->X || a->batch <= 0 || a->frames <= 0 || a->bins <= 0 ||
      a->rank <= 0 || a->iterations < 0) {
    p->err = "fb200_nmf_process: bad arguments";
    return FB200_ERR_INVALID;
  }
  FB_TRY(enter(p, a->mem));
  StageTimer t(p);
  t.mark(0);
  NmfDev d{};
  d.batch = (int) a->batch; d.F = (int) a->frames; d.B = (int) a->bins; d.K = a->rank;
  FB_TRY(alloc_nmf(p, d));
  const int64_t F = d.F, B = d.B, K = d.K;
  const size_t es = dsize(a->dtype);
  // X -> V (zero padded)
  const void* d_x;
  FB_TRY(to_device_raw(p, a->X, a->mem, es * (size_t) (a->batch * F * B), p->stage, &d_x));
  FB_CUDA(p, cudaMemsetAsync(d.V, 0, sizeof(float) * (size_t) d.batch * d.Fp * d.Bp, p->stream));
  launch_copy3d(p, d_x, a->dtype, F * B, B, d.V, FB200_F32, (int64_t) d.Fp * d.Bp, d.Bp, a->batch, F, B, nullptr, 0);
  // seeds / W0 / H0
  const float* dW0 = nullptr; const float* dH0 = nullptr; const float* U = nullptr; const float* U_h = nullptr;
  int64_t u_stride = std::max(B * K, K * F);
  if (a->W0) {
    const void* raw;
    FB_TRY(to_device_raw(p, a->W0, a->mem, es * (size_t) (a->batch * K * B), p->out_a, &raw));
    FB_CUDA(p, p->cspec.ensure(sizeof(float) * (size_t) (a->batch * K * B)));
    launch_copy3d(p, raw, a->dtype, K * B, B, p->cspec.p, FB200_F32, K * B, B, a->batch, K, B, nullptr, 0);
    dW0 = p->cspec.as<float>();
  }
  if (a->H0) {
    const void* raw;
    FB_TRY(to_device_raw(p, a->H0, a->mem, es * (size_t) (a->batch * F * K), p->out_b, &raw));
    FB_CUDA(p, p->frames.ensure(sizeof(float) * (size_t) (a->batch * F * K)));
    launch_copy3d(p, raw, a->dtype, F * K, K, p->frames.p, FB200_F32, F * K, K, a->batch, F, K, nullptr, 0);
    dH0 = p->frames.as<float>();
  }
  if (!a->W0 || !a->H0) {
    bool any_random = false;
    FB_TRY(upload_seeds(p, a->seeds, a->batch, &any_random));
    FB_TRY(draw_uniforms(p, a->batch, u_stride, any_random, &U, &U_h));
  }
  t.mark(1);
  launch_nmf_init(p, d, U, U_h, u_stride, dW0, dH0, 0);
  t.mark(2);
  int32_t st = run_nmf_loop(p, d, a->iterations, a->update_w != 0, a->update_h != 0, a->progress, a->progress_user,
                            a->progress_stride);
  if (st < 0) return st;
  const bool cancelled = st == FB200_CANCELLED;
  t.mark(3);
  // outputs (NMF.hpp:127-133); cancelled -> V1 = X (:176 skips :182)
  auto export_arr = [&](const float* src, int64_t s_b, int64_t s_r, void* user, int64_t rows, int64_t cols, DevBuf& tmp) -> int32_t {
    size_t bytes = es * (size_t) (a->batch * rows * cols);
    void* dst = user;
    if (a->mem == FB200_HOST) { FB_CUDA(p, tmp.ensure(bytes)); dst = tmp.p; }
    launch_copy3d(p, src, FB200_F32, s_b, s_r, dst, a->dtype, rows * cols, cols, a->batch, rows, cols, nullptr, 0);
    if (a->mem == FB200_HOST) FB_CUDA(p, cudaMemcpyAsync(user, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
    return FB200_OK;
  };
  if (a->W1) FB_TRY(export_arr(d.W, (int64_t) d.KP * d.Bp, d.Bp, a->W1, K, B, p->out_a));
  if (a->H1) FB_TRY(export_arr(d.H, (int64_t) d.Fp * d.KP, d.KP, a->H1, F, K, p->out_b));
  if (a->V1) {
    size_t bytes = es * (size_t) (a->batch * F * B);
    if (cancelled) {
      if (a->V1 != a->X) {
        if (a->mem == FB200_HOST) std::memcpy(a->V1, a->X, bytes);
        else FB_CUDA(p, cudaMemcpyAsync(a->V1, a->X, bytes, cudaMemcpyDeviceToDevice, p->stream));
      }
    } else {
      void* dst = a->V1;
      if (a->mem == FB200_HOST) { FB_CUDA(p, p->stage.ensure(bytes)); dst = p->stage.p; }
      launch_vhat(p, d, dst, a->dtype);
      if (a->mem == FB200_HOST) FB_CUDA(p, cudaMemcpyAsync(a->V1, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
    }
  }
  t.mark(4);
  FB_TRY(finish(p, t, 4));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_init = t.ms(1, 2); p->stats.ms_nmf = t.ms(2, 3); p->stats.ms_d2h = t.ms(3, 4);
  return cancelled ? FB200_CANCELLED : FB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int32_t fb200_nmf_process_frames(fb200_plan* p, const fb200_frames_args* a)
{
  if (!p) return FB200_ERR_INVALID;
  if (!a || a->struct_size != sizeof(fb200_frames_args) || !a->X || !a->W0 || a->frames <= 0 || a->bins <= 0 ||
      a->rank <= 0 || a->iterations < 0 || (!a->H && !a->V)) {
    p->err = "fb200_nmf_process_frames: bad arguments";
    return FB200_ERR_INVALID;
  }
  FB_TRY(enter(p, a->mem));
  StageTimer t(p);
  t.mark(0);
  NmfDev d{};
  d.batch = 1; d.F = (int) a->frames; d.B = (int) a->bins; d.K = a->rank; d.clamp_v = 1;
  FB_TRY(alloc_nmf(p, d));
  const int64_t F = d.F, B = d.B, K = d.K;
  const size_t es = dsize(a->dtype);
  const void* d_x;
  FB_TRY(to_device_raw(p, a->X, a->mem, es * (size_t) (F * B), p->stage, &d_x));
  FB_CUDA(p, cudaMemsetAsync(d.V, 0, sizeof(float) * (size_t) d.Fp * d.Bp, p->stream));
  launch_copy3d(p, d_x, a->dtype, F * B, B, d.V, FB200_F32, (int64_t) d.Fp * d.Bp, d.Bp, 1, F, B, nullptr, 0);
  const void* raw_w;
  FB_TRY(to_device_raw(p, a->W0, a->mem, es * (size_t) (K * B), p->out_a, &raw_w));
  FB_CUDA(p, p->cspec.ensure(sizeof(float) * (size_t) (K * B)));
  launch_copy3d(p, raw_w, a->dtype, K * B, B, p->cspec.p, FB200_F32, K * B, B, 1, K, B, nullptr, 0);
  const float* U0 = nullptr;
  int per_frame = 0;
  FB_TRY(draw_frame_h0(p, a->seed, F, K, &U0, &per_frame));
  t.mark(1);
  launch_nmf_init(p, d, U0, U0, K, p->cspec.as<float>(), nullptr, 1 + per_frame); // NMF.hpp:55-64
  t.mark(2);
  if (a->iterations > 0) FB_TRY(run_h_only(p, d, a->iterations));                // NMF.hpp:72-83
  t.mark(3);
  auto export_arr = [&](const float* src, int64_t s_r, void* user, int64_t rows, int64_t cols, DevBuf& tmp) -> int32_t {
    size_t bytes = es * (size_t) (rows * cols);
    void* dst = user;
    if (a->mem == FB200_HOST) { FB_CUDA(p, tmp.ensure(bytes)); dst = tmp.p; }
    launch_copy3d(p, src, FB200_F32, 0, s_r, dst, a->dtype, 0, cols, 1, rows, cols, nullptr, 0);
    if (a->mem == FB200_HOST) FB_CUDA(p, cudaMemcpyAsync(user, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
    return FB200_OK;
  };
  if (a->H) FB_TRY(export_arr(d.H, d.KP, a->H, F, K, p->out_b));
  if (a->W_norm) FB_TRY(export_arr(d.W, d.Bp, a->W_norm, K, B, p->out_a));
  if (a->V) {
    size_t bytes = es * (size_t) (F * B);
    void* dst = a->V;
    if (a->mem == FB200_HOST) { FB_CUDA(p, p->stage.ensure(bytes)); dst = p->stage.p; }
    launch_vhat(p, d, dst, a->dtype);
    if (a->mem == FB200_HOST) FB_CUDA(p, cudaMemcpyAsync(a->V, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
  }
  t.mark(4);
  FB_TRY(finish(p, t, 4));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_init = t.ms(1, 2); p->stats.ms_nmf = t.ms(2, 3); p->stats.ms_d2h = t.ms(3, 4);
  return FB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int32_t fb200_bufnmf(fb200_plan* p, const fb200_bufnmf_args* a)
{
  if (!p) return FB200_ERR_INVALID;
  if (!a || a->struct_size != sizeof(fb200_bufnmf_args) || !a->audio || a->batch <= 0 || a->n_samples <= 0 ||
      a->rank <= 0 || a->iterations < 0 || a->bases_mode < 0 || a->bases_mode > 2 || a->acts_mode < 0 || a->acts_mode > 2) {
    p->err = "fb200_bufnmf: bad arguments";
    return FB200_ERR_INVALID;
  }
  // NMFClient.hpp:134-136, 164-167: seed/fixed modes need the corresponding buffer
  if ((a->bases_mode > 0 && !a->bases_in) || (a->acts_mode > 0 && !a->acts_in)) {
    p->err = "Bases/Activations mode set to Seed or Fix, but no buffer supplied";
    return FB200_ERR_INVALID;
  }
  const bool fix_w = a->bases_mode == 2, fix_h = a->acts_mode == 2;
  const bool needs_analysis = !(fix_w && fix_h);                        // :141
  const bool resynth = a->resynth_out != nullptr;
  if (!needs_analysis && !resynth) {                                     // :143-145
    p->err = "Bases and Activations buffers both fixed, but resynthesis disabled: no work to do";
    return FB200_WARN_NO_WORK;
  }
  FB_TRY(enter(p, a->mem));
  StageTimer t(p);
  t.mark(0);
  const int64_t batch = a->batch, n = a->n_samples;
  const int64_t F = fb200_num_frames(n, p->win, p->hop);
  NmfDev d{};
  d.batch = (int) batch; d.F = (int) F; d.B = p->bins; d.K = a->rank;
  FB_TRY(alloc_nmf(p, d));
  const int64_t B = d.B, K = d.K;
  const int host = a->mem == FB200_HOST;

  // audio in (float32, NMFClient.hpp:240)
  const void* d_audio = a->audio;
  if (host) { // uploaded wave by wave inside run_stft, overlapped with the STFT kernels
    FB_CUDA(p, p->audio.ensure(sizeof(float) * (size_t) (batch * n)));
    d_audio = p->audio.p;
  }
  const float* dW0 = nullptr; const float* dH0 = nullptr; const float* U = nullptr;
  if (a->bases_mode > 0) {                                               // :248-252
    const void* raw;
    FB_TRY(to_device_raw(p, a->bases_in, a->mem, sizeof(float) * (size_t) (batch * K * B), p->out_a, &raw));
    dW0 = (const float*) raw;
  }
  if (a->acts_mode > 0) {                                                // :253-257
    const void* raw;
    FB_TRY(to_device_raw(p, a->acts_in, a->mem, sizeof(float) * (size_t) (batch * F * K), p->out_b, &raw));
    dH0 = (const float*) raw;
  }
  const int64_t u_stride = std::max(B * K, K * F);
  bool any_random = false;
  if (!dW0 || !dH0) FB_TRY(upload_seeds(p, a->seeds, batch, &any_random));
  t.mark(1);
  // STFT + |X|  (:241-242); k_magnitude writes every interior element, only the pads need zeros
  launch_zero_pads(p, d.V, d.batch, d.F, d.Fp, d.B, d.Bp);
  float2* spec_all = nullptr;
  if (resynth) {
    FB_CUDA(p, p->spec.ensure(sizeof(float2) * (size_t) (batch * F * B)));
    spec_all = p->spec.as<float2>();
  }
  // A call whose audio sits in host memory and that runs on the resident tensor-core engine is split: the first
  // sm_count buffers (one per CTA: exactly one round of the persistent kernel) are uploaded, transformed and started,
  // and the upload of the rest runs under that launch instead of in front of everything (config 2: 536 MB = 9.7 ms of
  // PCIe time, of which 1.4 ms stay exposed).  The number of rounds is unchanged for every batch > sm: 1 + ceil((batch - sm) / sm)
  // = ceil(batch / sm).
  const int be = p->cfg.backend;
  const bool async_or_none = !a->progress || a->progress_stride == FB200_PROGRESS_ASYNC;
  const bool can_tc1 = (be == FB200_BACKEND_AUTO || be == FB200_BACKEND_TCGEN05) && tc_eligible(d);
  const bool can_tc2 = !can_tc1 && be != FB200_BACKEND_SIMT && tcs_eligible(d) && !fix_w && !a->progress; // the streamed engine has no in-kernel progress
  const bool split = host && needs_analysis && a->iterations > 0 && async_or_none && batch > (int64_t) p->sm_count &&
                     (can_tc1 || can_tc2);
  const float* U_h = nullptr;
  int32_t st = FB200_OK;
  if (!split) {
    FB_TRY(run_stft(p, (const float*) d_audio, batch, n, F, d.V, d.Fp, d.Bp, spec_all, p->win / 2,
                    host ? (const float*) a->audio : nullptr));
    t.mark(2);
    if (!dW0 || !dH0) FB_TRY(draw_uniforms(p, batch, u_stride, any_random, &U, &U_h));
    launch_nmf_init(p, d, U, U_h, u_stride, dW0, dH0, 0);
    t.mark(3);
    st = run_nmf_loop(p, d, a->iterations * (needs_analysis ? 1 : 0), !fix_w, !fix_h, a->progress,
                      a->progress_user, a->progress_stride);            // :268-271
  } else {
    t.mark(2);
    if (!dW0 || !dH0) FB_TRY(draw_uniforms(p, batch, u_stride, any_random, &U, &U_h));
    launch_nmf_init(p, d, U, U_h, u_stride, dW0, dH0, 0); // does not depend on |X|
    t.mark(3);
    const int64_t head = p->sm_count;
    NmfDev parts[2] = {d, d};
    parts[0].batch = (int) head;
    parts[1].batch = (int) (batch - head);
    parts[1].V = d.V + head * d.Fp * d.Bp;
    parts[1].W = d.W + head * d.KP * d.Bp;
    parts[1].H = d.H + head * d.Fp * d.KP;
    parts[1].hden = d.hden + head * d.KP;
    auto front = [&](int i) -> int32_t {
      const int64_t b0 = i ? head : 0, nb = i ? batch - head : head;
      return run_stft(p, (const float*) d_audio + b0 * n, nb, n, F, d.V + b0 * d.Fp * d.Bp, d.Fp, d.Bp,
                      spec_all ? spec_all + b0 * F * B : nullptr, p->win / 2, (const float*) a->audio + b0 * n, 0, i == 0);
    };
    if (can_tc1) {
      p->backend_used = FB200_BACKEND_TCGEN05;
      st = run_tc_parts(p, parts, 2, front, a->iterations, !fix_w, !fix_h, a->progress, a->progress_user);
    } else {
      p->backend_used = FB200_BACKEND_TCGEN05_STREAMED;
      parts[0].op_total = parts[1].op_total = (int) batch;
      parts[0].op_first = 0; parts[1].op_first = (int) head;
      for (int i = 0; i < 2 && st == FB200_OK; i++) {
        st = front(i);
        if (st == FB200_OK) st = tcs_run(p, parts[i], a->iterations, !fix_w, !fix_h);
      }
    }
  }
  if (st < 0) return st;
  t.mark(4);
  if (st == FB200_CANCELLED) {                                           // :273-274
    FB_TRY(finish(p, t, 4));
    return FB200_CANCELLED;
  }
  // bases / activations out  (:277-300).  The seed uploads in out_a/out_b are consumed by now.
  if (a->bases_out && !fix_w) {
    size_t bytes = sizeof(float) * (size_t) (batch * K * B);
    void* dst = a->bases_out;
    if (host) { FB_CUDA(p, p->out_a.ensure(bytes)); dst = p->out_a.p; }
    launch_copy3d(p, d.W, FB200_F32, (int64_t) d.KP * d.Bp, d.Bp, dst, FB200_F32, K * B, B, batch, K, B, nullptr, 0);
    if (host) FB_CUDA(p, cudaMemcpyAsync(a->bases_out, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
  }
  if (a->acts_out && !fix_h) {
    size_t bytes = sizeof(float) * (size_t) (batch * F * K);
    FB_CUDA(p, p->scale.ensure(sizeof(float) * (size_t) batch));
    launch_h_max_scale(p, d, p->scale.as<float>());
    void* dst = a->acts_out;
    if (host) { FB_CUDA(p, p->out_b.ensure(bytes)); dst = p->out_b.p; }
    launch_copy3d(p, d.H, FB200_F32, (int64_t) d.Fp * d.KP, d.KP, dst, FB200_F32, F * K, K, batch, F, K, p->scale.as<float>(), 0);
    if (host) FB_CUDA(p, cudaMemcpyAsync(a->acts_out, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
  }
  t.mark(5);
  // resynthesis (:302-333): rank masked spectra per channel -> ISTFT
  if (resynth) {
    int64_t wave = std::max<int64_t>(1, wave_size(p, F, batch * K, 1) / K);
    wave = std::min(wave, batch);
    // fused path: the masks are applied while the inverse kernel loads its spectrum rows, the masked spectra are never
    // written; else mask kernel -> cuFFT C2R -> overlap-add
    const bool fused_inv = istft_fused_eligible(p, wave * K, F, n, p->win / 2);
    if (fused_inv) FB_CUDA(p, p->frames.ensure(sizeof(float) * (size_t) (wave * F * B)));
    else FB_CUDA(p, p->cspec.ensure(sizeof(float2) * (size_t) (wave * K * F * B)));
    float* dst_all = a->resynth_out;
    if (host) { FB_CUDA(p, p->stage.ensure(sizeof(float) * (size_t) (wave * K * n))); }
    for (int64_t b0 = 0; b0 < batch; b0 += wave) {
      int64_t nb = std::min(wave, batch - b0);
      float* dst = host ? p->stage.as<float>() : dst_all + b0 * K * n;
      int32_t stf = fused_inv ? launch_masked_istft_fused(p, d, spec_all, b0, nb, n, dst, p->win / 2, p->frames.as<float>()) : FB200_ERR_UNSUPPORTED;
      if (stf == FB200_ERR_UNSUPPORTED) {
        FB_CUDA(p, p->cspec.ensure(sizeof(float2) * (size_t) (wave * K * F * B)));
        launch_mask(p, d, spec_all, b0, nb, p->cspec.as<float2>());
        FB_TRY(run_istft(p, p->cspec.as<float2>(), nb * K, F, n, dst, p->win / 2));
      } else if (stf < 0) return stf;
      if (host) {
        FB_CUDA(p, cudaMemcpyAsync(a->resynth_out + b0 * K * n, dst, sizeof(float) * (size_t) (nb * K * n), cudaMemcpyDeviceToHost, p->stream));
        FB_CUDA(p, cudaStreamSynchronize(p->stream)); // staging buffer is reused by the next wave
      }
    }
  }
  t.mark(6);
  FB_TRY(finish(p, t, 6));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_stft = t.ms(1, 2); p->stats.ms_init = t.ms(2, 3); p->stats.ms_nmf = t.ms(3, 4);
  p->stats.ms_post = t.ms(4, 5); p->stats.ms_resynth = t.ms(5, 6);
  return FB200_OK;
}

// NMFFilter (and, with out == NULL, NMFMatch) over a mono stream that starts from reset state.
// NMFFilterClient.hpp:98-117 / NMFMatchClient.hpp:106-118 under STFTBufferedProcess (BufferedProcess.hpp:49-93,187-241):
// frame f (f*hop < n) sees stream[f*hop - win, f*hop); its `rank` masked resyntheses are overlap-added at [f*hop, f*hop+win).
// Frames are independent (fixed bases, identical h0), so the stream is cut into chunks of frames sized for the cuFFT
// scratch; a chunk re-computes the (win-1)/hop frames before its first output sample instead of carrying OLA state.
int32_t fb200_nmf_filter(fb200_plan* p, const fb200_filter_args* a)
{
  if (!p) return FB200_ERR_INVALID;
  if (!a || a->struct_size != sizeof(fb200_filter_args) || !a->audio || !a->bases || a->n_samples <= 0 || a->rank <= 0 ||
      a->iterations < 0 || (!a->out && !a->acts_out)) {
    p->err = "fb200_nmf_filter: bad arguments";
    return FB200_ERR_INVALID;
  }
  FB_TRY(enter(p, a->mem));
  StageTimer t(p);
  t.mark(0);
  const int64_t n = a->n_samples, K = a->rank, B = p->bins, hop = p->hop, win = p->win;
  const int64_t frames_total = (n + hop - 1) / hop;                      // BufferedProcess.hpp:57
  const int64_t ov = a->out ? (win - 1) / hop : 0;                       // earlier frames that still reach a sample
  // frames per chunk: bounded by the cuFFT scratch of the K masked resyntheses, or, for activations only (NMFMatch),
  // by that of the forward transform alone (K times larger chunks: fewer, better filled launches)
  int64_t chunk = (((int64_t) 1 << 27) / std::max<int64_t>(1, (a->out ? K : 1) * p->fft)) / 128 * 128;
  chunk = std::max<int64_t>(chunk, (ov / 128 + 1) * 128);
  chunk = std::min<int64_t>(chunk, (frames_total + ov + 127) / 128 * 128);
  const int64_t fresh = chunk - ov;                                      // new frames per chunk
  const int host = a->mem == FB200_HOST;

  const void* raw;
  FB_TRY(to_device_raw(p, a->audio, a->mem, sizeof(float) * (size_t) n, p->audio, &raw));
  const float* d_audio = (const float*) raw;
  FB_TRY(to_device_raw(p, a->bases, a->mem, sizeof(float) * (size_t) (K * B), p->out_a, &raw));
  const float* d_bases = (const float*) raw;
  const float* U0 = nullptr;
  int per_frame = 0;
  FB_TRY(draw_frame_h0(p, a->seed, frames_total, K, &U0, &per_frame)); // NMF.hpp:55: same h0 for every frame when seeded
  FB_CUDA(p, p->spec.ensure(sizeof(float2) * (size_t) (chunk * B)));
  if (a->out) FB_CUDA(p, p->cspec.ensure(sizeof(float2) * (size_t) (K * chunk * B)));
  if (host && a->out) FB_CUDA(p, p->stage.ensure(sizeof(float) * (size_t) (K * fresh * hop)));
  if (host && a->acts_out) FB_CUDA(p, p->out_b.ensure(sizeof(float) * (size_t) (fresh * K)));
  t.mark(1);
  for (int64_t f_new = 0; f_new < frames_total; f_new += fresh) {
    const int64_t f_lo = std::max<int64_t>(0, f_new - ov);
    const int64_t f_hi = std::min(frames_total, f_new + fresh);
    NmfDev d{};
    d.batch = 1; d.F = (int) (f_hi - f_lo); d.B = (int) B; d.K = (int) K; d.clamp_v = 1;
    FB_TRY(alloc_nmf(p, d));
    FB_CUDA(p, cudaMemsetAsync(d.V, 0, sizeof(float) * (size_t) d.Fp * d.Bp, p->stream));
    FB_TRY(run_stft(p, d_audio, 1, n, d.F, d.V, d.Fp, d.Bp, p->spec.as<float2>(), win - f_lo * hop));
    launch_nmf_init(p, d, U0 + (per_frame ? f_lo * K : 0), U0 + (per_frame ? f_lo * K : 0), K, d_bases, nullptr, 1 + per_frame); // NMF.hpp:55-64
    if (a->iterations > 0) FB_TRY(run_h_only(p, d, a->iterations));      // NMF.hpp:72-83
    if (a->acts_out) {
      const int64_t rows = f_hi - f_new;
      const float* src = d.H + (f_new - f_lo) * d.KP;
      float* dst = host ? p->out_b.as<float>() : a->acts_out + f_new * K;
      launch_copy3d(p, src, FB200_F32, 0, d.KP, dst, FB200_F32, 0, K, 1, rows, K, nullptr, 0);
      if (host)
        FB_CUDA(p, cudaMemcpyAsync(a->acts_out + f_new * K, dst, sizeof(float) * (size_t) (rows * K), cudaMemcpyDeviceToHost,
                                   p->stream));
    }
    if (a->out) {
      const int64_t t_lo = f_new * hop, t_hi = std::min(n, f_hi * hop), len = t_hi - t_lo;
      launch_mask(p, d, p->spec.as<float2>(), 0, 1, p->cspec.as<float2>()); // NMFFilterClient.hpp:104-113
      float* dst = host ? p->stage.as<float>() : a->out + t_lo;
      const int64_t stride = host ? len : n;
      FB_TRY(run_istft(p, p->cspec.as<float2>(), K, d.F, len, dst, t_lo - f_lo * hop, stride, 1));
      if (host)
        FB_CUDA(p, cudaMemcpy2DAsync(a->out + t_lo, sizeof(float) * (size_t) n, dst, sizeof(float) * (size_t) len,
                                     sizeof(float) * (size_t) len, (size_t) K, cudaMemcpyDeviceToHost, p->stream));
    }
    if (host) FB_CUDA(p, cudaStreamSynchronize(p->stream)); // staging buffers are reused by the next chunk
  }
  t.mark(2);
  FB_TRY(finish(p, t, 2));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_nmf = t.ms(1, 2);
  return FB200_OK;
}

// The per-frame body of the streaming clients for a batch of frames that a host-side BufferedProcess has already cut:
// STFT::processFrame (window + rFFT) -> |X| -> NMF::processFrame -> [NMFFilter only: estimate + RatioMask per component ->
// ISTFT::processFrame].  NMFFilterClient.hpp:98-117, NMFMatchClient.hpp:111-118, BufferedProcess.hpp:200-216.
int32_t fb200_nmf_filter_frames(fb200_plan* p, const fb200_filter_frames_args* a)
{
  if (!p) return FB200_ERR_INVALID;
  if (!a || a->struct_size != sizeof(fb200_filter_frames_args) || !a->in || !a->bases || a->frames <= 0 || a->rank <= 0 ||
      a->iterations < 0 || (!a->out && !a->acts_out)) {
    p->err = "fb200_nmf_filter_frames: bad arguments";
    return FB200_ERR_INVALID;
  }
  FB_TRY(enter(p, a->mem));
  StageTimer t(p);
  t.mark(0);
  const int64_t F = a->frames, K = a->rank, B = p->bins, win = p->win;
  const int host = a->mem == FB200_HOST;
  const void* raw;
  FB_TRY(to_device_raw(p, a->in, a->mem, sizeof(float) * (size_t) (F * win), p->audio, &raw));
  const float* d_in = (const float*) raw;
  FB_TRY(to_device_raw(p, a->bases, a->mem, sizeof(float) * (size_t) (K * B), p->out_a, &raw));
  const float* d_bases = (const float*) raw;
  const float* U0 = nullptr;
  int per_frame = 0;
  FB_TRY(draw_frame_h0(p, a->seed, F, K, &U0, &per_frame));
  NmfDev d{};
  d.batch = 1; d.F = (int) F; d.B = (int) B; d.K = (int) K; d.clamp_v = 1;
  FB_TRY(alloc_nmf(p, d));
  FB_CUDA(p, p->spec.ensure(sizeof(float2) * (size_t) (F * B)));
  t.mark(1);
  FB_CUDA(p, cudaMemsetAsync(d.V, 0, sizeof(float) * (size_t) d.Fp * d.Bp, p->stream));
  // the frames are one contiguous "signal" cut with hop = win and no padding: frame f = in[f*win, (f+1)*win)
  FB_TRY(run_stft(p, d_in, 1, F * win, F, d.V, d.Fp, d.Bp, p->spec.as<float2>(), 0, nullptr, (int) win));
  launch_nmf_init(p, d, U0, U0, K, d_bases, nullptr, 1 + per_frame);     // NMF.hpp:55-64
  if (a->iterations > 0) FB_TRY(run_h_only(p, d, a->iterations));        // NMF.hpp:72-83
  if (a->acts_out) {
    float* dst = a->acts_out;
    if (host) { FB_CUDA(p, p->out_b.ensure(sizeof(float) * (size_t) (F * K))); dst = p->out_b.as<float>(); }
    launch_copy3d(p, d.H, FB200_F32, 0, d.KP, dst, FB200_F32, 0, K, 1, F, K, nullptr, 0);
    if (host) FB_CUDA(p, cudaMemcpyAsync(a->acts_out, dst, sizeof(float) * (size_t) (F * K), cudaMemcpyDeviceToHost, p->stream));
  }
  if (a->out) {
    FB_CUDA(p, p->cspec.ensure(sizeof(float2) * (size_t) (K * F * B)));
    launch_mask(p, d, p->spec.as<float2>(), 0, 1, p->cspec.as<float2>());  // NMFFilterClient.hpp:104-113
    FB_CUDA(p, p->frames.ensure(sizeof(float) * (size_t) (K * F * p->fft)));
    cufftHandle h;
    FB_TRY(get_fft_plan(p, CUFFT_C2R, K * F, &h));
    FB_CUFFT(p, cufftExecC2R(h, reinterpret_cast<cufftComplex*>(p->cspec.p), p->frames.as<float>()));
    p->launches++;
    float* dst = a->out;
    if (host) { FB_CUDA(p, p->stage.ensure(sizeof(float) * (size_t) (F * K * win))); dst = p->stage.as<float>(); }
    launch_window_frames(p, p->frames.as<float>(), K, F, dst);
    if (host) FB_CUDA(p, cudaMemcpyAsync(a->out, dst, sizeof(float) * (size_t) (F * K * win), cudaMemcpyDeviceToHost, p->stream));
  }
  t.mark(2);
  FB_TRY(finish(p, t, 2));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_nmf = t.ms(1, 2);
  return FB200_OK;
}

// GriffinLim::process (GriffinLim.hpp:29-54) on one spectrogram spec[F][B], in place.  Work arrays: x1 (phase), x2 (mag),
// x4 / x5 (estimate / previous estimate), cspec (mag * phase, consumed by the inverse transform), audio (time signal).
static int32_t run_griffinlim(Plan* p, float2* spec, int64_t F, int64_t n, int iters, int64_t seed)
{
  const int64_t B = p->bins, cnt = F * B;
  bool neg = false;
  FB_TRY(upload_seeds(p, &seed, 1, &neg));
  FB_CUDA(p, p->rnd.ensure(sizeof(float) * (size_t) cnt));
  launch_mt_uniform(p, p->seeds.as<int64_t>(), 1, cnt, p->rnd.as<float>());
  FB_CUDA(p, p->x1.ensure(sizeof(float2) * (size_t) cnt));
  FB_CUDA(p, p->x2.ensure(sizeof(float) * (size_t) cnt));
  FB_CUDA(p, p->x4.ensure(sizeof(float2) * (size_t) cnt));
  FB_CUDA(p, p->x5.ensure(sizeof(float2) * (size_t) cnt));
  FB_CUDA(p, p->cspec.ensure(sizeof(float2) * (size_t) cnt));
  FB_CUDA(p, p->audio.ensure(sizeof(float) * (size_t) n));
  float2* phase = p->x1.as<float2>();
  float* mag = p->x2.as<float>();
  float2* est = p->x4.as<float2>();
  float2* prev = p->x5.as<float2>();
  launch_gl_init(p, spec, p->rnd.as<float>(), (int) F, (int) B, mag, phase);                  // :39-41
  FB_CUDA(p, cudaMemsetAsync(est, 0, sizeof(float2) * (size_t) cnt, p->stream));              // :42-43
  for (int i = 0; i < iters; i++) {                                                            // :44-52
    std::swap(est, prev);                                                                      // prev = estimate
    launch_gl_apply(p, mag, phase, cnt, p->cspec.as<float2>());                                // magnitude * phase
    FB_TRY(run_istft(p, p->cspec.as<float2>(), 1, F, n, p->audio.as<float>(), p->win / 2));
    FB_TRY(run_stft(p, p->audio.as<float>(), 1, n, F, nullptr, F, B, est, p->win / 2));
    launch_gl_phase(p, est, prev, cnt, phase);
  }
  launch_gl_apply(p, mag, phase, cnt, spec);                                                   // :53
  return FB200_OK;
}

// BufNMFCross: NMFCrossClient::process (clients/nrt/NMFCrossClient.hpp:85-185), channel 0 of source and target.
int32_t fb200_bufnmfcross(fb200_plan* p, const fb200_nmfcross_args* a)
{
  if (!p) return FB200_ERR_INVALID;
  if (!a || a->struct_size != sizeof(fb200_nmfcross_args) || !a->source || !a->target || (!a->out && !a->acts_out) ||
      a->iterations < 1 || a->time_sparsity < 1 || a->polyphony < 1 || a->continuity < 1 || a->griffinlim_iterations < 0) {
    p->err = "fb200_bufnmfcross: bad arguments";
    return FB200_ERR_INVALID;
  }
  if (a->n_source <= 0) { p->err = "Empty source buffer"; return FB200_ERR_INVALID; }               // :110
  if (a->n_target <= 0) { p->err = "Empty target buffer"; return FB200_ERR_INVALID; }               // :112
  const int64_t ns = a->n_source, nt = a->n_target, B = p->bins;
  const int64_t Fs = fb200_num_frames(ns, p->win, p->hop), Ft = fb200_num_frames(nt, p->win, p->hop); // :103-108
  if (a->time_sparsity > Ft) { p->err = "Time Sparsity is larger than target frames"; return FB200_ERR_INVALID; } // :114-116
  if (a->continuity > Ft) { p->err = "Continuity is larger than target frames"; return FB200_ERR_INVALID; }       // :117-119
  FB_TRY(enter(p, a->mem));
  StageTimer t(p);
  t.mark(0);
  const int host = a->mem == FB200_HOST;
  const int64_t R = Fs;                                                                           // rank = source frames (:140)
  const int poly = (int) std::min<int64_t>(Fs, a->polyphony);                                     // :160
  const void* raw;
  FB_TRY(to_device_raw(p, a->source, a->mem, sizeof(float) * (size_t) ns, p->audio, &raw));
  const float* d_src = (const float*) raw;
  FB_TRY(to_device_raw(p, a->target, a->mem, sizeof(float) * (size_t) nt, p->stage, &raw));
  const float* d_tgt = (const float*) raw;
  t.mark(1);
  // source: spectrum S[R][B] (kept for the resynthesis) and W = |S|; target: V = |T|   (:134-139)
  FB_CUDA(p, p->spec.ensure(sizeof(float2) * (size_t) (Fs * B)));
  FB_CUDA(p, p->W.ensure(sizeof(float) * (size_t) (R * B)));
  FB_CUDA(p, p->V.ensure(sizeof(float) * (size_t) (Ft * B)));
  FB_CUDA(p, p->cspec.ensure(sizeof(float2) * (size_t) (Ft * B)));
  FB_TRY(run_stft(p, d_src, 1, ns, Fs, p->W.as<float>(), Fs, B, p->spec.as<float2>(), p->win / 2));
  FB_TRY(run_stft(p, d_tgt, 1, nt, Ft, p->V.as<float>(), Ft, B, p->cspec.as<float2>(), p->win / 2));
  t.mark(2);
  // NMFCross::process (:60-76): H = U(R x F) column-major == H[F][R] in storage order; W clamped at eps
  FB_CUDA(p, p->hden.ensure(sizeof(float) * (size_t) (2 * R)));
  float* energy = p->hden.as<float>();
  float* hden = energy + R;
  launch_cross_prepare(p, p->W.as<float>(), (int) R, (int) B, energy, hden);                      // :156-159, :169
  FB_CUDA(p, p->x0.ensure(sizeof(float) * (size_t) (Ft * R)));
  FB_CUDA(p, p->x1.ensure(sizeof(float) * (size_t) (Ft * R)));
  FB_CUDA(p, p->x2.ensure(sizeof(float) * (size_t) (Ft * B)));
  float* H = p->x0.as<float>();
  float* T1 = p->x1.as<float>();
  float* ratio = p->x2.as<float>();
  {
    bool neg = false;
    int64_t seed = a->seed;
    FB_TRY(upload_seeds(p, &seed, 1, &neg));
    launch_mt_uniform(p, p->seeds.as<int64_t>(), 1, R * Ft, H);                                    // :72-73
  }
  t.mark(3);
  p->backend_used = FB200_BACKEND_SIMT;
  bool cancelled = false;
  for (int i = 0; i < a->iterations && !cancelled; i++) {                                          // :160-176
    // 1 - ((i + 1) / iterations) in INTEGER arithmetic (:119, :136): 1 until the last iteration, 0 in the last
    const float factor = 1.0f - (float) ((i + 1) / a->iterations);
    if (factor != 1.0f) {
      launch_cross_sparseness(p, H, T1, (int) Ft, (int) R, a->time_sparsity, factor);              // :163
      launch_cross_polyphony(p, T1, H, (int) Ft, (int) R, energy, poly, factor);                   // :164
    }
    launch_cross_continuity(p, H, T1, (int) Ft, (int) R, a->continuity);                           // :165
    launch_cross_ratio(p, T1, p->W.as<float>(), p->V.as<float>(), ratio, (int) Ft, (int) B, (int) R);          // :167-168
    launch_cross_update(p, ratio, p->W.as<float>(), T1, H, hden, (int) Ft, (int) B, (int) R);      // :168-170
    if (a->progress) {
      FB_CUDA(p, cudaStreamSynchronize(p->stream));
      if (!a->progress(a->progress_user, i + 1)) cancelled = true;                                 // :174-175
    }
  }
  t.mark(4);
  if (cancelled) {
    FB_TRY(finish(p, t, 4));
    return FB200_CANCELLED;
  }
  if (a->acts_out) {
    float* dst = a->acts_out;
    const size_t bytes = sizeof(float) * (size_t) (Ft * R);
    if (host) FB_CUDA(p, cudaMemcpyAsync(dst, H, bytes, cudaMemcpyDeviceToHost, p->stream));
    else FB_CUDA(p, cudaMemcpyAsync(dst, H, bytes, cudaMemcpyDeviceToDevice, p->stream));
  }
  if (a->out) {
    // NMFCross::synthesize (:50-58): result[F][B] = H[F][R] * S[R][B]; S complex == [R][2B] real
    FB_CUDA(p, p->x3.ensure(sizeof(float2) * (size_t) (Ft * B)));
    launch_sgemm_nn(p, H, reinterpret_cast<const float*>(p->spec.p), p->x3.as<float>(), (int) Ft, (int) (2 * B), (int) R);
    if (a->progress && !a->progress(a->progress_user, a->iterations + 1)) cancelled = true;        // :168-169
    if (!cancelled) {
      FB_TRY(run_griffinlim(p, p->x3.as<float2>(), Ft, nt, a->griffinlim_iterations, a->seed));    // :171-173
      if (a->progress && !a->progress(a->progress_user, a->iterations + 2)) cancelled = true;
    }
    if (!cancelled) {
      float* d_out = a->out;
      if (host) { FB_CUDA(p, p->out_a.ensure(sizeof(float) * (size_t) nt)); d_out = p->out_a.as<float>(); }
      FB_TRY(run_istft(p, p->x3.as<float2>(), 1, Ft, nt, d_out, p->win / 2));                      // :178
      if (host) FB_CUDA(p, cudaMemcpyAsync(a->out, d_out, sizeof(float) * (size_t) nt, cudaMemcpyDeviceToHost, p->stream));
      if (a->progress && !a->progress(a->progress_user, a->iterations + 3)) cancelled = true;
    }
  }
  t.mark(5);
  FB_TRY(finish(p, t, 5));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_stft = t.ms(1, 2); p->stats.ms_init = t.ms(2, 3); p->stats.ms_nmf = t.ms(3, 4);
  p->stats.ms_resynth = t.ms(4, 5);
  return cancelled ? FB200_CANCELLED : FB200_OK;
}

// NMFSeed: NMFSeedClient::process (clients/nrt/NMFSeedClient.hpp:74-133) = STFT -> magnitude -> NNDSVD::process
// (algorithms/public/NNDSVD.hpp:30-131) -> bases rows / activations scaled by 1 / max
int32_t fb200_nmfseed(fb200_plan* p, const fb200_nmfseed_args* a)
{
  if (!p) return FB200_ERR_INVALID;
  if (!a || a->struct_size != sizeof(fb200_nmfseed_args) || (!a->audio && !a->mags) || !a->rank_out || a->min_rank < 0 ||
      a->max_rank < 1 || a->min_rank > a->max_rank || a->method < 0 || a->method > 3 || a->coverage < 0 || a->coverage > 1 ||
      !(a->coverage > 0 || a->min_rank > 0)) {   // NNDSVD.hpp:39 assert(amount > 0 || minRank > 0)
    p->err = "fb200_nmfseed: bad arguments";
    return FB200_ERR_INVALID;
  }
  FB_TRY(enter(p, a->mem));
  StageTimer t(p);
  t.mark(0);
  const int64_t B = p->bins;
  const int host = a->mem == FB200_HOST;
  int64_t F = a->frames;
  const float* d_mags = nullptr;
  if (a->audio) {
    if (a->n_samples <= 0) { p->err = "fb200_nmfseed: audio given without n_samples"; return FB200_ERR_INVALID; }
    F = fb200_num_frames(a->n_samples, p->win, p->hop);                                       // :84-86
    const void* raw;
    FB_TRY(to_device_raw(p, a->audio, a->mem, sizeof(float) * (size_t) a->n_samples, p->audio, &raw));
    FB_CUDA(p, p->V.ensure(sizeof(float) * (size_t) (F * B)));
    FB_TRY(run_stft(p, (const float*) raw, 1, a->n_samples, F, p->V.as<float>(), F, B, nullptr, p->win / 2)); // :99-100
    d_mags = p->V.as<float>();
  } else {
    if (F <= 0) { p->err = "fb200_nmfseed: frames must be positive"; return FB200_ERR_INVALID; }
    const void* raw;
    FB_TRY(to_device_raw(p, a->mags, a->mem, sizeof(float) * (size_t) (F * B), p->stage, &raw));
    d_mags = (const float*) raw;
  }
  t.mark(1);
  int n = 0;
  FB_TRY(run_seed_svd(p, d_mags, (int) F, (int) B, &n));
  // singular values to the host: order, coverage -> rank  (NNDSVD.hpp:46-58)
  std::vector<double> nrm((size_t) n);
  FB_CUDA(p, cudaMemcpyAsync(nrm.data(), p->x3.p, sizeof(double) * (size_t) n, cudaMemcpyDeviceToHost, p->stream));
  FB_CUDA(p, cudaStreamSynchronize(p->stream));
  std::vector<int> ord((size_t) n);
  for (int i = 0; i < n; i++) ord[(size_t) i] = i;
  std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return nrm[(size_t) x] > nrm[(size_t) y]; });
  const int64_t r = std::min<int64_t>(B, F);
  int64_t k = 0;
  if (a->coverage == 0) k = a->min_rank;
  else {
    double cur = 0, total = 0;
    for (int64_t i = 0; i < r; i++) total += nrm[(size_t) ord[(size_t) i]];
    while ((cur / total) < a->coverage && k < r) cur += nrm[(size_t) ord[(size_t) k++]];
  }
  if (k < a->min_rank) k = a->min_rank;
  if (k > a->max_rank) k = a->max_rank;
  if (k > r) k = r;
  *a->rank_out = (int32_t) k;
  if (a->singular_values)
    for (int64_t i = 0; i < r; i++) a->singular_values[i] = nrm[(size_t) ord[(size_t) i]];
  const int64_t R = a->max_rank;
  FB_CUDA(p, p->W.ensure(sizeof(float) * (size_t) (R * B)));
  FB_CUDA(p, p->H.ensure(sizeof(float) * (size_t) (F * R)));
  FB_CUDA(p, cudaMemsetAsync(p->W.p, 0, sizeof(float) * (size_t) (R * B), p->stream));      // the client allocates zero-filled tensors (:93-94)
  FB_CUDA(p, cudaMemsetAsync(p->H.p, 0, sizeof(float) * (size_t) (F * R), p->stream));
  FB_CUDA(p, p->x4.ensure(sizeof(int) * (size_t) n));
  FB_CUDA(p, cudaMemcpyAsync(p->x4.p, ord.data(), sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, p->stream));
  launch_seed_build(p, p->x4.as<int>(), (int) k, (int) F, (int) B, n, (int) R, a->method, p->W.as<float>(), p->H.as<float>());
  if (a->method == 1 || a->method == 2) {
    const float* U = nullptr;
    if (a->method == 1) { // both generators restart from the seed (:108-113); one stream long enough for the larger matrix
      bool neg = false;
      int64_t seed = a->seed;
      FB_TRY(upload_seeds(p, &seed, 1, &neg));
      const int64_t cnt = std::max(R * B, F * R);
      FB_CUDA(p, p->rnd.ensure(sizeof(float) * (size_t) cnt));
      launch_mt_uniform(p, p->seeds.as<int64_t>(), 1, cnt, p->rnd.as<float>());
      U = p->rnd.as<float>();
    }
    launch_seed_fill(p, p->W.as<float>(), p->H.as<float>(), (int) F, (int) B, (int) R, a->method, d_mags, U);
  }
  FB_CUDA(p, cudaStreamSynchronize(p->stream)); // `ord` dies at scope exit
  t.mark(2);
  // outputs: bases [max_rank][B]; activations [F][max_rank], scaled by 1 / max over ALL of H as the client does (:119-128)
  if (a->bases) {
    const size_t bytes = sizeof(float) * (size_t) (R * B);
    FB_CUDA(p, cudaMemcpyAsync(a->bases, p->W.p, bytes, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, p->stream));
  }
  if (a->acts) {
    NmfDev d{};
    d.batch = 1; d.F = (int) F; d.Fp = (int) F; d.K = (int) R; d.KP = (int) R; d.H = p->H.as<float>();
    const float* scale = nullptr;
    if (a->scale_acts) {
      FB_CUDA(p, p->scale.ensure(sizeof(float)));
      launch_h_max_scale(p, d, p->scale.as<float>());
      scale = p->scale.as<float>();
    }
    float* dst = a->acts;
    const size_t bytes = sizeof(float) * (size_t) (F * R);
    if (host) { FB_CUDA(p, p->out_b.ensure(bytes)); dst = p->out_b.as<float>(); }
    launch_copy3d(p, p->H.p, FB200_F32, 0, R, dst, FB200_F32, 0, R, 1, F, R, scale, 0);
    if (host) FB_CUDA(p, cudaMemcpyAsync(a->acts, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
  }
  t.mark(3);
  FB_TRY(finish(p, t, 3));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_init = t.ms(1, 2); p->stats.ms_d2h = t.ms(2, 3);
  return FB200_OK;
}

// MelBands over a frame sequence: MelBands::init (MelBands.hpp:43-80, evaluated on the host in fp64) + processFrame (:82-101)
int32_t fb200_melbands(fb200_plan* p, const fb200_melbands_args* a)
{
  if (!p) return FB200_ERR_INVALID;
  if (!a || a->struct_size != sizeof(fb200_melbands_args) || a->batch <= 0 || !a->bands || (!a->mags && !a->audio) ||
      a->n_bands < 2 || !(a->hi > a->lo) || a->sample_rate <= 0) {   // asserts of init(): hi > lo, nBands > 1
    p->err = "fb200_melbands: bad arguments";
    return FB200_ERR_INVALID;
  }
  FB_TRY(enter(p, a->mem));
  StageTimer t(p);
  t.mark(0);
  const int64_t B = p->bins, nb = a->n_bands;
  const int host = a->mem == FB200_HOST;
  int64_t F = a->frames;
  const float* d_mags = nullptr;
  if (a->audio) { // STFT::process + magnitude first (STFT.hpp:90-108, 61-66)
    if (a->n_samples <= 0) { p->err = "fb200_melbands: audio given without n_samples"; return FB200_ERR_INVALID; }
    F = fb200_num_frames(a->n_samples, p->win, p->hop);
    const void* raw;
    FB_TRY(to_device_raw(p, a->audio, a->mem, sizeof(float) * (size_t) (a->batch * a->n_samples), p->audio, &raw));
    FB_CUDA(p, p->V.ensure(sizeof(float) * (size_t) (a->batch * F * B)));
    FB_TRY(run_stft(p, (const float*) raw, a->batch, a->n_samples, F, p->V.as<float>(), F, B, nullptr, p->win / 2));
    d_mags = p->V.as<float>();
  } else {
    if (F <= 0) { p->err = "fb200_melbands: frames must be positive"; return FB200_ERR_INVALID; }
    const void* raw;
    FB_TRY(to_device_raw(p, a->mags, a->mem, sizeof(float) * (size_t) (a->batch * F * B), p->stage, &raw));
    d_mags = (const float*) raw;
  }
  // MelBands::init: triangular filters between nBands + 2 points equally spaced on the mel scale
  std::vector<float> filt((size_t) (nb * B));
  const double fft = 2.0 * (double) (B - 1);
  const double scale1 = 1.0 / ((double) p->win / 4.0), scale2 = 1.0 / (2.0 * fft / (double) p->win);   // :50-53
  {
    auto hz2mel = [](double x) { return 1127.01048 * std::log(x / 700.0 + 1.0); };                   // :38-41
    const double mlo = hz2mel(a->lo), mhi = hz2mel(a->hi);
    auto lin = [](int64_t n, double lo, double hi, int64_t i) {  // Eigen LinSpaced for doubles (evaluated from the end nearer zero)
      if (n == 1) return hi;
      const double step = (hi - lo) / (double) (n - 1);
      if (std::fabs(hi) < std::fabs(lo)) return i == 0 ? lo : hi - (double) (n - 1 - i) * step;
      return i == n - 1 ? hi : lo + (double) i * step;
    };
    std::vector<double> mel((size_t) (nb + 2));
    for (int64_t i = 0; i < nb + 2; i++) mel[(size_t) i] = 700.0 * (std::exp(lin(nb + 2, mlo, mhi, i) / 1127.01048) - 1.0); // :55-56
    for (int64_t i = 0; i < nb; i++) {
      const double d0 = std::fabs(mel[(size_t) i] - mel[(size_t) i + 1]), d1 = std::fabs(mel[(size_t) i + 1] - mel[(size_t) i + 2]);
      for (int64_t b = 0; b < B; b++) {
        const double f = lin(B, 0.0, a->sample_rate / 2.0, b);                                         // :60
        const double lower = -(mel[(size_t) i] - f) / d0, upper = (mel[(size_t) i + 2] - f) / d1;      // :72-73
        filt[(size_t) (i * B + b)] = (float) std::max(0.0, std::min(lower, upper));                    // :74
      }
    }
  }
  FB_CUDA(p, p->x0.ensure(sizeof(float) * filt.size()));
  FB_CUDA(p, cudaMemcpyAsync(p->x0.p, filt.data(), sizeof(float) * filt.size(), cudaMemcpyHostToDevice, p->stream));
  FB_CUDA(p, cudaStreamSynchronize(p->stream)); // `filt` dies at scope exit
  t.mark(1);
  float* d_out = a->bands;
  const size_t obytes = sizeof(float) * (size_t) (a->batch * F * nb);
  if (host) { FB_CUDA(p, p->out_a.ensure(obytes)); d_out = p->out_a.as<float>(); }
  launch_melbands(p, d_mags, p->x0.as<float>(), a->batch * F, (int) B, (int) nb, (float) scale1, (float) scale2, a->flags, d_out);
  if (host) FB_CUDA(p, cudaMemcpyAsync(a->bands, d_out, obytes, cudaMemcpyDeviceToHost, p->stream));
  t.mark(2);
  FB_TRY(finish(p, t, 2));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_post = t.ms(1, 2);
  return FB200_OK;
}

// HPSS::processFrame over frame sequences from init() state (HPSS.hpp:47-162): spectrum [batch][F][B] -> out [batch][3][F][B]
int32_t fb200_hpss(fb200_plan* p, const fb200_hpss_args* a)
{
  if (!p) return FB200_ERR_INVALID;
  if (!a || a->struct_size != sizeof(fb200_hpss_args) || a->batch <= 0 || a->frames <= 0 || !a->spectrum || !a->out ||
      a->v_size < 3 || a->h_size < 3 || !(a->v_size & 1) || !(a->h_size & 1) || a->v_size > 129 || a->h_size > 129 ||
      a->v_size > p->bins || a->mode < 0 || a->mode > 2) {   // MedianFilter::init asserts size >= 3 and odd; processFrame vSize <= bins
    p->err = "fb200_hpss: bad arguments (filter sizes must be odd, 3 .. 129, percussive size <= bins; mode 0 .. 2)";
    return FB200_ERR_INVALID;
  }
  FB_TRY(enter(p, a->mem));
  StageTimer t(p);
  t.mark(0);
  const int64_t B = p->bins, F = a->frames, cnt = a->batch * F * B;
  const int host = a->mem == FB200_HOST;
  const void* raw;
  FB_TRY(to_device_raw(p, a->spectrum, a->mem, sizeof(float2) * (size_t) cnt, p->stage, &raw));
  // makeThreshold (:164-181), fp64 on the host
  std::vector<float> th((size_t) (2 * B), 1.0f);
  for (int which = 0; which < 2; which++) {
    const double x1 = a->thresholds[4 * which], y1 = a->thresholds[4 * which + 1], x2 = a->thresholds[4 * which + 2], y2 = a->thresholds[4 * which + 3];
    const int64_t ks = (int64_t) std::floor(x1 * (double) B), ke = (int64_t) std::floor(x2 * (double) B), kl = ke - ks;
    float* th1 = th.data() + which * B;
    for (int64_t i = 0; i < ks && i < B; i++) th1[i] = (float) std::pow(10.0, y1 / 20.0);
    for (int64_t i = 0; i < kl && ks + i < B; i++) {
      const double step = kl > 1 ? (y2 - y1) / (double) (kl - 1) : 0.0;
      const double y = kl == 1 ? y2 : (std::fabs(y2) < std::fabs(y1) ? (i == 0 ? y1 : y2 - (double) (kl - 1 - i) * step) : (i == kl - 1 ? y2 : y1 + (double) i * step));
      th1[ks + i] = (float) std::pow(10.0, y / 20.0);
    }
    for (int64_t i = std::max<int64_t>(ke, 0); i < B; i++) th1[i] = (float) std::pow(10.0, y2 / 20.0);
  }
  FB_CUDA(p, p->x0.ensure(sizeof(float) * th.size()));
  FB_CUDA(p, cudaMemcpyAsync(p->x0.p, th.data(), sizeof(float) * th.size(), cudaMemcpyHostToDevice, p->stream));
  FB_CUDA(p, cudaStreamSynchronize(p->stream));
  FB_CUDA(p, p->x1.ensure(sizeof(float) * (size_t) cnt));
  t.mark(1);
  float2* d_out = reinterpret_cast<float2*>(a->out);
  const size_t obytes = sizeof(float2) * (size_t) (3 * cnt);
  if (host) { FB_CUDA(p, p->cspec.ensure(obytes)); d_out = p->cspec.as<float2>(); }
  launch_hpss(p, (const float2*) raw, p->x1.as<float>(), a->batch, (int) F, (int) B, a->v_size, a->h_size, a->mode, p->x0.as<float>(),
              p->x0.as<float>() + B, d_out);
  if (host) FB_CUDA(p, cudaMemcpyAsync(a->out, d_out, obytes, cudaMemcpyDeviceToHost, p->stream));
  t.mark(2);
  FB_TRY(finish(p, t, 2));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_post = t.ms(1, 2);
  return FB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int32_t fb200_bufstft_sizes(int32_t win, int32_t hop, int32_t padding_mode, int32_t invert, int64_t count, int64_t* padding,
                            int64_t* out)
{
  if (win <= 0 || padding_mode < 0 || padding_mode > 2 || count < 0) return FB200_ERR_INVALID;
  if (hop <= 0) hop = win >> 1;
  if (hop <= 0) return FB200_ERR_INVALID;
  const int64_t pad = padding_mode == 0 ? 0 : (padding_mode == 1 ? (win >> 1) : win - hop); // ParameterTypes.hpp:315-323
  if (padding) *padding = pad;
  if (!invert) {
    int64_t padded = count + 2 * pad;                                    // BufSTFTClient.hpp:121-124
    if (padding_mode == 2) padded = (padded + hop - 1) / hop * hop;      // :126-128
    if (padded < win) return FB200_ERR_INVALID;
    if (out) *out = 1 + (padded - win) / hop;                            // :130-131
  } else {
    if (count < 1) return FB200_ERR_INVALID;
    const int64_t len = (count - 1) * hop + win - pad;                   // :241-242
    if (len <= 0) return FB200_ERR_INVALID;
    if (out) *out = len;
  }
  return FB200_OK;
}

int32_t fb200_bufstft(fb200_plan* p, const fb200_bufstft_args* a)
{
  if (!p) return FB200_ERR_INVALID;
  if (!a || a->struct_size != sizeof(fb200_bufstft_args) || a->batch <= 0 || a->frames <= 0) {
    p->err = "fb200_bufstft: bad arguments";
    return FB200_ERR_INVALID;
  }
  const int64_t batch = a->batch, F = a->frames, B = p->bins;
  const int host = a->mem == FB200_HOST;
  int64_t pad = 0, derived = 0;
  FB_TRY(enter(p, a->mem));
  StageTimer t(p);
  t.mark(0);
  const size_t fb_bytes = sizeof(float) * (size_t) (batch * F * B);
  if (!a->invert) {
    // ---- processFwd (:82-190)
    if (!a->audio) { p->err = "No input buffer supplied"; return FB200_ERR_INVALID; }                    // :86
    if (!a->mag && !a->phase) { p->err = "Neither magnitude nor phase buffer supplied"; return FB200_ERR_INVALID; } // :94-96
    if (fb200_bufstft_sizes(p->win, p->hop, a->padding_mode, 0, a->n_samples, &pad, &derived) != FB200_OK || derived != F) {
      p->err = "fb200_bufstft: frames does not match fb200_bufstft_sizes for this input length";
      return FB200_ERR_INVALID;
    }
    const int64_t n = a->n_samples;
    const void* raw;
    FB_TRY(to_device_raw(p, a->audio, a->mem, sizeof(float) * (size_t) (batch * n), p->audio, &raw));
    t.mark(1);
    FB_CUDA(p, p->spec.ensure(sizeof(float2) * (size_t) (batch * F * B)));
    float* d_mag = nullptr;
    if (a->mag) {
      d_mag = a->mag;
      if (host) { FB_CUDA(p, p->out_a.ensure(fb_bytes)); d_mag = p->out_a.as<float>(); }
    }
    // frame i = padded[i*hop, i*hop + win) with the audio at offset `pad` (:148-160); dense magnitudes [F][B]
    FB_TRY(run_stft(p, (const float*) raw, batch, n, F, d_mag, F, B, p->spec.as<float2>(), pad));
    if (a->phase) {
      float* d_ph = a->phase;
      if (host) { FB_CUDA(p, p->out_b.ensure(fb_bytes)); d_ph = p->out_b.as<float>(); }
      launch_phase(p, p->spec.as<float2>(), batch * F * B, d_ph);
      if (host) FB_CUDA(p, cudaMemcpyAsync(a->phase, d_ph, fb_bytes, cudaMemcpyDeviceToHost, p->stream));
    }
    if (a->mag && host) FB_CUDA(p, cudaMemcpyAsync(a->mag, d_mag, fb_bytes, cudaMemcpyDeviceToHost, p->stream));
    t.mark(2);
    FB_TRY(finish(p, t, 2));
    p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_stft = t.ms(1, 2);
    return FB200_OK;
  }
  // ---- processInverse (:192-279)
  if (!a->mag || !a->phase) { p->err = "Need both magnutude and phase buffers for inverse transform"; return FB200_ERR_INVALID; } // :201-203
  if (!a->resynth) { p->err = "No resynthesis buffer supplied"; return FB200_ERR_INVALID; }              // :207
  if (fb200_bufstft_sizes(p->win, p->hop, a->padding_mode, 1, F, &pad, &derived) != FB200_OK) {
    p->err = "fb200_bufstft: bad frame count";
    return FB200_ERR_INVALID;
  }
  const int64_t n_out = derived;
  const void* raw_m; const void* raw_p;
  FB_TRY(to_device_raw(p, a->mag, a->mem, fb_bytes, p->out_a, &raw_m));
  FB_TRY(to_device_raw(p, a->phase, a->mem, fb_bytes, p->out_b, &raw_p));
  t.mark(1);
  FB_CUDA(p, p->cspec.ensure(sizeof(float2) * (size_t) (batch * F * B)));
  launch_polar(p, (const float*) raw_m, (const float*) raw_p, batch * F, p->cspec.as<float2>());
  float* d_out = a->resynth;
  if (host) { FB_CUDA(p, p->audio.ensure(sizeof(float) * (size_t) (batch * n_out))); d_out = p->audio.as<float>(); }
  FB_TRY(run_istft(p, p->cspec.as<float2>(), batch, F, n_out, d_out, pad));                               // :254-275
  if (host) FB_CUDA(p, cudaMemcpyAsync(a->resynth, d_out, sizeof(float) * (size_t) (batch * n_out), cudaMemcpyDeviceToHost, p->stream));
  t.mark(2);
  FB_TRY(finish(p, t, 2));
  p->stats.ms_h2d = t.ms(0, 1); p->stats.ms_resynth = t.ms(1, 2);
  return FB200_OK;
}

int32_t fb200_selftest_tcgen05(fb200_plan* p, const float* in, int64_t n_in, float* out, int64_t n_out)
{
  if (!p) return FB200_ERR_INVALID;
  const int64_t need_in = 128 * 16 + 16 * 64 + 128 * 64 + 16 * 128 + 64 * 16 + 128 * 64 + 128 * 68;
  const int64_t need_out = 128 * 64 + 128 * 16 + 128 * 64 + 128 * 16 + 128 * 32;
  if (!in || !out || n_in != need_in || n_out != need_out) { p->err = "fb200_selftest_tcgen05: bad sizes"; return FB200_ERR_INVALID; }
  FB_TRY(enter(p, FB200_HOST));
  StageTimer t(p);
  t.mark(0);
  FB_CUDA(p, p->stage.ensure(sizeof(float) * (size_t) n_in));
  FB_CUDA(p, p->out_a.ensure(sizeof(float) * (size_t) n_out));
  FB_CUDA(p, cudaMemcpyAsync(p->stage.p, in, sizeof(float) * (size_t) n_in, cudaMemcpyHostToDevice, p->stream));
  FB_CUDA(p, cudaMemsetAsync(p->out_a.p, 0, sizeof(float) * (size_t) n_out, p->stream));
  FB_TRY(run_tc_selftest(p, p->stage.as<float>(), p->out_a.as<float>()));
  FB_CUDA(p, cudaMemcpyAsync(out, p->out_a.p, sizeof(float) * (size_t) n_out, cudaMemcpyDeviceToHost, p->stream));
  if (const char* e = getenv("FB200_MMA_TIMING")) { // developer aid: cycles per small tcgen05.mma, printed to stderr
    FB_CUDA(p, p->out_b.ensure(sizeof(long long) * 32));
    FB_TRY(run_tc_mma_timing(p, p->out_b.as<long long>(), atoi(e) > 0 ? atoi(e) : 256));
    long long h[20];
    FB_CUDA(p, cudaMemcpyAsync(h, p->out_b.p, sizeof(h), cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(p, cudaStreamSynchronize(p->stream));
    for (int v = 0; v < 10; v++) fprintf(stderr, "mma_timing variant %d: issue %.1f cyc, issue+exec %.1f cyc per MMA\n", v, h[2 * v] / 1000.0, h[2 * v + 1] / 1000.0);
  }
  t.mark(1);
  return finish(p, t, 1);
}

const fb200_api* fb200_get_api(uint32_t abi_version)
{
  static const fb200_api api = {FB200_ABI_VERSION, (uint32_t) sizeof(fb200_api), fb200_device_count, fb200_plan_create,
                                fb200_plan_destroy, fb200_last_error, fb200_num_frames, fb200_resolve_fft,
                                fb200_shard_range, fb200_stft, fb200_istft, fb200_nmf_process, fb200_nmf_process_frames,
                                fb200_bufnmf, fb200_nmf_filter, fb200_get_stats, fb200_bufstft_sizes, fb200_bufstft,
                                fb200_nmf_filter_frames, fb200_bufnmf_sharded, fb200_bufnmfcross, fb200_melbands, fb200_hpss, fb200_nmfseed};
  return abi_version == FB200_ABI_VERSION ? &api : nullptr;
}

} // extern "C"
