/*
 * flucoma_oracle.c -- CPU fp64 restatement of flucoma-core's STFT -> |X| -> NMF -> mask -> ISTFT path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (flucoma-core_b200/csrc + include/flucoma_b200.h) never links, loads or calls anything in oracle/.
 *
 * PARITY STATUS: "parity unpinned".  The reference cannot be compiled in this image (its arithmetic lives in
 * Eigen 3.4.0 and HISSTools_Library@f3292ad, fetched by CMake FetchContent, CMakeLists.txt:54-71; neither is on
 * disk) and the reference's own tests hold no numerical golden vector for this path -- only seed repeatability
 * (tests/algorithms/public/TestNMF.cpp:11-73, tests/algorithms/util/TestEigenRandom.cpp:92-114) and the Hann
 * overlap-add identity (tests/clients/common/TestBufferedProcess.cpp:20-70).  Those properties are replicated in
 * tests/test_oracle.py; everything numerical is additionally cross-checked against an independent numpy/pocketfft
 * restatement (oracle/np_oracle.py) and analytic known-answer vectors.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference/include/flucoma).
 * All arithmetic is IEEE fp64, single-threaded per buffer, like the reference.  Layouts are the reference's
 * FluidTensor row-major layouts: X[F][B], W[K][B], H[F][K], V[F][B], spectrum [F][B] interleaved (re,im).
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=x86-64-v3 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FO_EPS DBL_EPSILON /* algorithms/util/AlgorithmUtils.hpp:19 */
#define FO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * Size rules.  clients/common/ParameterTypes.hpp:295-313 (FFTParams) and clients/nrt/NMFClient.hpp:111-113
 * ---------------------------------------------------------------------------------------------- */
FO_EXPORT int64_t fo_next_pow2(int64_t x)
{ /* ParameterTypes.hpp:323-335, up=true */
  if (x <= 0) return 0;
  uint32_t v = (uint32_t) x;
  --v;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return (int64_t) (v + 1);
}

FO_EXPORT void fo_fft_params(int64_t win, int64_t hop, int64_t fft, int64_t* out_win, int64_t* out_hop,
                             int64_t* out_fft, int64_t* out_bins)
{
  int64_t f = fft < 0 ? fo_next_pow2(win) : fft; /* :295-300 */
  int64_t h = hop > 0 ? hop : (win >> 1);       /* :306-309 */
  *out_win = win; *out_hop = h; *out_fft = f; *out_bins = (f >> 1) + 1; /* :312 */
}

/* nWindows of the BufNMF client, NMFClient.hpp:111-113; equals STFT::process's own count STFT.hpp:94-99 */
FO_EXPORT int64_t fo_stft_num_frames(int64_t n_samples, int64_t win, int64_t hop)
{
  int64_t padded = n_samples + win + hop;
  return (padded - win) / hop;
}

/* ------------------------------------------------------------------------------------------------
 * Hann window.  algorithms/public/WindowFuncs.hpp:41-45
 * ---------------------------------------------------------------------------------------------- */
FO_EXPORT void fo_hann(int64_t size, double* out)
{
  const double pi = 3.14159265358979323846;
  for (int64_t i = 0; i < size; i++) out[i] = 0.5 - 0.5 * cos((pi * 2 * (double) i) / (double) size);
}

/* ------------------------------------------------------------------------------------------------
 * FFT conventions.  algorithms/util/FFT.hpp:92-108 (forward: true unnormalised DFT of the zero-padded input,
 * bins 0..fft/2, Im X[0] = Im X[fft/2] = 0) and :149-163 (inverse: Im of DC/Nyquist discarded, result = fft * x).
 * HISSTools is absent; this is a plain iterative radix-2 complex FFT with directly evaluated twiddles.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int64_t n; double* cs; double* sn; int64_t* rev; } fo_fft_plan;

static fo_fft_plan* fo_fft_plan_new(int64_t n)
{
  fo_fft_plan* p = (fo_fft_plan*) malloc(sizeof(fo_fft_plan));
  p->n = n;
  p->cs = (double*) malloc(sizeof(double) * (size_t) (n / 2 + 1));
  p->sn = (double*) malloc(sizeof(double) * (size_t) (n / 2 + 1));
  p->rev = (int64_t*) malloc(sizeof(int64_t) * (size_t) n);
  const double pi = 3.14159265358979323846;
  for (int64_t k = 0; k < n / 2 + 1; k++) {
    p->cs[k] = cos(2 * pi * (double) k / (double) n);
    p->sn[k] = sin(2 * pi * (double) k / (double) n);
  }
  int bits = 0;
  while (((int64_t) 1 << bits) < n) bits++;
  for (int64_t i = 0; i < n; i++) {
    int64_t r = 0;
    for (int b = 0; b < bits; b++) if (i & ((int64_t) 1 << b)) r |= (int64_t) 1 << (bits - 1 - b);
    p->rev[i] = r;
  }
  return p;
}

static void fo_fft_plan_free(fo_fft_plan* p)
{
  if (!p) return;
  free(p->cs); free(p->sn); free(p->rev); free(p);
}

/* in-place complex FFT, sign = -1 forward, +1 inverse (unnormalised) */
static void fo_cfft(const fo_fft_plan* p, double* re, double* im, int sign)
{
  int64_t n = p->n;
  for (int64_t i = 0; i < n; i++) {
    int64_t j = p->rev[i];
    if (j > i) {
      double t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
  for (int64_t len = 2; len <= n; len <<= 1) {
    int64_t half = len >> 1, step = n / len;
    for (int64_t s = 0; s < n; s += len) {
      for (int64_t k = 0; k < half; k++) {
        double wr = p->cs[k * step], wi = sign * p->sn[k * step];
        double xr = re[s + k + half], xi = im[s + k + half];
        double tr = xr * wr - xi * wi, ti = xr * wi + xi * wr;
        re[s + k + half] = re[s + k] - tr; im[s + k + half] = im[s + k] - ti;
        re[s + k] += tr; im[s + k] += ti;
      }
    }
  }
}

/* FFT.hpp:92-108. x has n_in <= fft samples (zero-padded). out = interleaved (re,im) [fft/2+1]. */
static void fo_rfft_plan(const fo_fft_plan* p, const double* x, int64_t n_in, double* out, double* re, double* im)
{
  int64_t n = p->n;
  for (int64_t i = 0; i < n; i++) { re[i] = i < n_in ? x[i] : 0.0; im[i] = 0.0; }
  fo_cfft(p, re, im, -1);
  for (int64_t k = 0; k <= n / 2; k++) { out[2 * k] = re[k]; out[2 * k + 1] = im[k]; }
  out[1] = 0.0;         /* FFT.hpp:101 */
  out[2 * (n / 2) + 1] = 0.0; /* FFT.hpp:100 */
}

/* FFT.hpp:149-163: y[n] = sum over the Hermitian-extended spectrum, = fft * x; Im(DC), Im(Nyquist) ignored */
static void fo_irfft_plan(const fo_fft_plan* p, const double* in, double* y, double* re, double* im)
{
  int64_t n = p->n;
  re[0] = in[0]; im[0] = 0.0;
  re[n / 2] = in[2 * (n / 2)]; im[n / 2] = 0.0;
  for (int64_t k = 1; k < n / 2; k++) {
    re[k] = in[2 * k]; im[k] = in[2 * k + 1];
    re[n - k] = in[2 * k]; im[n - k] = -in[2 * k + 1];
  }
  fo_cfft(p, re, im, +1);
  for (int64_t i = 0; i < n; i++) y[i] = re[i];
}

FO_EXPORT void fo_rfft(const double* x, int64_t n_in, int64_t fft, double* out)
{
  fo_fft_plan* p = fo_fft_plan_new(fft);
  double* re = (double*) malloc(sizeof(double) * (size_t) fft * 2);
  fo_rfft_plan(p, x, n_in, out, re, re + fft);
  free(re); fo_fft_plan_free(p);
}

FO_EXPORT void fo_irfft_unnorm(const double* in, int64_t fft, double* y)
{
  fo_fft_plan* p = fo_fft_plan_new(fft);
  double* re = (double*) malloc(sizeof(double) * (size_t) fft * 2);
  fo_irfft_plan(p, in, y, re, re + fft);
  free(re); fo_fft_plan_free(p);
}

/* ------------------------------------------------------------------------------------------------
 * STFT::process  algorithms/public/STFT.hpp:90-108  (+ ctor :36-47: Hann table of `win`)
 * spec: interleaved complex [F][B], F = fo_stft_num_frames(n, win, hop), B = fft/2+1
 * ---------------------------------------------------------------------------------------------- */
FO_EXPORT void fo_stft(const double* audio, int64_t n, int64_t win, int64_t fft, int64_t hop, double* spec)
{
  int64_t half = win / 2;                           /* :92 */
  int64_t plen = n + win + hop;                     /* :94 */
  int64_t nframes = (plen - win) / hop;             /* :98-99 */
  int64_t bins = fft / 2 + 1;
  double* padded = (double*) calloc((size_t) plen, sizeof(double)); /* :95 */
  memcpy(padded + half, audio, sizeof(double) * (size_t) n);        /* :96-97 */
  double* w = (double*) malloc(sizeof(double) * (size_t) win);
  fo_hann(win, w);
  fo_fft_plan* p = fo_fft_plan_new(fft);
  double* scratch = (double*) malloc(sizeof(double) * (size_t) (2 * fft + win));
  double* frame = scratch + 2 * fft;
  for (int64_t i = 0; i < nframes; i++) {           /* :102-106 */
    for (int64_t j = 0; j < win; j++) frame[j] = padded[i * hop + j] * w[j];
    fo_rfft_plan(p, frame, win, spec + 2 * i * bins, scratch, scratch + fft);
  }
  free(scratch); fo_fft_plan_free(p); free(w); free(padded);
}

/* STFT::processFrame STFT.hpp:110-118: one already-cut frame of `win` samples -> B complex bins */
FO_EXPORT void fo_stft_frame(const double* frame_in, int64_t win, int64_t fft, double* out)
{
  double* w = (double*) malloc(sizeof(double) * (size_t) win * 2);
  fo_hann(win, w);
  for (int64_t j = 0; j < win; j++) w[win + j] = frame_in[j] * w[j];
  fo_rfft(w + win, win, fft, out);
  free(w);
}

/* STFT::magnitude STFT.hpp:61-73 */
FO_EXPORT void fo_magnitude(const double* spec, int64_t count, double* mag)
{
  for (int64_t i = 0; i < count; i++) mag[i] = hypot(spec[2 * i], spec[2 * i + 1]);
}

/* ISTFT::process STFT.hpp:178-199 */
FO_EXPORT void fo_istft(const double* spec, int64_t nframes, int64_t win, int64_t fft, int64_t hop, double* audio,
                        int64_t n_out)
{
  int64_t half = win / 2;                                         /* :181 */
  int64_t bins = fft / 2 + 1;
  int64_t osz = win + (nframes - 1) * hop + win + hop;            /* :183-184 */
  if (osz < half + n_out) osz = half + n_out; /* guard only; the reference would read out of bounds here */
  double scale = 1.0 / (double) fft;                              /* :157 */
  double* out = (double*) calloc((size_t) osz, sizeof(double));
  double* nrm = (double*) calloc((size_t) osz, sizeof(double));
  double* w = (double*) malloc(sizeof(double) * (size_t) win);
  fo_hann(win, w);
  fo_fft_plan* p = fo_fft_plan_new(fft);
  double* scratch = (double*) malloc(sizeof(double) * (size_t) fft * 3);
  double* y = scratch + 2 * fft;
  for (int64_t i = 0; i < nframes; i++) {                         /* :189-195 */
    fo_irfft_plan(p, spec + 2 * i * bins, y, scratch, scratch + fft);
    for (int64_t j = 0; j < win; j++) {
      out[i * hop + j] += y[j] * scale * w[j];
      nrm[i * hop + j] += w[j] * w[j];
    }
  }
  for (int64_t i = 0; i < n_out; i++) {                           /* :196-198 */
    double d = nrm[half + i] > FO_EPS ? nrm[half + i] : FO_EPS;
    audio[i] = out[half + i] / d;
  }
  free(scratch); fo_fft_plan_free(p); free(w); free(nrm); free(out);
}

/* ------------------------------------------------------------------------------------------------
 * BufSTFT  clients/nrt/BufSTFTClient.hpp:82-190 (processFwd) and :192-279 (processInverse), one mono channel.
 * Size rules: FFTParams::padding (clients/common/ParameterTypes.hpp:315-323); paddedLength and numHops (:121-131);
 * inverse output length (:241-242).  float32 buffers in and out (BufferAdaptor), fp64 arithmetic in between.
 * ---------------------------------------------------------------------------------------------- */
FO_EXPORT int fo_bufstft_sizes(int64_t win, int64_t hop, int64_t mode, int invert, int64_t count, int64_t* padding,
                               int64_t* out)
{
  if (win <= 0 || mode < 0 || mode > 2 || count < 0) return -1;
  if (hop <= 0) hop = win >> 1;
  if (hop <= 0) return -1;
  int64_t pad = mode == 0 ? 0 : (mode == 1 ? (win >> 1) : win - hop);   /* ParameterTypes.hpp:319-321 */
  if (padding) *padding = pad;
  if (!invert) {
    int64_t padded = count + 2 * pad;                                  /* :121-124 */
    if (mode == 2) padded = (padded + hop - 1) / hop * hop;            /* :126-128 ceil to a hop multiple */
    if (padded < win) return -1;                                       /* the reference would read out of bounds */
    if (out) *out = 1 + (padded - win) / hop;                          /* :130-131 */
  } else {
    if (count < 1) return -1;
    if (out) *out = (count - 1) * hop + win - pad;                     /* :241-242 */
  }
  return 0;
}

/* mag / phase: [numHops][bins] float32 (either may be NULL) */
FO_EXPORT int fo_bufstft_fwd(const float* audio, int64_t n, int64_t win, int64_t fft, int64_t hop, int64_t mode, float* mag,
                             float* phase)
{
  int64_t pad, hops;
  if (fo_bufstft_sizes(win, hop, mode, 0, n, &pad, &hops)) return -1;
  int64_t bins = fft / 2 + 1;
  int64_t padded_len = (hops - 1) * hop + win;                         /* the part of paddedInput the frames touch */
  double* padded = (double*) calloc((size_t) padded_len, sizeof(double));   /* :150 */
  for (int64_t i = 0; i < n && pad + i < padded_len; i++) padded[pad + i] = (double) audio[i]; /* :152-153 */
  double* w = (double*) malloc(sizeof(double) * (size_t) win);
  fo_hann(win, w);
  fo_fft_plan* p = fo_fft_plan_new(fft);
  double* scratch = (double*) malloc(sizeof(double) * (size_t) (2 * fft + win + 2 * bins));
  double* frame = scratch + 2 * fft;
  double* spec = frame + win;
  for (int64_t i = 0; i < hops; i++) {                                 /* :162-164 STFT::processFrame */
    for (int64_t j = 0; j < win; j++) frame[j] = padded[i * hop + j] * w[j];
    fo_rfft_plan(p, frame, win, spec, scratch, scratch + fft);
    for (int64_t b = 0; b < bins; b++) {
      if (mag) mag[i * bins + b] = (float) hypot(spec[2 * b], spec[2 * b + 1]);    /* :166-170 */
      if (phase) phase[i * bins + b] = (float) atan2(spec[2 * b + 1], spec[2 * b]); /* :172-176, STFT.hpp:75-87 */
    }
  }
  free(scratch); fo_fft_plan_free(p); free(w); free(padded);
  return 0;
}

/* mag, phase [frames][bins] float32 -> out [(frames-1)*hop + win - padding] float32 */
FO_EXPORT int fo_bufstft_inv(const float* mag, const float* phase, int64_t frames, int64_t win, int64_t fft, int64_t hop,
                             int64_t mode, float* out)
{
  int64_t pad, n_out;
  if (fo_bufstft_sizes(win, hop, mode, 1, frames, &pad, &n_out)) return -1;
  int64_t bins = fft / 2 + 1;
  int64_t plen = (frames - 1) * hop + win;                             /* :241 */
  double scale = 1.0 / (double) fft;                                   /* ISTFT ctor, STFT.hpp:157 */
  double* acc = (double*) calloc((size_t) plen, sizeof(double));
  double* nrm = (double*) calloc((size_t) plen, sizeof(double));
  double* w = (double*) malloc(sizeof(double) * (size_t) win);
  fo_hann(win, w);
  fo_fft_plan* p = fo_fft_plan_new(fft);
  double* scratch = (double*) malloc(sizeof(double) * (size_t) (3 * fft + 2 * bins));
  double* y = scratch + 2 * fft;
  double* spec = y + fft;
  for (int64_t i = 0; i < frames; i++) {                               /* :262-268 */
    for (int64_t b = 0; b < bins; b++) {                               /* std::polar, :236-239 */
      double m = (double) mag[i * bins + b], ph = (double) phase[i * bins + b];
      spec[2 * b] = m * cos(ph); spec[2 * b + 1] = m * sin(ph);
    }
    fo_irfft_plan(p, spec, y, scratch, scratch + fft);                  /* ISTFT::processFrame STFT.hpp:201-214 */
    for (int64_t j = 0; j < win; j++) {
      acc[i * hop + j] += y[j] * scale * w[j];
      nrm[i * hop + j] += w[j] * w[j];
    }
  }
  for (int64_t t = 0; t < n_out; t++) {                                /* :270-277 */
    double d = nrm[pad + t] > FO_EPS ? nrm[pad + t] : FO_EPS;
    out[t] = (float) (acc[pad + t] / d);
  }
  free(scratch); fo_fft_plan_free(p); free(w); free(nrm); free(acc);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * EigenRandom.  algorithms/util/EigenRandom.hpp:73-110 on libstdc++:
 *   std::mt19937_64 g{seed}; std::uniform_real_distribution<double>{0,1}(g) == generate_canonical<double,53>
 *   == double(raw u64) / 2^64 (one draw; a result of exactly 1.0 is replaced by nextafter(1,0)).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t mt[312]; int idx; } fo_mt64;

static void fo_mt64_seed(fo_mt64* s, uint64_t seed)
{
  s->mt[0] = seed;
  for (int i = 1; i < 312; i++) s->mt[i] = 6364136223846793005ULL * (s->mt[i - 1] ^ (s->mt[i - 1] >> 62)) + (uint64_t) i;
  s->idx = 312;
}

static uint64_t fo_mt64_next(fo_mt64* s)
{
  if (s->idx >= 312) {
    const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL, MA = 0xB5026F5AA96619E9ULL;
    for (int i = 0; i < 312; i++) {
      uint64_t x = (s->mt[i] & UM) | (s->mt[(i + 1) % 312] & LM);
      s->mt[i] = s->mt[(i + 156) % 312] ^ (x >> 1) ^ ((x & 1ULL) ? MA : 0ULL);
    }
    s->idx = 0;
  }
  uint64_t x = s->mt[s->idx++];
  x ^= (x >> 29) & 0x5555555555555555ULL;
  x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
  x ^= (x << 37) & 0xFFF7EEE000000000ULL;
  x ^= (x >> 43);
  return x;
}

static double fo_mt64_uniform(fo_mt64* s)
{
  double r = (double) fo_mt64_next(s) / 18446744073709551616.0;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}

/* raw stream: out[i] = i-th draw of uniform[0,1) from mt19937_64(seed) */
FO_EXPORT void fo_random_uniform(int64_t seed, int64_t count, double* out)
{
  fo_mt64 g; fo_mt64_seed(&g, (uint64_t) seed);
  for (int64_t i = 0; i < count; i++) out[i] = fo_mt64_uniform(&g);
}

/* NMF.hpp:104-105 / :116-117: Eigen NullaryExpr fills a column-major MatrixXd in storage order, and W and H
 * each restart from state(seed).  Column-major (B x K) W  == our W[K][B] memory; column-major (K x F) H == H[F][K]. */
FO_EXPORT void fo_nmf_random_init(int64_t seed, int64_t bins, int64_t rank, int64_t frames, double* W, double* H)
{
  if (W) fo_random_uniform(seed, bins * rank, W);
  if (H) fo_random_uniform(seed, rank * frames, H);
}

/* ------------------------------------------------------------------------------------------------
 * NMF::multiplicativeUpdates.  algorithms/public/NMF.hpp:144-183
 * Memory layouts (== Eigen column-major B x K, K x F, B x F):  W[K][B], H[F][K], V[F][B].
 * faithful != 0 executes the reference's redundant work too (ones-matrix GEMMs :160,:169 and the dead R=WH :173)
 * so that CPU timings represent what the reference really runs.
 * progress: optional callback(iteration 1..n) -> nonzero to continue (NMF.hpp:175-176).
 * returns 1 if cancelled (V is then left untouched == X, because :182 is skipped), else 0.
 * ---------------------------------------------------------------------------------------------- */
typedef int (*fo_progress_cb)(void* user, int64_t iter);

/* The three small GEMMs below are blocked over two frames (and four components) so that every loaded W / num
 * element feeds two (eight) multiply-adds: the plain one-frame loops ran at 0.8x the single-thread rate of
 * OpenBLAS on the same shapes, which would have flattered the GPU/CPU ratio. */
static void fo_wh(const double* W, const double* H, int64_t B, int64_t F, int64_t K, double* P, int clamp)
{ /* P[f][b] = sum_k H[f][k] W[k][b] */
  int64_t f = 0;
  for (; f + 2 <= F; f += 2) {
    double *p = P + f * B, *q = p + B;
    const double *ha = H + f * K, *hb = ha + K;
    for (int64_t b = 0; b < B; b++) { p[b] = 0.0; q[b] = 0.0; }
    int64_t k = 0;
    for (; k + 4 <= K; k += 4) {
      const double a0 = ha[k], a1 = ha[k + 1], a2 = ha[k + 2], a3 = ha[k + 3];
      const double c0 = hb[k], c1 = hb[k + 1], c2 = hb[k + 2], c3 = hb[k + 3];
      const double *w0 = W + k * B, *w1 = w0 + B, *w2 = w1 + B, *w3 = w2 + B;
#pragma omp simd
      for (int64_t b = 0; b < B; b++) {
        const double x0 = w0[b], x1 = w1[b], x2 = w2[b], x3 = w3[b];
        p[b] += a0 * x0 + a1 * x1 + a2 * x2 + a3 * x3;
        q[b] += c0 * x0 + c1 * x1 + c2 * x2 + c3 * x3;
      }
    }
    for (; k < K; k++) {
      const double a = ha[k], c = hb[k];
      const double* w = W + k * B;
#pragma omp simd
      for (int64_t b = 0; b < B; b++) { p[b] += a * w[b]; q[b] += c * w[b]; }
    }
  }
  for (; f < F; f++) {
    double* p = P + f * B;
    for (int64_t b = 0; b < B; b++) p[b] = 0.0;
    for (int64_t k = 0; k < K; k++) {
      const double h = H[f * K + k];
      const double* w = W + k * B;
#pragma omp simd
      for (int64_t b = 0; b < B; b++) p[b] += h * w[b];
    }
  }
  if (clamp) {
#pragma omp simd
    for (int64_t i = 0; i < F * B; i++) P[i] = P[i] > FO_EPS ? P[i] : FO_EPS;
  }
}

/* num[k][b] = sum_f R[f][b] * H[f][k]   (R * H^T) */
static void fo_r_ht(const double* R, const double* H, int64_t B, int64_t F, int64_t K, double* num)
{
  memset(num, 0, sizeof(double) * (size_t) (K * B));
  int64_t f = 0;
  for (; f + 2 <= F; f += 2) {
    const double *r = R + f * B, *s = r + B;
    const double *ha = H + f * K, *hb = ha + K;
    int64_t k = 0;
    for (; k + 4 <= K; k += 4) {
      const double a0 = ha[k], a1 = ha[k + 1], a2 = ha[k + 2], a3 = ha[k + 3];
      const double c0 = hb[k], c1 = hb[k + 1], c2 = hb[k + 2], c3 = hb[k + 3];
      double *o0 = num + k * B, *o1 = o0 + B, *o2 = o1 + B, *o3 = o2 + B;
#pragma omp simd
      for (int64_t b = 0; b < B; b++) {
        const double rb = r[b], sb = s[b];
        o0[b] += a0 * rb + c0 * sb; o1[b] += a1 * rb + c1 * sb; o2[b] += a2 * rb + c2 * sb; o3[b] += a3 * rb + c3 * sb;
      }
    }
    for (; k < K; k++) {
      const double a = ha[k], c = hb[k];
      double* o = num + k * B;
#pragma omp simd
      for (int64_t b = 0; b < B; b++) o[b] += a * r[b] + c * s[b];
    }
  }
  for (; f < F; f++) {
    const double* r = R + f * B;
    for (int64_t k = 0; k < K; k++) {
      const double h = H[f * K + k];
      double* o = num + k * B;
#pragma omp simd
      for (int64_t b = 0; b < B; b++) o[b] += h * r[b];
    }
  }
}

/* num[f][k] = sum_b W[k][b] * R[f][b]   (W^T * R) */
static void fo_wt_r(const double* W, const double* R, int64_t B, int64_t F, int64_t K, double* num)
{
  int64_t f = 0;
  for (; f + 2 <= F; f += 2) {
    const double *r = R + f * B, *s = r + B;
    int64_t k = 0;
    for (; k + 4 <= K; k += 4) {
      const double *w0 = W + k * B, *w1 = w0 + B, *w2 = w1 + B, *w3 = w2 + B;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma omp simd reduction(+ : s0, s1, s2, s3, t0, t1, t2, t3)
      for (int64_t b = 0; b < B; b++) {
        const double rb = r[b], sb = s[b];
        s0 += w0[b] * rb; s1 += w1[b] * rb; s2 += w2[b] * rb; s3 += w3[b] * rb;
        t0 += w0[b] * sb; t1 += w1[b] * sb; t2 += w2[b] * sb; t3 += w3[b] * sb;
      }
      num[f * K + k] = s0; num[f * K + k + 1] = s1; num[f * K + k + 2] = s2; num[f * K + k + 3] = s3;
      num[(f + 1) * K + k] = t0; num[(f + 1) * K + k + 1] = t1; num[(f + 1) * K + k + 2] = t2; num[(f + 1) * K + k + 3] = t3;
    }
    for (; k < K; k++) {
      const double* w = W + k * B;
      double s0 = 0.0, t0 = 0.0;
#pragma omp simd reduction(+ : s0, t0)
      for (int64_t b = 0; b < B; b++) { s0 += w[b] * r[b]; t0 += w[b] * s[b]; }
      num[f * K + k] = s0; num[(f + 1) * K + k] = t0;
    }
  }
  for (; f < F; f++) {
    const double* r = R + f * B;
    for (int64_t k = 0; k < K; k++) {
      const double* w = W + k * B;
      double s0 = 0.0;
#pragma omp simd reduction(+ : s0)
      for (int64_t b = 0; b < B; b++) s0 += w[b] * r[b];
      num[f * K + k] = s0;
    }
  }
}

static void fo_normalize_w_rows(double* W, int64_t B, int64_t K)
{ /* W.colwise().normalize() on B x K == each memory row W[k][:] / ||.||2   (NMF.hpp:152,162) */
  for (int64_t k = 0; k < K; k++) {
    double s = 0.0;
    for (int64_t b = 0; b < B; b++) s += W[k * B + b] * W[k * B + b];
    double nrm = sqrt(s);
    /* Eigen's normalize() divides only when the squared norm is > 0 */
    if (s > 0.0) for (int64_t b = 0; b < B; b++) W[k * B + b] /= nrm;
  }
}

static int fo_mu(double* V, double* W, double* H, int64_t B, int64_t F, int64_t K, int64_t n_iter, int update_w,
                 int update_h, int faithful, fo_progress_cb cb, void* user)
{
  double* P = (double*) malloc(sizeof(double) * (size_t) (F * B));
  double* R = (double*) malloc(sizeof(double) * (size_t) (F * B));
  double* ones = NULL;
  double* wnum = (double*) malloc(sizeof(double) * (size_t) (K * B));
  double* wden = (double*) malloc(sizeof(double) * (size_t) (K * B));
  double* hnum = (double*) malloc(sizeof(double) * (size_t) (F * K));
  double* hden = (double*) malloc(sizeof(double) * (size_t) (F * K));
  if (faithful) {
    ones = (double*) malloc(sizeof(double) * (size_t) (F * B)); /* :149 */
    for (int64_t i = 0; i < F * B; i++) ones[i] = 1.0;
  }
  for (int64_t i = 0; i < F * K; i++) H[i] = H[i] > FO_EPS ? H[i] : FO_EPS; /* :150 */
  for (int64_t i = 0; i < K * B; i++) W[i] = W[i] > FO_EPS ? W[i] : FO_EPS; /* :151 */
  fo_normalize_w_rows(W, B, K);                                             /* :152 */
  for (int64_t k = 0; k < K; k++) { /* :153 H.rowwise().normalize(): row k of (K x F) == column k of H[F][K] */
    double s = 0.0;
    for (int64_t f = 0; f < F; f++) s += H[f * K + k] * H[f * K + k];
    double nrm = sqrt(s);
    if (s > 0.0) for (int64_t f = 0; f < F; f++) H[f * K + k] /= nrm;
  }
  int cancelled = 0;
  for (int64_t it = 0; it < n_iter; ++it) {
    if (update_w) {
      fo_wh(W, H, B, F, K, P, 1);                                           /* :158 */
      for (int64_t i = 0; i < F * B; i++) R[i] = V[i] / P[i];
      fo_r_ht(R, H, B, F, K, wnum);                                         /* :159 */
      if (faithful) {
        fo_r_ht(ones, H, B, F, K, wden);                                    /* :160 */
      } else {
        for (int64_t k = 0; k < K; k++) {
          double s = 0.0;
          for (int64_t f = 0; f < F; f++) s += H[f * K + k];
          for (int64_t b = 0; b < B; b++) wden[k * B + b] = s;
        }
      }
      double wmax = -INFINITY;
      for (int64_t i = 0; i < K * B; i++) {                                 /* :161 */
        double d = wden[i] > FO_EPS ? wden[i] : FO_EPS;
        W[i] = W[i] * wnum[i] / d;
        if (W[i] > wmax) wmax = W[i];
      }
      if (wmax > FO_EPS) fo_normalize_w_rows(W, B, K);                      /* :162 */
    }
    fo_wh(W, H, B, F, K, P, 1);                                             /* :165 */
    if (update_h) {
      for (int64_t i = 0; i < F * B; i++) R[i] = V[i] / P[i];
      fo_wt_r(W, R, B, F, K, hnum);                                         /* :168 */
      if (faithful) {
        fo_wt_r(W, ones, B, F, K, hden);                                    /* :169 */
      } else {
        for (int64_t k = 0; k < K; k++) {
          double s = 0.0;
          for (int64_t b = 0; b < B; b++) s += W[k * B + b];
          for (int64_t f = 0; f < F; f++) hden[f * K + k] = s;
        }
      }
      for (int64_t i = 0; i < F * K; i++) {                                 /* :170 */
        double d = hden[i] > FO_EPS ? hden[i] : FO_EPS;
        H[i] = H[i] * hnum[i] / d;
      }
    }
    if (faithful) fo_wh(W, H, B, F, K, P, 1);                               /* :173-174 dead R */
    if (cb && !cb(user, it + 1)) { cancelled = 1; break; }                  /* :175-176 */
  }
  if (!cancelled) fo_wh(W, H, B, F, K, V, 0);                               /* :182 */
  free(P); free(R); free(ones); free(wnum); free(wden); free(hnum); free(hden);
  return cancelled;
}

/* NMF::process  NMF.hpp:91-134.
 * X[F][B] in; W0[K][B] / H0[F][K] optional seeds (NULL => random from `seed`, each restarting the stream);
 * outputs W1[K][B], H1[F][K], V1[F][B] (any may be NULL).  Returns 1 when a progress callback cancelled. */
FO_EXPORT int fo_nmf_process(const double* X, int64_t F, int64_t B, int64_t K, int64_t n_iter, int update_w,
                             int update_h, int64_t seed, const double* W0, const double* H0, double* W1, double* H1,
                             double* V1, int faithful, fo_progress_cb cb, void* user)
{
  double* V = (double*) malloc(sizeof(double) * (size_t) (F * B));
  double* W = (double*) malloc(sizeof(double) * (size_t) (K * B));
  double* H = (double*) malloc(sizeof(double) * (size_t) (F * K));
  memcpy(V, X, sizeof(double) * (size_t) (F * B));                          /* :125 */
  if (W0) memcpy(W, W0, sizeof(double) * (size_t) (K * B));                 /* :111 */
  else fo_nmf_random_init(seed, B, K, F, W, NULL);                          /* :104-105 */
  if (H0) memcpy(H, H0, sizeof(double) * (size_t) (F * K));                 /* :123 */
  else fo_nmf_random_init(seed, B, K, F, NULL, H);                          /* :116-117 */
  int cancelled = fo_mu(V, W, H, B, F, K, n_iter, update_w, update_h, faithful, cb, user); /* :126 */
  if (V1) memcpy(V1, V, sizeof(double) * (size_t) (F * B));                 /* :131 */
  if (W1) memcpy(W1, W, sizeof(double) * (size_t) (K * B));                 /* :132 */
  if (H1) memcpy(H1, H, sizeof(double) * (size_t) (F * K));                 /* :133 */
  free(V); free(W); free(H);
  return cancelled;
}

/* NMF::processFrame  NMF.hpp:45-89.  W0[K][B] IS MUTATED (eps clamp :58, row normalise :63-64), as in the reference.
 * out[K] and v[B] may be NULL (:85,:88). */
FO_EXPORT void fo_nmf_process_frame(const double* x, double* W0, int64_t B, int64_t K, int64_t n_iter, int64_t seed,
                                    double* out, double* v)
{
  double* h = (double*) malloc(sizeof(double) * (size_t) (K * 3));
  double* hnum = h + K; double* hden = h + 2 * K;
  double* v0 = (double*) malloc(sizeof(double) * (size_t) (B * 3));
  double* v1 = v0 + B; double* r = v0 + 2 * B;
  fo_random_uniform(seed, K, h);                                            /* :55 */
  for (int64_t i = 0; i < K * B; i++) W0[i] = W0[i] > FO_EPS ? W0[i] : FO_EPS; /* :58 */
  for (int64_t k = 0; k < K; k++) h[k] = h[k] > FO_EPS ? h[k] : FO_EPS;     /* :59 */
  for (int64_t b = 0; b < B; b++) v0[b] = x[b] > FO_EPS ? x[b] : FO_EPS;    /* :60 */
  for (int64_t k = 0; k < K; k++) {                                         /* :63-64 */
    double s = 0.0;
    for (int64_t b = 0; b < B; b++) s += W0[k * B + b] * W0[k * B + b];
    double nrm = sqrt(s);
    for (int64_t b = 0; b < B; b++) W0[k * B + b] /= nrm;
  }
  while (n_iter-- > 0) {                                                    /* :72-83 */
    for (int64_t b = 0; b < B; b++) v1[b] = 0.0;
    for (int64_t k = 0; k < K; k++)
      for (int64_t b = 0; b < B; b++) v1[b] += W0[k * B + b] * h[k];         /* :74 */
    for (int64_t b = 0; b < B; b++) {
      double p = v1[b] > FO_EPS ? v1[b] : FO_EPS;                           /* :75 */
      r[b] = v0[b] / p;                                                     /* :76 */
    }
    for (int64_t k = 0; k < K; k++) {
      double sn = 0.0, sd = 0.0;
      for (int64_t b = 0; b < B; b++) { sn += W0[k * B + b] * r[b]; sd += W0[k * B + b]; } /* :77-78 */
      hnum[k] = sn; hden[k] = sd;
    }
    for (int64_t k = 0; k < K; k++) h[k] = h[k] * hnum[k] / (hden[k] > FO_EPS ? hden[k] : FO_EPS); /* :79 */
  }
  if (out) for (int64_t k = 0; k < K; k++) out[k] = h[k];                   /* :85-86 */
  if (v) {                                                                  /* :88 */
    for (int64_t b = 0; b < B; b++) v[b] = 0.0;
    for (int64_t k = 0; k < K; k++)
      for (int64_t b = 0; b < B; b++) v[b] += W0[k * B + b] * h[k];
  }
  free(h); free(v0);
}

/* NMF::estimate NMF.hpp:33-42:  E[f][b] = H[f][idx] * W[idx][b] */
FO_EXPORT void fo_nmf_estimate(const double* W, const double* H, int64_t F, int64_t B, int64_t K, int64_t idx, double* E)
{
  for (int64_t f = 0; f < F; f++)
    for (int64_t b = 0; b < B; b++) E[f * B + b] = H[f * K + idx] * W[idx * B + b];
}

/* RatioMask::init + process, exponent 1.  algorithms/public/RatioMask.hpp:33-57
 * out = mixture * min(1, target * (1 / max(denominator, eps)))   (complex interleaved) */
FO_EXPORT void fo_ratio_mask(const double* mixture, const double* target, const double* denominator, int64_t count,
                             double* out)
{
  for (int64_t i = 0; i < count; i++) {
    double mult = 1.0 / (denominator[i] > FO_EPS ? denominator[i] : FO_EPS); /* :39-40 */
    double m = target[i] * mult;                                            /* :53-55 (pow 1) */
    if (m > 1.0) m = 1.0;                                                   /* :56 */
    out[2 * i] = mixture[2 * i] * m;
    out[2 * i + 1] = mixture[2 * i + 1] * m;
  }
}

/* ------------------------------------------------------------------------------------------------
 * BufNMF per channel.  clients/nrt/NMFClient.hpp:233-335 with float32 host buffers (BufferAdaptor).
 * audio: float[n] one channel.  bases_out float[K][B], acts_out float[F][K] (scaled by 1/max(H), :289-298),
 * resynth_out float[K][n] (NULL = no resynthesis).  Optional seeds bases_in[K][B] / acts_in[F][K] (float) with
 * modes 0 none / 1 seed / 2 fixed (:116-118, :139-141).  When a mode is 2 the corresponding *_out is not written
 * (:277, :286).  Raw (unscaled) double W/H/V can be captured through dbg_W/dbg_H/dbg_V (NULL ok).
 * ---------------------------------------------------------------------------------------------- */
FO_EXPORT int fo_bufnmf_channel(const float* audio, int64_t n, int64_t win, int64_t fft, int64_t hop, int64_t K,
                                int64_t iters, int64_t seed, int bases_mode, const float* bases_in, int acts_mode,
                                const float* acts_in, float* bases_out, float* acts_out, float* resynth_out,
                                int faithful, double* dbg_W, double* dbg_H, double* dbg_V)
{
  int64_t B = fft / 2 + 1;
  int64_t F = fo_stft_num_frames(n, win, hop);
  int fix_w = bases_mode == 2, fix_h = acts_mode == 2;
  int needs_analysis = !(fix_w && fix_h);                                   /* :141 */
  double* tmp = (double*) malloc(sizeof(double) * (size_t) n);
  for (int64_t i = 0; i < n; i++) tmp[i] = (double) audio[i];               /* :240 */
  double* spec = (double*) malloc(sizeof(double) * (size_t) (2 * F * B));
  double* mag = (double*) malloc(sizeof(double) * (size_t) (F * B));
  fo_stft(tmp, n, win, fft, hop, spec);                                     /* :241 */
  fo_magnitude(spec, F * B, mag);                                           /* :242 */
  double* W0 = NULL; double* H0 = NULL;
  if (bases_mode > 0) {                                                     /* :248-252 */
    W0 = (double*) malloc(sizeof(double) * (size_t) (K * B));
    for (int64_t i = 0; i < K * B; i++) W0[i] = (double) bases_in[i];
  }
  if (acts_mode > 0) {                                                      /* :253-257 */
    H0 = (double*) malloc(sizeof(double) * (size_t) (F * K));
    for (int64_t i = 0; i < F * K; i++) H0[i] = (double) acts_in[i];
  }
  double* W = (double*) malloc(sizeof(double) * (size_t) (K * B));
  double* H = (double*) malloc(sizeof(double) * (size_t) (F * K));
  double* Vh = (double*) malloc(sizeof(double) * (size_t) (F * B));
  fo_nmf_process(mag, F, B, K, iters * needs_analysis, !fix_w, !fix_h, seed, W0, H0, W, H, Vh, faithful, NULL,
                 NULL);                                                     /* :268-271 */
  if (dbg_W) memcpy(dbg_W, W, sizeof(double) * (size_t) (K * B));
  if (dbg_H) memcpy(dbg_H, H, sizeof(double) * (size_t) (F * K));
  if (dbg_V) memcpy(dbg_V, Vh, sizeof(double) * (size_t) (F * B));
  if (bases_out && !fix_w)                                                  /* :277-283 */
    for (int64_t i = 0; i < K * B; i++) bases_out[i] = (float) W[i];
  if (acts_out && !fix_h) {                                                 /* :286-300 */
    double maxh = H[0];
    for (int64_t i = 1; i < F * K; i++) if (H[i] > maxh) maxh = H[i];
    float scale = (float) (1.0 / maxh);
    for (int64_t i = 0; i < F * K; i++) { float x = (float) H[i]; x *= scale; acts_out[i] = x; }
  }
  if (resynth_out) {                                                        /* :302-333 */
    double* est = (double*) malloc(sizeof(double) * (size_t) (F * B));
    double* msp = (double*) malloc(sizeof(double) * (size_t) (2 * F * B));
    double* aud = (double*) malloc(sizeof(double) * (size_t) n);
    for (int64_t j = 0; j < K; j++) {
      fo_nmf_estimate(W, H, F, B, K, j, est);                               /* :319 */
      fo_ratio_mask(spec, est, Vh, F * B, msp);                             /* :306, :324 */
      fo_istft(msp, F, win, fft, hop, aud, n);                              /* :328 */
      for (int64_t i = 0; i < n; i++) resynth_out[j * n + i] = (float) aud[i]; /* :329 */
    }
    free(est); free(msp); free(aud);
  }
  free(tmp); free(spec); free(mag); free(W0); free(H0); free(W); free(H); free(Vh);
  return 0;
}

/* Batch driver used only by the CPU baseline: `batch` independent single-channel buffers of equal length, one
 * OpenMP thread per buffer (models several concurrent BufNMF jobs; the reference itself is single-threaded per job,
 * clients/common/FluidNRTClientWrapper.hpp:1048).  seeds[b] per buffer. */
FO_EXPORT int fo_bufnmf_batch(const float* audio, int64_t batch, int64_t n, int64_t win, int64_t fft, int64_t hop,
                              int64_t K, int64_t iters, const int64_t* seeds, float* bases_out, float* acts_out,
                              float* resynth_out, int faithful, int threads)
{
  int64_t B = fft / 2 + 1;
  int64_t F = fo_stft_num_frames(n, win, hop);
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void) threads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < batch; b++) {
    fo_bufnmf_channel(audio + b * n, n, win, fft, hop, K, iters, seeds ? seeds[b] : b, 0, NULL, 0, NULL,
                      bases_out ? bases_out + b * K * B : NULL, acts_out ? acts_out + b * F * K : NULL,
                      resynth_out ? resynth_out + b * K * n : NULL, faithful, NULL, NULL, NULL);
  }
  return 0;
}

/* Streaming NMFMatch equivalent: clients/rt/NMFMatchClient.hpp:106-118 applied to `nframes` magnitude frames.
 * Every frame: W = float bases copied fresh (:106-107), processFrame with n_iter iterations (10 in NMFMatch, :115),
 * h0 from `seed` (identical for every frame when seed >= 0).  mags[nframes][B] -> acts[nframes][K]. */
FO_EXPORT void fo_nmfmatch_frames(const double* mags, int64_t nframes, const double* W_in, int64_t B, int64_t K,
                                  int64_t n_iter, int64_t seed, double* acts, int threads)
{
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void) threads;
#endif
#pragma omp parallel
  {
    double* W = (double*) malloc(sizeof(double) * (size_t) (K * B));
#pragma omp for schedule(static)
    for (int64_t f = 0; f < nframes; f++) {
      memcpy(W, W_in, sizeof(double) * (size_t) (K * B));
      fo_nmf_process_frame(mags + f * B, W, B, K, n_iter, seed, acts + f * K, NULL);
    }
    free(W);
  }
}

/* Streaming frame cutter: the frames a BufferedProcess-driven client sees (clients/common/BufferedProcess.hpp:75-93,
 * FluidSource.hpp:68-89): frame f covers samples [f*hop - win, f*hop) of the stream, zeros before the start. */
FO_EXPORT void fo_stream_frames_mag(const double* audio, int64_t n, int64_t win, int64_t fft, int64_t hop,
                                    int64_t nframes, double* mags, double* spec_out)
{
  int64_t B = fft / 2 + 1;
  double* frame = (double*) malloc(sizeof(double) * (size_t) win);
  double* sp = (double*) malloc(sizeof(double) * (size_t) (2 * B));
  for (int64_t f = 0; f < nframes; f++) {
    for (int64_t j = 0; j < win; j++) {
      int64_t t = f * hop - win + j;
      frame[j] = (t >= 0 && t < n) ? audio[t] : 0.0;
    }
    fo_stft_frame(frame, win, fft, sp);
    fo_magnitude(sp, B, mags + f * B);
    if (spec_out) memcpy(spec_out + 2 * f * B, sp, sizeof(double) * (size_t) (2 * B));
  }
  free(frame); free(sp);
}

/* Streaming NMFFilter equivalent: clients/rt/NMFFilterClient.hpp:98-117 driven by STFTBufferedProcess<true>
 * (clients/common/BufferedProcess.hpp:187-241) over a mono stream of n samples that starts from reset state.
 *   frame f (f*hop < n) = stream[f*hop - win, f*hop) (FluidSource.hpp:68-89), windowed rFFT (STFT.hpp:110-118), |.|;
 *   processFrame with the bases copied fresh from the float filter buffer (:94-95; the copy happens once per host
 *   vector, we model host vectors <= hop), estimate v = W^T h (:104-106); mask k = h[k] W[k,:] / max(v, eps), min 1
 *   (:107-113); ISTFT::processFrame = irfft * 1/fft * window (STFT.hpp:154-164); overlap-add at [f*hop, f*hop+win)
 *   (FluidSink.hpp:49-68) next to window*window (BufferedProcess.hpp:219-224); output sample = x != 0 ?
 *   x / (g > 0 ? g : 1) : x (:231-237).  Latency = win samples (NMFFilterClient.hpp:64).
 * out[K][n] (NULL ok), acts[nframes][K] (NULL ok; what NMFMatch would emit with the same iteration count). */
FO_EXPORT void fo_nmffilter_stream(const double* audio, int64_t n, int64_t win, int64_t fft, int64_t hop,
                                   const double* W_in, int64_t K, int64_t n_iter, int64_t seed, double* out,
                                   double* acts)
{
  int64_t B = fft / 2 + 1;
  int64_t nframes = (n + hop - 1) / hop;
  double scale = 1.0 / (double) fft;
  double* w = (double*) malloc(sizeof(double) * (size_t) win);
  fo_hann(win, w);
  double* frame = (double*) malloc(sizeof(double) * (size_t) win);
  double* sp = (double*) malloc(sizeof(double) * (size_t) (2 * B));
  double* msp = (double*) malloc(sizeof(double) * (size_t) (2 * B));
  double* mag = (double*) malloc(sizeof(double) * (size_t) B);
  double* est = (double*) malloc(sizeof(double) * (size_t) B);
  double* src = (double*) malloc(sizeof(double) * (size_t) B);
  double* W = (double*) malloc(sizeof(double) * (size_t) (K * B));
  double* h = (double*) malloc(sizeof(double) * (size_t) K);
  double* nrm = (double*) calloc((size_t) (n + win), sizeof(double));
  double* acc = out ? (double*) calloc((size_t) (K * (n + win)), sizeof(double)) : NULL;
  fo_fft_plan* p = fo_fft_plan_new(fft);
  double* scratch = (double*) malloc(sizeof(double) * (size_t) fft * 3);
  double* y = scratch + 2 * fft;
  for (int64_t f = 0; f < nframes; f++) {
    for (int64_t j = 0; j < win; j++) {
      int64_t t = f * hop - win + j;
      frame[j] = (t >= 0 && t < n) ? audio[t] : 0.0;
    }
    fo_stft_frame(frame, win, fft, sp);
    fo_magnitude(sp, B, mag);
    memcpy(W, W_in, sizeof(double) * (size_t) (K * B));
    fo_nmf_process_frame(mag, W, B, K, n_iter, seed, h, est);
    if (acts) memcpy(acts + f * K, h, sizeof(double) * (size_t) K);
    for (int64_t j = 0; j < win; j++) nrm[f * hop + j] += w[j] * w[j];
    if (!out) continue;
    for (int64_t k = 0; k < K; k++) {
      fo_nmf_estimate(W, h, 1, B, K, k, src);
      fo_ratio_mask(sp, src, est, B, msp);
      fo_irfft_plan(p, msp, y, scratch, scratch + fft);
      double* a = acc + k * (n + win);
      for (int64_t j = 0; j < win; j++) a[f * hop + j] += y[j] * scale * w[j];
    }
  }
  if (out)
    for (int64_t k = 0; k < K; k++)
      for (int64_t t = 0; t < n; t++) {
        double x = acc[k * (n + win) + t], g = nrm[t];
        if (x != 0) x /= (g > 0) ? g : 1;
        out[k * n + t] = x;
      }
  free(scratch); fo_fft_plan_free(p); free(acc); free(nrm); free(h); free(W); free(src); free(est); free(mag);
  free(msp); free(sp); free(frame); free(w);
}

/* ------------------------------------------------------------------------------------------------
 * NMFCross  algorithms/public/NMFCross.hpp:60-185  (BufNMFCross: resynthesise a target out of a source's frames)
 * Layouts as everywhere here: X[F][B] target magnitudes, W0[R][B] source magnitudes (rank R = source frames),
 * H[F][R] (== Eigen column-major R x F).  H starts as U(0,1) from mt19937_64(seed), column-major fill (:72-73), and is
 * NOT clamped or normalised; W is clamped at eps (:156) and never updated.
 * Every iteration i (:161-176): H <- sparseness(H, r, i); H <- polyphony(H, p, i); H <- continuity(H, c); then the KL
 * H-update with plain column sums as denominator.  NOTE (:119, :136): the attenuation factor is written
 * 1 - ((iteration + 1) / mIterations) with INTEGER operands, i.e. 1 for every iteration but the last and 0 for the
 * last: sparseness and polyphony are no-ops until the final iteration, where they zero every entry that is not a local
 * temporal maximum / not among the p strongest of its frame.  Restated as written.
 * ---------------------------------------------------------------------------------------------- */
static void fo_cross_sparseness(const double* H, int64_t F, int64_t R, int64_t size, double factor, double* out)
{ /* :104-127: entry (k, f) keeps its value iff the FIRST maximum of H[k][f-half .. f-half+size) (zero padded) is itself */
  int64_t half = (size - 1) / 2;
  for (int64_t f = 0; f < F; f++)
    for (int64_t k = 0; k < R; k++) {
      int64_t arg = 0;
      double best = -INFINITY;
      for (int64_t t = 0; t < size; t++) {
        int64_t ff = f + t - half;
        double v = (ff >= 0 && ff < F) ? H[ff * R + k] : 0.0;
        if (v > best) { best = v; arg = t; }
      }
      out[f * R + k] = arg != half ? H[f * R + k] * factor : H[f * R + k];
    }
}

typedef struct { double v; int64_t i; } fo_vi;
static int fo_vi_desc(const void* a, const void* b)
{
  double x = ((const fo_vi*) a)->v, y = ((const fo_vi*) b)->v;
  if (x > y) return -1;
  if (x < y) return 1;
  int64_t i = ((const fo_vi*) a)->i, j = ((const fo_vi*) b)->i; /* ties: the reference's std::sort order is unspecified */
  return i < j ? -1 : (i > j ? 1 : 0);
}
static void fo_cross_polyphony(const double* H, int64_t F, int64_t R, const double* energy, int64_t p, double factor,
                               double* out)
{ /* :130-143: per frame, the p components with the largest H * energy keep their value, the rest are attenuated */
  fo_vi* v = (fo_vi*) malloc(sizeof(fo_vi) * (size_t) R);
  for (int64_t f = 0; f < F; f++) {
    for (int64_t k = 0; k < R; k++) { v[k].v = H[f * R + k] * energy[k]; v[k].i = k; out[f * R + k] = H[f * R + k] * factor; }
    qsort(v, (size_t) R, sizeof(fo_vi), fo_vi_desc);
    for (int64_t t = 0; t < p && t < R; t++) out[f * R + v[t].i] = H[f * R + v[t].i];
  }
  free(v);
}

static void fo_cross_continuity(const double* H, int64_t F, int64_t R, int64_t size, double* out)
{ /* :86-102: sum along the diagonal (component and frame advance together), zero padded */
  int64_t half = (size - 1) / 2;
  for (int64_t f = 0; f < F; f++)
    for (int64_t k = 0; k < R; k++) {
      double s = 0.0;
      for (int64_t d = 0; d < size; d++) {
        int64_t kk = k + d - half, ff = f + d - half;
        if (kk >= 0 && kk < R && ff >= 0 && ff < F) s += H[ff * R + kk];
      }
      out[f * R + k] = s;
    }
}

FO_EXPORT int fo_nmfcross_process(const double* X, int64_t F, int64_t B, const double* W0, int64_t R, int64_t n_iter,
                                  int64_t r, int64_t p, int64_t c, int64_t seed, double* H1, fo_progress_cb cb, void* user)
{
  double* W = (double*) malloc(sizeof(double) * (size_t) (R * B));
  double* H = (double*) malloc(sizeof(double) * (size_t) (F * R));
  double* T = (double*) malloc(sizeof(double) * (size_t) (F * R));
  double* P = (double*) malloc(sizeof(double) * (size_t) (F * B));
  double* num = (double*) malloc(sizeof(double) * (size_t) (F * R));
  double* energy = (double*) malloc(sizeof(double) * (size_t) R);
  double* hden = (double*) malloc(sizeof(double) * (size_t) R);
  fo_random_uniform(seed, R * F, H);                                          /* :72-73 */
  for (int64_t i = 0; i < R * B; i++) W[i] = W0[i] > FO_EPS ? W0[i] : FO_EPS;  /* :156 */
  for (int64_t k = 0; k < R; k++) {
    double e = 0.0, s = 0.0;
    for (int64_t b = 0; b < B; b++) { e += W[k * B + b] * W[k * B + b]; s += W[k * B + b]; }
    energy[k] = e;                                                            /* :159 */
    hden[k] = s > FO_EPS ? s : FO_EPS;                                        /* :169-170 */
  }
  int cancelled = 0;
  for (int64_t i = 0; i < n_iter; i++) {
    double factor = 1.0 - (double) ((i + 1) / n_iter);                        /* integer division, see above */
    fo_cross_sparseness(H, F, R, r, factor, T);                               /* :163 */
    fo_cross_polyphony(T, F, R, energy, p, factor, H);                        /* :164 */
    fo_cross_continuity(H, F, R, c, T);                                       /* :165 */
    memcpy(H, T, sizeof(double) * (size_t) (F * R));
    fo_wh(W, H, B, F, R, P, 1);                                               /* :167 */
    for (int64_t e = 0; e < F * B; e++) P[e] = X[e] / P[e];
    fo_wt_r(W, P, B, F, R, num);                                              /* :168 */
    for (int64_t f = 0; f < F; f++)
      for (int64_t k = 0; k < R; k++) H[f * R + k] = H[f * R + k] * num[f * R + k] / hden[k]; /* :170 */
    if (cb && !cb(user, i + 1)) { cancelled = 1; break; }                     /* :174-175 */
  }
  memcpy(H1, H, sizeof(double) * (size_t) (F * R));
  free(W); free(H); free(T); free(P); free(num); free(energy); free(hden);
  return cancelled;
}

/* NMFCross::synthesize :50-58: out[F][B] (complex) = H[F][R] * S[R][B] (complex source spectrogram) */
FO_EXPORT void fo_nmfcross_synthesize(const double* H, const double* S, int64_t F, int64_t R, int64_t B, double* out)
{
  memset(out, 0, sizeof(double) * (size_t) (2 * F * B));
  for (int64_t f = 0; f < F; f++)
    for (int64_t k = 0; k < R; k++) {
      double h = H[f * R + k];
      const double* s = S + 2 * k * B;
      double* o = out + 2 * f * B;
      for (int64_t b = 0; b < 2 * B; b++) o[b] += h * s[b];
    }
}

/* GriffinLim::process  algorithms/public/GriffinLim.hpp:29-54.  spec[F][B] complex, in place.
 * Random phase: EigenRandomPhase (EigenRandom.hpp:146-160): polar(1, U(0, 2 pi)) from mt19937_64(seed), filling a
 * column-major F x B array in storage order: entry (f, b) is draw b * F + f; libstdc++'s uniform_real_distribution(a, b)
 * is generate_canonical * (b - a) + a. */
FO_EXPORT void fo_griffinlim(double* spec, int64_t F, int64_t B, int64_t n_samples, int64_t n_iter, int64_t win,
                             int64_t fft, int64_t hop, int64_t seed)
{
  const double momentum = 0.9, twopi = 2.0 * 3.14159265358979323846; /* 2 * M_PI */
  int64_t cnt = F * B;
  double* mag = (double*) malloc(sizeof(double) * (size_t) cnt);
  double* phase = (double*) malloc(sizeof(double) * (size_t) (2 * cnt));
  double* est = (double*) calloc((size_t) (2 * cnt), sizeof(double));
  double* prev = (double*) calloc((size_t) (2 * cnt), sizeof(double));
  double* sg = (double*) malloc(sizeof(double) * (size_t) (2 * cnt));
  double* tmp = (double*) calloc((size_t) n_samples, sizeof(double));
  for (int64_t e = 0; e < cnt; e++) mag[e] = hypot(spec[2 * e], spec[2 * e + 1]);   /* :39 */
  {
    fo_mt64 g; fo_mt64_seed(&g, (uint64_t) seed);
    for (int64_t b = 0; b < B; b++)
      for (int64_t f = 0; f < F; f++) {
        double th = fo_mt64_uniform(&g) * (twopi - 0.0) + 0.0;
        phase[2 * (f * B + b)] = cos(th);
        phase[2 * (f * B + b) + 1] = sin(th);
      }
  }
  for (int64_t i = 0; i < n_iter; i++) {                                            /* :44-52 */
    memcpy(prev, est, sizeof(double) * (size_t) (2 * cnt));
    for (int64_t e = 0; e < cnt; e++) { sg[2 * e] = mag[e] * phase[2 * e]; sg[2 * e + 1] = mag[e] * phase[2 * e + 1]; }
    fo_istft(sg, F, win, fft, hop, tmp, n_samples);
    fo_stft(tmp, n_samples, win, fft, hop, est);
    for (int64_t e = 0; e < cnt; e++) {
      double re = est[2 * e] - (momentum / (1 + momentum)) * prev[2 * e];
      double im = est[2 * e + 1] - (momentum / (1 + momentum)) * prev[2 * e + 1];
      double a = hypot(re, im) + FO_EPS;
      phase[2 * e] = re / a; phase[2 * e + 1] = im / a;
    }
  }
  for (int64_t e = 0; e < cnt; e++) { spec[2 * e] = mag[e] * phase[2 * e]; spec[2 * e + 1] = mag[e] * phase[2 * e + 1]; } /* :53 */
  free(mag); free(phase); free(est); free(prev); free(sg); free(tmp);
}

/* BufNMFCross  clients/nrt/NMFCrossClient.hpp:85-185: mono source and target (channel 0), output [n_target] float.
 * H_out (optional, [Ft][Fs]) returns the activations for tests. */
FO_EXPORT int fo_bufnmfcross(const float* source, int64_t ns, const float* target, int64_t nt, int64_t win, int64_t fft,
                             int64_t hop, int64_t time_sparsity, int64_t polyphony, int64_t continuity, int64_t n_iter,
                             int64_t seed, int64_t gl_iter, float* out, double* H_out)
{
  int64_t B = fft / 2 + 1;
  int64_t Fs = fo_stft_num_frames(ns, win, hop), Ft = fo_stft_num_frames(nt, win, hop);   /* :103-108 */
  if (ns <= 0 || nt <= 0 || time_sparsity > Ft || continuity > Ft) return -1;             /* :110-119 */
  double* sa = (double*) malloc(sizeof(double) * (size_t) ns);
  double* ta = (double*) malloc(sizeof(double) * (size_t) nt);
  for (int64_t i = 0; i < ns; i++) sa[i] = (double) source[i];
  for (int64_t i = 0; i < nt; i++) ta[i] = (double) target[i];
  double* S = (double*) malloc(sizeof(double) * (size_t) (2 * Fs * B));
  double* T = (double*) malloc(sizeof(double) * (size_t) (2 * Ft * B));
  double* W = (double*) malloc(sizeof(double) * (size_t) (Fs * B));
  double* M = (double*) malloc(sizeof(double) * (size_t) (Ft * B));
  double* H = (double*) malloc(sizeof(double) * (size_t) (Ft * Fs));
  double* res = (double*) malloc(sizeof(double) * (size_t) (2 * Ft * B));
  double* audio = (double*) malloc(sizeof(double) * (size_t) nt);
  fo_stft(sa, ns, win, fft, hop, S); fo_magnitude(S, Fs * B, W);                          /* :134-136 */
  fo_stft(ta, nt, win, fft, hop, T); fo_magnitude(T, Ft * B, M);                          /* :137-139 */
  int64_t p = polyphony < Fs ? polyphony : Fs;                                            /* :160 */
  fo_nmfcross_process(M, Ft, B, W, Fs, n_iter, time_sparsity, p, continuity, seed, H, NULL, NULL); /* :158-161 */
  if (H_out) memcpy(H_out, H, sizeof(double) * (size_t) (Ft * Fs));
  fo_nmfcross_synthesize(H, S, Ft, Fs, B, res);                                           /* :166 */
  fo_griffinlim(res, Ft, B, nt, gl_iter, win, fft, hop, seed);                            /* :171-173 (50 iterations) */
  fo_istft(res, Ft, win, fft, hop, audio, nt);                                            /* :178 */
  for (int64_t i = 0; i < nt; i++) out[i] = (float) audio[i];                             /* :183 */
  free(sa); free(ta); free(S); free(T); free(W); free(M); free(H); free(res); free(audio);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * MelBands  algorithms/public/MelBands.hpp:35-101  (an epilogue on |X|: SURVEY 8f rank 4)
 * Eigen's LinSpaced(n, a, b) for doubles (Eigen 3.4 linspaced_op_impl, non-integer): step = (b - a) / (n - 1); when
 * |b| < |a| the sequence is evaluated from the top (b - (n-1-i) step, first element exactly a), else from the bottom
 * (a + i step, last element exactly b).
 * ---------------------------------------------------------------------------------------------- */
static double fo_linspaced(int64_t n, double a, double b, int64_t i)
{
  if (n == 1) return b;
  double step = (b - a) / (double) (n - 1);
  int flip = fabs(b) < fabs(a);
  if (flip) return i == 0 ? a : b - (double) (n - 1 - i) * step;
  return i == n - 1 ? b : a + (double) i * step;
}

/* init :43-80: filters[nBands][nBins] triangular, scale1 / scale2 */
FO_EXPORT void fo_melbands_init(double lo, double hi, int64_t n_bands, int64_t n_bins, double sample_rate, int64_t win,
                                double* filters, double* scale1, double* scale2)
{
  int64_t fft = 2 * (n_bins - 1);
  *scale1 = 1.0 / ((double) win / 4.0);                                   /* :50 */
  *scale2 = 1.0 / (2.0 * (double) fft / (double) win);                    /* :53 */
  double mlo = 1127.01048 * log(lo / 700.0 + 1.0), mhi = 1127.01048 * log(hi / 700.0 + 1.0); /* :38-41 */
  double* mel = (double*) malloc(sizeof(double) * (size_t) (n_bands + 2));
  for (int64_t i = 0; i < n_bands + 2; i++) mel[i] = 700.0 * (exp(fo_linspaced(n_bands + 2, mlo, mhi, i) / 1127.01048) - 1.0); /* :55-56 */
  for (int64_t i = 0; i < n_bands; i++) {
    double d0 = fabs(mel[i] - mel[i + 1]), d1 = fabs(mel[i + 1] - mel[i + 2]);                  /* :62-64 */
    for (int64_t b = 0; b < n_bins; b++) {
      double f = fo_linspaced(n_bins, 0.0, sample_rate / 2.0, b);                               /* :60 */
      double lower = -(mel[i] - f) / d0, upper = (mel[i + 2] - f) / d1;                          /* :72-73 */
      double v = lower < upper ? lower : upper;
      filters[i * n_bins + b] = v > 0.0 ? v : 0.0;                                               /* :74 */
    }
  }
  free(mel);
}

/* processFrame :82-101.  NOTE: the reference scales / squares the caller's frame IN PLACE (the Eigen map aliases `in`);
 * `frame` is mutated here in the same way. */
FO_EXPORT void fo_melbands_frame(const double* filters, int64_t n_bands, int64_t n_bins, double scale1, double scale2,
                                 double* frame, double* out, int mag_norm, int use_power, int log_output)
{
  if (mag_norm) for (int64_t b = 0; b < n_bins; b++) frame[b] *= scale1;     /* :90 */
  double energy = 0.0;
  for (int64_t b = 0; b < n_bins; b++) energy += frame[b];
  energy *= scale2;                                                           /* :91 */
  if (use_power) for (int64_t b = 0; b < n_bins; b++) frame[b] *= frame[b];  /* :92 */
  double sum = 0.0;
  for (int64_t i = 0; i < n_bands; i++) {                                     /* :94-95 */
    double s = 0.0;
    for (int64_t b = 0; b < n_bins; b++) s += filters[i * n_bins + b] * frame[b];
    out[i] = s;
    sum += s;
  }
  if (mag_norm) { double d = sum > FO_EPS ? sum : FO_EPS; for (int64_t i = 0; i < n_bands; i++) out[i] = out[i] * energy / d; } /* :97 */
  if (log_output) for (int64_t i = 0; i < n_bands; i++) out[i] = 20.0 * log10(out[i] > FO_EPS ? out[i] : FO_EPS);              /* :99 */
}

/* mags[F][B] -> bands[F][nBands]: init + processFrame per frame (the client's call: MelBandsClient.hpp:96-113) */
FO_EXPORT void fo_melbands(const double* mags, int64_t F, int64_t n_bins, double lo, double hi, int64_t n_bands,
                           double sample_rate, int64_t win, int mag_norm, int use_power, int log_output, double* bands)
{
  double* filt = (double*) malloc(sizeof(double) * (size_t) (n_bands * n_bins));
  double* frame = (double*) malloc(sizeof(double) * (size_t) n_bins);
  double s1, s2;
  fo_melbands_init(lo, hi, n_bands, n_bins, sample_rate, win, filt, &s1, &s2);
  for (int64_t f = 0; f < F; f++) {
    memcpy(frame, mags + f * n_bins, sizeof(double) * (size_t) n_bins);
    fo_melbands_frame(filt, n_bands, n_bins, s1, s2, frame, bands + f * n_bands, mag_norm, use_power, log_output);
  }
  free(filt); free(frame);
}

/* ------------------------------------------------------------------------------------------------
 * HPSS  algorithms/public/HPSS.hpp:47-162 + algorithms/util/MedianFilter.hpp:36-57, frame by frame from init() state.
 * spec[F][B] complex in -> out[3][F][B] complex (harmonic, percussive, residual), exactly what processFrame emits per
 * frame: the output of frame t belongs to input frame t - (hSize - 1) (zeros while the delay line fills).
 * Restated literally, including the placement quirks: the vertical median of bin b runs over bins b .. b + vSize - 1
 * (padded.segment(v2 * 3, nBins) of a causal filter over an array padded by v2 at the front, :79-89), the horizontal
 * median is written to column h2 + 1 of its delay line (:91-93) and therefore reaches column 0 h2 + 1 frames later.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double* unsorted; double* sorted; int64_t n; } fo_median;
static void fo_median_init(fo_median* m, int64_t n) { m->n = n; memset(m->unsorted, 0, sizeof(double) * (size_t) n); memset(m->sorted, 0, sizeof(double) * (size_t) n); }
static double fo_median_push(fo_median* m, double val)
{ /* MedianFilter::processSample :47-57 */
  int64_t n = m->n;
  double old = m->unsorted[0];
  memmove(m->unsorted, m->unsorted + 1, sizeof(double) * (size_t) (n - 1));
  m->unsorted[n - 1] = val;
  int64_t lo = 0;                       /* lower_bound(old): first element >= old */
  while (lo < n && m->sorted[lo] < old) lo++;
  memmove(m->sorted + lo, m->sorted + lo + 1, sizeof(double) * (size_t) (n - 1 - lo));
  int64_t up = 0;                       /* upper_bound(val) in the n-1 remaining: first element > val */
  while (up < n - 1 && !(val < m->sorted[up])) up++;
  memmove(m->sorted + up + 1, m->sorted + up, sizeof(double) * (size_t) (n - 1 - up));
  m->sorted[up] = val;
  return m->sorted[n / 2];
}

static void fo_hpss_threshold(int64_t n_bins, double x1, double y1, double x2, double y2, double* th)
{ /* makeThreshold :164-181 */
  int64_t ks = (int64_t) floor(x1 * (double) n_bins), ke = (int64_t) floor(x2 * (double) n_bins), kl = ke - ks;
  for (int64_t i = 0; i < n_bins; i++) th[i] = 1.0;
  for (int64_t i = 0; i < ks; i++) th[i] = pow(10.0, y1 / 20.0);
  for (int64_t i = 0; i < kl; i++) th[ks + i] = pow(10.0, fo_linspaced(kl, y1, y2, i) / 20.0);
  for (int64_t i = ke; i < n_bins; i++) th[i] = pow(10.0, y2 / 20.0);
}

FO_EXPORT void fo_hpss(const double* spec, int64_t F, int64_t B, int64_t v_size, int64_t h_size, int64_t mode, double hx1,
                       double hy1, double hx2, double hy2, double px1, double py1, double px2, double py2, double* out)
{
  int64_t h2 = (h_size - 1) / 2, v2 = (v_size - 1) / 2;
  double* v = (double*) calloc((size_t) (B * h_size), sizeof(double));
  double* h = (double*) calloc((size_t) (B * h_size), sizeof(double));
  double* buf = (double*) calloc((size_t) (2 * B * h_size), sizeof(double));
  double* padded = (double*) malloc(sizeof(double) * (size_t) (2 * v_size + B));
  double* hstore = (double*) calloc((size_t) (2 * B * h_size), sizeof(double));
  fo_median* hf = (fo_median*) malloc(sizeof(fo_median) * (size_t) B);
  for (int64_t b = 0; b < B; b++) { hf[b].unsorted = hstore + 2 * b * h_size; hf[b].sorted = hf[b].unsorted + h_size; fo_median_init(&hf[b], h_size); } /* :57-60 */
  double* vstore = (double*) malloc(sizeof(double) * (size_t) (2 * v_size));
  fo_median vf; vf.unsorted = vstore; vf.sorted = vstore + v_size;
  double* th_h = (double*) malloc(sizeof(double) * (size_t) B);
  double* th_p = (double*) malloc(sizeof(double) * (size_t) B);
  fo_hpss_threshold(B, hx1, hy1, hx2, hy2, th_h);
  fo_hpss_threshold(B, px1, py1, px2, py2, th_p);
  for (int64_t t = 0; t < F; t++) {
    const double* in = spec + 2 * t * B;
    for (int64_t b = 0; b < B; b++) {                                       /* :74-76 shift the delay lines */
      memmove(v + b * h_size, v + b * h_size + 1, sizeof(double) * (size_t) (h_size - 1));
      memmove(h + b * h_size, h + b * h_size + 1, sizeof(double) * (size_t) (h_size - 1));
      memmove(buf + 2 * b * h_size, buf + 2 * b * h_size + 2, sizeof(double) * (size_t) (2 * (h_size - 1)));
    }
    for (int64_t i = 0; i < 2 * v_size + B; i++) padded[i] = 0.0;           /* :78-79 */
    for (int64_t b = 0; b < B; b++) padded[v2 + b] = hypot(in[2 * b], in[2 * b + 1]); /* :81 */
    fo_median_init(&vf, v_size);                                            /* :82 */
    for (int64_t i = 0; i < 2 * v_size + B; i++) padded[i] = fo_median_push(&vf, padded[i]); /* :83-86 */
    for (int64_t b = 0; b < B; b++) {
      v[b * h_size + h_size - 1] = padded[v2 * 3 + b];                      /* :88 */
      buf[2 * (b * h_size + h_size - 1)] = in[2 * b];                       /* :89 */
      buf[2 * (b * h_size + h_size - 1) + 1] = in[2 * b + 1];
      h[b * h_size + h2 + 1] = fo_median_push(&hf[b], hypot(in[2 * b], in[2 * b + 1])); /* :90-92 */
    }
    for (int64_t b = 0; b < B; b++) {
      double h0 = h[b * h_size], v0 = v[b * h_size];
      double hm, pm, rm = mode == 2 ? 1.0 : 0.0;                            /* :97-98 */
      if (mode == 0) {                                                      /* :101-107 */
        double d = h0 + v0;
        double mult = 1.0 / (d > FO_EPS ? d : FO_EPS);
        hm = h0 * mult; pm = v0 * mult;
      } else if (mode == 1) {                                               /* :108-115 */
        hm = (h0 / v0) > th_h[b] ? 1.0 : 0.0;
        pm = 1.0 - hm;
      } else {                                                              /* :116-135 */
        hm = (h0 / v0) > th_h[b] ? 1.0 : 0.0;
        pm = (v0 / h0) > th_p[b] ? 1.0 : 0.0;
        rm = rm * (1.0 - hm);
        rm = rm * (1.0 - pm);
        double nrm = 1.0 / (hm + pm + rm);
        nrm = nrm > FO_EPS ? nrm : FO_EPS;
        hm *= nrm; pm *= nrm; rm *= nrm;
      }
      double re = buf[2 * b * h_size], im = buf[2 * b * h_size + 1];
      double m0 = hm < 1.0 ? hm : 1.0, m1 = pm < 1.0 ? pm : 1.0, m2 = rm < 1.0 ? rm : 1.0; /* :138-140 */
      out[2 * ((0 * F + t) * B + b)] = re * m0; out[2 * ((0 * F + t) * B + b) + 1] = im * m0;
      out[2 * ((1 * F + t) * B + b)] = re * m1; out[2 * ((1 * F + t) * B + b) + 1] = im * m1;
      out[2 * ((2 * F + t) * B + b)] = re * m2; out[2 * ((2 * F + t) * B + b) + 1] = im * m2;
    }
  }
  free(v); free(h); free(buf); free(padded); free(hstore); free(hf); free(vstore); free(th_h); free(th_p);
}

/* ------------------------------------------------------------------------------------------------
 * NNDSVD  algorithms/public/NNDSVD.hpp:30-131  (NMFSeed: SVD-based seeds for W and H)
 * The reference calls Eigen::BDCSVD (thin U, V) on X^T (bins x frames).  The SVD is restated as a one-sided Jacobi
 * iteration in fp64 (singular values and, up to the sign of each (u, v) pair, singular vectors are unique for distinct
 * singular values).  Method 0 takes absolute values, so the sign does not matter.  Methods 1-3 split u, v into positive
 * and negative parts, and -- because of the reference's `yNNorm = xN.norm()` (:84), restated as written -- their result
 * DOES depend on the sign BDCSVD happens to return, which no public contract fixes.  This restatement (and the GPU path)
 * fix it by a convention: the entry of u of largest magnitude is positive.
 * X[F][B] -> W[max_rank][B], H[F][max_rank] (both must come in zero-filled, as the client allocates them); returns k.
 * ---------------------------------------------------------------------------------------------- */
static void fo_svd_jacobi(const double* X, int64_t F, int64_t B, double* U /*[r][B] rows = left vectors of X^T*/,
                          double* s /*[r]*/, double* V /*[r][F]*/)
{ /* columns of A = X (F x B) are orthogonalised: A J = Q Sigma  =>  X^T = J Sigma Q^T */
  double* A = (double*) malloc(sizeof(double) * (size_t) (B * F)); /* A^T: row b = column b of X */
  double* J = (double*) calloc((size_t) (B * B), sizeof(double));  /* J^T: row b = column b of J */
  for (int64_t b = 0; b < B; b++) { J[b * B + b] = 1.0; for (int64_t f = 0; f < F; f++) A[b * F + f] = X[f * B + b]; }
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0.0;
    for (int64_t i = 0; i < B - 1; i++)
      for (int64_t j = i + 1; j < B; j++) {
        double a = 0.0, bb = 0.0, c = 0.0;
        const double *x = A + i * F, *y = A + j * F;
        for (int64_t f = 0; f < F; f++) { a += x[f] * x[f]; bb += y[f] * y[f]; c += x[f] * y[f]; }
        if (c == 0.0 || a == 0.0 || bb == 0.0) continue;
        double rel = fabs(c) / sqrt(a * bb);
        if (rel > off) off = rel;
        if (rel < 1e-15) continue;
        double zeta = (bb - a) / (2.0 * c);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        double *xi = A + i * F, *yj = A + j * F;
        for (int64_t f = 0; f < F; f++) { double p = xi[f], q = yj[f]; xi[f] = cs * p - sn * q; yj[f] = sn * p + cs * q; }
        double *ji = J + i * B, *jj = J + j * B;
        for (int64_t k = 0; k < B; k++) { double p = ji[k], q = jj[k]; ji[k] = cs * p - sn * q; jj[k] = sn * p + cs * q; }
      }
    if (off < 1e-14) break;
  }
  int64_t r = B < F ? B : F;
  double* nrm = (double*) malloc(sizeof(double) * (size_t) B);
  int64_t* ord = (int64_t*) malloc(sizeof(int64_t) * (size_t) B);
  for (int64_t b = 0; b < B; b++) { double a = 0.0; for (int64_t f = 0; f < F; f++) a += A[b * F + f] * A[b * F + f]; nrm[b] = sqrt(a); ord[b] = b; }
  for (int64_t i = 1; i < B; i++) { /* insertion sort, descending */
    int64_t o = ord[i]; int64_t k = i - 1;
    while (k >= 0 && nrm[ord[k]] < nrm[o]) { ord[k + 1] = ord[k]; k--; }
    ord[k + 1] = o;
  }
  for (int64_t c = 0; c < r; c++) {
    int64_t b = ord[c];
    s[c] = nrm[b];
    double sign = 1.0, big = 0.0;
    for (int64_t k = 0; k < B; k++) if (fabs(J[b * B + k]) > big) { big = fabs(J[b * B + k]); sign = J[b * B + k] < 0 ? -1.0 : 1.0; }
    for (int64_t k = 0; k < B; k++) U[c * B + k] = sign * J[b * B + k];
    for (int64_t f = 0; f < F; f++) V[c * F + f] = nrm[b] > 0 ? sign * A[b * F + f] / nrm[b] : 0.0;
  }
  free(A); free(J); free(nrm); free(ord);
}

FO_EXPORT int64_t fo_nndsvd(const double* X, int64_t F, int64_t B, int64_t min_rank, int64_t max_rank, double amount,
                            int64_t method, int64_t seed, double* W, double* H, double* sv_out)
{
  int64_t r = B < F ? B : F;
  double* U = (double*) malloc(sizeof(double) * (size_t) (r * B));
  double* V = (double*) malloc(sizeof(double) * (size_t) (r * F));
  double* s = (double*) malloc(sizeof(double) * (size_t) r);
  fo_svd_jacobi(X, F, B, U, s, V);
  if (sv_out) memcpy(sv_out, s, sizeof(double) * (size_t) r);
  int64_t k = 0;
  if (amount == 0) k = min_rank;                                              /* :49-50 */
  else {
    double cur = 0.0, total = 0.0;
    for (int64_t i = 0; i < r; i++) total += s[i];
    while ((cur / total) < amount && k < r) cur += s[k++];                    /* :53-55 (k < r: guard only) */
  }
  if (k < min_rank) k = min_rank;                                             /* :57 */
  if (k > max_rank) k = max_rank;                                             /* :58 */
  if (k > r) k = r;
  if (method == 0) {                                                          /* :60-65 */
    for (int64_t j = 0; j < k; j++) {
      for (int64_t b = 0; b < B; b++) W[j * B + b] = fabs(U[j * B + b]);
      for (int64_t f = 0; f < F; f++) H[f * max_rank + j] = fabs(s[j] * V[j * F + f]);
    }
  } else {
    for (int64_t b = 0; b < B; b++) W[b] = fabs(U[b]);                        /* :69 */
    for (int64_t f = 0; f < F; f++) H[f * max_rank] = sqrt(s[0]) * fabs(V[f]); /* :70 */
    for (int64_t j = 1; j < k; j++) {                                         /* :72-104 */
      const double *x = U + j * B, *y = V + j * F;
      double xp = 0, yp = 0, xn = 0;
      for (int64_t b = 0; b < B; b++) { double p = x[b] > 0 ? x[b] : 0, q = x[b] < 0 ? -x[b] : 0; xp += p * p; xn += q * q; }
      for (int64_t f = 0; f < F; f++) { double p = y[f] > 0 ? y[f] : 0; yp += p * p; }
      double xPNorm = sqrt(xp), yPNorm = sqrt(yp), xNNorm = sqrt(xn);
      double yNNorm = xNNorm;                                                 /* :84  "double yNNorm = xN.norm();" */
      double mP = xPNorm * yPNorm, mN = xNNorm * yNNorm;
      double sigma;
      if (mP > mN) {
        sigma = mP;
        for (int64_t b = 0; b < B; b++) W[j * B + b] = (x[b] > 0 ? x[b] : 0) / xPNorm;
        double lbd = sqrt(s[j] * sigma);
        for (int64_t f = 0; f < F; f++) H[f * max_rank + j] = lbd * ((y[f] > 0 ? y[f] : 0) / yPNorm);
      } else {
        sigma = mN;
        for (int64_t b = 0; b < B; b++) W[j * B + b] = (x[b] < 0 ? -x[b] : 0) / xNNorm;
        double lbd = sqrt(s[j] * sigma);
        for (int64_t f = 0; f < F; f++) H[f * max_rank + j] = lbd * ((y[f] < 0 ? -y[f] : 0) / yNNorm);
      }
    }
    double mean = 0.0;
    for (int64_t e = 0; e < F * B; e++) mean += X[e];
    mean /= (double) (F * B);                                                 /* :106 */
    if (method == 1) {                                                        /* :107-117 */
      fo_mt64 g; double lo = FO_EPS, hi = mean * 0.001;
      fo_mt64_seed(&g, (uint64_t) seed);
      for (int64_t j = 0; j < max_rank; j++)       /* column-major B x maxRank: W^T(b, j) = draw j * B + b */
        for (int64_t b = 0; b < B; b++) { double u = fo_mt64_uniform(&g) * (hi - lo) + lo; if (W[j * B + b] < FO_EPS) W[j * B + b] = u; }
      fo_mt64_seed(&g, (uint64_t) seed);
      for (int64_t f = 0; f < F; f++)              /* column-major maxRank x F: H^T(j, f) = draw f * maxRank + j */
        for (int64_t j = 0; j < max_rank; j++) { double u = fo_mt64_uniform(&g) * (hi - lo) + lo; if (H[f * max_rank + j] < FO_EPS) H[f * max_rank + j] = u; }
    } else if (method == 2) {                                                 /* :118-125 */
      for (int64_t e = 0; e < max_rank * B; e++) if (W[e] < FO_EPS) W[e] = mean;
      for (int64_t e = 0; e < F * max_rank; e++) if (H[e] < FO_EPS) H[e] = mean;
    }
  }
  free(U); free(V); free(s);
  return k;
}

FO_EXPORT int fo_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
