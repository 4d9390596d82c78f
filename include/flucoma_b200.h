/*
 * flucoma_b200.h -- C ABI of libflucoma_b200.so: the B200 (sm_100a) implementation of flucoma-core's
 * STFT -> |X| -> NMF multiplicative updates -> ratio-mask -> ISTFT hot path, batched over many buffers.
 *
 * This is the drop-in boundary.  Plain C, plain pointers and sizes, no C++/torch types.
 * Every entry point names the reference interface it replaces; paths are relative to
 * /root/reference/include/flucoma.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - All matrices are dense row-major in the reference's own FluidTensor orientation:
 *       X / V / magnitude [F][B], spectrum [F][B] interleaved (re,im), W / bases [K][B], H / activations [F][K],
 *     with a leading batch axis (buffers or channels).  F = frames (nWindows), B = bins = fft/2+1, K = rank.
 *   - dtype: host/device element type of the caller's arrays (FB200_F32 or FB200_F64).  The device always computes
 *     in fp32 (tensor-core operands are split-bf16 pairs, accumulate fp32); FB200_F64 arrays are converted on the
 *     device.  The reference computes in fp64; agreement is to 1e-4 relative (tests/test_gpu_parity.py).
 *   - mem: where the caller's arrays live.  FB200_HOST pointers are copied by the library (pinned staging inside
 *     the plan); FB200_DEVICE pointers must be on the plan's device and are used in place.
 *   - Every call is synchronous: outputs are complete when it returns.
 *   - Return value: 0 ok, >0 warning, <0 error (fb200_status).  fb200_last_error() gives the text.  No exceptions
 *     or CUDA/cuFFT errors escape (reference convention: Result{Status,msg}, clients/common/Result.hpp:21-81).
 *   - A plan binds one device and one stream; calls on distinct plans are thread-safe, one plan is not re-entrant
 *     (same rule as one client instance: clients/common/FluidNRTClientWrapper.hpp:868-872).
 *   - The library never retains caller pointers after a call returns.
 */
#ifndef FLUCOMA_B200_H
#define FLUCOMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB200_ABI_VERSION 2u

#if defined(_WIN32)
#define FB200_API __declspec(dllexport)
#else
#define FB200_API __attribute__((visibility("default")))
#endif

typedef enum fb200_status {
  FB200_OK = 0,
  FB200_WARN_NO_WORK = 1,       /* NMFClient.hpp:143-145 "no work to do" */
  FB200_CANCELLED = 2,          /* Result::Status::kCancelled (progress callback returned 0) */
  FB200_ERR_INVALID = -1,       /* bad argument / shape (the reference asserts or returns kError) */
  FB200_ERR_CUDA = -2,
  FB200_ERR_CUFFT = -3,
  FB200_ERR_NOMEM = -4,
  FB200_ERR_UNSUPPORTED = -5,
  FB200_ERR_NO_DEVICE = -6
} fb200_status;

typedef enum fb200_dtype { FB200_F32 = 0, FB200_F64 = 1 } fb200_dtype;
typedef enum fb200_mem { FB200_HOST = 0, FB200_DEVICE = 1 } fb200_mem;

/* NMF update engine.  AUTO picks a tensor-core engine when the shape qualifies and the call has enough independent
 * buffers / frame tiles to fill the SMs, else SIMT.  fb200_stats.backend_used reports the engine that ran. */
typedef enum fb200_backend {
  FB200_BACKEND_AUTO = 0,
  FB200_BACKEND_SIMT = 1,             /* fp32 CUDA-core engine: any shape */
  FB200_BACKEND_TCGEN05 = 2,          /* tensor-core engines; as a request: whichever of the two takes the shape */
  FB200_BACKEND_TCGEN05_STREAMED = 3  /* tensor-core engine with W/H streamed from L2 (rank 9..64, any bins = 128 m + 1, any
                                         frame count, fixed-dictionary frame streams); as a request: force it */
} fb200_backend;

typedef struct fb200_plan fb200_plan; /* opaque; owns device memory, cuFFT plans, stream, pinned staging */

/* Replaces the constructor arguments of algorithm::STFT / ISTFT (algorithms/public/STFT.hpp:36-47,154-164) and the
 * FFTParams size rules (clients/common/ParameterTypes.hpp:295-313): hop <= 0 -> win/2, fft < 0 -> nextPow2(win). */
typedef struct fb200_config {
  uint32_t struct_size;   /* = sizeof(fb200_config) */
  int32_t device;         /* CUDA ordinal */
  int32_t win, hop, fft;  /* window, hop, fft size (fft must be a power of two >= win) */
  int32_t max_rank;       /* largest K any call will use */
  int64_t max_batch;      /* largest number of buffers/channels per call */
  int64_t max_samples;    /* largest samples per buffer (or, for frame APIs, max_frames = fb200_num_frames()) */
  int32_t backend;        /* fb200_backend */
  int32_t reserved;
} fb200_config;

/* progress callback: NMF::addProgressCallback (algorithms/public/NMF.hpp:31,136-139,175-176).
 * Called on the calling thread with iteration 1..n, each exactly once and in order; return 0 to cancel.
 * `progress_stride` of the argument structs selects how the update loop is polled:
 *   s >= 1 (0 means 1, the reference's cadence): exact mode.  The device runs s iterations, then the callbacks of those
 *       iterations are replayed; on cancel W/H hold the state after the last iteration of that group -- at s = 1 exactly
 *       the reference's NMF.hpp:175-176.  Cancellation granularity is s iterations.  Both engines support it (the
 *       tensor-core engine runs one persistent launch per group; its state round-trips exactly through fp32 W/H).
 *   FB200_PROGRESS_ASYNC: the loop is never interrupted.  The calling thread polls device-written pass counters while the
 *       persistent kernel runs and reports iteration i once the batch as a whole has done the work of i iterations (the
 *       reference's FluidTask progress is channel-major in the same way, NMFClient.hpp:233-267, FluidTask.hpp:22-39).
 *       Returning 0 raises a device-visible cancel word: every buffer in flight finishes the iteration it is in, buffers
 *       not yet started keep their initial state, the call returns FB200_CANCELLED.  This is the mode for BufNMF, which
 *       discards all outputs of a cancelled job (NMFClient.hpp:273-274). */
typedef int (*fb200_progress_fn)(void* user, int64_t iteration);
#define FB200_PROGRESS_ASYNC (-1)

/* ---- plan management ------------------------------------------------------------------------------------- */
FB200_API uint32_t fb200_abi_version(void);
FB200_API int32_t fb200_device_count(void);
FB200_API int32_t fb200_plan_create(const fb200_config* cfg, fb200_plan** out);
FB200_API void fb200_plan_destroy(fb200_plan* plan);
FB200_API const char* fb200_last_error(const fb200_plan* plan); /* plan may be NULL: last create error */

/* size rules: ParameterTypes.hpp:295-313; NMFClient.hpp:111-113 / STFT.hpp:94-99 */
FB200_API int64_t fb200_num_frames(int64_t n_samples, int32_t win, int32_t hop);
FB200_API int32_t fb200_resolve_fft(int32_t win, int32_t hop, int32_t fft, int32_t* out_hop, int32_t* out_fft,
                                    int32_t* out_bins);

/* contiguous shard of `total` independent buffers owned by `rank` of `world` */
FB200_API void fb200_shard_range(int64_t total, int32_t world, int32_t rank, int64_t* begin, int64_t* count);

/* ---- STFT::process + STFT::magnitude  (STFT.hpp:90-108, 61-66) ---------------------------------------------- */
/* audio [batch][n_samples] -> spectrum [batch][F][B] complex (may be NULL) and/or magnitude [batch][F][B] (may be NULL) */
FB200_API int32_t fb200_stft(fb200_plan* plan, const void* audio, int64_t batch, int64_t n_samples, void* spectrum,
                             void* magnitude, int32_t dtype, int32_t mem);

/* ---- ISTFT::process  (STFT.hpp:178-199) -------------------------------------------------------------------- */
/* spectrum [batch][n_frames][B] complex -> audio [batch][n_samples] */
FB200_API int32_t fb200_istft(fb200_plan* plan, const void* spectrum, int64_t batch, int64_t n_frames, void* audio,
                              int64_t n_samples, int32_t dtype, int32_t mem);

/* ---- NMF::process  (algorithms/public/NMF.hpp:91-134, 144-183) ---------------------------------------------- */
typedef struct fb200_nmf_args {
  uint32_t struct_size;
  int32_t dtype, mem;
  int64_t batch, frames, bins;  /* X is [batch][frames][bins] */
  int32_t rank, iterations;
  int32_t update_w, update_h;   /* NMF.hpp:92-93 */
  const void* X;                /* magnitudes, >= 0 */
  const int64_t* seeds;         /* [batch] host array; seed < 0 -> nondeterministic (EigenRandom.hpp:14-17,80); NULL -> all -1 */
  const void* W0;               /* optional [batch][rank][bins]  (NMF.hpp:94, 102-112); NULL -> random */
  const void* H0;               /* optional [batch][frames][rank] (NMF.hpp:95, 114-124); NULL -> random */
  void* W1;                     /* out [batch][rank][bins]   (may be NULL) */
  void* H1;                     /* out [batch][frames][rank] (may be NULL) */
  void* V1;                     /* out [batch][frames][bins] = W*H (may be NULL); untouched copy of X if cancelled */
  fb200_progress_fn progress;   /* optional */
  void* progress_user;
  int32_t progress_stride;      /* see fb200_progress_fn: >= 1 exact (0 -> 1), FB200_PROGRESS_ASYNC */
  int32_t reserved;
} fb200_nmf_args;
FB200_API int32_t fb200_nmf_process(fb200_plan* plan, const fb200_nmf_args* args);

/* ---- NMF::processFrame over many frames, fixed dictionary  (NMF.hpp:45-89; NMFMatchClient.hpp:106-118) ------ */
typedef struct fb200_frames_args {
  uint32_t struct_size;
  int32_t dtype, mem;
  int64_t frames, bins;
  int32_t rank, iterations;     /* NMFMatch hard-codes 10 (NMFMatchClient.hpp:115) */
  int64_t seed;                 /* h0 = U(K) from this seed, identical for every frame when >= 0 */
  const void* X;                /* [frames][bins] magnitudes */
  const void* W0;               /* [rank][bins]; used as given (clamped + row-normalised internally, NMF.hpp:58,63-64) */
  void* W_norm;                 /* optional out [rank][bins]: the mutated W0 the reference leaves behind */
  void* H;                      /* out [frames][rank] */
  void* V;                      /* optional out [frames][bins] = W^T h (NMF.hpp:88) */
} fb200_frames_args;
FB200_API int32_t fb200_nmf_process_frames(fb200_plan* plan, const fb200_frames_args* args);

/* ---- BufNMF: NMFClient::process per channel  (clients/nrt/NMFClient.hpp:233-335) ---------------------------- */
typedef struct fb200_bufnmf_args {
  uint32_t struct_size;
  int32_t mem;                  /* audio & outputs are float32, like BufferAdaptor (BufferAdaptor.hpp:49-66) */
  int64_t batch, n_samples;     /* `batch` channels/buffers of equal length, planar [batch][n_samples] */
  int32_t rank, iterations;
  int32_t bases_mode, acts_mode; /* 0 none / 1 seed / 2 fixed (NMFClient.hpp:63-67) */
  const float* audio;
  const int64_t* seeds;         /* [batch] host; NULL -> -1 */
  const float* bases_in;        /* [batch][rank][bins] when bases_mode > 0 */
  const float* acts_in;         /* [batch][frames][rank] when acts_mode > 0 */
  float* bases_out;             /* [batch][rank][bins]; not written when bases_mode == 2 (NMFClient.hpp:277) */
  float* acts_out;              /* [batch][frames][rank], scaled by 1/max(H) per channel (:289-298); not written when acts_mode == 2 */
  float* resynth_out;           /* optional [batch][rank][n_samples] (:302-333) */
  fb200_progress_fn progress;
  void* progress_user;
  int32_t progress_stride;
  int32_t reserved;
} fb200_bufnmf_args;
FB200_API int32_t fb200_bufnmf(fb200_plan* plan, const fb200_bufnmf_args* args);

/* ---- NMFFilter over a stream  (clients/rt/NMFFilterClient.hpp:98-117 + BufferedProcess.hpp:187-241) ---------- */
/* audio [n_samples]: a mono stream from reset state.  Frame f (f*hop < n_samples) covers stream[f*hop - win, f*hop)
 * (FluidSource.hpp:68-89); out [rank][n_samples] are the `rank` masked resyntheses exactly as the streaming client emits
 * them, i.e. delayed by its latency of `win` samples (NMFFilterClient.hpp:64) -- append `win` zeros to flush the tail.
 * out == NULL with acts_out != NULL is NMFMatch on the same stream (clients/rt/NMFMatchClient.hpp:106-118, which
 * hard-codes 10 iterations).  Independent of the host vector size the client would have been driven with. */
typedef struct fb200_filter_args {
  uint32_t struct_size;
  int32_t mem;
  int64_t n_samples;
  int32_t rank, iterations;
  int64_t seed;
  const float* audio;
  const float* bases;           /* [rank][bins] */
  float* out;                   /* [rank][n_samples] */
  float* acts_out;              /* optional [frames][rank] (what NMFMatch would output) */
} fb200_filter_args;
FB200_API int32_t fb200_nmf_filter(fb200_plan* plan, const fb200_filter_args* args);

/* ---- BufNMFCross: NMFCrossClient::process  (clients/nrt/NMFCrossClient.hpp:85-185) -------------------------------------- */
/* Resynthesises `target` out of the frames of `source` (channel 0 of each, as the client does): STFT of both, NMFCross::process
 * (algorithms/public/NMFCross.hpp:60-185: activation-only KL updates with the source magnitude spectrogram as dictionary,
 * rank = source frames, and temporal-sparseness / polyphony / continuity post-processing of H every iteration),
 * NMFCross::synthesize (H times the complex source spectrogram), GriffinLim::process (algorithms/public/GriffinLim.hpp:29-54,
 * the client hard-codes 50 iterations) and ISTFT::process.  The attenuation factor of the sparseness / polyphony steps is
 * computed as the reference writes it, with integer operands (NMFCross.hpp:119,136).
 * progress is called with 1..iterations during the updates, then iterations+1 .. +3 after synthesis, Griffin-Lim and the
 * inverse transform (the client's progressTotal = iterations + 3, :152). */
typedef struct fb200_nmfcross_args {
  uint32_t struct_size;
  int32_t mem;
  int64_t n_source, n_target;
  int32_t time_sparsity, polyphony, continuity;   /* defaults 7, 11, 7 (odd) */
  int32_t iterations;                              /* default 50 */
  int64_t seed;                                    /* H init and Griffin-Lim phases; < 0 -> nondeterministic */
  int32_t griffinlim_iterations;                   /* 50 in the client (:172) */
  int32_t reserved;
  const float* source;                             /* [n_source] */
  const float* target;                             /* [n_target] */
  float* out;                                      /* [n_target] (may be NULL when only the activations are wanted) */
  float* acts_out;                                 /* optional [target frames][source frames] */
  fb200_progress_fn progress;
  void* progress_user;
} fb200_nmfcross_args;
FB200_API int32_t fb200_bufnmfcross(fb200_plan* plan, const fb200_nmfcross_args* args);

/* ---- NMFSeed: NMFSeedClient::process  (clients/nrt/NMFSeedClient.hpp:74-133; NNDSVD.hpp:30-131) --------------------------- */
/* SVD-based seeds for BufNMF's bases / activations (feed them to fb200_bufnmf with bases_mode = acts_mode = 1).  Input: mono
 * audio [n_samples] (STFT::process + magnitude run first) or magnitudes [frames][bins].  The thin SVD of X^T is computed on
 * the device (one-sided Jacobi, fp64).  *rank_out = number of components that cover `coverage` of the singular-value sum,
 * clamped to [min_rank, max_rank] (:46-58).  bases [max_rank][bins] and acts [frames][max_rank] have their first *rank_out
 * rows / columns filled (the rest: zeros, or the fill of methods 1 / 2, exactly as the reference's full-size matrices).
 * method 0 NMF-SVD, 1 NNDSVDar, 2 NNDSVDa, 3 NNDSVD.  Methods 1-3 depend on the sign of each singular pair through the
 * reference's `yNNorm = xN.norm()` (:84, kept as written); Eigen's BDCSVD sign is unspecified, this library fixes it by making
 * the largest-magnitude entry of every left vector positive.  scale_acts != 0 multiplies acts by 1 / max(acts) (:121-128). */
typedef struct fb200_nmfseed_args {
  uint32_t struct_size;
  int32_t mem;
  int64_t n_samples, frames;
  int32_t min_rank, max_rank;      /* client defaults 1, 200 */
  double coverage;                 /* 0.5 */
  int32_t method, scale_acts;
  int64_t seed;                    /* method 1 */
  const float* audio;              /* or NULL */
  const float* mags;               /* or NULL */
  float* bases;                    /* optional out */
  float* acts;                     /* optional out */
  double* singular_values;         /* optional out, HOST memory, [min(bins, frames)] descending */
  int32_t* rank_out;               /* HOST memory */
} fb200_nmfseed_args;
FB200_API int32_t fb200_nmfseed(fb200_plan* plan, const fb200_nmfseed_args* args);

/* ---- epilogues on the STFT output: MelBands and HPSS over whole frame sequences ------------------------------ */
/* MelBands::init + processFrame (algorithms/public/MelBands.hpp:43-101; MelBandsClient.hpp:96-113 calls it with usePower =
 * false).  Input either magnitudes mags [batch][frames][bins] or audio [batch][n_samples] (then STFT::process + magnitude run
 * first, frames = fb200_num_frames(n_samples)); bands [batch][frames][n_bands].  The window size of the scale factors is the
 * plan's.  flags: 1 magNorm (client: normalize), 2 usePower, 4 logOutput (client: scale = dB). */
typedef struct fb200_melbands_args {
  uint32_t struct_size;
  int32_t mem;
  int64_t batch, frames, n_samples;
  int32_t n_bands, flags;
  double lo, hi, sample_rate;       /* client defaults: 20, 20000 Hz; 40 bands */
  const float* mags;                /* or NULL */
  const float* audio;               /* or NULL */
  float* bands;
} fb200_melbands_args;
FB200_API int32_t fb200_melbands(fb200_plan* plan, const fb200_melbands_args* args);

/* HPSS::processFrame (algorithms/public/HPSS.hpp:66-162, median filters algorithms/util/MedianFilter.hpp:36-57) applied to
 * the frames of spectrum [batch][frames][bins] (complex, interleaved) in order from init() state.  out [batch][3][frames][bins]:
 * harmonic, percussive, residual spectra exactly as processFrame emits them, i.e. output frame t belongs to input frame
 * t - (h_size - 1) (zeros while the delay line fills; HPSSClient::latency adds one window, HPSSClient.hpp:80-84).
 * v_size = percussive (vertical, across bins) filter size, h_size = harmonic (horizontal, across frames) filter size, both
 * odd; mode 0 classic soft masks, 1 coupled, 2 advanced; thresholds = {hX1, hY1, hX2, hY2, pX1, pY1, pX2, pY2} (x as a
 * fraction of the bins, y in dB; used by modes 1 and 2, makeThreshold :164-181). */
typedef struct fb200_hpss_args {
  uint32_t struct_size;
  int32_t mem;
  int64_t batch, frames;
  int32_t v_size, h_size, mode, reserved;
  double thresholds[8];
  const float* spectrum;
  float* out;
} fb200_hpss_args;
FB200_API int32_t fb200_hpss(fb200_plan* plan, const fb200_hpss_args* args);

/* ---- BufNMF over several devices of one process ------------------------------------------------------------ */
/* The channels / buffers of a BufNMF job are independent (NMFClient.hpp:233 loops over them), so the job is cut into
 * contiguous shards (fb200_shard_range), one per plan; every shard runs the whole pipeline on its plan's device from its
 * own host thread and writes straight into the job's (host) arrays.  There is no collective on the data path.  With
 * gathered_acts != NULL the final activations of ALL buffers are additionally gathered onto EVERY device with one
 * ncclAllGather (libnccl.so.2 is dlopen'ed on first use): gathered_acts[d] is a device pointer on plans[d]'s device to
 * float [n_devices * per][frames][rank], per = ceil(batch / n_devices); shard r occupies rows [r * per, r * per + count_r).
 * A progress callback sees the minimum iteration over the shards, each iteration once; a cancel stops every shard. */
typedef struct fb200_sharded_args {
  uint32_t struct_size;
  int32_t n_devices;
  fb200_plan* const* plans;         /* [n_devices] plans with identical FFT settings on distinct devices */
  const fb200_bufnmf_args* job;     /* the whole job, mem == FB200_HOST */
  float* const* gathered_acts;      /* optional [n_devices] device pointers (see above) */
} fb200_sharded_args;
FB200_API int32_t fb200_bufnmf_sharded(const fb200_sharded_args* args);

/* ---- the per-frame body of NMFFilter / NMFMatch for frames a host-side BufferedProcess has cut ------------------------ */
/* For hosts that keep the reference's streaming structure (FluidSource / FluidSink ring buffers on the audio thread's
 * side, clients/common/BufferedProcess.hpp:49-93) and hand the frames that fall due to the device in one batch:
 * in [frames][win] are raw time-domain frames exactly as FluidSource::pull delivers them (FluidSource.hpp:68-89); every
 * frame goes through STFT::processFrame (window, rFFT; STFT.hpp:110-118), STFT::magnitude, NMF::processFrame (NMF.hpp:45-89)
 * and, when `out` is given, NMF::estimate + RatioMask::process per component and ISTFT::processFrame (STFT.hpp:201-208):
 * out [frames][rank][win] are the windowed time-domain frames a FluidSink overlap-adds (the Normalise channel window^2 is
 * host-side data, BufferedProcess.hpp:219-224).  acts_out [frames][rank] are the activations (NMFMatch). */
typedef struct fb200_filter_frames_args {
  uint32_t struct_size;
  int32_t mem;
  int64_t frames;
  int32_t rank, iterations;     /* NMFMatch hard-codes 10 (NMFMatchClient.hpp:115) */
  int64_t seed;                 /* >= 0: the same h0 for every frame; < 0: fresh draws per frame (statistically equivalent) */
  const float* in;              /* [frames][win] */
  const float* bases;           /* [rank][bins] */
  float* out;                   /* optional [frames][rank][win] */
  float* acts_out;              /* optional [frames][rank] */
} fb200_filter_frames_args;
FB200_API int32_t fb200_nmf_filter_frames(fb200_plan* plan, const fb200_filter_frames_args* args);

/* ---- BufSTFT: BufferSTFTClient::processFwd / processInverse  (clients/nrt/BufSTFTClient.hpp:82-190, 192-279) ----- */
/* Size rules of the client.  padding = FFTParams::padding (clients/common/ParameterTypes.hpp:315-323): mode 0 -> 0,
 * 1 -> win/2, 2 -> win-hop.  Forward (invert = 0, count = samples): padded = count + 2*padding, rounded up to a multiple of
 * hop in mode 2 (:126-128); *out = numHops = 1 + (padded - win)/hop (:130-131).  Inverse (invert = 1, count = frames):
 * *out = (count-1)*hop + win - padding (:241-242).  Returns FB200_ERR_INVALID when the padded input is shorter than
 * one window (the reference would read past its buffer). */
FB200_API int32_t fb200_bufstft_sizes(int32_t win, int32_t hop, int32_t padding_mode, int32_t invert, int64_t count,
                                      int64_t* padding, int64_t* out);
typedef struct fb200_bufstft_args {
  uint32_t struct_size;
  int32_t mem;                  /* float32 buffers, like BufferAdaptor */
  int32_t invert;               /* 0 forward: audio -> mag and/or phase; 1 inverse: mag + phase -> resynth */
  int32_t padding_mode;         /* 0 / 1 / 2 as above (the client's default is 1) */
  int64_t batch;                /* independent mono buffers per call (the client does one) */
  int64_t n_samples;            /* forward: samples per buffer */
  int64_t frames;               /* frames per buffer in mag / phase: forward = numHops, inverse = given */
  const float* audio;           /* forward in  [batch][n_samples] */
  float* mag;                   /* forward out (may be NULL) / inverse in: [batch][frames][bins]  (:171-175 / :236-239) */
  float* phase;                 /* forward out (may be NULL) / inverse in: [batch][frames][bins], radians (:177-181) */
  float* resynth;               /* inverse out [batch][(frames-1)*hop + win - padding] (:243-275) */
} fb200_bufstft_args;
FB200_API int32_t fb200_bufstft(fb200_plan* plan, const fb200_bufstft_args* args);

/* ---- instrumentation (bench.py roofline) ------------------------------------------------------------------- */
typedef struct fb200_stats {
  float ms_h2d, ms_stft, ms_init, ms_nmf, ms_post, ms_resynth, ms_d2h, ms_total; /* CUDA-event times of the last call */
  int64_t launches_total;   /* kernels of this library launched by the last call (cuFFT launches counted as 1 per exec) */
  int64_t launches_nmf;     /* ... of which NMF update-loop kernels */
  int32_t backend_used;     /* fb200_backend actually run */
  int32_t update_kernel_launches; /* launches of the dominant NMF update kernel (tile kernel) in the last call */
  float ms_update_kernel;   /* sum of their device durations, each bracketed by CUDA events on the plan's stream */
  float reserved;
} fb200_stats;
FB200_API int32_t fb200_get_stats(const fb200_plan* plan, fb200_stats* out);

/* ---- diagnostics: exercises the tcgen05/TMEM/TMA building blocks of the tensor-core engine on small known matrices.
 * in  (host float): H1[128x16] W1[16x64] R1[128x64] W2[16x128] H2[64x16] R2[128x64] V[128x68]   (n_in  = 31488)
 * out (host float): H1*W1 [128x64], R1*W1^T [128x16], W2^T*H2^T [128x64], R2*H2 [128x16], V[:,32:64] [128x32] (n_out = 24576) */
FB200_API int32_t fb200_selftest_tcgen05(fb200_plan* plan, const float* in, int64_t n_in, float* out, int64_t n_out);

/* ---- dlsym'd function table (north_star: "one .so, dlsym'd function table") -------------------------------- */
typedef struct fb200_api {
  uint32_t abi_version;
  uint32_t struct_size;
  int32_t (*device_count)(void);
  int32_t (*plan_create)(const fb200_config*, fb200_plan**);
  void (*plan_destroy)(fb200_plan*);
  const char* (*last_error)(const fb200_plan*);
  int64_t (*num_frames)(int64_t, int32_t, int32_t);
  int32_t (*resolve_fft)(int32_t, int32_t, int32_t, int32_t*, int32_t*, int32_t*);
  void (*shard_range)(int64_t, int32_t, int32_t, int64_t*, int64_t*);
  int32_t (*stft)(fb200_plan*, const void*, int64_t, int64_t, void*, void*, int32_t, int32_t);
  int32_t (*istft)(fb200_plan*, const void*, int64_t, int64_t, void*, int64_t, int32_t, int32_t);
  int32_t (*nmf_process)(fb200_plan*, const fb200_nmf_args*);
  int32_t (*nmf_process_frames)(fb200_plan*, const fb200_frames_args*);
  int32_t (*bufnmf)(fb200_plan*, const fb200_bufnmf_args*);
  int32_t (*nmf_filter)(fb200_plan*, const fb200_filter_args*);
  int32_t (*get_stats)(const fb200_plan*, fb200_stats*);
  int32_t (*bufstft_sizes)(int32_t, int32_t, int32_t, int32_t, int64_t, int64_t*, int64_t*);
  int32_t (*bufstft)(fb200_plan*, const fb200_bufstft_args*);
  int32_t (*nmf_filter_frames)(fb200_plan*, const fb200_filter_frames_args*);
  int32_t (*bufnmf_sharded)(const fb200_sharded_args*);
  int32_t (*bufnmfcross)(fb200_plan*, const fb200_nmfcross_args*);
  int32_t (*melbands)(fb200_plan*, const fb200_melbands_args*);
  int32_t (*hpss)(fb200_plan*, const fb200_hpss_args*);
  int32_t (*nmfseed)(fb200_plan*, const fb200_nmfseed_args*);
} fb200_api;
/* returns NULL when abi_version is not supported */
FB200_API const fb200_api* fb200_get_api(uint32_t abi_version);

#ifdef __cplusplus
}
#endif
#endif /* FLUCOMA_B200_H */
