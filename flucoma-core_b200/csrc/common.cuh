// Internal declarations shared by the translation units of libflucoma_b200.so.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/flucoma_b200.h"

namespace fb200 {

constexpr float kEps = 2.220446049250313e-16f; // algorithms/util/AlgorithmUtils.hpp:19 (DBL_EPSILON), rounded to fp32

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int rank_pad(int k)
{ // compile-time rank variants of the update kernels; pad rows of W / columns of H are exact zeros
  int kp = 4;
  while (kp < k) kp <<= 1;
  return kp;
}

// growable device allocation
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  // grow without losing the first `keep` bytes
  cudaError_t ensure_keep(size_t bytes, size_t keep)
  {
    if (bytes <= cap) return cudaSuccess;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) return e;
    if (p && keep) e = cudaMemcpy(q, p, keep < cap ? keep : cap, cudaMemcpyDeviceToDevice);
    if (p) cudaFree(p);
    p = q; cap = bytes;
    return e;
  }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostBuf { // pinned staging
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release()
  {
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
  }
};

// Device-side problem descriptor of one batched NMF factorisation (all fp32, zero padded).
//   V [batch][Fp][Bp]   magnitudes, frames-major (== the reference's X[F][B] and its column-major B x F "V")
//   W [batch][KP][Bp]   dictionary rows  (== column-major B x K "W" == output W1[K][B])
//   H [batch][Fp][KP]   activations     (== column-major K x F "H" == output H1[F][K])
struct NmfDev {
  float* V; float* W; float* H;
  float* hden;      // [batch][KP]       sum_b W[k][b]      (NMF.hpp:169)
  float* wnum_part; // [batch][ctas][KP][Bp] per-CTA partials of (V/WH) H^T  (NMF.hpp:159)
  float* wden_part; // [batch][ctas][KP]     per-CTA partials of sum_f H[f][k] (NMF.hpp:160)
  int* ticket;      // [batch] zero between launches: the last tile CTA of a buffer runs the W finalisation (null: separate kernel)
  int batch, F, B, K;
  int Fp, Bp, KP;
  int ctas_per_buf;   // grid.x of the tile kernel
  int tiles_per_cta;  // consecutive 128-frame tiles handled by one CTA
  int clamp_v;        // processFrame clamps the input magnitudes at eps (NMF.hpp:60)
  int shared_w;       // 1: a single W (batch stride 0) shared by all buffers (processFrame path)
  int op_first, op_total; // streamed engine: this NmfDev is buffers [op_first, op_first + batch) of a call of op_total buffers
                          // (split calls: the operand arrays are sized for the whole call, each part uses its own slice); 0, 0 = whole call
};

struct Plan {
  fb200_config cfg{};
  int win = 0, hop = 0, fft = 0, bins = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr; // host -> device uploads that overlap with kernels on `stream`
  cudaStream_t poll_stream = nullptr; // the asynchronous progress mode reads the device counters here (never behind an upload)
  std::vector<cudaEvent_t> cev;       // one event per upload wave (+ one fence)
  cudaEvent_t ev[10]{};
  cudaEvent_t ev_async = nullptr;     // completion of the persistent launch the asynchronous progress mode polls
  std::string err;
  fb200_stats stats{};
  int64_t launches = 0, launches_nmf = 0;
  int sm_count = 148;
  std::vector<cudaEvent_t> kev; // event pairs around every update-kernel launch of the current call
  size_t kev_used = 0;
  int backend_used = FB200_BACKEND_SIMT; // engine that ran the update loop of the current call
  uint32_t attr_mask = 0; // which k_nmf_tile<KP> variants already have their dynamic-smem attribute set

  DevBuf window;   // float[win]
  DevBuf audio;    // float [batch][n]
  DevBuf stage;    // raw upload/download staging (caller dtype)
  DevBuf frames;   // float [wave*F][fft]  cuFFT real side
  DevBuf spec;     // float2 [batch][F][B] (kept when resynthesis is requested) or [wave][F][B]
  DevBuf cspec;    // float2 [wave][K][F][B] masked component spectra
  DevBuf V, W, H, hden, wnum_part, wden_part, ticket, rnd, seeds, scale, out_a, out_b;
  HostBuf pin_a, pin_b;
  HostBuf ctrl;    // pinned read-back of the progress counters of the asynchronous progress mode
  DevBuf ctrl_dev; // device control words: [0] cancel request, [1 + cta] finished (buffer, pass) units
  DevBuf twiddle;  // float2 [fft/2 + fft/2 + 1] twiddles of the fused STFT kernel
  DevBuf x0, x1, x2, x3, x4, x5; // BufNMFCross / Griffin-Lim work arrays
  DevBuf wop_buf, hop_buf; // split-bf16 operand copies of W / H for the streamed tensor-core engine
  std::map<std::pair<int, int64_t>, cufftHandle> fft_plans; // (type, batch) -> handle, bounded LRU
  std::map<std::pair<int, int64_t>, uint64_t> fft_lru;
  uint64_t fft_tick = 0;

  bool fail(int code, const std::string& msg) { err = msg; last_code = code; return false; }
  int last_code = 0;
};

// error plumbing ------------------------------------------------------------------------------------------------
#define FB_CUDA(plan, expr)                                                                                  \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess) {                                                                                 \
      (plan)->err = std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr;                     \
      return FB200_ERR_CUDA;                                                                                 \
    }                                                                                                        \
  } while (0)

#define FB_CUFFT(plan, expr)                                                                                 \
  do {                                                                                                       \
    cufftResult _r = (expr);                                                                                 \
    if (_r != CUFFT_SUCCESS) {                                                                               \
      (plan)->err = std::string("cuFFT error ") + std::to_string((int) _r) + " at " #expr;                   \
      return FB200_ERR_CUFFT;                                                                                \
    }                                                                                                        \
  } while (0)

#define FB_TRY(expr)                                                                                         \
  do {                                                                                                       \
    int32_t _s = (expr);                                                                                     \
    if (_s < 0) return _s;                                                                                   \
  } while (0)

// kernels_util.cu -----------------------------------------------------------------------------------------------
// dst[b][i][j] = scale_b * src[b][i][j] for i < rows, j < cols; element strides given per array. Converts dtype.
void launch_copy3d(Plan* p, const void* src, int src_dtype, int64_t s_b, int64_t s_r, void* dst, int dst_dtype,
                   int64_t d_b, int64_t d_r, int64_t batch, int64_t rows, int64_t cols, const float* scale_per_batch,
                   int clamp_eps);
// U[b][i] = i-th draw of uniform[0,1) from mt19937_64(seeds[b])  (EigenRandom.hpp:73-110 on libstdc++)
void launch_mt_uniform(Plan* p, const int64_t* d_seeds, int64_t batch, int64_t count, float* U);
// W/H initialisation: random or seeded, eps clamp, W row / H column L2 normalisation, hden (NMF.hpp:101-124,150-153)
// U_w / U_h: uniform draws for W and for H (the same array unless the two streams are seeded independently).
// frame_mode 0: NMF::process; 1: processFrame, every frame starts from h0 = U_h[0..K); 2: processFrame, frame f from U_h[f*K..]
void launch_nmf_init(Plan* p, const NmfDev& d, const float* U_w, const float* U_h, int64_t u_stride, const float* W0,
                     const float* H0, int frame_mode);
// per-buffer max over the real H entries -> scale[b] = 1/max  (NMFClient.hpp:289-291)
void launch_h_max_scale(Plan* p, const NmfDev& d, float* scale);

// kernels_nmf_simt.cu -------------------------------------------------------------------------------------------
int32_t simt_configure(Plan* p, NmfDev& d); // chooses tiling, sizes partial buffers
void simt_launch_tile(Plan* p, const NmfDev& d, int do_h, int do_w, int h_iters);
void simt_launch_w_finalize(Plan* p, const NmfDev& d);
// Vhat[b][f][bin] = sum_k H W  -> dst (dtype, dense [batch][F][B])
void launch_vhat(Plan* p, const NmfDev& d, void* dst, int dst_dtype);

// kernels_stft.cu -----------------------------------------------------------------------------------------------
void launch_hann(Plan* p);
void launch_frame_window(Plan* p, const float* audio, int64_t n, int64_t nbuf, int64_t F, float* frames, int64_t half,
                         int hop_override = 0);
// y [K][F][fft] (C2R output) -> out [F][K][win]: first `win` samples * window / fft  (ISTFT::processFrame, STFT.hpp:201-208)
void launch_window_frames(Plan* p, const float* y, int64_t K, int64_t F, float* out);
// exact zeros in the pads of V[batch][Fp][Bp] (bins >= B, frames >= F) when the interior is about to be overwritten
void launch_zero_pads(Plan* p, float* V, int64_t batch, int64_t F, int64_t Fp, int64_t B, int64_t Bp);
// spec [nbuf*F][B] complex -> V[nbuf][Fp][Bp] magnitudes (+ zero imag of DC/Nyquist in place, FFT.hpp:99-101)
void launch_magnitude(Plan* p, float2* spec, int64_t nbuf, int64_t F, float* V, int64_t Fp, int64_t Bp);
// phase[e] = arg(spec[e]) (STFT.hpp:75-87); spec = polar(mag, phase) (BufSTFTClient.hpp:236-239)
void launch_phase(Plan* p, const float2* spec, int64_t count, float* phase);
void launch_polar(Plan* p, const float* mag, const float* phase, int64_t rows, float2* spec);
// masked component spectra for buffers [b0, b0+nb): cspec[nb][K][F][B]  (NMF.hpp:33-42 + RatioMask.hpp:33-57)
void launch_mask(Plan* p, const NmfDev& d, const float2* spec, int64_t b0, int64_t nb, float2* cspec);
// overlap-add + normalise + trim (STFT.hpp:178-199): y [nsig][F][fft] -> out [nsig][n]
void launch_ola(Plan* p, const float* y, int64_t nsig, int64_t F, int64_t n, float* out, int64_t half, int64_t out_stride,
                int stream_norm);

// kernels_stft_fused.cu -----------------------------------------------------------------------------------------
bool stft_fused_eligible(const Plan* p, const float* audio, int64_t n, int64_t batch, int hop);
bool istft_fused_eligible(const Plan* p, int64_t nsig, int64_t F, int64_t n, int64_t half);
int32_t launch_masked_istft_fused(Plan* p, const NmfDev& d, const float2* spec, int64_t b0, int64_t nb, int64_t n, float* out, int64_t half,
                                  float* scratch);
int32_t launch_istft_fused(Plan* p, const float2* spec, int64_t nsig, int64_t F, int64_t n, float* out, int64_t half, int64_t out_stride,
                           int stream_norm);
// audio [batch][n] -> V[batch][Fp][Bp] magnitudes (may be null) and/or spec [batch][F][B] (may be null); one kernel
int32_t launch_stft_fused(Plan* p, const float* audio, int64_t batch, int64_t n, int64_t F, float* V, int64_t Fp, int64_t Bp,
                          float2* spec, int64_t half, int hop);

// kernels_seed.cu (NNDSVD.hpp:30-131) ---------------------------------------------------------------------------------
// one-sided Jacobi SVD of X^T in fp64: leaves A (rotated rows) in x1, J (accumulated rotations) in x2, row norms in x3
int32_t run_seed_svd(Plan* p, const float* mags, int F, int B, int* n_out);
void launch_seed_build(Plan* p, const int* d_ord, int k, int F, int B, int n, int max_rank, int method, float* W, float* H);
void launch_seed_fill(Plan* p, float* W, float* H, int F, int B, int max_rank, int method, const float* mags, const float* U);

// kernels_spectral.cu (MelBands.hpp:35-101, HPSS.hpp:47-162) ------------------------------------------------------
void launch_melbands(Plan* p, const float* mags, const float* filt, int64_t frames, int B, int nb, float scale1, float scale2, int flags,
                     float* bands);
void launch_hpss(Plan* p, const float2* spec, float* mag, int64_t batch, int F, int B, int vsize, int hsize, int mode, const float* th_h,
                 const float* th_p, float2* out);

// kernels_cross.cu (BufNMFCross: NMFCross.hpp:60-185, GriffinLim.hpp:29-54) -----------------------------------------
void launch_sgemm_nn(Plan* p, const float* A, const float* B, float* C, int M, int N, int K); // C[M][N] = A[M][K] B[K][N]
void launch_cross_prepare(Plan* p, float* W, int R, int B, float* energy, float* hden);
void launch_cross_ratio(Plan* p, const float* H, const float* W, const float* V, float* ratio, int F, int B, int R);
void launch_cross_update(Plan* p, const float* ratio, const float* W, const float* H_in, float* H_out, const float* hden, int F, int B, int R);
void launch_cross_sparseness(Plan* p, const float* H, float* out, int F, int R, int size, float factor);
void launch_cross_polyphony(Plan* p, const float* H, float* out, int F, int R, const float* energy, int poly, float factor);
void launch_cross_continuity(Plan* p, const float* H, float* out, int F, int R, int size);
void launch_gl_init(Plan* p, const float2* spec, const float* U, int F, int B, float* mag, float2* phase);
void launch_gl_apply(Plan* p, const float* mag, const float2* phase, int64_t total, float2* out);
void launch_gl_phase(Plan* p, const float2* est, const float2* prev, int64_t total, float2* phase);

// kernels_tc_selftest.cu ---------------------------------------------------------------------------------------
int32_t make_v_tensor_map(Plan* p, void* tmap_out, const float* V, int64_t Bp, int64_t Fp, int64_t batch, int box_rows);
int32_t run_tc_selftest(Plan* p, const float* in, float* out);
int32_t run_tc_mma_timing(Plan* p, long long* d_out, int reps);

// kernels_nmf_tc.cu ---------------------------------------------------------------------------------------------
bool tc_eligible(const NmfDev& d);
// ctrl != nullptr: device control words (ctrl[0] cancel request, ctrl[1 + cta] finished passes), see Ctl in the .cu
int32_t tc_run(Plan* p, const NmfDev& d, int iters, bool upd_w, bool upd_h, unsigned int* ctrl = nullptr,
               const unsigned int* host_cancel = nullptr);
int tc_grid(const Plan* p, const NmfDev& d);

// kernels_nmf_tcs.cu --------------------------------------------------------------------------------------------
bool tcs_eligible(const NmfDev& d);
int32_t tcs_run(Plan* p, const NmfDev& d, int iters, bool upd_w, bool upd_h);

} // namespace fb200
