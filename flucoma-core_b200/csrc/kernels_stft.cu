// STFT / ISTFT side kernels: Hann table, frame gather + window, |X|, ratio masks, overlap-add.
// cuFFT (R2C / C2R, batched) sits between them and is the only library call on the path.
#include "common.cuh"

namespace fb200 {

// WindowFuncs.hpp:41-45 -- periodic Hann evaluated in fp64, rounded once to fp32
__global__ void k_hann(float* __restrict__ w, int win)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < win) w[i] = (float) (0.5 - 0.5 * cos((3.14159265358979323846 * 2.0 * (double) i) / (double) win));
}

void launch_hann(Plan* p)
{
  k_hann<<<(p->win + 255) / 256, 256, 0, p->stream>>>(p->window.as<float>(), p->win);
  p->launches++;
}

// STFT.hpp:92-105: frame i of buffer b = padded[i*hop : i*hop+win] * window, padded = [win/2 zeros | audio | zeros];
// written zero-extended to `fft` samples (FFT.hpp:97: htl::rfft zero-pads a short input).
// `half` is the left padding: win/2 for STFT::process, win for the streaming clients (BufferedProcess.hpp:75-93).
// VEC: hop, half, n and win are multiples of 4 and the audio is 16-byte aligned, so a frame quad inside the signal is
// one 16-byte load (the window quad always is); the 64-bit index split is done once per quad.
template <bool VEC>
__global__ void __launch_bounds__(256) k_frame_window(const float* __restrict__ audio, int64_t n, int64_t nbuf, int64_t F,
                                                      const float* __restrict__ window, int win, int fft, int hop,
                                                      int64_t half, float* __restrict__ frames)
{
  const int q4 = fft / 4;
  const int64_t total = nbuf * F * (int64_t) q4;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    const int j4 = (int) (e % q4);
    const int64_t fr = e / q4;
    const int64_t i = fr % F, b = fr / F;
    const float* a = audio + b * n;
    const int j0 = 4 * j4;
    const int64_t t0 = i * hop + j0 - half;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (VEC && j0 + 3 < win && t0 >= 0 && t0 + 3 < n) {
      const float4 x = *reinterpret_cast<const float4*>(a + t0);
      const float4 w = *reinterpret_cast<const float4*>(window + j0);
      o = make_float4(x.x * w.x, x.y * w.y, x.z * w.z, x.w * w.w);
    } else if (j0 < win) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int j = j0 + q;
        const int64_t t = t0 + q;
        v[q] = (j < win && t >= 0 && t < n) ? a[t] * window[j] : 0.f;
      }
      o = make_float4(v[0], v[1], v[2], v[3]);
    }
    *reinterpret_cast<float4*>(frames + fr * fft + j0) = o;
  }
}

void launch_frame_window(Plan* p, const float* audio, int64_t n, int64_t nbuf, int64_t F, float* frames, int64_t half,
                         int hop_override)
{
  int64_t total = nbuf * F * (int64_t) (p->fft / 4);
  if (total <= 0) return;
  const int hop = hop_override > 0 ? hop_override : p->hop;
  int grid = (int) std::min<int64_t>((total + 255) / 256, (int64_t) p->sm_count * 32);
  const bool vec = (hop % 4 == 0) && (half % 4 == 0) && (n % 4 == 0) && (p->win % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(audio) % 16 == 0);
  if (vec)
    k_frame_window<true><<<grid, 256, 0, p->stream>>>(audio, n, nbuf, F, p->window.as<float>(), p->win, p->fft, hop, half, frames);
  else
    k_frame_window<false><<<grid, 256, 0, p->stream>>>(audio, n, nbuf, F, p->window.as<float>(), p->win, p->fft, hop, half, frames);
  p->launches++;
}

// |X| buffers are padded to [Fp][Bp]; the update kernels rely on exact zeros in the pads (bins >= B, frames >= F).
// Zeroing only the pads replaces a memset of the whole array (1.1 GB at config 2) by a few MB of stores.
__global__ void __launch_bounds__(256) k_zero_pads(float* __restrict__ V, int64_t rows_total, int F, int Fp, int B, int Bp)
{
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t) gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows_total; row += warps) {
    const int f = (int) (row % Fp);
    float* r = V + row * Bp;
    if (f < F) {
      if (lane < Bp - B) r[B + lane] = 0.f;
    } else {
      for (int j = lane; j < Bp; j += 32) r[j] = 0.f;
    }
  }
}

void launch_zero_pads(Plan* p, float* V, int64_t batch, int64_t F, int64_t Fp, int64_t B, int64_t Bp)
{
  const int64_t rows = batch * Fp;
  if (rows <= 0 || (Bp == B && Fp == F)) return;
  if (Bp - B > 32) { // not a layout this library produces; keep it simple and correct
    cudaMemsetAsync(V, 0, sizeof(float) * (size_t) (rows * Bp), p->stream);
    return;
  }
  int grid = (int) std::min<int64_t>((rows + 7) / 8, (int64_t) p->sm_count * 32);
  k_zero_pads<<<grid, 256, 0, p->stream>>>(V, rows, (int) F, (int) Fp, (int) B, (int) Bp);
  p->launches++;
}

// STFT.hpp:61-66 |X|, with Im(DC) = Im(Nyquist) = 0 as FFT.hpp:99-101 leaves them.  spec [nbuf*F][B] -> V [nbuf][Fp][Bp]
// HBM-bound (8 B in, 4 B out per bin): flat, fully coalesced indexing; the bin count is a template parameter for the
// usual FFT sizes so that the per-element index split is a multiply-shift instead of 64-bit divisions (BT = 0: generic).
template <int BT>
__global__ void __launch_bounds__(256) k_magnitude(float2* __restrict__ spec, uint32_t total, uint32_t F, int Brt,
                                                   float* __restrict__ V, uint32_t Fp, uint32_t Bp)
{
  const uint32_t B = BT ? (uint32_t) BT : (uint32_t) Brt;
  auto one = [&](uint32_t e, float2 x) {
    const uint32_t fr = e / B, bin = e - fr * B;
    if (bin == 0 || bin == B - 1) {
      x.y = 0.f;
      spec[e] = x;
    }
    if (V) {
      const uint32_t b = fr / F, f = fr - b * F;
      V[((size_t) b * Fp + f) * Bp + bin] = hypotf(x.x, x.y);
    }
  };
  // two bins (16 bytes) per load and two loads in flight per thread: the kernel is latency bound otherwise
  const uint32_t head = (reinterpret_cast<uintptr_t>(spec) & 8) ? 1u : 0u; // a wave may start on an odd bin: 8-byte aligned only
  const uint32_t pairs = (total - head) >> 1, stride = gridDim.x * blockDim.x;
  const float4* __restrict__ s4 = reinterpret_cast<const float4*>(spec + head);
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < pairs; i += 2 * stride) {
    const float4 a = s4[i], c = s4[i + stride];
    one(head + 2 * i, make_float2(a.x, a.y)); one(head + 2 * i + 1, make_float2(a.z, a.w));
    one(head + 2 * (i + stride), make_float2(c.x, c.y)); one(head + 2 * (i + stride) + 1, make_float2(c.z, c.w));
  }
  if (i < pairs) {
    const float4 a = s4[i];
    one(head + 2 * i, make_float2(a.x, a.y)); one(head + 2 * i + 1, make_float2(a.z, a.w));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (head) one(0, spec[0]);
    if ((total - head) & 1) one(total - 1, spec[total - 1]);
  }
}

static void launch_magnitude_rows(Plan* p, float2* sp, uint32_t total, uint32_t F, float* v, uint32_t Fp, uint32_t Bp)
{
  const int grid = (int) std::min<int64_t>(((int64_t) total / 2 + 255) / 256 + 1, (int64_t) p->sm_count * 16);
  switch (p->bins) {
  case 129: k_magnitude<129><<<grid, 256, 0, p->stream>>>(sp, total, F, p->bins, v, Fp, Bp); break;
  case 257: k_magnitude<257><<<grid, 256, 0, p->stream>>>(sp, total, F, p->bins, v, Fp, Bp); break;
  case 513: k_magnitude<513><<<grid, 256, 0, p->stream>>>(sp, total, F, p->bins, v, Fp, Bp); break;
  case 1025: k_magnitude<1025><<<grid, 256, 0, p->stream>>>(sp, total, F, p->bins, v, Fp, Bp); break;
  case 2049: k_magnitude<2049><<<grid, 256, 0, p->stream>>>(sp, total, F, p->bins, v, Fp, Bp); break;
  default: k_magnitude<0><<<grid, 256, 0, p->stream>>>(sp, total, F, p->bins, v, Fp, Bp); break;
  }
  p->launches++;
}

void launch_magnitude(Plan* p, float2* spec, int64_t nbuf, int64_t F, float* V, int64_t Fp, int64_t Bp)
{
  if (nbuf <= 0 || F <= 0) return;
  const int64_t max_rows = std::max<int64_t>(1, (((int64_t) 1 << 31) - 1) / p->bins); // 32-bit flat index per launch
  if (F <= max_rows) { // whole buffers per launch
    const int64_t per = std::max<int64_t>(1, max_rows / F);
    for (int64_t b0 = 0; b0 < nbuf; b0 += per) {
      const int64_t nb = std::min(per, nbuf - b0);
      launch_magnitude_rows(p, spec + b0 * F * p->bins, (uint32_t) (nb * F * p->bins), (uint32_t) F, V ? V + b0 * Fp * Bp : nullptr,
                            (uint32_t) Fp, (uint32_t) Bp);
    }
  } else { // a single buffer longer than 2^31 bins: row ranges of one buffer at a time
    for (int64_t b = 0; b < nbuf; b++)
      for (int64_t f0 = 0; f0 < F; f0 += max_rows) {
        const int64_t nr = std::min(max_rows, F - f0);
        launch_magnitude_rows(p, spec + (b * F + f0) * p->bins, (uint32_t) (nr * p->bins), (uint32_t) nr,
                              V ? V + (b * Fp + f0) * Bp : nullptr, (uint32_t) nr, (uint32_t) Bp);
      }
  }
}

// STFT::phase (STFT.hpp:75-87): arg of every bin, dense [count].  Im(DC) = Im(Nyquist) = 0 was applied by k_magnitude.
__global__ void __launch_bounds__(256) k_phase(const float2* __restrict__ spec, int64_t count, float* __restrict__ phase)
{
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < count; e += (int64_t) gridDim.x * blockDim.x) {
    const float2 x = spec[e];
    phase[e] = atan2f(x.y, x.x);
  }
}

void launch_phase(Plan* p, const float2* spec, int64_t count, float* phase)
{
  if (count <= 0) return;
  int grid = (int) std::min<int64_t>((count + 255) / 256, (int64_t) p->sm_count * 32);
  k_phase<<<grid, 256, 0, p->stream>>>(spec, count, phase);
  p->launches++;
}

// std::polar(m, p) per bin (BufSTFTClient.hpp:236-239) -> spec [rows][B]; the imaginary parts of DC and Nyquist are
// dropped, as IFFT::process does when it packs the spectrum (FFT.hpp:151-158).
__global__ void __launch_bounds__(256) k_polar(const float* __restrict__ mag, const float* __restrict__ phase, int64_t rows, int B,
                                               float2* __restrict__ spec)
{
  const int64_t count = rows * B;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < count; e += (int64_t) gridDim.x * blockDim.x) {
    const int bin = (int) (e % B);
    float sn, cs;
    sincosf(phase[e], &sn, &cs);
    const float m = mag[e];
    spec[e] = make_float2(m * cs, (bin == 0 || bin == B - 1) ? 0.f : m * sn);
  }
}

void launch_polar(Plan* p, const float* mag, const float* phase, int64_t rows, float2* spec)
{
  if (rows <= 0) return;
  int grid = (int) std::min<int64_t>((rows * p->bins + 255) / 256, (int64_t) p->sm_count * 32);
  k_polar<<<grid, 256, 0, p->stream>>>(mag, phase, rows, p->bins, spec);
  p->launches++;
}

// NMF::estimate (NMF.hpp:33-42) + RatioMask::init/process with exponent 1 (RatioMask.hpp:33-57), all components at once:
//   out_k[f][b] = S[f][b] * min(1, H[f][k] W[k][b] * (1 / max(sum_j H[f][j] W[j][b], eps)))
// spec holds buffers [0, batch); this launch handles [b0, b0+nb) and writes cspec[nb][K][F][B].
template <int KMAX>
__global__ void __launch_bounds__(256) k_mask(NmfDev d, const float2* __restrict__ spec, int64_t b0, int64_t nb,
                                              float2* __restrict__ cspec)
{
  extern __shared__ float hs[]; // [8][KP] activations of the 8 frames of this block
  const int B = d.B, K = d.K, KP = d.KP;
  const int64_t F = d.F;
  int64_t bl = blockIdx.y;           // buffer within the wave
  int64_t buf = b0 + bl;
  int64_t f0 = (int64_t) blockIdx.x * 8;
  const float* __restrict__ W = d.W + (d.shared_w ? (int64_t) 0 : buf * KP * d.Bp);
  const float* __restrict__ H = d.H + (buf * d.Fp + f0) * KP;
  int nf = (int) min((int64_t) 8, F - f0);
  for (int e = threadIdx.x; e < 8 * KP; e += 256) hs[e] = e < nf * KP ? H[e] : 0.f;
  __syncthreads();
  for (int bin = threadIdx.x; bin < B; bin += 256) {
    float w[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) w[k] = k < K ? W[(int64_t) k * d.Bp + bin] : 0.f;
    for (int i = 0; i < nf; i++) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; k++) v = fmaf(hs[i * KP + (k < KP ? k : 0)], w[k], v);
      float mult = 1.0f / fmaxf(v, kEps);
      float2 s = spec[(buf * F + f0 + i) * B + bin];
#pragma unroll
      for (int k = 0; k < KMAX; k++) {
        if (k < K) {
          float m = fminf(1.0f, hs[i * KP + k] * w[k] * mult);
          cspec[((bl * K + k) * F + f0 + i) * B + bin] = make_float2(s.x * m, s.y * m);
        }
      }
    }
  }
}

void launch_mask(Plan* p, const NmfDev& d, const float2* spec, int64_t b0, int64_t nb, float2* cspec)
{
  dim3 grid((unsigned) ((d.F + 7) / 8), (unsigned) nb);
  size_t smem = sizeof(float) * 8 * d.KP;
  if (d.KP <= 4) k_mask<4><<<grid, 256, smem, p->stream>>>(d, spec, b0, nb, cspec);
  else if (d.KP <= 8) k_mask<8><<<grid, 256, smem, p->stream>>>(d, spec, b0, nb, cspec);
  else if (d.KP <= 16) k_mask<16><<<grid, 256, smem, p->stream>>>(d, spec, b0, nb, cspec);
  else if (d.KP <= 32) k_mask<32><<<grid, 256, smem, p->stream>>>(d, spec, b0, nb, cspec);
  else k_mask<64><<<grid, 256, smem, p->stream>>>(d, spec, b0, nb, cspec);
  p->launches++;
}

// ISTFT::process (STFT.hpp:178-199) as a gather: sample t of signal s sums the <= ceil(win/hop) frames covering
// padded position t+half, each * (1/fft) * window, then / max(sum window^2, eps).  y [nsig][F][fft] (cuFFT C2R output).
// stream_norm selects the streaming clients' rule instead (BufferedProcess.hpp:231-237): x != 0 ? x / (g > 0 ? g : 1) : x.
// Signal s writes out[s * out_stride + t], t in [0, n).
__global__ void __launch_bounds__(256) k_ola(const float* __restrict__ y, int64_t nsig, int64_t F, int64_t n,
                                             const float* __restrict__ window, int win, int fft, int hop, int64_t half,
                                             float* __restrict__ out, int64_t out_stride, int stream_norm)
{
  int64_t total = nsig * n;
  const float scale = 1.0f / (float) fft;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    int64_t t = e % n, s = e / n;
    int64_t pos = t + half;
    int64_t i_hi = pos / hop;
    if (i_hi > F - 1) i_hi = F - 1;
    int64_t i_lo = pos - win + 1 <= 0 ? 0 : (pos - win + hop) / hop; // ceil((pos-win+1)/hop)
    float acc = 0.f, nrm = 0.f;
    const float* ys = y + s * F * fft;
    for (int64_t i = i_lo; i <= i_hi; i++) {
      int j = (int) (pos - i * hop);
      float w = window[j];
      acc = fmaf(ys[i * fft + j] * scale, w, acc);
      nrm = fmaf(w, w, nrm);
    }
    float r = stream_norm ? (acc != 0.f ? acc / (nrm > 0.f ? nrm : 1.f) : acc) : acc / fmaxf(nrm, kEps);
    out[s * out_stride + t] = r;
  }
}

// ISTFT::processFrame (STFT.hpp:201-208) for every frame of every component: y [K][F][fft] (cuFFT C2R output) ->
// out [F][K][win] = y[k][f][0:win] * window / fft, the frames a FluidSink overlap-adds (BufferedProcess.hpp:208-216)
__global__ void __launch_bounds__(256) k_window_frames(const float* __restrict__ y, int64_t K, int64_t F,
                                                       const float* __restrict__ window, int win, int fft,
                                                       float* __restrict__ out)
{
  const int64_t total = K * F * win;
  const float scale = 1.0f / (float) fft;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    const int j = (int) (e % win);
    const int64_t fk = e / win;
    const int64_t k = fk % K, f = fk / K;
    out[e] = y[(k * F + f) * fft + j] * scale * window[j];
  }
}

void launch_window_frames(Plan* p, const float* y, int64_t K, int64_t F, float* out)
{
  const int64_t total = K * F * p->win;
  if (total <= 0) return;
  int grid = (int) std::min<int64_t>((total + 255) / 256, (int64_t) p->sm_count * 32);
  k_window_frames<<<grid, 256, 0, p->stream>>>(y, K, F, p->window.as<float>(), p->win, p->fft, out);
  p->launches++;
}

void launch_ola(Plan* p, const float* y, int64_t nsig, int64_t F, int64_t n, float* out, int64_t half, int64_t out_stride,
                int stream_norm)
{
  int64_t total = nsig * n;
  if (total <= 0) return;
  int grid = (int) std::min<int64_t>((total + 255) / 256, (int64_t) p->sm_count * 32);
  k_ola<<<grid, 256, 0, p->stream>>>(y, nsig, F, n, p->window.as<float>(), p->win, p->fft, p->hop, half, out,
                                           out_stride, stream_norm);
  p->launches++;
}

} // namespace fb200
