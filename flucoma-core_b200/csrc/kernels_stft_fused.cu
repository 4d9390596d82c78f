// Fused forward and inverse STFT.
// Forward: audio -> (window, real FFT, |X|) in ONE kernel, no frame materialisation, no complex round trip.
//   STFT::process + STFT::magnitude, algorithms/public/STFT.hpp:90-108, 61-66; FFT conventions algorithms/util/FFT.hpp:92-108.
// A CTA takes a run of consecutive frames of one buffer.  The samples they cover -- (frames - 1) * hop + win of them,
// each shared by win / hop frames -- are brought into shared memory ONCE by TMA (cp.async.bulk.tensor over a 2-D map
// [batch][n]; the zero padding of STFT.hpp:92-97 is the tensor map's out-of-bounds fill, also for negative sample
// indices), the Hann window is applied on the way into the transform, a real FFT of `fft` points runs per frame in shared
// memory (complex Stockham FFT of fft / 2 points, radix 8 with a radix 2 / 4 tail, one butterfly per thread and stage,
// followed by the real-input split), and |X| goes straight into the padded [Fp][Bp] layout the NMF engines read (and,
// when resynthesis needs it, the complex spectrum into [F][B]).  HBM sees hop samples in and B magnitudes out per frame:
// the algorithmic 4 * hop + 4 * B bytes, where the cuFFT pipeline (k_frame_window -> cuFFT R2C -> k_magnitude)
// moved six times that (profiles/r01e_stft_kernels.csv).  cuFFT remains the fallback for sizes this kernel does not take.
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda.h>

namespace fb200 {
using namespace tc;

namespace stf {
__device__ __forceinline__ int pad(int i) { return i + (i >> 5); } // one float2 of padding per 32: conflict-free strided stores
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); } // a * (-i)

// forward DFTs in natural output order
__device__ __forceinline__ void dft2(float2& a, float2& b)
{
  const float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3)
{
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = mul_mi(csub(a1, a3));
  a0 = cadd(s02, s13); a2 = csub(s02, s13); a1 = cadd(d02, d13); a3 = csub(d02, d13);
}
__device__ __forceinline__ void dft8(float2 (&v)[8])
{
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4(e0, e1, e2, e3);
  dft4(o0, o1, o2, o3);
  const float h = 0.70710678118654752440f;
  o1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));   // * (1 - i) / sqrt 2
  o2 = mul_mi(o2);                                          // * (-i)
  o3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));  // * (-1 - i) / sqrt 2
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}
} // namespace stf
using namespace stf;

// twiddles of a plan: tw[k] = exp(-2 pi i k / NC), k < NC, then sp[k] = exp(-2 pi i k / (2 NC)), k <= NC; fp64 evaluation
__global__ void k_stft_twiddles(float2* __restrict__ tw, int NC)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < NC) {
    double s, c;
    sincospi(-2.0 * (double) i / (double) NC, &s, &c);
    tw[i] = make_float2((float) c, (float) s);
  }
  if (i <= NC) {
    double s, c;
    sincospi(-(double) i / (double) NC, &s, &c);
    tw[NC + i] = make_float2((float) c, (float) s);
  }
}

// NC = fft / 2 (complex transform length); threads per frame = NC / 8, frames per round G = 2048 / NC, 256 threads.
template <int NC>
__global__ void __launch_bounds__(256, 3) k_stft_fused(const __grid_constant__ CUtensorMap amap, const float* __restrict__ window,
                                                    const float2* __restrict__ tw, int win, int hop, int half, int n, int F, int fpb,
                                                    int tile_floats, float* __restrict__ V, int Fp, int Bp, float2* __restrict__ spec)
{
  constexpr int TPF = NC / 8, G = 256 / TPF, NP = NC + NC / 32 + 1;
  extern __shared__ __align__(128) uint8_t smem[];
  float* tile = reinterpret_cast<float*>(smem);                         // [tile_floats]
  float2* bufA = reinterpret_cast<float2*>(tile + tile_floats);         // [G][NP]
  float2* bufB = bufA + G * NP;                                         // [G][NP]
  uint64_t* bar = reinterpret_cast<uint64_t*>(bufB + G * NP);
  const int tid = threadIdx.x, g = tid / TPF, t = tid % TPF;
  const int buf = blockIdx.y, f_tile = blockIdx.x * fpb;
  const int nf = min(fpb, F - f_tile);
  // first sample of the tile; the copy starts at the 16-byte boundary below it (the copy engine faults on a box whose
  // first element is not 16-byte aligned in global memory), `delta` samples earlier
  const int xs = f_tile * hop - half;
  const int delta = ((xs % 4) + 4) % 4;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  { // the tile's samples: 256-wide boxes; the tensor map zero-fills the parts of a box outside [0, n).  A box that lies
    // ENTIRELY outside is not requested (the copy engine faulted on those): the threads zero it themselves.
    const int need = (nf - 1) * hop + win + delta;
    const int nbox = (need + 255) / 256;
    const int x0 = xs - delta;
    if (tid == 0) {
      int live_boxes = 0;
      for (int b = 0; b < nbox; b++) live_boxes += (x0 + 256 * b < n && x0 + 256 * b + 256 > 0) ? 1 : 0;
      if (live_boxes) mbar_arrive_expect_tx(bar, (uint32_t) live_boxes * 1024u);
      else mbar_arrive(bar);
      for (int b = 0; b < nbox; b++)
        if (x0 + 256 * b < n && x0 + 256 * b + 256 > 0)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                           smem_u32(tile + 256 * b)),
                       "l"((uint64_t) &amap), "r"(x0 + 256 * b), "r"(buf), "r"(smem_u32(bar))
                       : "memory");
    }
    for (int b = 0; b < nbox; b++)
      if (!(x0 + 256 * b < n && x0 + 256 * b + 256 > 0)) tile[256 * b + tid] = 0.f;
  }
  mbar_wait(bar, 0);
  __syncthreads();

  float2* A = bufA + g * NP;
  float2* Bf = bufB + g * NP;
  const float2* sp = tw + NC;
  // everything that depends on the thread but not on the frame lives in registers: the thread's 16 window samples and
  // the twiddles of its butterfly in the second and third radix-8 stage (the L1 / shared-memory data pipe was the
  // limiter of the first version: 77 % busy, a third of it re-loading these tables for every frame)
  float2 wv[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int j = 2 * (t + r * TPF);
    wv[r].x = j < win ? window[j] : 0.f;
    wv[r].y = j + 1 < win ? window[j + 1] : 0.f;
  }
  constexpr int NST = (NC >= 512) ? 2 : 1; // radix-8 stages after the first (NC = 128, 256: one; 512 .. 2048: two)
  float2 twr[NST][7];
  {
    int Ns = 8;
#pragma unroll
    for (int s = 0; s < NST; s++, Ns *= 8) {
      const int k = t & (Ns - 1), tstep = NC / (Ns * 8);
#pragma unroll
      for (int r = 1; r < 8; r++) twr[s][r - 1] = tw[r * k * tstep];
    }
  }
  for (int f0 = 0; f0 < nf; f0 += G) {
    const int fl = f0 + g;
    const bool live = fl < nf;
    // ---- window + pack: z[m] = (x[2m] w[2m], x[2m+1] w[2m+1]); thread t holds m = t + r * TPF: butterfly t of stage 1
    float2 v[8];
    if (live) {
      const float* x = tile + delta + fl * hop;
      if (((delta + fl * hop) & 1) == 0) { // 8-byte aligned frame start: one LDS.64 per complex input
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const float2 xx = *reinterpret_cast<const float2*>(x + 2 * (t + r * TPF));
          v[r] = make_float2(xx.x * wv[r].x, xx.y * wv[r].y);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const int j = 2 * (t + r * TPF);
          v[r] = make_float2(x[j] * wv[r].x, x[j + 1] * wv[r].y);
        }
      }
      dft8(v); // stage 1: Ns = 1, no twiddles; outputs to j0 + r with j0 = 8 t
#pragma unroll
      for (int r = 0; r < 8; r++) A[pad(8 * t + r)] = v[r];
    }
    __syncthreads();
    float2* in = A;
    float2* out = Bf;
    int Ns = 8;
#pragma unroll
    for (int s = 0; s < NST; s++, Ns *= 8) { // further radix-8 stages
      if (live) {
        const int k = t & (Ns - 1);
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = in[pad(t + r * TPF)];
#pragma unroll
        for (int r = 1; r < 8; r++) v[r] = cmul(v[r], twr[s][r - 1]);
        dft8(v);
        const int j0 = (t - k) * 8 + k;
#pragma unroll
        for (int r = 0; r < 8; r++) out[pad(j0 + r * Ns)] = v[r];
      }
      __syncthreads();
      float2* sw = in; in = out; out = sw;
    }
    constexpr int LOG = NC == 128 ? 7 : (NC == 256 ? 8 : (NC == 512 ? 9 : (NC == 1024 ? 10 : 11)));
    constexpr int TAIL = 1 << (LOG % 3); // 1: none, 2, 4
    if (TAIL > 1) {
      if (live) {
        constexpr int NB = NC / TAIL; // butterflies of the tail stage
        for (int j = t; j < NB; j += TPF) {
          const int k = j & (Ns - 1); // Ns * TAIL == NC: twiddle step 1
          float2 u[4];
#pragma unroll
          for (int r = 0; r < TAIL; r++) u[r] = in[pad(j + r * NB)];
#pragma unroll
          for (int r = 1; r < TAIL; r++) u[r] = cmul(u[r], tw[r * k]);
          if (TAIL == 2) dft2(u[0], u[1]);
          else dft4(u[0], u[1], u[2], u[3]);
          const int j0 = (j - k) * TAIL + k;
#pragma unroll
          for (int r = 0; r < TAIL; r++) out[pad(j0 + r * Ns)] = u[r];
        }
      }
      __syncthreads();
      float2* sw = in; in = out; out = sw;
    }
    // ---- real-input split: X[k] = a - i q with a = (Z[k] + conj Z[N-k]) / 2, q = e^{-2 pi i k / fft} (Z[k] - conj Z[N-k]) / 2,
    //      and from the same pair X[N-k] = conj(a) - i conj(q): k = 0 .. NC / 2 yields all NC + 1 bins
    if (live) {
      const int f = f_tile + fl;
      float* vrow = V ? V + ((int64_t) buf * Fp + f) * Bp : nullptr;
      float2* srow = spec ? spec + ((int64_t) buf * F + f) * (NC + 1) : nullptr;
      for (int k = t; k <= NC / 2; k += TPF) {
        const float2 z = in[pad(k)];
        float2 zc = in[pad((NC - k) & (NC - 1))];
        zc.y = -zc.y;
        const float2 a = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y + zc.y));
        const float2 b = make_float2(0.5f * (z.x - zc.x), 0.5f * (z.y - zc.y));
        const float2 q = cmul(sp[k], b);
        float2 X0 = make_float2(a.x + q.y, a.y - q.x);   // bin k
        float2 X1 = make_float2(a.x - q.y, -a.y - q.x);  // bin NC - k
        if (k == 0) { X0.y = 0.f; X1.y = 0.f; } // exact by construction up to rounding; FFT.hpp:99-101 leaves them at zero
        if (srow) srow[k] = X0;
        if (vrow) vrow[k] = sqrtf(fmaf(X0.x, X0.x, X0.y * X0.y)); // hypotf's range scaling is not needed for audio magnitudes and cost 9 % of the kernel
        if (2 * k != NC) {
          if (srow) srow[NC - k] = X1;
          if (vrow) vrow[NC - k] = sqrtf(fmaf(X1.x, X1.x, X1.y * X1.y));
        }
      }
    }
    __syncthreads(); // the buffers are reused by the next round
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused inverse STFT: complex spectra -> (inverse real FFT, window, overlap-add, / sum w^2) in ONE kernel.
//   ISTFT::process, algorithms/public/STFT.hpp:178-199 (and the streaming normalisation of BufferedProcess.hpp:231-237).
// A CTA owns a range of S = fpb * hop padded output positions of one signal and transforms every frame that reaches into
// it (its fpb frames plus the ceil(win / hop) - 1 frames before them: recomputed, not exchanged -- no atomics, fixed
// summation order, bitwise repeatable).  Per frame: the spectrum row (B = fft / 2 + 1 complex values, the only HBM read)
// goes to shared memory, the inverse of the real-input split builds the fft / 2 complex points whose inverse transform
// is the frame (x[2m] + i x[2m+1]), the inverse runs as conj(FFT(conj Z)) on the forward Stockham stages above, the
// windowed samples of the round's frames are parked in shared memory and gathered, frame by frame in ascending order,
// into the CTA's output tile.  The cuFFT pipeline this replaces wrote and re-read an fft-sample frame per spectrum row
// (8 KB at fft 1024) and needed a cuFFT plan per (frames x signals) shape -- tens of milliseconds on first use.
// MASK: the rows are not read from `spec` as they are but as component k of buffer b of a BufNMF resynthesis,
//   X_k[f][b] = X[f][b] * min(1, H[f][k] W[k][b] / max(sum_j H[f][j] W[j][b], eps))      (NMF::estimate + RatioMask, exponent 1)
// with signal index = buffer * K + k; 1 / max(sum, eps) comes precomputed per (frame, bin) from k_inv_estimate, so the masked
// spectra (K complex spectrograms per buffer) are never written.
struct InvMask {
  const float* W;    // [buffers | 1][KP][Bp]
  const float* H;    // [buffers][Fp][KP]
  const float* invV; // [nb][F][B], buffers b0 .. b0 + nb
  int64_t b0;
  int K, KP, Bp, Fp, shared_w;
};

// invV[bl][f][bin] = 1 / max(sum_k H[f][k] W[k][bin], eps), the same fmaf chain as k_mask
template <int KMAX>
__global__ void __launch_bounds__(256) k_inv_estimate(InvMask mk, int F, int B, float* __restrict__ invV)
{
  extern __shared__ float hs8[]; // [8][KP]
  const int KP = mk.KP, K = mk.K;
  const int64_t bl = blockIdx.y, buf = mk.b0 + bl;
  const int f0 = blockIdx.x * 8;
  const float* __restrict__ W = mk.W + (mk.shared_w ? (int64_t) 0 : buf * KP * mk.Bp);
  const float* __restrict__ H = mk.H + (buf * mk.Fp + f0) * KP;
  const int nf = min(8, F - f0);
  for (int e = threadIdx.x; e < 8 * KP; e += 256) hs8[e] = e < nf * KP ? H[e] : 0.f;
  __syncthreads();
  for (int bin = threadIdx.x; bin < B; bin += 256) {
    float w[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) w[k] = k < K ? W[(int64_t) k * mk.Bp + bin] : 0.f;
    for (int i = 0; i < nf; i++) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; k++) v = fmaf(hs8[i * KP + (k < KP ? k : 0)], w[k], v);
      invV[(bl * F + f0 + i) * B + bin] = 1.0f / fmaxf(v, kEps);
    }
  }
}

template <int NC, bool MASK>
__global__ void __launch_bounds__(256, 3) k_istft_fused(const float2* __restrict__ spec, const float* __restrict__ window,
                                                     const float2* __restrict__ tw, int win, int hop, int half, int64_t n, int F, int fpb,
                                                     int nct, float* __restrict__ out, int64_t out_stride, int stream_norm, InvMask mk)
{
  constexpr int TPF = NC / 8, G = 256 / TPF, NP = NC + NC / 32 + 1;
  extern __shared__ __align__(128) uint8_t smem[];
  const int S = fpb * hop;
  float* acc = reinterpret_cast<float*>(smem);                          // [S] output tile (padded positions p0 .. p0 + S)
  float* ys = acc + S;                                                  // [G][win] windowed frames of the current round
  float2* bufA = reinterpret_cast<float2*>(ys + G * win);               // [G][NP]
  float2* bufB = bufA + G * NP;                                         // [G][NP]
  const int tid = threadIdx.x, g = tid / TPF, t = tid % TPF;
  const int c = blockIdx.x % nct;
  const int64_t sig = blockIdx.x / nct;
  const int p0 = c * S; // all positions fit 31 bits (istft_fused_eligible)
  const int i_hi = min(F - 1, (p0 + S - 1) / hop);
  const int i_lo = p0 - win + 1 <= 0 ? 0 : (p0 - win + hop) / hop; // ceil((p0 - win + 1) / hop)
  for (int i = tid; i < S; i += 256) acc[i] = 0.f;
  float2* A = bufA + g * NP;
  float2* Bf = bufB + g * NP;
  const float2* sp = tw + NC;
  const float scale = 1.0f / (float) (2 * NC);
  constexpr int NST = (NC >= 512) ? 2 : 1;
  float2 twr[NST][7];
  {
    int Ns = 8;
#pragma unroll
    for (int s = 0; s < NST; s++, Ns *= 8) {
      const int k = t & (Ns - 1), tstep = NC / (Ns * 8);
#pragma unroll
      for (int r = 1; r < 8; r++) twr[s][r - 1] = tw[r * k * tstep];
    }
  }
  __syncthreads();
  for (int i0 = i_lo; i0 <= i_hi; i0 += G) {
    const int fi = i0 + g;
    const bool live = fi <= i_hi;
    if (live) { // the spectrum row of this frame -> shared memory (all loads of a thread in flight together)
      float2 x[8];
      float2 xl = make_float2(0.f, 0.f);
      if constexpr (MASK) {
        const int64_t bl = sig / mk.K, buf = mk.b0 + bl;
        const int k = (int) (sig - bl * mk.K);
        const float2* row = spec + (buf * F + fi) * (int64_t) (NC + 1);
        const float* wrow = mk.W + (mk.shared_w ? (int64_t) 0 : buf * mk.KP * mk.Bp) + (int64_t) k * mk.Bp;
        const float* irow = mk.invV + (bl * F + fi) * (int64_t) (NC + 1);
        const float h = mk.H[(buf * mk.Fp + fi) * mk.KP + k];
        float wv[8], iv[8];
#pragma unroll
        for (int r = 0; r < 8; r++) { x[r] = row[t + r * TPF]; wv[r] = wrow[t + r * TPF]; iv[r] = irow[t + r * TPF]; }
        float wl = 0.f, il = 0.f;
        if (t == 0) { xl = row[NC]; wl = wrow[NC]; il = irow[NC]; }
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const float m = fminf(1.0f, h * wv[r] * iv[r]);
          x[r] = make_float2(x[r].x * m, x[r].y * m);
        }
        const float ml = fminf(1.0f, h * wl * il);
        xl = make_float2(xl.x * ml, xl.y * ml);
      } else {
        const float2* row = spec + (sig * F + fi) * (int64_t) (NC + 1);
#pragma unroll
        for (int r = 0; r < 8; r++) x[r] = row[t + r * TPF];
        if (t == 0) xl = row[NC];
      }
      if (t == 0) { x[0].y = 0.f; xl.y = 0.f; } // a real signal's DC and Nyquist bins are real: the C2R convention ignores what is there
#pragma unroll
      for (int r = 0; r < 8; r++) Bf[pad(t + r * TPF)] = x[r];
      if (t == 0) Bf[pad(NC)] = xl;
    }
    __syncthreads();
    float2 v[8];
    if (live) {
      // inverse of the real-input split (cf. the forward kernel): with a = X[m] + conj X[NC-m], b = X[m] - conj X[NC-m],
      // Z[m] = a + i e^{+2 pi i m / fft} b is the spectrum of z[j] = x[2j] + i x[2j+1] scaled by fft (the unnormalised C2R
      // convention); the inverse transform is taken as conj(FFT(conj Z))
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int m = t + r * TPF;
        const float2 xk = Bf[pad(m)];
        float2 xn = Bf[pad(NC - m)];
        xn.y = -xn.y;
        const float2 a = cadd(xk, xn), b = csub(xk, xn);
        float2 e = sp[m];
        e.y = -e.y;                               // e^{+2 pi i m / fft}
        const float2 eb = cmul(e, b);
        const float2 z = make_float2(a.x - eb.y, a.y + eb.x); // a + i (e b)
        v[r] = make_float2(z.x, -z.y);
      }
      dft8(v);
#pragma unroll
      for (int r = 0; r < 8; r++) A[pad(8 * t + r)] = v[r];
    }
    __syncthreads(); // also: every thread has read the spectrum row out of Bf before the next stage writes there
    float2* in = A;
    float2* outb = Bf;
    int Ns = 8;
#pragma unroll
    for (int s = 0; s < NST; s++, Ns *= 8) {
      if (live) {
        const int k = t & (Ns - 1);
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = in[pad(t + r * TPF)];
#pragma unroll
        for (int r = 1; r < 8; r++) v[r] = cmul(v[r], twr[s][r - 1]);
        dft8(v);
        const int j0 = (t - k) * 8 + k;
#pragma unroll
        for (int r = 0; r < 8; r++) outb[pad(j0 + r * Ns)] = v[r];
      }
      __syncthreads();
      float2* sw = in; in = outb; outb = sw;
    }
    constexpr int LOG = NC == 128 ? 7 : (NC == 256 ? 8 : (NC == 512 ? 9 : (NC == 1024 ? 10 : 11)));
    constexpr int TAIL = 1 << (LOG % 3);
    if (TAIL > 1) {
      if (live) {
        constexpr int NB = NC / TAIL;
        for (int j = t; j < NB; j += TPF) {
          const int k = j & (Ns - 1);
          float2 u[4];
#pragma unroll
          for (int r = 0; r < TAIL; r++) u[r] = in[pad(j + r * NB)];
#pragma unroll
          for (int r = 1; r < TAIL; r++) u[r] = cmul(u[r], tw[r * k]);
          if (TAIL == 2) dft2(u[0], u[1]);
          else dft4(u[0], u[1], u[2], u[3]);
          const int j0 = (j - k) * TAIL + k;
#pragma unroll
          for (int r = 0; r < TAIL; r++) outb[pad(j0 + r * Ns)] = u[r];
        }
      }
      __syncthreads();
      float2* sw = in; in = outb; outb = sw;
    }
    // ---- window: x[2m] = Re w[m], x[2m+1] = -Im w[m]; the frame's first `win` samples * window / fft (STFT.hpp:189-193)
    if (live) {
      float* y = ys + g * win;
      for (int m = t; 2 * m < win; m += TPF) {
        const float2 w = in[pad(m)];
        const int j = 2 * m;
        y[j] = w.x * scale * window[j];
        if (j + 1 < win) y[j + 1] = -w.y * scale * window[j + 1];
      }
    }
    __syncthreads();
    { // ---- gather the round's frames into the tile, ascending frame order per position
      const int gl = min(G, i_hi - i0 + 1);
      const int r0 = i0 * hop; // first position of the round's first frame
      const int lo = max(p0, r0), hi = min(p0 + S, r0 + (gl - 1) * hop + win);
      for (int pp = lo + tid; pp < hi; pp += 256) {
        const int rel = pp - r0; // >= 0
        const int g_hi = min(gl - 1, rel / hop);
        const int g_lo = rel - win + 1 <= 0 ? 0 : (rel - win + hop) / hop;
        float a = acc[pp - p0];
        for (int gg = g_lo; gg <= g_hi; gg++) a += ys[gg * win + rel - gg * hop];
        acc[pp - p0] = a;
      }
    }
    __syncthreads();
  }
  // ---- normalise and write the owned samples: t = pos - half in [0, n)
  for (int i = tid; i < S; i += 256) {
    const int pos = p0 + i, ts = pos - half;
    if (ts < 0 || ts >= n) continue;
    const int f_hi = min(F - 1, pos / hop);
    const int f_lo = pos - win + 1 <= 0 ? 0 : (pos - win + hop) / hop;
    float nrm = 0.f;
    for (int f = f_lo; f <= f_hi; f++) {
      const float w = window[pos - f * hop];
      nrm = fmaf(w, w, nrm);
    }
    const float a = acc[i];
    out[sig * out_stride + ts] = stream_norm ? (a != 0.f ? a / (nrm > 0.f ? nrm : 1.f) : a) : a / fmaxf(nrm, kEps);
  }
}

// true when the fused inverse takes this call (else cuFFT C2R + k_ola)
bool istft_fused_eligible(const Plan* p, int64_t nsig, int64_t F, int64_t n, int64_t half)
{
  const int fft = p->fft;
  if (fft != 256 && fft != 512 && fft != 1024 && fft != 2048 && fft != 4096) return false;
  if (p->hop > p->win || nsig <= 0 || F <= 0 || n <= 0 || half < 0) return false;
  if (n + half >= ((int64_t) 1 << 31) || F >= ((int64_t) 1 << 30)) return false;
  return getenv("FB200_ISTFT_CUFFT") == nullptr;
}

template <int NC>
static int32_t launch_inv_t(Plan* p, const float2* spec, int64_t nsig, int64_t F, int64_t n, float* out, int64_t half, int64_t out_stride,
                            int stream_norm, const InvMask* mk)
{
  constexpr int TPF = NC / 8, G = 256 / TPF, NP = NC + NC / 32 + 1;
  // positions per CTA: up to 6144 (24 KB of accumulators: with the 80 registers of the kernel three CTAs share an SM; 8192 with
  // two CTAs measured 12 % slower), at least one hop;
  // the frames before the range are recomputed
  static const int s_max = getenv("FB200_ISTFT_S") ? atoi(getenv("FB200_ISTFT_S")) : 6144;
  int fpb = std::max(1, s_max / p->hop);
  const int64_t span = n + half; // padded positions that produce output
  fpb = (int) std::min<int64_t>(fpb, (span + p->hop - 1) / p->hop);
  const int S = fpb * p->hop;
  const int nct = (int) ((span + S - 1) / S);
  const size_t smem = sizeof(float) * (size_t) (S + G * p->win) + 2 * sizeof(float2) * (size_t) (G * NP);
  if (smem > 200 * 1024) return FB200_ERR_UNSUPPORTED;
  if ((int64_t) nct * nsig >= ((int64_t) 1 << 31)) return FB200_ERR_UNSUPPORTED;
  if (mk) {
    FB_CUDA(p, cudaFuncSetAttribute(k_istft_fused<NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); // per device
    k_istft_fused<NC, true><<<(unsigned) (nct * nsig), 256, smem, p->stream>>>(spec, p->window.as<float>(), p->twiddle.as<float2>(), p->win, p->hop,
                                                                               (int) half, n, (int) F, fpb, nct, out, out_stride, stream_norm, *mk);
  } else {
    FB_CUDA(p, cudaFuncSetAttribute(k_istft_fused<NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); // per device
    k_istft_fused<NC, false><<<(unsigned) (nct * nsig), 256, smem, p->stream>>>(spec, p->window.as<float>(), p->twiddle.as<float2>(), p->win, p->hop,
                                                                                (int) half, n, (int) F, fpb, nct, out, out_stride, stream_norm, InvMask{});
  }
  p->launches++;
  FB_CUDA(p, cudaGetLastError());
  return FB200_OK;
}

static int32_t ensure_twiddles(Plan* p)
{
  const int NC = p->fft / 2;
  if (!p->twiddle.p) {
    FB_CUDA(p, p->twiddle.ensure(sizeof(float2) * (size_t) (2 * NC + 1)));
    k_stft_twiddles<<<(NC + 256) / 256, 256, 0, p->stream>>>(p->twiddle.as<float2>(), NC);
    p->launches++;
  }
  return FB200_OK;
}

static int32_t launch_inv(Plan* p, const float2* spec, int64_t nsig, int64_t F, int64_t n, float* out, int64_t half, int64_t out_stride,
                          int stream_norm, const InvMask* mk)
{
  FB_TRY(ensure_twiddles(p));
  switch (p->fft / 2) {
  case 128: return launch_inv_t<128>(p, spec, nsig, F, n, out, half, out_stride, stream_norm, mk);
  case 256: return launch_inv_t<256>(p, spec, nsig, F, n, out, half, out_stride, stream_norm, mk);
  case 512: return launch_inv_t<512>(p, spec, nsig, F, n, out, half, out_stride, stream_norm, mk);
  case 1024: return launch_inv_t<1024>(p, spec, nsig, F, n, out, half, out_stride, stream_norm, mk);
  default: return launch_inv_t<2048>(p, spec, nsig, F, n, out, half, out_stride, stream_norm, mk);
  }
}

int32_t launch_istft_fused(Plan* p, const float2* spec, int64_t nsig, int64_t F, int64_t n, float* out, int64_t half, int64_t out_stride,
                           int stream_norm)
{
  return launch_inv(p, spec, nsig, F, n, out, half, out_stride, stream_norm, nullptr);
}

// BufNMF resynthesis of buffers [b0, b0 + nb): rank masked spectra per buffer -> out[(bl * K + k) * n ..], without writing them.
// spec = the unmasked spectra of ALL buffers [batch][F][B]; scratch = nb * F * B floats.
int32_t launch_masked_istft_fused(Plan* p, const NmfDev& d, const float2* spec, int64_t b0, int64_t nb, int64_t n, float* out, int64_t half,
                                  float* scratch)
{
  InvMask mk{};
  mk.W = d.W; mk.H = d.H; mk.invV = scratch; mk.b0 = b0;
  mk.K = d.K; mk.KP = d.KP; mk.Bp = d.Bp; mk.Fp = d.Fp; mk.shared_w = d.shared_w;
  dim3 grid((unsigned) ((d.F + 7) / 8), (unsigned) nb);
  const size_t sm = sizeof(float) * 8 * d.KP;
  if (d.KP <= 4) k_inv_estimate<4><<<grid, 256, sm, p->stream>>>(mk, d.F, d.B, scratch);
  else if (d.KP <= 8) k_inv_estimate<8><<<grid, 256, sm, p->stream>>>(mk, d.F, d.B, scratch);
  else if (d.KP <= 16) k_inv_estimate<16><<<grid, 256, sm, p->stream>>>(mk, d.F, d.B, scratch);
  else if (d.KP <= 32) k_inv_estimate<32><<<grid, 256, sm, p->stream>>>(mk, d.F, d.B, scratch);
  else k_inv_estimate<64><<<grid, 256, sm, p->stream>>>(mk, d.F, d.B, scratch);
  p->launches++;
  return launch_inv(p, spec, nb * d.K, d.F, n, out, half, n, 0, &mk);
}

static int32_t make_audio_tensor_map(Plan* p, CUtensorMap* tmap, const float* audio, int64_t n, int64_t batch)
{
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      p->err = "cuTensorMapEncodeTiled not available";
      return FB200_ERR_CUDA;
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  cuuint64_t dims[2] = {(cuuint64_t) n, (cuuint64_t) batch};
  cuuint64_t strides[1] = {(cuuint64_t) n * 4};
  cuuint32_t box[2] = {256, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(audio), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    p->err = "cuTensorMapEncodeTiled (audio) failed: " + std::to_string((int) r);
    return FB200_ERR_CUDA;
  }
  return FB200_OK;
}

// true when the fused kernel takes this call (else the caller runs the cuFFT pipeline)
bool stft_fused_eligible(const Plan* p, const float* audio, int64_t n, int64_t batch, int hop)
{
  const int fft = p->fft;
  if (fft != 256 && fft != 512 && fft != 1024 && fft != 2048 && fft != 4096) return false;
  if (n < 4 || (n % 4) != 0 || (reinterpret_cast<uintptr_t>(audio) % 16) != 0) return false; // TMA: 16-byte rows
  if (batch > 65535 || n >= ((int64_t) 1 << 31)) return false;
  if (hop > p->win) return false; // gaps between frames: nothing to share, and the tile bound below assumes overlap
  return getenv("FB200_STFT_CUFFT") == nullptr;
}

template <int NC>
static int32_t launch_t(Plan* p, const CUtensorMap& amap, int hop, int64_t half, int64_t n, int64_t batch, int64_t F, float* V,
                        int64_t Fp, int64_t Bp, float2* spec)
{
  constexpr int TPF = NC / 8, G = 256 / TPF, NP = NC + NC / 32 + 1;
  // frames per CTA: as many as keep the sample tile within ~32 KB, a multiple of the G frames transformed per round
  int fpb = (8192 - p->win) / hop + 1;
  fpb = std::max(G, std::min(64, fpb) / G * G);
  const int tile_floats = (((fpb - 1) * hop + p->win + 3 + 255) / 256) * 256; // + 3: alignment slack, see `delta`
  const size_t smem = sizeof(float) * (size_t) tile_floats + 2 * sizeof(float2) * (size_t) (G * NP) + 16;
  if (smem > 200 * 1024) return FB200_ERR_UNSUPPORTED;
  FB_CUDA(p, cudaFuncSetAttribute(k_stft_fused<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); // per device
  dim3 grid((unsigned) ((F + fpb - 1) / fpb), (unsigned) batch);
  k_stft_fused<NC><<<grid, 256, smem, p->stream>>>(amap, p->window.as<float>(), p->twiddle.as<float2>(), p->win, hop, (int) half, (int) n, (int) F,
                                                   fpb, tile_floats, V, (int) Fp, (int) Bp, spec);
  p->launches++;
  FB_CUDA(p, cudaGetLastError());
  return FB200_OK;
}

int32_t launch_stft_fused(Plan* p, const float* audio, int64_t batch, int64_t n, int64_t F, float* V, int64_t Fp, int64_t Bp,
                          float2* spec, int64_t half, int hop)
{
  const int NC = p->fft / 2;
  FB_TRY(ensure_twiddles(p));
  alignas(64) CUtensorMap amap;
  FB_TRY(make_audio_tensor_map(p, &amap, audio, n, batch));
  switch (NC) {
  case 128: return launch_t<128>(p, amap, hop, half, n, batch, F, V, Fp, Bp, spec);
  case 256: return launch_t<256>(p, amap, hop, half, n, batch, F, V, Fp, Bp, spec);
  case 512: return launch_t<512>(p, amap, hop, half, n, batch, F, V, Fp, Bp, spec);
  case 1024: return launch_t<1024>(p, amap, hop, half, n, batch, F, V, Fp, Bp, spec);
  default: return launch_t<2048>(p, amap, hop, half, n, batch, F, V, Fp, Bp, spec);
  }
}

} // namespace fb200
