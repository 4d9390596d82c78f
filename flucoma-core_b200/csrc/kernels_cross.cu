// BufNMFCross kernels (algorithms/public/NMFCross.hpp:60-185, GriffinLim.hpp:29-54): the activation-only KL update with
// the SOURCE SPECTROGRAM as dictionary (rank = number of source frames, hundreds to thousands), its three per-iteration
// post-processing steps on H, the complex resynthesis H * S and the Griffin-Lim phase iteration.
// With rank in the hundreds the two products of the update are ordinary GEMMs; they run as a 64 x 64 x 16 register-tiled
// fp32 kernel with the element-wise part of the update fused into the epilogue (fp32 operands keep the discontinuous
// selection steps -- local maxima, top-p -- on the same side of their thresholds as the fp64 reference).
#include "common.cuh"

namespace fb200 {

// ---------------------------------------------------------------------------------------------------------------
// C[M][N] = epilogue(A[M][K] * B), B given as [K][N] (TRANSB = 0) or as [N][K] (TRANSB = 1).  All row-major, fp32.
//   EPI 0: C = acc                      (synthesis H * S)
//   EPI 1: C = X / max(acc, eps)        (ratio V / max(W H, eps), NMFCross.hpp:167-168)
//   EPI 2: C = X * acc / d[n]           (H <- H * hnum / max(hden, eps), :170; d already clamped)
// ---------------------------------------------------------------------------------------------------------------
template <int TRANSB, int EPI>
__global__ void __launch_bounds__(256) k_sgemm(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C,
                                               int M, int N, int K, const float* __restrict__ X, const float* __restrict__ dvec)
{
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    // A tile: 64 rows x 16 k  (thread: row = tid / 4, 4 consecutive k)
    {
      const int r = tid >> 2, kk = (tid & 3) * 4;
      const int m = m0 + r;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int k = k0 + kk + q;
        As[kk + q][r] = (m < M && k < K) ? A[(int64_t) m * K + k] : 0.f;
      }
    }
    if (TRANSB) { // B as [N][K]: 64 n-rows x 16 k
      const int r = tid >> 2, kk = (tid & 3) * 4;
      const int n = n0 + r;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int k = k0 + kk + q;
        Bs[kk + q][r] = (n < N && k < K) ? Bm[(int64_t) n * K + k] : 0.f;
      }
    } else { // B as [K][N]: 16 k-rows x 64 n
      const int kk = tid >> 4, c = (tid & 15) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int n = n0 + c + q;
        Bs[kk][c + q] = (k < K && n < N) ? Bm[(int64_t) k * N + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const int64_t e = (int64_t) m * N + n;
      float v = acc[i][j];
      if (EPI == 1) v = X[e] / fmaxf(v, kEps);
      if (EPI == 2) v = X[e] * v / dvec[n];
      C[e] = v;
    }
  }
}

template <int TRANSB, int EPI>
static void sgemm(Plan* p, const float* A, const float* B, float* C, int M, int N, int K, const float* X, const float* d)
{
  dim3 grid((unsigned) ((N + 63) / 64), (unsigned) ((M + 63) / 64));
  k_sgemm<TRANSB, EPI><<<grid, 256, 0, p->stream>>>(A, B, C, M, N, K, X, d);
  p->launches++;
}
void launch_sgemm_nn(Plan* p, const float* A, const float* B, float* C, int M, int N, int K) { sgemm<0, 0>(p, A, B, C, M, N, K, nullptr, nullptr); }
void launch_cross_ratio(Plan* p, const float* H, const float* W, const float* V, float* ratio, int F, int B, int R)
{ // ratio[F][B] = V / max(H[F][R] * W[R][B], eps)
  sgemm<0, 1>(p, H, W, ratio, F, B, R, V, nullptr);
}
void launch_cross_update(Plan* p, const float* ratio, const float* W, const float* H_in, float* H_out, const float* hden, int F, int B, int R)
{ // H_out[F][R] = H_in * (ratio[F][B] * W[R][B]^T) / hden[r]
  sgemm<1, 2>(p, ratio, W, H_out, F, R, B, H_in, hden);
  p->launches_nmf++;
}

// W <- max(W, eps); energy[r] = sum_b W^2 (:159); hden[r] = max(sum_b W, eps) (:169-170).  One warp per source frame.
__global__ void __launch_bounds__(256) k_cross_prepare(float* __restrict__ W, int R, int B, float* __restrict__ energy, float* __restrict__ hden)
{
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  float e = 0.f, s = 0.f;
  for (int b = lane; b < B; b += 32) {
    const float w = fmaxf(W[(int64_t) r * B + b], kEps);
    W[(int64_t) r * B + b] = w;
    e = fmaf(w, w, e);
    s += w;
  }
  for (int o = 16; o; o >>= 1) { e += __shfl_xor_sync(0xffffffffu, e, o); s += __shfl_xor_sync(0xffffffffu, s, o); }
  if (lane == 0) { energy[r] = e; hden[r] = fmaxf(s, kEps); }
}
void launch_cross_prepare(Plan* p, float* W, int R, int B, float* energy, float* hden)
{
  k_cross_prepare<<<(R + 7) / 8, 256, 0, p->stream>>>(W, R, B, energy, hden);
  p->launches++;
}

// enforceTemporalSparseness (:104-127): H[f][k] keeps its value iff the FIRST maximum of the zero-padded window
// H[f-half .. f-half+size)[k] is the centre; otherwise it is multiplied by `factor`.
__global__ void __launch_bounds__(256) k_cross_sparseness(const float* __restrict__ H, float* __restrict__ out, int F, int R, int size, float factor)
{
  const int64_t total = (int64_t) F * R;
  const int half = (size - 1) / 2;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    const int k = (int) (e % R), f = (int) (e / R);
    int arg = 0;
    float best = -INFINITY;
    for (int t = 0; t < size; t++) {
      const int ff = f + t - half;
      const float v = (ff >= 0 && ff < F) ? H[(int64_t) ff * R + k] : 0.f;
      if (v > best) { best = v; arg = t; }
    }
    out[e] = arg != half ? H[e] * factor : H[e];
  }
}
// restrictPolyphony (:130-143): per frame the p components with the largest H * energy keep their value (ties: the lower
// index first), all others are multiplied by `factor`.  One CTA per frame, p rounds of a block-wide arg-max.
__global__ void __launch_bounds__(256) k_cross_polyphony(const float* __restrict__ H, float* __restrict__ out, int R, const float* __restrict__ energy,
                                                         int p, float factor)
{
  extern __shared__ float sc[]; // [R] scores, then 2 x 8 reduction slots
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ int winner;
  const int f = blockIdx.x, tid = threadIdx.x;
  const float* h = H + (int64_t) f * R;
  float* o = out + (int64_t) f * R;
  for (int k = tid; k < R; k += 256) { sc[k] = h[k] * energy[k]; o[k] = h[k] * factor; }
  __syncthreads();
  for (int round = 0; round < p; round++) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = tid; k < R; k += 256) {
      const float v = sc[k];
      if (v > bv || (v == bv && k < bi)) { bv = v; bi = k; }
    }
    for (int off = 16; off; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { rv[tid >> 5] = bv; ri[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
      float v = rv[0];
      int i = ri[0];
      for (int w = 1; w < 8; w++)
        if (rv[w] > v || (rv[w] == v && ri[w] < i)) { v = rv[w]; i = ri[w]; }
      winner = i;
      if (i < R) { o[i] = h[i]; sc[i] = -INFINITY; }
    }
    __syncthreads();
    (void) winner;
  }
}
// promoteContinuity (:86-102): out[f][k] = sum_d H[f + d - half][k + d - half] (zero padded): a diagonal box filter.
__global__ void __launch_bounds__(256) k_cross_continuity(const float* __restrict__ H, float* __restrict__ out, int F, int R, int size)
{
  const int64_t total = (int64_t) F * R;
  const int half = (size - 1) / 2;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    const int k = (int) (e % R), f = (int) (e / R);
    float s = 0.f;
    for (int d = 0; d < size; d++) {
      const int kk = k + d - half, ff = f + d - half;
      if (kk >= 0 && kk < R && ff >= 0 && ff < F) s += H[(int64_t) ff * R + kk];
    }
    out[e] = s;
  }
}
void launch_cross_sparseness(Plan* p, const float* H, float* out, int F, int R, int size, float factor)
{
  int grid = (int) std::min<int64_t>(((int64_t) F * R + 255) / 256, (int64_t) p->sm_count * 16);
  k_cross_sparseness<<<grid, 256, 0, p->stream>>>(H, out, F, R, size, factor);
  p->launches++;
}
void launch_cross_polyphony(Plan* p, const float* H, float* out, int F, int R, const float* energy, int poly, float factor)
{
  k_cross_polyphony<<<F, 256, sizeof(float) * (size_t) R, p->stream>>>(H, out, R, energy, poly, factor);
  p->launches++;
}
void launch_cross_continuity(Plan* p, const float* H, float* out, int F, int R, int size)
{
  int grid = (int) std::min<int64_t>(((int64_t) F * R + 255) / 256, (int64_t) p->sm_count * 16);
  k_cross_continuity<<<grid, 256, 0, p->stream>>>(H, out, F, R, size);
  p->launches++;
}

// ---- Griffin-Lim (GriffinLim.hpp:29-54) -------------------------------------------------------------------------
// mag = |spec|, phase(f, b) = polar(1, 2 pi u[b * F + f])  (EigenRandomPhase fills a column-major F x B array, :146-160)
__global__ void __launch_bounds__(256) k_gl_init(const float2* __restrict__ spec, const float* __restrict__ U, int F, int B, float* __restrict__ mag,
                                                 float2* __restrict__ phase)
{
  const int64_t total = (int64_t) F * B;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    const int b = (int) (e % B), f = (int) (e / B);
    const float2 s = spec[e];
    mag[e] = hypotf(s.x, s.y);
    const double th = (double) U[(int64_t) b * F + f] * 6.283185307179586476925;
    phase[e] = make_float2((float) cos(th), (float) sin(th));
  }
}
// out = mag * phase
__global__ void __launch_bounds__(256) k_gl_apply(const float* __restrict__ mag, const float2* __restrict__ phase, int64_t total, float2* __restrict__ out)
{
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    const float m = mag[e];
    const float2 ph = phase[e];
    out[e] = make_float2(m * ph.x, m * ph.y);
  }
}
// phase = est - momentum / (1 + momentum) * prev; phase /= |phase| + eps  (:49-50)
__global__ void __launch_bounds__(256) k_gl_phase(const float2* __restrict__ est, const float2* __restrict__ prev, int64_t total, float2* __restrict__ phase)
{
  const float c = 0.9f / 1.9f;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    const float2 a = est[e], b = prev[e];
    const float re = a.x - c * b.x, im = a.y - c * b.y;
    const float n = hypotf(re, im) + kEps;
    phase[e] = make_float2(re / n, im / n);
  }
}
static int flat_grid(Plan* p, int64_t total) { return (int) std::min<int64_t>((total + 255) / 256, (int64_t) p->sm_count * 16); }
void launch_gl_init(Plan* p, const float2* spec, const float* U, int F, int B, float* mag, float2* phase)
{
  k_gl_init<<<flat_grid(p, (int64_t) F * B), 256, 0, p->stream>>>(spec, U, F, B, mag, phase);
  p->launches++;
}
void launch_gl_apply(Plan* p, const float* mag, const float2* phase, int64_t total, float2* out)
{
  k_gl_apply<<<flat_grid(p, total), 256, 0, p->stream>>>(mag, phase, total, out);
  p->launches++;
}
void launch_gl_phase(Plan* p, const float2* est, const float2* prev, int64_t total, float2* phase)
{
  k_gl_phase<<<flat_grid(p, total), 256, 0, p->stream>>>(est, prev, total, phase);
  p->launches++;
}

} // namespace fb200
