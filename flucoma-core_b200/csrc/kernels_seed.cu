// NMFSeed / NNDSVD (algorithms/public/NNDSVD.hpp:30-131, clients/nrt/NMFSeedClient.hpp:74-133): SVD-based seeds for W, H.
// The thin SVD of X^T (bins x frames) is a one-sided Jacobi iteration in fp64: the bins' rows of A = X^T are rotated pairwise
// until they are mutually orthogonal; their norms are the singular values, the normalised rows the right vectors (over
// frames), the accumulated rotations the left vectors (over bins).  A round-robin tournament makes the n / 2 pairs of a
// round independent: one CTA per pair, n - 1 rounds per sweep, a handful of sweeps.
#include "common.cuh"

namespace fb200 {

__global__ void __launch_bounds__(256) k_seed_load(const float* __restrict__ mags, int F, int B, int n, double* __restrict__ A, double* __restrict__ J)
{ // A [n][F] = X^T (rows >= B are zero padding of the tournament), J [n][n] = identity
  const int64_t ta = (int64_t) n * F, tj = (int64_t) n * n;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < ta + tj; e += (int64_t) gridDim.x * blockDim.x) {
    if (e < ta) {
      const int b = (int) (e / F), f = (int) (e % F);
      A[e] = b < B ? (double) mags[(int64_t) f * B + b] : 0.0;
    } else {
      const int64_t j = e - ta;
      J[j] = (j / n) == (j % n) ? 1.0 : 0.0;
    }
  }
}

__device__ __forceinline__ double block_sum(double v, double* red)
{
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < 8; w++) s += red[w];
  return s;
}

// one round of the tournament: CTA p rotates rows (i, j) of A and of J
__global__ void __launch_bounds__(256) k_seed_round(double* __restrict__ A, double* __restrict__ J, int F, int n, int round, unsigned long long* off)
{
  __shared__ double red[8];
  const int p = blockIdx.x, m = n - 1;
  int i, j;
  if (p == 0) { i = m; j = round % m; }
  else { i = (round + p) % m; j = (round - p + m) % m; }
  double* x = A + (int64_t) i * F;
  double* y = A + (int64_t) j * F;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int f = threadIdx.x; f < F; f += 256) { const double u = x[f], v = y[f]; a += u * u; b += v * v; c += u * v; }
  a = block_sum(a, red); b = block_sum(b, red); c = block_sum(c, red);
  if (c == 0.0 || a == 0.0 || b == 0.0) return;
  const double rel = fabs(c) / sqrt(a * b);
  if (threadIdx.x == 0) atomicMax(off, (unsigned long long) __double_as_longlong(rel)); // positive doubles order like integers
  if (rel < 1e-15) return;
  const double zeta = (b - a) / (2.0 * c);
  const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
  for (int f = threadIdx.x; f < F; f += 256) { const double u = x[f], v = y[f]; x[f] = cs * u - sn * v; y[f] = sn * u + cs * v; }
  double* ji = J + (int64_t) i * n;
  double* jj = J + (int64_t) j * n;
  for (int k = threadIdx.x; k < n; k += 256) { const double u = ji[k], v = jj[k]; ji[k] = cs * u - sn * v; jj[k] = sn * u + cs * v; }
}

__global__ void __launch_bounds__(256) k_seed_norms(const double* __restrict__ A, int F, double* __restrict__ nrm)
{
  __shared__ double red[8];
  const double* x = A + (int64_t) blockIdx.x * F;
  double a = 0.0;
  for (int f = threadIdx.x; f < F; f += 256) a += x[f] * x[f];
  a = block_sum(a, red);
  if (threadIdx.x == 0) nrm[blockIdx.x] = sqrt(a);
}

// component c (sorted position) = tournament row ord[c]: u = sign * J row (bins), v = sign * A row / s (frames), the sign
// chosen so that the largest-magnitude entry of u is positive.  Writes W[c][B] and H[f][max_rank] column c (fp32).
//   method 0 (:60-65): W = |u|, H = |s v|;  methods 1-3 (:66-104): component 0 W = |u|, H = sqrt(s) |v|, the others the
//   dominant of the positive / negative parts, scaled as the reference writes it (including yNNorm = ||xN||, :84).
__global__ void __launch_bounds__(256) k_seed_build(const double* __restrict__ A, const double* __restrict__ J, const double* __restrict__ nrm,
                                                    const int* __restrict__ ord, int F, int B, int n, int max_rank, int method,
                                                    float* __restrict__ W, float* __restrict__ H)
{
  __shared__ double red[8];
  __shared__ double sh_big;
  __shared__ int sh_idx;
  const int c = blockIdx.x, row = ord[c], tid = threadIdx.x;
  const double* jr = J + (int64_t) row * n;
  const double* ar = A + (int64_t) row * F;
  const double s = nrm[row];
  // sign convention: arg max |u|
  double big = -1.0;
  int idx = 0;
  for (int b = tid; b < B; b += 256) { const double v = fabs(jr[b]); if (v > big) { big = v; idx = b; } }
  for (int o = 16; o; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, big, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ob > big || (ob == big && oi < idx)) { big = ob; idx = oi; }
  }
  __shared__ double wb[8];
  __shared__ int wi[8];
  if ((tid & 31) == 0) { wb[tid >> 5] = big; wi[tid >> 5] = idx; }
  __syncthreads();
  if (tid == 0) {
    double bb = wb[0]; int bi = wi[0];
    for (int w = 1; w < 8; w++) if (wb[w] > bb || (wb[w] == bb && wi[w] < bi)) { bb = wb[w]; bi = wi[w]; }
    sh_big = bb; sh_idx = bi;
  }
  __syncthreads();
  (void) sh_big;
  const double sign = jr[sh_idx] < 0 ? -1.0 : 1.0;
  const double inv_s = s > 0 ? 1.0 / s : 0.0;
  if (method == 0 || c == 0) {
    const double hs = method == 0 ? s : sqrt(s);
    for (int b = tid; b < B; b += 256) W[(int64_t) c * B + b] = (float) fabs(jr[b]);
    for (int f = tid; f < F; f += 256) H[(int64_t) f * max_rank + c] = (float) (hs * fabs(ar[f] * inv_s));
    return;
  }
  double xp = 0.0, xn = 0.0, yp = 0.0;
  for (int b = tid; b < B; b += 256) { const double x = sign * jr[b]; if (x > 0) xp += x * x; else xn += x * x; }
  for (int f = tid; f < F; f += 256) { const double y = sign * ar[f] * inv_s; if (y > 0) yp += y * y; }
  xp = block_sum(xp, red); xn = block_sum(xn, red); yp = block_sum(yp, red);
  const double xPNorm = sqrt(xp), yPNorm = sqrt(yp), xNNorm = sqrt(xn), yNNorm = xNNorm; // :84 as written
  const double mP = xPNorm * yPNorm, mN = xNNorm * yNNorm;
  const bool pos = mP > mN;
  const double sigma = pos ? mP : mN, lbd = sqrt(s * sigma);
  const double xd = pos ? xPNorm : xNNorm, yd = pos ? yPNorm : yNNorm;
  for (int b = tid; b < B; b += 256) {
    const double x = sign * jr[b];
    const double part = pos ? (x > 0 ? x : 0.0) : (x < 0 ? -x : 0.0);
    W[(int64_t) c * B + b] = (float) (part / xd);
  }
  for (int f = tid; f < F; f += 256) {
    const double y = sign * ar[f] * inv_s;
    const double part = pos ? (y > 0 ? y : 0.0) : (y < 0 ? -y : 0.0);
    H[(int64_t) f * max_rank + c] = (float) (lbd * (part / yd));
  }
}

// methods 1 / 2 (:105-125): entries below eps are replaced by uniform draws in [eps, mean / 1000) (ar) or by the mean (a).
// U_w / U_h: uniform(0,1) draws; W^T is filled column-major (draw j * B + b), H^T column-major (draw f * max_rank + j).
__global__ void __launch_bounds__(256) k_seed_fill(float* __restrict__ W, float* __restrict__ H, int F, int B, int max_rank, int method, const double* mean_p,
                                                   const float* __restrict__ U)
{
  const double mean = *mean_p;
  const double lo = 2.220446049250313e-16, hi = mean * 0.001;
  const int64_t tw = (int64_t) max_rank * B, th = (int64_t) F * max_rank;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < tw + th; e += (int64_t) gridDim.x * blockDim.x) {
    float* dst = e < tw ? W + e : H + (e - tw);          // W[j][b]: e = j * B + b == the draw index; H[f][j]: e' = f * max_rank + j likewise
    if ((double) *dst < lo) {
      const double u = method == 1 ? (double) U[e < tw ? e : e - tw] * (hi - lo) + lo : mean;
      *dst = (float) u;
    }
  }
}

__global__ void __launch_bounds__(256) k_seed_mean(const float* __restrict__ mags, int64_t count, double* __restrict__ out)
{ // single CTA, fixed order: repeatable
  __shared__ double red[8];
  double a = 0.0;
  for (int64_t e = threadIdx.x; e < count; e += 256) a += (double) mags[e];
  a = block_sum(a, red);
  if (threadIdx.x == 0) *out = a / (double) count;
}

// Returns the thin SVD pieces in p->x1 (A), p->x2 (J), p->x3 (norms) for n = even(B) tournament rows.
int32_t run_seed_svd(Plan* p, const float* mags, int F, int B, int* n_out)
{
  const int n = (B + 1) & ~1;
  FB_CUDA(p, p->x1.ensure(sizeof(double) * (size_t) n * F));
  FB_CUDA(p, p->x2.ensure(sizeof(double) * (size_t) n * n));
  FB_CUDA(p, p->x3.ensure(sizeof(double) * (size_t) n + 16));
  double* A = p->x1.as<double>();
  double* J = p->x2.as<double>();
  unsigned long long* off = reinterpret_cast<unsigned long long*>(p->x3.as<double>() + n);
  const int64_t total = (int64_t) n * F + (int64_t) n * n;
  k_seed_load<<<(int) std::min<int64_t>((total + 255) / 256, p->sm_count * 32), 256, 0, p->stream>>>(mags, F, B, n, A, J);
  p->launches++;
  for (int sweep = 0; sweep < 40; sweep++) {
    FB_CUDA(p, cudaMemsetAsync(off, 0, sizeof(unsigned long long), p->stream));
    for (int r = 0; r < n - 1; r++) k_seed_round<<<n / 2, 256, 0, p->stream>>>(A, J, F, n, r, off);
    p->launches += n - 1;
    unsigned long long h = 0;
    FB_CUDA(p, cudaMemcpyAsync(&h, off, sizeof(h), cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(p, cudaStreamSynchronize(p->stream));
    double rel;
    memcpy(&rel, &h, sizeof(rel));
    if (rel < 1e-13) break;
  }
  k_seed_norms<<<n, 256, 0, p->stream>>>(A, F, p->x3.as<double>());
  p->launches++;
  *n_out = n;
  return FB200_OK;
}

void launch_seed_build(Plan* p, const int* d_ord, int k, int F, int B, int n, int max_rank, int method, float* W, float* H)
{
  if (k <= 0) return;
  k_seed_build<<<k, 256, 0, p->stream>>>(p->x1.as<double>(), p->x2.as<double>(), p->x3.as<double>(), d_ord, F, B, n, max_rank, method, W, H);
  p->launches++;
}
void launch_seed_fill(Plan* p, float* W, float* H, int F, int B, int max_rank, int method, const float* mags, const float* U)
{
  double* mean = p->x3.as<double>() + (((B + 1) & ~1) + 1);
  k_seed_mean<<<1, 256, 0, p->stream>>>(mags, (int64_t) F * B, mean);
  const int64_t total = (int64_t) max_rank * B + (int64_t) F * max_rank;
  k_seed_fill<<<(int) std::min<int64_t>((total + 255) / 256, p->sm_count * 32), 256, 0, p->stream>>>(W, H, F, B, max_rank, method, mean, U);
  p->launches += 2;
}

} // namespace fb200
