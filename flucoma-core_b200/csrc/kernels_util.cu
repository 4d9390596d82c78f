// Utility kernels: strided dtype-converting copies, device-side mt19937_64, W/H initialisation, activation scaling.
#include "common.cuh"

namespace fb200 {

// ------------------------------------------------------------------------------------------------------------
// copy3d
// ------------------------------------------------------------------------------------------------------------
template <class S, class D>
__global__ void __launch_bounds__(256) k_copy3d(const S* __restrict__ src, int64_t s_b, int64_t s_r, D* __restrict__ dst,
                                                int64_t d_b, int64_t d_r, int64_t batch, int64_t rows, int64_t cols,
                                                const float* __restrict__ scale, int clamp_eps)
{
  int64_t total = batch * rows * cols;
  for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < total; i += (int64_t) gridDim.x * blockDim.x) {
    int64_t j = i % cols;
    int64_t r = (i / cols) % rows;
    int64_t b = i / (cols * rows);
    float v = (float) src[b * s_b + r * s_r + j];
    if (clamp_eps) v = fmaxf(v, kEps);
    if (scale) v *= scale[b];
    dst[b * d_b + r * d_r + j] = (D) v;
  }
}

// Row-wise variant for rows of at least a warp's width: one warp per row, the (batch, row) decomposition is done once per
// row instead of three 64-bit divisions per element (the element-wise kernel above moved the 2 GB |X| stream of config 5
// at 1.4 TB/s).
template <class S, class D>
__global__ void __launch_bounds__(256) k_copy_rows(const S* __restrict__ src, int64_t s_b, int64_t s_r, D* __restrict__ dst,
                                                   int64_t d_b, int64_t d_r, int64_t batch, int64_t rows, int cols,
                                                   const float* __restrict__ scale, int clamp_eps)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < batch * rows; row += nwarps) {
    const int64_t b = row / rows, r = row - b * rows;
    const S* s = src + b * s_b + r * s_r;
    D* d = dst + b * d_b + r * d_r;
    const float sc = scale ? scale[b] : 1.0f;
#pragma unroll 4
    for (int j = lane; j < cols; j += 32) {
      float v = (float) s[j];
      if (clamp_eps) v = fmaxf(v, kEps);
      if (scale) v *= sc;
      d[j] = (D) v;
    }
  }
}

void launch_copy3d(Plan* p, const void* src, int src_dtype, int64_t s_b, int64_t s_r, void* dst, int dst_dtype,
                   int64_t d_b, int64_t d_r, int64_t batch, int64_t rows, int64_t cols, const float* scale, int clamp_eps)
{
  int64_t total = batch * rows * cols;
  if (total <= 0) return;
  const bool by_rows = cols >= 32 && cols < (int64_t) 1 << 30;
  int grid = by_rows ? (int) std::min<int64_t>((batch * rows + 7) / 8, (int64_t) p->sm_count * 16)
                     : (int) std::min<int64_t>((total + 255) / 256, (int64_t) p->sm_count * 16);
#define FB_COPY(S, D)                                                                                                             \
  do {                                                                                                                            \
    if (by_rows) k_copy_rows<S, D><<<grid, 256, 0, p->stream>>>((const S*) src, s_b, s_r, (D*) dst, d_b, d_r, batch, rows, (int) cols, scale, clamp_eps); \
    else k_copy3d<S, D><<<grid, 256, 0, p->stream>>>((const S*) src, s_b, s_r, (D*) dst, d_b, d_r, batch, rows, cols, scale, clamp_eps); \
  } while (0)
  if (src_dtype == FB200_F32 && dst_dtype == FB200_F32) FB_COPY(float, float);
  else if (src_dtype == FB200_F64 && dst_dtype == FB200_F32) FB_COPY(double, float);
  else if (src_dtype == FB200_F32 && dst_dtype == FB200_F64) FB_COPY(float, double);
  else FB_COPY(double, double);
#undef FB_COPY
  p->launches++;
}

// ------------------------------------------------------------------------------------------------------------
// mt19937_64 + libstdc++ uniform_real_distribution<double>(0,1): value = double(raw) / 2^64, 1.0 -> nextafter(1,0).
// algorithms/util/EigenRandom.hpp:73-110.  One WARP per seed, the 312-word state in shared memory.  The twist of a
// state block is two data-parallel halves: words 0..155 depend on old words only, words 156..311 on old words and on
// the new first half (word 311 also on the new word 0), exactly as the sequential recurrence sees them.  Each half
// reads everything it needs before it writes (the sequential loop reads mt[i+1] before overwriting it).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mt_uniform(const int64_t* __restrict__ seeds, int64_t batch, int64_t count,
                                                    float* __restrict__ U)
{
  __shared__ unsigned long long state[4][312];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t b = (int64_t) blockIdx.x * 4 + w;
  if (b >= batch) return;
  unsigned long long* mt = state[w];
  if (lane == 0) {
    unsigned long long x = (unsigned long long) seeds[b];
    mt[0] = x;
    for (int i = 1; i < 312; i++) {
      x = 6364136223846793005ULL * (x ^ (x >> 62)) + (unsigned long long) i;
      mt[i] = x;
    }
  }
  __syncwarp();
  float* out = U + b * count;
  for (int64_t base = 0; base < count; base += 312) {
#pragma unroll
    for (int half = 0; half < 2; half++) {
      unsigned long long nv[5];
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const int i = half * 156 + lane + 32 * j;
        if (lane + 32 * j < 156) {
          const int i1 = i + 1 == 312 ? 0 : i + 1;
          const int im = i + 156 >= 312 ? i - 156 : i + 156;
          const unsigned long long x = (mt[i] & 0xFFFFFFFF80000000ULL) | (mt[i1] & 0x7FFFFFFFULL);
          nv[j] = mt[im] ^ (x >> 1) ^ ((x & 1ULL) ? 0xB5026F5AA96619E9ULL : 0ULL);
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 5; j++)
        if (lane + 32 * j < 156) mt[half * 156 + lane + 32 * j] = nv[j];
      __syncwarp();
    }
    for (int i = lane; i < 312 && base + i < count; i += 32) {
      unsigned long long x = mt[i];
      x ^= (x >> 29) & 0x5555555555555555ULL;
      x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
      x ^= (x << 37) & 0xFFF7EEE000000000ULL;
      x ^= (x >> 43);
      double r = __ull2double_rn(x) * 5.42101086242752217e-20; // 2^-64
      if (r >= 1.0) r = 0.99999999999999989;
      out[base + i] = (float) r;
    }
    __syncwarp();
  }
}

void launch_mt_uniform(Plan* p, const int64_t* d_seeds, int64_t batch, int64_t count, float* U)
{
  if (batch <= 0 || count <= 0) return;
  k_mt_uniform<<<(unsigned) ((batch + 3) / 4), 128, 0, p->stream>>>(d_seeds, batch, count, U);
  p->launches++;
}

// ------------------------------------------------------------------------------------------------------------
// NMF initialisation.  NMF.hpp:101-124 (random: column-major fill, W and H both restart the stream; or seeds),
// :150-153 (eps clamp, W column / H row L2 normalise -- in our layouts: rows of W[KP][Bp], columns of H[Fp][KP]).
// Pad entries (k >= K, b >= B, f >= F) are exact zeros and stay zero under the updates.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_init_w(NmfDev d, const float* __restrict__ U, int64_t u_stride,
                                                const float* __restrict__ W0)
{
  int buf = blockIdx.x;
  float* W = d.W + (int64_t) buf * d.KP * d.Bp;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < d.KP; k += 8) {
    float* row = W + (int64_t) k * d.Bp;
    float ss = 0.f;
    for (int b = lane; b < d.Bp; b += 32) {
      float w = 0.f;
      if (k < d.K && b < d.B) {
        w = W0 ? W0[((int64_t) buf * d.K + k) * d.B + b] : U[(int64_t) buf * u_stride + (int64_t) k * d.B + b];
        w = fmaxf(w, kEps);
      }
      row[b] = w;
      ss += w * w;
    }
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    float inv = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    float sum = 0.f;
    for (int b = lane; b < d.Bp; b += 32) {
      float w = row[b] * inv;
      row[b] = w;
      sum += w;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) d.hden[(int64_t) buf * d.KP + k] = sum;
  }
}

__global__ void __launch_bounds__(256) k_init_h(NmfDev d, const float* __restrict__ U, int64_t u_stride,
                                                const float* __restrict__ H0)
{
  __shared__ float red[256];
  __shared__ float inv[64];
  int buf = blockIdx.x;
  float* H = d.H + (int64_t) buf * d.Fp * d.KP;
  int64_t total = (int64_t) d.Fp * d.KP;
  int k = threadIdx.x % d.KP; // 256 % KP == 0, so a thread always sees the same column
  float ss = 0.f;
  for (int64_t e = threadIdx.x; e < total; e += 256) {
    int64_t f = e / d.KP;
    float h = 0.f;
    if (k < d.K && f < d.F) {
      h = H0 ? H0[((int64_t) buf * d.F + f) * d.K + k] : U[(int64_t) buf * u_stride + f * d.K + k];
      h = fmaxf(h, kEps);
    }
    H[e] = h;
    ss += h * h;
  }
  red[threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < d.KP) {
    float s = 0.f;
    for (int t = threadIdx.x; t < 256; t += d.KP) s += red[t];
    inv[threadIdx.x] = s > 0.f ? 1.0f / sqrtf(s) : 0.f;
  }
  __syncthreads();
  float sc = inv[k];
  for (int64_t e = threadIdx.x; e < total; e += 256) H[e] *= sc;
}

// processFrame: every frame starts from the same h0 = max(U(K), eps), not normalised (NMF.hpp:55,59)
__global__ void __launch_bounds__(256) k_init_h_frames(NmfDev d, const float* __restrict__ U, int per_frame)
{
  int64_t total = (int64_t) d.batch * d.Fp * d.KP;
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    int k = (int) (e % d.KP);
    int64_t f = (e / d.KP) % d.Fp;
    d.H[e] = (k < d.K && f < d.F) ? fmaxf(U[per_frame ? f * d.K + k : k], kEps) : 0.f;
  }
}

void launch_nmf_init(Plan* p, const NmfDev& d, const float* U, const float* U_h, int64_t u_stride, const float* W0,
                     const float* H0, int frame_mode)
{
  int nW = d.shared_w ? 1 : d.batch;
  k_init_w<<<nW, 256, 0, p->stream>>>(d, U, u_stride, W0);
  p->launches++;
  if (frame_mode) {
    int64_t total = (int64_t) d.batch * d.Fp * d.KP;
    int grid = (int) std::min<int64_t>((total + 255) / 256, (int64_t) p->sm_count * 16);
    k_init_h_frames<<<grid, 256, 0, p->stream>>>(d, U_h, frame_mode == 2);
  } else {
    k_init_h<<<d.batch, 256, 0, p->stream>>>(d, U_h, u_stride, H0);
  }
  p->launches++;
}

// ------------------------------------------------------------------------------------------------------------
// scale[b] = 1 / max_{f<F,k<K} H   (NMFClient.hpp:289-291)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_h_max_scale(NmfDev d, float* __restrict__ scale)
{
  __shared__ float red[256];
  int buf = blockIdx.x;
  const float* H = d.H + (int64_t) buf * d.Fp * d.KP;
  int64_t total = (int64_t) d.F * d.KP;
  float m = -INFINITY;
  for (int64_t e = threadIdx.x; e < total; e += 256) {
    int k = (int) (e % d.KP);
    if (k < d.K) m = fmaxf(m, H[e]);
  }
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) scale[buf] = (float) (1.0 / (double) red[0]);
}

void launch_h_max_scale(Plan* p, const NmfDev& d, float* scale)
{
  k_h_max_scale<<<d.batch, 256, 0, p->stream>>>(d, scale);
  p->launches++;
}

} // namespace fb200
