// Epilogues on the STFT output: MelBands (algorithms/public/MelBands.hpp:35-101) and HPSS
// (algorithms/public/HPSS.hpp:47-162 with algorithms/util/MedianFilter.hpp:36-57) for whole frame sequences at once.
#include "common.cuh"

namespace fb200 {

// ---- MelBands::processFrame (:82-101) for every frame: one CTA per frame ------------------------------------------------
// mags [frames][B] -> bands [frames][nb].  flags: 1 magNorm, 2 usePower, 4 logOutput.  filt [nb][B] from MelBands::init.
__global__ void __launch_bounds__(256) k_melbands(const float* __restrict__ mags, const float* __restrict__ filt, int B, int nb, float scale1,
                                                  float scale2, int flags, float* __restrict__ bands)
{
  extern __shared__ float sm[]; // [B] frame, [nb] results, [8] reduction
  float* frame = sm;
  float* res = sm + B;
  float* red = res + nb;
  const int64_t f = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool norm = flags & 1, power = flags & 2, logo = flags & 4;
  float part = 0.f;
  for (int b = tid; b < B; b += 256) {
    float x = mags[f * B + b];
    if (norm) x *= scale1;                  // :90
    part += x;
    frame[b] = power ? x * x : x;           // :92
  }
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  float energy = 0.f;
  for (int w = 0; w < 8; w++) energy += red[w];
  energy *= scale2;                         // :91
  for (int i = warp; i < nb; i += 8) {      // :94-95
    const float* fr = filt + (int64_t) i * B;
    float s = 0.f;
    for (int b = lane; b < B; b += 32) s = fmaf(fr[b], frame[b], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) res[i] = s;
  }
  __syncthreads();
  float sum = 0.f;
  if (norm) for (int i = 0; i < nb; i++) sum += res[i];
  for (int i = tid; i < nb; i += 256) {
    float r = res[i];
    if (norm) r = r * energy / fmaxf(kEps, sum);        // :97
    if (logo) r = 20.0f * log10f(fmaxf(r, kEps));       // :99
    bands[f * nb + i] = r;
  }
}

void launch_melbands(Plan* p, const float* mags, const float* filt, int64_t frames, int B, int nb, float scale1, float scale2, int flags,
                     float* bands)
{
  if (frames <= 0) return;
  k_melbands<<<(unsigned) frames, 256, sizeof(float) * (size_t) (B + nb + 8), p->stream>>>(mags, filt, B, nb, scale1, scale2, flags, bands);
  p->launches++;
}

// ---- HPSS --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_abs(const float2* __restrict__ spec, int64_t count, float* __restrict__ mag)
{
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < count; e += (int64_t) gridDim.x * blockDim.x) {
    const float2 s = spec[e];
    mag[e] = hypotf(s.x, s.y);
  }
}

// median of n values (n odd): the element with exactly n / 2 predecessors in the stable order (value, index)
__device__ __forceinline__ float median_of(const float* w, int n)
{
  float med = 0.f;
  for (int i = 0; i < n; i++) {
    const float x = w[i];
    int c = 0;
    for (int j = 0; j < n; j++) c += (w[j] < x || (w[j] == x && j < i)) ? 1 : 0;
    if (c == n / 2) med = x;
  }
  return med;
}

// What the delay lines of HPSS::processFrame hold when frame t is emitted (closed form of the streaming recursion):
//   X0 = X[u], u = t - (hSize - 1);  v0 = median(|X[u]|[b .. b + vSize - 1]) (zeros past the last bin);
//   h0 = median(|X[w - hSize + 1 .. w]|[b]), w = t - (h2 + 1) (zeros before the stream);  all zero while the lines fill.
// mode 0 classic soft masks (:101-107), 1 coupled binary (:108-115), 2 advanced with residual (:116-135).
__global__ void __launch_bounds__(128) k_hpss(const float2* __restrict__ spec, const float* __restrict__ mag, int F, int B, int vsize, int hsize,
                                              int mode, const float* __restrict__ th_h, const float* __restrict__ th_p, float2* __restrict__ out)
{
  const int64_t total = (int64_t) F * B;
  const int h2 = (hsize - 1) / 2;
  float w[129];
  for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    const int b = (int) (e % B), t = (int) (e / B);
    const int64_t buf = blockIdx.y;
    const float2* S = spec + buf * total;
    const float* A = mag + buf * total;
    const int u = t - (hsize - 1), wf = t - (h2 + 1);
    float2 x0 = make_float2(0.f, 0.f);
    float v0 = 0.f, h0 = 0.f;
    if (u >= 0) {
      x0 = S[(int64_t) u * B + b];
      for (int i = 0; i < vsize; i++) w[i] = (b + i < B) ? A[(int64_t) u * B + b + i] : 0.f;
      v0 = median_of(w, vsize);
    }
    if (wf >= 0) {
      for (int i = 0; i < hsize; i++) {
        const int ff = wf - hsize + 1 + i;
        w[i] = ff >= 0 ? A[(int64_t) ff * B + b] : 0.f;
      }
      h0 = median_of(w, hsize);
    }
    float hm, pm, rm = mode == 2 ? 1.f : 0.f;
    if (mode == 0) {
      const float mult = 1.0f / fmaxf(h0 + v0, kEps);
      hm = h0 * mult; pm = v0 * mult;
    } else if (mode == 1) {
      hm = (h0 / v0) > th_h[b] ? 1.f : 0.f;
      pm = 1.f - hm;
    } else {
      hm = (h0 / v0) > th_h[b] ? 1.f : 0.f;
      pm = (v0 / h0) > th_p[b] ? 1.f : 0.f;
      rm = rm * (1.f - hm);
      rm = rm * (1.f - pm);
      const float nrm = fmaxf(1.0f / (hm + pm + rm), kEps);
      hm *= nrm; pm *= nrm; rm *= nrm;
    }
    hm = fminf(hm, 1.f); pm = fminf(pm, 1.f); rm = fminf(rm, 1.f);            // :138-140
    float2* o = out + buf * 3 * total;
    o[e] = make_float2(x0.x * hm, x0.y * hm);
    o[total + e] = make_float2(x0.x * pm, x0.y * pm);
    o[2 * total + e] = make_float2(x0.x * rm, x0.y * rm);
  }
}

void launch_hpss(Plan* p, const float2* spec, float* mag, int64_t batch, int F, int B, int vsize, int hsize, int mode, const float* th_h,
                 const float* th_p, float2* out)
{
  const int64_t total = (int64_t) F * B;
  int g = (int) std::min<int64_t>((batch * total + 255) / 256, (int64_t) p->sm_count * 32);
  k_abs<<<g, 256, 0, p->stream>>>(spec, batch * total, mag);
  dim3 grid((unsigned) std::min<int64_t>((total + 127) / 128, (int64_t) p->sm_count * 16), (unsigned) batch);
  k_hpss<<<grid, 128, 0, p->stream>>>(spec, mag, F, B, vsize, hsize, mode, th_h, th_p, out);
  p->launches += 2;
}

} // namespace fb200
