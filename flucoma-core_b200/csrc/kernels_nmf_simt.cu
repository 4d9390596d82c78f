// SIMT fp32 engine for the NMF multiplicative updates (algorithms/public/NMF.hpp:144-183), any shape.
//
// One launch of k_nmf_tile performs, for a tile of 128 frames of one buffer:
//   phase 1 (do_h): H-update of these frames   P = H W, R = V / max(P,eps), H <- H * (R W^T) / max(sum_b W, eps)   (:165-170)
//   phase 2 (do_w): this tile's share of the NEXT W-update with the new H:  wnum_part = (V / max(H W, eps))^T H,
//                   wden_part = sum_f H                                                                      (:158-160)
// k_w_finalize then reduces the partials in a fixed order (bitwise repeatable, like the reference's seed tests
// require: tests/algorithms/public/TestNMF.cpp:31-39), applies W <- W * wnum / max(wden,eps), the conditional
// column normalisation (:161-162) and refreshes hden = sum_b W.
// Fusing "H-update of iteration i" with "W-numerator of iteration i+1" is legal because both use W(i) and the latter
// only needs H(i+1) of the same frames; it halves the passes over V.
//
// Bins are streamed in chunks of 64 so that the working set is independent of the FFT size; W chunks are staged in
// shared memory (double buffered), the ratio tile R goes through shared memory between the two small GEMMs.
#include "common.cuh"

namespace fb200 {

constexpr int TF = 128; // frames per tile
constexpr int BC = 64;  // bins per chunk
constexpr int WCS = BC + 4; // padded row stride of the staged W chunk
constexpr int NT = 256;

template <int KP>
struct TileSmem {
  static constexpr int ht = 0;                    // [KP][TF]
  static constexpr int hs = ht + KP * TF;         // [TF][KP]
  static constexpr int wc = hs + TF * KP;         // [2][KP][WCS]
  static constexpr int rt = wc + 2 * KP * WCS;    // [TF][BC]
  static constexpr int hden = rt + TF * BC;       // [KP]
  static constexpr int total = hden + KP;
};

template <int KP>
__device__ __forceinline__ void load_w_chunk_regs(const float* __restrict__ W, int Bp, int c, int tid, float4 (&wr)[(KP * 16 + NT - 1) / NT])
{
  // KP rows x 16 float4 per chunk
#pragma unroll
  for (int i = 0; i < (KP * 16 + NT - 1) / NT; i++) {
    int e = tid + i * NT;
    int k = e >> 4, g = e & 15;
    int bin = c * BC + 4 * g;
    wr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < KP && bin < Bp) wr[i] = *reinterpret_cast<const float4*>(W + (int64_t) k * Bp + bin);
  }
}

template <int KP>
__device__ __forceinline__ void store_w_chunk(float* wc, int tid, const float4 (&wr)[(KP * 16 + NT - 1) / NT])
{
#pragma unroll
  for (int i = 0; i < (KP * 16 + NT - 1) / NT; i++) {
    int e = tid + i * NT;
    int k = e >> 4, g = e & 15;
    if (k < KP) *reinterpret_cast<float4*>(wc + k * WCS + 4 * g) = wr[i];
  }
}

// ratio micro-tile: frames 8*tf..+7, bins 4*tb..+3 of chunk c.  R = V / max(H W, eps) -> Rt
template <int KP>
__device__ __forceinline__ void ratio_tile(const float* __restrict__ Vt, int Bp, int c, int tf, int tb,
                                           const float* __restrict__ Ht, const float* __restrict__ wc, float* __restrict__ Rt,
                                           int clamp_v)
{
  float p[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int q = 0; q < 4; q++) p[i][q] = 0.f;
#pragma unroll 8
  for (int k = 0; k < KP; k++) {
    float4 w = *reinterpret_cast<const float4*>(wc + k * WCS + 4 * tb);
    float4 h0 = *reinterpret_cast<const float4*>(Ht + k * TF + 8 * tf);
    float4 h1 = *reinterpret_cast<const float4*>(Ht + k * TF + 8 * tf + 4);
    float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      p[i][0] = fmaf(h[i], w.x, p[i][0]);
      p[i][1] = fmaf(h[i], w.y, p[i][1]);
      p[i][2] = fmaf(h[i], w.z, p[i][2]);
      p[i][3] = fmaf(h[i], w.w, p[i][3]);
    }
  }
  int bin = c * BC + 4 * tb;
  bool inb = bin < Bp;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inb) v = __ldg(reinterpret_cast<const float4*>(Vt + (int64_t) (8 * tf + i) * Bp + bin));
    if (clamp_v) { v.x = fmaxf(v.x, kEps); v.y = fmaxf(v.y, kEps); v.z = fmaxf(v.z, kEps); v.w = fmaxf(v.w, kEps); }
    float4 r;
    r.x = __fdividef(v.x, fmaxf(p[i][0], kEps));
    r.y = __fdividef(v.y, fmaxf(p[i][1], kEps));
    r.z = __fdividef(v.z, fmaxf(p[i][2], kEps));
    r.w = __fdividef(v.w, fmaxf(p[i][3], kEps));
    *reinterpret_cast<float4*>(Rt + (8 * tf + i) * BC + 4 * tb) = r;
  }
}

// W <- W * wnum / max(wden, eps); if max(W) > eps normalise each W row (Eigen column) to unit L2; hden = sum_b W.
// NMF.hpp:161-162, :169.  One CTA per buffer, one warp per component row.  The partials come from other CTAs (ld.cg: an L1
// line of an earlier iteration may still be around when this runs inside the tile kernel).
template <int KP>
__device__ __forceinline__ void w_finalize_body(const NmfDev& d, int buf)
{
  __shared__ float wden[KP];
  __shared__ float inv_norm[KP];
  __shared__ float wmax[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nc = d.ctas_per_buf, Bp = d.Bp;
  float* __restrict__ W = d.W + (int64_t) buf * KP * Bp;
  const float* __restrict__ part = d.wnum_part + (int64_t) buf * nc * KP * Bp;
  if (tid < KP) {
    float s = 0.f;
    for (int c = 0; c < nc; c++) s += __ldcg(d.wden_part + ((int64_t) buf * nc + c) * KP + tid);
    wden[tid] = fmaxf(s, kEps);
  }
  __syncthreads();
  float mx = 0.f;
  for (int k = warp; k < KP; k += 8) {
    float ss = 0.f;
    float den = wden[k];
    for (int b = lane; b < Bp; b += 32) {
      float wn = 0.f;
      for (int c = 0; c < nc; c++) wn += __ldcg(part + ((int64_t) c * KP + k) * Bp + b);
      float w = W[(int64_t) k * Bp + b] * wn / den;
      W[(int64_t) k * Bp + b] = w;
      ss = fmaf(w, w, ss);
      mx = fmaxf(mx, w);
    }
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) inv_norm[k] = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
  }
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) wmax[warp] = mx;
  __syncthreads();
  float gmax = wmax[0];
#pragma unroll
  for (int i = 1; i < 8; i++) gmax = fmaxf(gmax, wmax[i]);
  const bool norm = gmax > kEps;
  for (int k = warp; k < KP; k += 8) {
    float sc = norm ? inv_norm[k] : 1.0f;
    float sum = 0.f;
    for (int b = lane; b < Bp; b += 32) {
      float w = W[(int64_t) k * Bp + b];
      if (norm) { w *= sc; W[(int64_t) k * Bp + b] = w; }
      sum += w;
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) d.hden[(int64_t) buf * KP + k] = sum;
  }
}

template <int KP>
__global__ void __launch_bounds__(NT) k_nmf_tile(NmfDev d, int do_h, int do_w, int h_iters)
{
  extern __shared__ __align__(16) float smem[];
  using L = TileSmem<KP>;
  float* Ht = smem + L::ht;
  float* Hs = smem + L::hs;
  float* Wc = smem + L::wc;
  float* Rt = smem + L::rt;
  float* hden_s = smem + L::hden;

  constexpr int KG = KP / 4;   // groups of 4 components
  constexpr int BS = 64 / KP;  // bin splits in the hnum GEMM   (KG * BS == 16)
  constexpr int FS = 16 / KG;  // frame splits in the wnum GEMM (KG * FS == 16), FS == BS
  constexpr int NW = (KP * 16 + NT - 1) / NT;

  const int tid = threadIdx.x;
  const int buf = blockIdx.y;
  const int tile = blockIdx.x;
  const int f0 = tile * TF;
  const int Bp = d.Bp;
  const float* __restrict__ Vt = d.V + ((int64_t) buf * d.Fp + f0) * Bp;
  const float* __restrict__ W = d.W + (d.shared_w ? (int64_t) 0 : (int64_t) buf * KP * Bp);
  float* __restrict__ Hg = d.H + ((int64_t) buf * d.Fp + f0) * KP;
  const int nchunks = (Bp + BC - 1) / BC;

  // H tile in both orientations
  for (int e = tid; e < TF * KP; e += NT) {
    float h = Hg[e];
    Hs[e] = h;
    int f = e / KP, k = e - f * KP;
    Ht[k * TF + f] = h;
  }
  if (tid < KP) hden_s[tid] = d.hden[(d.shared_w ? 0 : (int64_t) buf * KP) + tid];

  const int tb = tid & 15, tf = tid >> 4;
  float4 wr[NW];

  if (do_h) {
    const int bs = tid % BS, kg = (tid / BS) % KG, fg = tid >> 4; // fg == tf: R rows come from this half-warp
    for (int it = 0; it < h_iters; it++) {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int q = 0; q < 4; q++) acc[i][q] = 0.f;
      load_w_chunk_regs<KP>(W, Bp, 0, tid, wr);
      store_w_chunk<KP>(Wc, tid, wr);
      __syncthreads(); // H tile (first pass: loaded, later: updated) + chunk 0 visible
      for (int c = 0; c < nchunks; c++) {
        const float* wc = Wc + (c & 1) * KP * WCS;
        if (c + 1 < nchunks) load_w_chunk_regs<KP>(W, Bp, c + 1, tid, wr);
        ratio_tile<KP>(Vt, Bp, c, tf, tb, Ht, wc, Rt, d.clamp_v);
        __syncwarp();
        // hnum[f][k] += sum_b R[f][b] W[k][b]   (NMF.hpp:168)
#pragma unroll 1
        for (int g = bs; g < 16; g += BS) {
          float4 w4[4];
#pragma unroll
          for (int kk = 0; kk < 4; kk++) w4[kk] = *reinterpret_cast<const float4*>(wc + (4 * kg + kk) * WCS + 4 * g);
#pragma unroll
          for (int i = 0; i < 8; i++) {
            float4 r = *reinterpret_cast<const float4*>(Rt + (8 * fg + i) * BC + 4 * g);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              float a = acc[i][kk];
              a = fmaf(r.x, w4[kk].x, a);
              a = fmaf(r.y, w4[kk].y, a);
              a = fmaf(r.z, w4[kk].z, a);
              a = fmaf(r.w, w4[kk].w, a);
              acc[i][kk] = a;
            }
          }
        }
        if (c + 1 < nchunks) store_w_chunk<KP>(Wc + ((c + 1) & 1) * KP * WCS, tid, wr);
        __syncthreads();
      }
      // reduce over the bin splits (adjacent lanes), then H <- H * hnum / max(hden, eps)   (NMF.hpp:170)
#pragma unroll
      for (int o = 1; o < BS; o <<= 1)
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int kk = 0; kk < 4; kk++) acc[i][kk] += __shfl_xor_sync(0xffffffffu, acc[i][kk], o);
      if (bs == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
            int f = 8 * fg + i, k = 4 * kg + kk;
            float hn = Hs[f * KP + k] * acc[i][kk] / fmaxf(hden_s[k], kEps);
            Hs[f * KP + k] = hn;
            Ht[k * TF + f] = hn;
          }
      }
      __syncthreads();
    }
    for (int e = tid; e < TF * KP; e += NT) Hg[e] = Hs[e];
  } else {
    __syncthreads();
  }

  if (do_w) {
    const int cta = buf * gridDim.x + tile;
    if (tid < KP) { // wden partial: sum_f H[f][k] over this tile   (NMF.hpp:160)
      float s = 0.f;
      for (int f = 0; f < TF; f++) s += Hs[f * KP + tid];
      d.wden_part[(int64_t) cta * KP + tid] = s;
    }
    // thread bit-fields (LSB first): bg_lo[3] fsl[log2 FSL] bg_hi[1] fsh[log2 FSH] kg[log2 KG].  Eight consecutive lanes
    // read eight consecutive float4 of one R row (conflict free); up to four frame splits sit at lane strides 8/16 so
    // their reduction is a warp shuffle; further splits (KP <= 8) are combined through shared memory.
    constexpr int FSL = FS < 4 ? FS : 4;
    constexpr int FSH = FS / FSL;
    constexpr int LSH = FSL == 4 ? 2 : (FSL == 2 ? 1 : 0);
    const int bg = (tid & 7) | (((tid >> (3 + LSH)) & 1) << 3);
    const int fsl = (tid >> 3) & (FSL - 1);
    const int rest = tid >> (4 + LSH);
    const int fsh = rest % FSH, kg = rest / FSH;
    const int fs = fsh * FSL + fsl;
    float* __restrict__ part = d.wnum_part + (int64_t) cta * KP * Bp;
    load_w_chunk_regs<KP>(W, Bp, 0, tid, wr);
    store_w_chunk<KP>(Wc, tid, wr);
    __syncthreads();
    for (int c = 0; c < nchunks; c++) {
      const float* wc = Wc + (c & 1) * KP * WCS;
      if (c + 1 < nchunks) load_w_chunk_regs<KP>(W, Bp, c + 1, tid, wr);
      ratio_tile<KP>(Vt, Bp, c, tf, tb, Ht, wc, Rt, d.clamp_v);
      __syncthreads();
      // wnum[k][b] = sum_f R[f][b] H[f][k] over the frames of this split   (NMF.hpp:159)
      float a2[4][4];
#pragma unroll
      for (int kk = 0; kk < 4; kk++)
#pragma unroll
        for (int q = 0; q < 4; q++) a2[kk][q] = 0.f;
      constexpr int FPER = TF / FS;
#pragma unroll 4
      for (int j = 0; j < FPER; j++) {
        int f = fs * FPER + j;
        float4 h = *reinterpret_cast<const float4*>(Hs + f * KP + 4 * kg);
        float4 r = *reinterpret_cast<const float4*>(Rt + f * BC + 4 * bg);
        float hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          a2[kk][0] = fmaf(hh[kk], r.x, a2[kk][0]);
          a2[kk][1] = fmaf(hh[kk], r.y, a2[kk][1]);
          a2[kk][2] = fmaf(hh[kk], r.z, a2[kk][2]);
          a2[kk][3] = fmaf(hh[kk], r.w, a2[kk][3]);
        }
      }
#pragma unroll
      for (int o = 1; o < FSL; o <<= 1)
#pragma unroll
        for (int kk = 0; kk < 4; kk++)
#pragma unroll
          for (int q = 0; q < 4; q++) a2[kk][q] += __shfl_xor_sync(0xffffffffu, a2[kk][q], o * 8);
      int bin = c * BC + 4 * bg;
      if (FSH > 1) { // combine the cross-warp frame splits through shared memory (the R tile is free after the sync)
        __syncthreads();
        float* red = Rt; // [FSH][KG][16 bg][16]
        if (fsl == 0) {
#pragma unroll
          for (int kk = 0; kk < 4; kk++)
            *reinterpret_cast<float4*>(red + ((fsh * KG + kg) * 16 + bg) * 16 + 4 * kk) = make_float4(a2[kk][0], a2[kk][1], a2[kk][2], a2[kk][3]);
        }
        __syncthreads();
        if (fs == 0) {
          for (int h = 1; h < FSH; h++)
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              float4 o4 = *reinterpret_cast<const float4*>(red + ((h * KG + kg) * 16 + bg) * 16 + 4 * kk);
              a2[kk][0] += o4.x; a2[kk][1] += o4.y; a2[kk][2] += o4.z; a2[kk][3] += o4.w;
            }
        }
      }
      if (fs == 0 && bin < Bp) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++)
          *reinterpret_cast<float4*>(part + (int64_t) (4 * kg + kk) * Bp + bin) = make_float4(a2[kk][0], a2[kk][1], a2[kk][2], a2[kk][3]);
      }
      if (c + 1 < nchunks) store_w_chunk<KP>(Wc + ((c + 1) & 1) * KP * WCS, tid, wr);
      __syncthreads();
    }
    // The CTA that finishes last for this buffer reduces the partials and updates W (fixed summation order, whoever runs it):
    // one launch per iteration instead of two -- a single-buffer call is launch-bound (config 1: 101 + 100 launches).
    if (d.ticket) {
      __shared__ int s_last;
      __threadfence(); // this CTA's partials are visible device-wide before its ticket is
      __syncthreads();
      if (tid == 0) {
        const int tk = atomicAdd(d.ticket + buf, 1);
        s_last = tk == (int) gridDim.x - 1;
        if (s_last) d.ticket[buf] = 0; // ready for the next launch
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        w_finalize_body<KP>(d, buf);
      }
    }
  }
}

// stand-alone form of the W finalisation (kept for callers that have no ticket array)
template <int KP>
__global__ void __launch_bounds__(NT) k_w_finalize(NmfDev d)
{
  w_finalize_body<KP>(d, blockIdx.x);
}

// Vhat = W H  (NMF.hpp:182 / :88), written dense [batch][F][B] in the caller's dtype
template <class D>
__global__ void __launch_bounds__(NT) k_vhat(NmfDev d, D* __restrict__ out)
{
  extern __shared__ __align__(16) float smem[];
  const int buf = blockIdx.y, f0 = blockIdx.x * 32, KP = d.KP;
  const float* __restrict__ W = d.W + (d.shared_w ? (int64_t) 0 : (int64_t) buf * KP * d.Bp);
  const float* __restrict__ H = d.H + ((int64_t) buf * d.Fp + f0) * KP;
  int nf = min(32, d.F - f0);
  for (int e = threadIdx.x; e < 32 * KP; e += NT) smem[e] = (e < nf * KP) ? H[e] : 0.f;
  __syncthreads();
  for (int b = threadIdx.x; b < d.B; b += NT) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0.f;
    for (int k = 0; k < KP; k++) {
      float w = W[(int64_t) k * d.Bp + b];
#pragma unroll
      for (int i = 0; i < 32; i++) acc[i] = fmaf(smem[i * KP + k], w, acc[i]);
    }
    for (int i = 0; i < nf; i++) out[((int64_t) buf * d.F + f0 + i) * d.B + b] = (D) acc[i];
  }
}

// ------------------------------------------------------------------------------------------------------------
template <int KP>
static void launch_tile_t(Plan* p, const NmfDev& d, int do_h, int do_w, int h_iters)
{
  size_t smem = sizeof(float) * TileSmem<KP>::total;
  if (!(p->attr_mask & (uint32_t) KP)) { // once per plan (== per device)
    cudaFuncSetAttribute(k_nmf_tile<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    p->attr_mask |= (uint32_t) KP;
  }
  dim3 grid(d.ctas_per_buf, d.batch);
  k_nmf_tile<KP><<<grid, NT, smem, p->stream>>>(d, do_h, do_w, h_iters);
}

int32_t simt_configure(Plan* p, NmfDev& d)
{
  d.KP = rank_pad(d.K);
  if (d.KP > 64) { p->err = "rank > 64 is not supported by the SIMT engine"; return FB200_ERR_UNSUPPORTED; }
  d.Bp = (int) round_up(d.B, 4);
  d.Fp = (int) round_up(d.F, TF);
  d.ctas_per_buf = d.Fp / TF;
  d.tiles_per_cta = 1;
  if (d.batch > 65535) { p->err = "batch > 65535 per call"; return FB200_ERR_UNSUPPORTED; }
  return FB200_OK;
}

static cudaEvent_t next_kernel_event(Plan* p)
{
  if (p->kev_used == p->kev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    p->kev.push_back(e);
  }
  return p->kev[p->kev_used++];
}

void simt_launch_tile(Plan* p, const NmfDev& d, int do_h, int do_w, int h_iters)
{
  cudaEventRecord(next_kernel_event(p), p->stream);
  switch (d.KP) {
  case 4: launch_tile_t<4>(p, d, do_h, do_w, h_iters); break;
  case 8: launch_tile_t<8>(p, d, do_h, do_w, h_iters); break;
  case 16: launch_tile_t<16>(p, d, do_h, do_w, h_iters); break;
  case 32: launch_tile_t<32>(p, d, do_h, do_w, h_iters); break;
  default: launch_tile_t<64>(p, d, do_h, do_w, h_iters); break;
  }
  cudaEventRecord(next_kernel_event(p), p->stream);
  p->launches++; p->launches_nmf++;
}

void simt_launch_w_finalize(Plan* p, const NmfDev& d)
{
  switch (d.KP) {
  case 4: k_w_finalize<4><<<d.batch, NT, 0, p->stream>>>(d); break;
  case 8: k_w_finalize<8><<<d.batch, NT, 0, p->stream>>>(d); break;
  case 16: k_w_finalize<16><<<d.batch, NT, 0, p->stream>>>(d); break;
  case 32: k_w_finalize<32><<<d.batch, NT, 0, p->stream>>>(d); break;
  default: k_w_finalize<64><<<d.batch, NT, 0, p->stream>>>(d); break;
  }
  p->launches++; p->launches_nmf++;
}

void launch_vhat(Plan* p, const NmfDev& d, void* dst, int dst_dtype)
{
  dim3 grid((d.F + 31) / 32, d.batch);
  size_t smem = sizeof(float) * 32 * d.KP;
  if (dst_dtype == FB200_F64) k_vhat<double><<<grid, NT, smem, p->stream>>>(d, (double*) dst);
  else k_vhat<float><<<grid, NT, smem, p->stream>>>(d, (float*) dst);
  p->launches++;
}

} // namespace fb200
