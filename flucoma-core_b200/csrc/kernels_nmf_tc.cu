// tcgen05 / TMEM / TMA engine for the NMF multiplicative updates (algorithms/public/NMF.hpp:144-183), rank 16.
//
// One persistent CTA owns one buffer for ALL iterations: W and H live in shared memory (fp32 masters + split-bf16
// hi/lo operand copies in UMMA core-matrix layout), V = |X| is streamed by TMA (128B swizzle) exactly once per phase,
// WH is accumulated in TMEM, the ratio V / max(WH, eps) is computed by the epilogue warps straight out of TMEM and
// written back to TMEM as the A operand of the second MMA -- neither WH nor the ratio ever touch shared or global
// memory, and there is no inter-CTA communication at all.
//
//   phase 1 (H-update of a 128-frame tile t, NMF.hpp:165-170), per 64-bin chunk c:
//       P[f][b]   = H_t W_c            tcgen05.mma SS   A = H blocks (K-major)   B = W blocks (MN-major)   M128 N64 K16
//       R[f][b]   = V / max(P, eps)    epilogue: tcgen05.ld, swizzled LDS of the TMA tile, MUFU rcp, bf16 hi/lo, tcgen05.st
//       hnum[f][k] += R W_c^T          tcgen05.mma TS   A = R (TMEM)             B = W blocks (K-major)    M128 N16 K64
//     then H <- H * hnum / max(hden, eps) for the tile.
//   phase 2 (this tile's share of the next W-update, NMF.hpp:158-160), per 128-bin tile m and 64-frame half s:
//       P[b][f]   = W_m^T H_ts^T       SS   A = W blocks (MN-major)  B = H blocks (K-major)    M128 N64 K16
//       R[b][f]   = V / max(P, eps)
//       wnum[b][k] += R H_ts           TS   A = R (TMEM)             B = H blocks (MN-major)   M128 N16 K64
//   after the last tile: W <- W * wnum / max(wden, eps), conditional column normalisation (:161-162), hden = sum_b W.
// Every product is evaluated as hi*hi + hi*lo + lo*hi on bf16 pairs (x ~ hi + lo, 16 mantissa bits): plain bf16 misses
// the 1e-4 parity bar, the three-term split meets it with fp32-like margins.
// The Nyquist bin (B = 2^m + 1) does not fit the 128-wide tiles; its column is carried on the SIMT side of the
// epilogue (a 16-term dot product per frame), so the tensor tiles cover bins 0 .. B-2 exactly.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2-9 = two epilogue
// warpgroups that alternate steps (ping-pong on two P/R TMEM buffers).  All reductions are fixed-order: results are
// bitwise repeatable.
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda.h>

namespace fb200 {
using namespace tc;

namespace tcn {
constexpr int K = 16;
constexpr int KB = 2;            // 8-component blocks
constexpr int NS = 2;            // V ring stages of 32 KB
constexpr int STAGE = 32768;
constexpr int BT_MAX = 512;      // tensor bins (B - 1)
constexpr int FP_MAX = 512;      // padded frames
constexpr int WPITCH = BT_MAX + 4;
constexpr int NTHREADS = 320;

// TMEM columns
constexpr uint32_t TM_P = 0;     // + 64 g
constexpr uint32_t TM_R = 128;   // + 64 g : hi [0,32) lo [32,64)
constexpr uint32_t TM_HNUM = 256;
constexpr uint32_t TM_WNUM = 272; // + 16 m

// shared memory map (bytes)
constexpr int OFF_V = 0;
constexpr int OFF_WHI = OFF_V + NS * STAGE;
constexpr int OFF_WLO = OFF_WHI + K * BT_MAX * 2;
constexpr int OFF_HHI = OFF_WLO + K * BT_MAX * 2;
constexpr int OFF_HLO = OFF_HHI + FP_MAX * K * 2;
constexpr int OFF_WM = OFF_HLO + FP_MAX * K * 2;      // float [K][WPITCH]
constexpr int OFF_HM = OFF_WM + K * WPITCH * 4;       // float [FP_MAX][K]
constexpr int OFF_VN = OFF_HM + FP_MAX * K * 4;       // float [FP_MAX]  Nyquist column of V
constexpr int OFF_WN = OFF_VN + FP_MAX * 4;           // float [K]       Nyquist row of W
constexpr int OFF_HDEN = OFF_WN + K * 4;              // float [K]
constexpr int OFF_PART = OFF_HDEN + K * 4;            // float [4 tiles][8 warps][32]
constexpr int OFF_RED = OFF_PART + 4 * 8 * 32 * 4;    // float [8 warps][36]
constexpr int OFF_FIN = OFF_RED + 8 * 36 * 4;         // float [64]
constexpr int OFF_BAR = OFF_FIN + 64 * 4;             // mbarriers
constexpr int NBAR = 2 * NS + 2 + 2 + 1 + 3;
constexpr int OFF_SLOT = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_SLOT + 16;

__device__ __forceinline__ int wop_index(int k, int b) { return ((b >> 3) * KB + (k >> 3)) * 64 + (k & 7) * 8 + (b & 7); }
__device__ __forceinline__ int hop_index(int f, int k) { return ((f >> 3) * KB + (k >> 3)) * 64 + (f & 7) * 8 + (k & 7); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct Sched {
  int npass, both, upd_w, upd_h, iters;
  __device__ bool p1(int pass) const { return both ? pass > 0 : upd_h != 0; }
  __device__ bool p2(int pass) const { return both ? pass < iters : upd_w != 0; }
};
} // namespace tcn

using namespace tcn;

__global__ void __launch_bounds__(NTHREADS, 1)
k_nmf_tc(NmfDev d, const __grid_constant__ CUtensorMap tmap1, const __grid_constant__ CUtensorMap tmap2, int iters, int upd_w, int upd_h)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  __nv_bfloat16* whi = reinterpret_cast<__nv_bfloat16*>(smem + OFF_WHI);
  __nv_bfloat16* wlo = reinterpret_cast<__nv_bfloat16*>(smem + OFF_WLO);
  __nv_bfloat16* hhi = reinterpret_cast<__nv_bfloat16*>(smem + OFF_HHI);
  __nv_bfloat16* hlo = reinterpret_cast<__nv_bfloat16*>(smem + OFF_HLO);
  float* Wm = reinterpret_cast<float*>(smem + OFF_WM);
  float* Hm = reinterpret_cast<float*>(smem + OFF_HM);
  float* VN = reinterpret_cast<float*>(smem + OFF_VN);
  float* WN = reinterpret_cast<float*>(smem + OFF_WN);
  float* hden = reinterpret_cast<float*>(smem + OFF_HDEN);
  float* part = reinterpret_cast<float*>(smem + OFF_PART);
  float* red = reinterpret_cast<float*>(smem + OFF_RED);
  float* fin = reinterpret_cast<float*>(smem + OFF_FIN);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* v_full = bars;            // [NS]
  uint64_t* v_empty = bars + NS;      // [NS]
  uint64_t* p_full = bars + 2 * NS;   // [2]
  uint64_t* r_full = p_full + 2;      // [2]
  uint64_t* acc_full = r_full + 2;    // [1]
  // three separate "operands ready" barriers so that two completions can never pile up unobserved on one of them
  uint64_t* buf_ready = acc_full + 1;   // buffer prologue done
  uint64_t* prep_ready = buf_ready + 1; // tile prep (H-update) done
  uint64_t* w_ready = prep_ready + 1;   // W-update done
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + OFF_SLOT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Fp = d.Fp, Bp = d.Bp, BT = d.B - 1;
  const int T = Fp / 128, MT = BT / 128, C1 = BT / 64;
  Sched sc;
  sc.both = upd_w && upd_h; sc.upd_w = upd_w; sc.upd_h = upd_h; sc.iters = iters;
  sc.npass = sc.both ? iters + 1 : iters;

  if (tid == 0) {
    for (int i = 0; i < NS; i++) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 4); }
    for (int i = 0; i < 2; i++) { mbar_init(&p_full[i], 1); mbar_init(&r_full[i], 4); }
    mbar_init(acc_full, 1);
    mbar_init(buf_ready, 8);
    mbar_init(prep_ready, 8);
    mbar_init(w_ready, 8);
    mbar_fence_init();
    tma_prefetch_desc(&tmap1);
    tma_prefetch_desc(&tmap2);
  }
  if (warp == 1) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *slot;

  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      uint32_t n = 0;
      for (int buf = blockIdx.x; buf < d.batch; buf += gridDim.x) {
        for (int pass = 0; pass < sc.npass; pass++) {
          const bool p1 = sc.p1(pass), p2 = sc.p2(pass);
          for (int t = 0; t < T; t++) {
            if (p1)
              for (int c = 0; c < C1; c++, n++) {
                const uint32_t st = n % NS, k = n / NS;
                mbar_wait(&v_empty[st], (k & 1) ^ 1);
                mbar_arrive_expect_tx(&v_full[st], STAGE);
                uint8_t* dst = smem + OFF_V + st * STAGE;
                tma_load_3d(dst, &tmap1, 64 * c, 128 * t, buf, &v_full[st]);
                tma_load_3d(dst + 16384, &tmap1, 64 * c + 32, 128 * t, buf, &v_full[st]);
              }
            if (p2)
              for (int m = 0; m < MT; m++)
                for (int s = 0; s < 2; s++, n++) {
                  const uint32_t st = n % NS, k = n / NS;
                  mbar_wait(&v_empty[st], (k & 1) ^ 1);
                  mbar_arrive_expect_tx(&v_full[st], STAGE);
                  uint8_t* dst = smem + OFF_V + st * STAGE;
#pragma unroll
                  for (int w = 0; w < 4; w++) tma_load_3d(dst + w * 8192, &tmap2, 128 * m + 32 * w, 128 * t + 64 * s, buf, &v_full[st]);
                }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================== MMA issuer =============================================
    if (lane == 0) {
      const uint32_t whi_a = smem_u32(whi), wlo_a = smem_u32(wlo), hhi_a = smem_u32(hhi), hlo_a = smem_u32(hlo);
      constexpr uint32_t ID_P1A = make_idesc_bf16(128, 64, 0, 1);
      constexpr uint32_t ID_P1B = make_idesc_bf16(128, 16, 0, 0);
      constexpr uint32_t ID_P2A = make_idesc_bf16(128, 64, 1, 0);
      constexpr uint32_t ID_P2B = make_idesc_bf16(128, 16, 0, 1);
      uint32_t n = 0, buf_cnt = 0, prep_cnt = 0, w_cnt = 0;
      // pending second-stage MMA (issued one step late so the next step's first MMA overlaps this step's epilogue)
      int pend_valid = 0, pend_phase = 0, pend_g = 0, pend_blk = 0, pend_acc = 0, pend_m = 0;
      uint32_t pend_k = 0;
      auto issue_b = [&]() {
        if (!pend_valid) return;
        mbar_wait(&r_full[pend_g], pend_k & 1);
        tc_fence_after();
        const uint32_t rhi = tbase + TM_R + 64 * pend_g, rlo = rhi + 32;
        if (pend_phase == 1) { // hnum += R W_c^T : B = W blocks K-major, K-step j = bin blocks pend_blk + 2j
          const uint32_t dacc = tbase + TM_HNUM;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t off = (uint32_t) (pend_blk + 2 * j) * 256;
            const uint64_t bh = make_smem_desc(whi_a + off, 256, 128), bl = make_smem_desc(wlo_a + off, 256, 128);
            mma_ts(dacc, rhi + 8 * j, bh, ID_P1B, (pend_acc || j) ? 1u : 0u);
            mma_ts(dacc, rhi + 8 * j, bl, ID_P1B, 1u);
            mma_ts(dacc, rlo + 8 * j, bh, ID_P1B, 1u);
          }
        } else { // wnum[m] += R H_ts : B = H blocks MN-major, K-step j = frame blocks pend_blk + 2j
          const uint32_t dacc = tbase + TM_WNUM + 16 * pend_m;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t off = (uint32_t) (pend_blk + 2 * j) * 256;
            const uint64_t bh = make_smem_desc(hhi_a + off, 256, 128), bl = make_smem_desc(hlo_a + off, 256, 128);
            mma_ts(dacc, rhi + 8 * j, bh, ID_P2B, (pend_acc || j) ? 1u : 0u);
            mma_ts(dacc, rhi + 8 * j, bl, ID_P2B, 1u);
            mma_ts(dacc, rlo + 8 * j, bh, ID_P2B, 1u);
          }
        }
        pend_valid = 0;
      };
      for (int buf = blockIdx.x; buf < d.batch; buf += gridDim.x) {
        mbar_wait(buf_ready, buf_cnt & 1); buf_cnt++; // operands of this buffer are in shared memory
        tc_fence_after();
        for (int pass = 0; pass < sc.npass; pass++) {
          const bool p1 = sc.p1(pass), p2 = sc.p2(pass);
          for (int t = 0; t < T; t++) {
            if (p1) {
              for (int c = 0; c < C1; c++, n++) {
                const int g = n & 1;
                const uint32_t dP = tbase + TM_P + 64 * g;
                const uint32_t aoff = (uint32_t) (16 * t) * 256, boff = (uint32_t) (8 * c) * 256;
                const uint64_t ah = make_smem_desc(hhi_a + aoff, 128, 256), al = make_smem_desc(hlo_a + aoff, 128, 256);
                const uint64_t bh = make_smem_desc(whi_a + boff, 128, 256), bl = make_smem_desc(wlo_a + boff, 128, 256);
                mma_ss(dP, ah, bh, ID_P1A, 0u);
                mma_ss(dP, ah, bl, ID_P1A, 1u);
                mma_ss(dP, al, bh, ID_P1A, 1u);
                mma_commit(&p_full[g]);
                issue_b();
                pend_valid = 1; pend_phase = 1; pend_g = g; pend_k = n >> 1; pend_blk = 8 * c; pend_acc = c > 0; pend_m = 0;
              }
              issue_b();
              mma_commit(acc_full); // hnum of tile t complete (and every MMA reading H_op(t) has retired)
            }
            mbar_wait(prep_ready, prep_cnt & 1); prep_cnt++; // tile prep done: H_op(t) updated
            tc_fence_after();
            if (p2) {
              for (int m = 0; m < MT; m++)
                for (int s = 0; s < 2; s++, n++) {
                  const int g = n & 1;
                  const uint32_t dP = tbase + TM_P + 64 * g;
                  const uint32_t aoff = (uint32_t) (16 * m) * 256, boff = (uint32_t) (16 * t + 8 * s) * 256;
                  const uint64_t ah = make_smem_desc(whi_a + aoff, 128, 256), al = make_smem_desc(wlo_a + aoff, 128, 256);
                  const uint64_t bh = make_smem_desc(hhi_a + boff, 128, 256), bl = make_smem_desc(hlo_a + boff, 128, 256);
                  mma_ss(dP, ah, bh, ID_P2A, 0u);
                  mma_ss(dP, ah, bl, ID_P2A, 1u);
                  mma_ss(dP, al, bh, ID_P2A, 1u);
                  mma_commit(&p_full[g]);
                  issue_b();
                  pend_valid = 1; pend_phase = 2; pend_g = g; pend_k = n >> 1; pend_blk = 16 * t + 8 * s;
                  pend_acc = (t > 0 || s > 0); pend_m = m;
                }
              if (!sc.p1(pass) || t == T - 1) { // nothing else will flush it before the accumulators are read
                issue_b();
              }
            }
          }
          if (p2) {
            issue_b();
            mma_commit(acc_full); // wnum complete
            mbar_wait(w_ready, w_cnt & 1); w_cnt++; // W-update done: W_op rewritten
            tc_fence_after();
          }
        }
      }
    }
  } else {
    // =========================================== epilogue warps =========================================
    const int et = tid - 64;               // 0..255
    const int wg = (warp - 2) >> 2;        // epilogue warpgroup 0/1
    const int ew = warp - 2;               // 0..7
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int r = 32 * q + lane;           // row (frame or bin) inside a 128-row tile
    const uint32_t lane_off = (uint32_t) (32 * q) << 16;
    const uint32_t tP = tbase + TM_P + 64 * wg + lane_off;
    const uint32_t tRhi = tbase + TM_R + 64 * wg + lane_off, tRlo = tRhi + 32;
    uint32_t n = 0, acc_cnt = 0;

    // ratio of 32 consecutive columns held in p[] against 32 values v[] -> packed bf16 hi/lo, stored to TMEM
    auto ratio_store = [&](const uint32_t (&p)[32], const float (&v)[32], int h) {
      uint32_t ph[16], pl[16];
#pragma unroll
      for (int j = 0; j < 16; j++) {
        float r0 = __fdividef(v[2 * j], fmaxf(__uint_as_float(p[2 * j]), kEps));
        float r1 = __fdividef(v[2 * j + 1], fmaxf(__uint_as_float(p[2 * j + 1]), kEps));
        uint32_t hi = pack_bf16x2(r0, r1);
        ph[j] = hi;
        pl[j] = pack_bf16x2(r0 - bf16lo_to_f(hi), r1 - bf16hi_to_f(hi));
      }
      tmem_st16(tRhi + 16 * h, ph);
      tmem_st16(tRlo + 16 * h, pl);
    };

    for (int buf = blockIdx.x; buf < d.batch; buf += gridDim.x) {
      // ---------------- buffer prologue: masters + operand copies -----------------------------------------------
      const float* gW = d.W + (int64_t) buf * K * Bp;
      const float* gH = d.H + (int64_t) buf * Fp * K;
      const float* gV = d.V + (int64_t) buf * Fp * Bp;
      for (int e = et; e < K * Bp; e += 256) {
        int k = e / Bp, b = e - k * Bp;
        float w = gW[e];
        if (b < BT) {
          Wm[k * WPITCH + b] = w;
          __nv_bfloat16 hi, lo;
          split_bf16(w, hi, lo);
          whi[wop_index(k, b)] = hi;
          wlo[wop_index(k, b)] = lo;
        } else if (b == BT) {
          WN[k] = w;
        }
      }
      for (int e = et; e < Fp * K; e += 256) {
        float h = gH[e];
        Hm[e] = h;
        __nv_bfloat16 hi, lo;
        split_bf16(h, hi, lo);
        int f = e >> 4, k = e & 15;
        hhi[hop_index(f, k)] = hi;
        hlo[hop_index(f, k)] = lo;
      }
      for (int f = et; f < Fp; f += 256) VN[f] = gV[(int64_t) f * Bp + BT];
      if (et < K) hden[et] = d.hden[(int64_t) buf * K + et];
      fence_proxy_async();
      epi_bar();
      if (lane == 0) mbar_arrive(buf_ready);

      for (int pass = 0; pass < sc.npass; pass++) {
        const bool p1 = sc.p1(pass), p2 = sc.p2(pass);
        for (int t = 0; t < T; t++) {
          // ---------------- phase 1 steps ------------------------------------------------------------------------
          if (p1) {
            for (int c = 0; c < C1; c++, n++) {
              if ((int) (n & 1) != wg) continue;
              const uint32_t st = n % NS;
              mbar_wait(&p_full[wg], (n >> 1) & 1);
              mbar_wait(&v_full[st], (n / NS) & 1);
              tc_fence_after();
              const uint8_t* vt = smem + OFF_V + st * STAGE;
#pragma unroll 1
              for (int h = 0; h < 2; h++) {
                uint32_t p[32];
                tmem_ld32(tP + 32 * h, p);
                float v[32];
                const uint8_t* row = vt + h * 16384 + r * 128;
#pragma unroll
                for (int c4 = 0; c4 < 8; c4++) {
                  float4 x = *reinterpret_cast<const float4*>(row + ((c4 ^ (r & 7)) << 4));
                  v[4 * c4] = x.x; v[4 * c4 + 1] = x.y; v[4 * c4 + 2] = x.z; v[4 * c4 + 3] = x.w;
                }
                tmem_wait_ld();
                ratio_store(p, v, h);
              }
              tmem_wait_st();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) { mbar_arrive(&r_full[wg]); mbar_arrive(&v_empty[st]); }
            }
            mbar_wait(acc_full, acc_cnt & 1); acc_cnt++;
            tc_fence_after();
          }
          // ---------------- tile prep: H-update (if p1), W-denominator / Nyquist partials (if p2) ----------------
          {
            uint32_t hn[16];
            if (p1) { tmem_ld16(tbase + TM_HNUM + lane_off, hn); tmem_wait_ld(); }
            const int f = 128 * t + r;
            float contrib[32];
#pragma unroll
            for (int j = 0; j < 32; j++) contrib[j] = 0.f;
            if ((lane & 1) == wg) { // the two warpgroups split the rows of the tile
              float h[16];
#pragma unroll
              for (int j4 = 0; j4 < 4; j4++) {
                float4 x = *reinterpret_cast<const float4*>(Hm + f * K + 4 * j4);
                h[4 * j4] = x.x; h[4 * j4 + 1] = x.y; h[4 * j4 + 2] = x.z; h[4 * j4 + 3] = x.w;
              }
              const float vn = VN[f];
              if (p1) {
                float pn = 0.f;
#pragma unroll
                for (int k = 0; k < K; k++) pn = fmaf(h[k], WN[k], pn);
                const float rn = vn / fmaxf(pn, kEps);
#pragma unroll
                for (int k = 0; k < K; k++) {
                  float num = fmaf(rn, WN[k], __uint_as_float(hn[k]));
                  h[k] = h[k] * num / fmaxf(hden[k], kEps);      // NMF.hpp:170
                }
                uint32_t ph[8], pl[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                  uint32_t hi = pack_bf16x2(h[2 * j], h[2 * j + 1]);
                  ph[j] = hi;
                  pl[j] = pack_bf16x2(h[2 * j] - bf16lo_to_f(hi), h[2 * j + 1] - bf16hi_to_f(hi));
                }
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++)
                  *reinterpret_cast<float4*>(Hm + f * K + 4 * j4) = make_float4(h[4 * j4], h[4 * j4 + 1], h[4 * j4 + 2], h[4 * j4 + 3]);
#pragma unroll
                for (int kb = 0; kb < KB; kb++) { // one 16-byte row of each core matrix
                  const int idx = ((f >> 3) * KB + kb) * 64 + (f & 7) * 8;
                  *reinterpret_cast<uint4*>(hhi + idx) = make_uint4(ph[4 * kb], ph[4 * kb + 1], ph[4 * kb + 2], ph[4 * kb + 3]);
                  *reinterpret_cast<uint4*>(hlo + idx) = make_uint4(pl[4 * kb], pl[4 * kb + 1], pl[4 * kb + 2], pl[4 * kb + 3]);
                }
              }
              if (p2) {
                float pn = 0.f;
#pragma unroll
                for (int k = 0; k < K; k++) pn = fmaf(h[k], WN[k], pn);
                const float rn = vn / fmaxf(pn, kEps);
#pragma unroll
                for (int k = 0; k < K; k++) { contrib[k] = h[k]; contrib[16 + k] = rn * h[k]; } // wden, Nyquist wnum
              }
            }
            if (p2) {
#pragma unroll
              for (int j = 0; j < 32; j++) {
                float x = contrib[j];
#pragma unroll
                for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                contrib[j] = x;
              }
              if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 32; j++) part[(t * 8 + ew) * 32 + j] = contrib[j];
              }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(prep_ready);
          }
          // ---------------- phase 2 steps ------------------------------------------------------------------------
          if (p2) {
            for (int m = 0; m < MT; m++)
              for (int s = 0; s < 2; s++, n++) {
                if ((int) (n & 1) != wg) continue;
                const uint32_t st = n % NS;
                mbar_wait(&p_full[wg], (n >> 1) & 1);
                mbar_wait(&v_full[st], (n / NS) & 1);
                tc_fence_after();
                const uint8_t* vt = smem + OFF_V + st * STAGE + q * 8192; // box q: bins 128m + 32q .. +31, rows = 64 frames
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                  uint32_t p[32];
                  tmem_ld32(tP + 32 * h, p);
                  float v[32];
#pragma unroll
                  for (int j = 0; j < 32; j++) {
                    const int fr = 32 * h + j;
                    v[j] = *reinterpret_cast<const float*>(vt + fr * 128 + ((((lane >> 2) ^ (fr & 7))) << 4) + ((lane & 3) << 2));
                  }
                  tmem_wait_ld();
                  ratio_store(p, v, h);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&r_full[wg]); mbar_arrive(&v_empty[st]); }
              }
          }
        }
        // ---------------- W-update (NMF.hpp:161-162) + hden refresh (:169) --------------------------------------
        if (p2) {
          mbar_wait(acc_full, acc_cnt & 1); acc_cnt++;
          tc_fence_after();
          epi_bar(); // all partials of all tiles written
          if (et < 32) {
            float s = 0.f;
            for (int i = 0; i < T * 8; i++) s += part[i * 32 + et];
            fin[et] = s; // [0,16) wden, [16,32) Nyquist wnum
          }
          epi_bar();
          float wnew[2][16];
          float ss[16], sm[16], mx = 0.f;
#pragma unroll
          for (int k = 0; k < K; k++) { ss[k] = 0.f; sm[k] = 0.f; }
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const int m = wg + 2 * i;
            if (m < MT) {
              uint32_t wn[16];
              tmem_ld16(tbase + TM_WNUM + 16 * m + lane_off, wn);
              tmem_wait_ld();
              const int b = 128 * m + r;
#pragma unroll
              for (int k = 0; k < K; k++) {
                float w = Wm[k * WPITCH + b] * __uint_as_float(wn[k]) / fmaxf(fin[k], kEps);
                wnew[i][k] = w;
                ss[k] = fmaf(w, w, ss[k]);
                sm[k] += w;
                mx = fmaxf(mx, w);
              }
            }
          }
          float wnq[16]; // Nyquist row, carried by one thread
          if (et == 0) {
#pragma unroll
            for (int k = 0; k < K; k++) {
              float w = WN[k] * fin[16 + k] / fmaxf(fin[k], kEps);
              wnq[k] = w;
              ss[k] = fmaf(w, w, ss[k]);
              sm[k] += w;
              mx = fmaxf(mx, w);
            }
          }
#pragma unroll
          for (int k = 0; k < K; k++) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
              ss[k] += __shfl_xor_sync(0xffffffffu, ss[k], o);
              sm[k] += __shfl_xor_sync(0xffffffffu, sm[k], o);
            }
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K; k++) { red[ew * 36 + k] = ss[k]; red[ew * 36 + 16 + k] = sm[k]; }
            red[ew * 36 + 32] = mx;
          }
          epi_bar();
          if (et < K) {
            float s2 = 0.f, s1 = 0.f, gm = 0.f;
            for (int w = 0; w < 8; w++) { s2 += red[w * 36 + et]; s1 += red[w * 36 + 16 + et]; gm = fmaxf(gm, red[w * 36 + 32]); }
            const bool norm = gm > kEps;                                   // NMF.hpp:162
            const float inv = norm ? (s2 > 0.f ? 1.0f / sqrtf(s2) : 0.f) : 1.0f;
            fin[32 + et] = inv;
            hden[et] = s1 * inv;                                           // sum_b W after normalisation
          }
          epi_bar();
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const int m = wg + 2 * i;
            if (m < MT) {
              const int b = 128 * m + r;
#pragma unroll
              for (int k = 0; k < K; k++) {
                float w = wnew[i][k] * fin[32 + k];
                Wm[k * WPITCH + b] = w;
                __nv_bfloat16 hi, lo;
                split_bf16(w, hi, lo);
                whi[wop_index(k, b)] = hi;
                wlo[wop_index(k, b)] = lo;
              }
            }
          }
          if (et == 0) {
#pragma unroll
            for (int k = 0; k < K; k++) WN[k] = wnq[k] * fin[32 + k];
          }
          fence_proxy_async();
          tc_fence_before();
          epi_bar();
          if (lane == 0) mbar_arrive(w_ready);
        }
      }
      // ---------------- buffer epilogue: masters back to global --------------------------------------------------
      epi_bar();
      float* oW = d.W + (int64_t) buf * K * Bp;
      float* oH = d.H + (int64_t) buf * Fp * K;
      for (int e = et; e < K * Bp; e += 256) {
        int k = e / Bp, b = e - k * Bp;
        oW[e] = b < BT ? Wm[k * WPITCH + b] : (b == BT ? WN[k] : 0.f);
      }
      for (int e = et; e < Fp * K; e += 256) oH[e] = Hm[e];
      epi_bar(); // masters are overwritten by the next buffer's prologue
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------------------------
bool tc_eligible(const NmfDev& d)
{
  const int BT = d.B - 1;
  return d.K == 16 && d.KP == 16 && BT >= 128 && BT <= BT_MAX && (BT % 128) == 0 && d.Fp <= FP_MAX && (d.Fp % 128) == 0 &&
         d.Bp == d.B + 3 && !d.clamp_v && !d.shared_w;
}

int32_t tc_run(Plan* p, const NmfDev& d, int iters, bool upd_w, bool upd_h)
{
  alignas(64) CUtensorMap tmap1, tmap2;
  FB_TRY(make_v_tensor_map(p, &tmap1, d.V, d.Bp, d.Fp, d.batch, 128));
  FB_TRY(make_v_tensor_map(p, &tmap2, d.V, d.Bp, d.Fp, d.batch, 64));
  if (!(p->attr_mask & 0x10000u)) {
    FB_CUDA(p, cudaFuncSetAttribute(k_nmf_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    p->attr_mask |= 0x10000u;
  }
  int grid = std::min(d.batch, p->sm_count);
  if (p->kev.size() < p->kev_used + 2) {
    while (p->kev.size() < p->kev_used + 2) { cudaEvent_t e; cudaEventCreate(&e); p->kev.push_back(e); }
  }
  cudaEventRecord(p->kev[p->kev_used++], p->stream);
  k_nmf_tc<<<grid, NTHREADS, SMEM_BYTES, p->stream>>>(d, tmap1, tmap2, iters, upd_w ? 1 : 0, upd_h ? 1 : 0);
  cudaEventRecord(p->kev[p->kev_used++], p->stream);
  p->launches++; p->launches_nmf++;
  FB_CUDA(p, cudaGetLastError());
  return FB200_OK;
}

} // namespace fb200
