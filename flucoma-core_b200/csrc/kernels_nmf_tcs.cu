// Streamed tcgen05 / TMEM / TMA engine for the NMF multiplicative updates (algorithms/public/NMF.hpp:144-183 and the
// fixed-dictionary activation solve NMF.hpp:45-89), rank 9..64 (padded to 16, 32 or 64), any bins = 128 m + 1, any frame count.
//
// Same arithmetic as kernels_nmf_tc.cu (W.H from bf16 split operands with the terms down to 2^-16, ratio V / max(WH, eps)
// computed by the epilogue warps straight out of TMEM and fed back as the TMEM A operand of the second MMA, every step's
// second MMA into a FRESH accumulator that the epilogue sums with round-to-nearest adds), but no operand is resident:
// the exact fp32 state (d.W, d.H) lives in global memory (L2) next to its two-part bf16 split in UMMA core-matrix layout
//     Wop [BT/8 block rows][2 parts][K/8][8 k][8 b] bf16      Hop [Fp/8 block rows][2 parts][K/8][8 f][8 k] bf16
// and one persistent CTA per buffer streams the operands through shared memory with bulk TMA copies.
// Every half-iteration is a sequence of JOBS that all look alike:
//     job(H, t): H-update of the 128-frame tile t  (NMF.hpp:165-170)   stationary = H rows of t,  stream = 64-bin W chunks
//     job(W, m): W-update of the 128-bin  tile m  (NMF.hpp:158-161)   stationary = W rows of m,  stream = 64-frame H chunks
//   step (64 columns):  P = stationary x chunk        tcgen05.mma SS  M128 N64, K/16 k-steps x 3 split terms (hh, hm, mh)
//                       R = V / max(P, eps)           epilogue warps: tcgen05.ld, swizzled V tile, rcp, 2-way split, tcgen05.st
//                       num += R x chunk^T            tcgen05.mma TS  M128 N = 2K (R_hi [X_hi | X_mid]) + N = K (R_lo X_hi), 4 k-steps
//   the fresh partial of every step is added (round-to-nearest) to running sums that live in TMEM as well;
//   end of job:  H-tile / W-tile update from the fp32 state, written back as fp32 + split operand.
// After the last W tile of a half-iteration: column normalisation over all bins (NMF.hpp:162), rescale sweep, hden (:169).
//
// TMEM (512 columns): P 2 x 64 | R 2 x 64 | accumulators | running sums.
//   rank 16 / 32: one accumulator (3K columns) and one set of sums (K) per epilogue warpgroup; a warpgroup collects its own steps.
//   rank 64:      ONE accumulator (192 columns) and one set of sums (64): both warpgroups collect EVERY step, each its 32
//                 components, and the single second-MMA issuer waits for both before it reuses the accumulator.
//
// Fixed dictionary (update_w == 0: NMFMatch / NMFFilter / BufNMF with fixed bases): a CTA takes PAIRS of frame tiles and
// runs all iterations for them (job(H, a), job(H, b), job(H, a), ...), so the |X| tile is read from HBM once and stays in
// L2 for the remaining iterations, and the update of tile a overlaps the steps of tile b.
//
// Warp roles (384 threads) as in kernels_nmf_tc.cu: warp 0 producer (V tensor tiles + operand bulk copies), warp 1 first
// MMA issuer + TMEM owner, warps 2/3 second-MMA issuers (one per epilogue warpgroup; rank 64: warp 2 alone), warps 4-11 two
// epilogue warpgroups on alternating steps.  All sums are fixed-order: results are bitwise repeatable.
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace fb200 {
using namespace tc;

namespace tcs {
constexpr int NS = 3;             // V ring stages
constexpr int STAGE = 32768;      // 128 x 64 fp32
constexpr int NSO = 3;            // streamed-operand ring stages
constexpr int NTHREADS = 384;

template <int K>
struct Cfg {
  static constexpr int KB = K / 8;                      // 8-component blocks
  static constexpr int KS = K / 16;                     // k-steps of the first MMA
  static constexpr uint32_t ROWB = 2 * KB * 128;        // bytes per 8-row block row (2 parts: hi, mid)
  static constexpr uint32_t CHUNK = 8 * ROWB;           // streamed chunk: 64 rows
  static constexpr uint32_t TILE = 16 * ROWB;           // stationary tile: 128 rows
  // TMEM columns
  static constexpr bool ONE_ACC = K == 64;              // a single accumulator / set of sums shared by both warpgroups
#ifndef FB200_TCS_GROUP
#define FB200_TCS_GROUP 2
#endif
  // rank 64: consecutive steps whose second MMAs accumulate in TMEM before the epilogue collects them.  Collecting is 192
  // columns per thread and stalls the one accumulator; the tensor core's truncating accumulate is the price (a chain of
  // 4 GROUP k-steps instead of 4): measured errors in profiles/r02k_experiments.txt.
  static constexpr int GROUP = ONE_ACC ? FB200_TCS_GROUP : 1;
  static constexpr uint32_t TM_P = 0;                   // + 64 g
  static constexpr uint32_t TM_R = 128;                 // + 64 g : hi [0,32) lo [32,64)
  static constexpr uint32_t ACOLS = 3 * K;              // [0,K) R_hi X_hi | [K,2K) R_hi X_mid | [2K,3K) R_lo X_hi
  static constexpr uint32_t TM_ACC = 256;               // + ACOLS g (rank < 64)
  static constexpr uint32_t ASTRIDE = K == 16 ? 64 : ACOLS; // accumulator g starts at TM_ACC + ASTRIDE g
  static constexpr uint32_t TM_SUM = TM_ACC + ACOLS;    // rank 64 only: running sums, K columns
  static_assert(ONE_ACC ? TM_SUM + K <= 512 : TM_ACC + ASTRIDE + ACOLS <= 512, "TMEM budget");
  // shared memory map
  static constexpr int OFF_V = 0;
  static constexpr int OFF_ST = OFF_V + NS * STAGE;                 // 2 stationary tiles
  static constexpr int OFF_O = OFF_ST + 2 * (int) TILE;             // NSO streamed chunks
  // fp32 state of the stationary tile (the tile update multiplies the exact old values): staged by the producer next to
  // the operand tile where shared memory allows it; rank 64 has no room and reads the rows from L2 instead (its jobs are
  // 32 steps long, the round trip is 1-2 % of a job; at rank 16 a job is 8 steps and the round trip cost 5-15 %)
  static constexpr bool HAS_F32 = K <= 32;
  static constexpr uint32_t F32TILE = 128 * K * 4;
  static constexpr int OFF_F32 = OFF_O + NSO * (int) CHUNK;         // 2 x float [128 frames][K] (H job) / [K][128 bins] (W job)
  // running sums of the partial numerators: rank <= 32 in shared memory (float [K/4][2 wg][128][4], one 16-byte slot per
  // thread and component quad: conflict-free), rank 64 in TMEM (no shared memory left; measured on rank 16 / 32 the TMEM
  // variant is 3-7 % slower: the epilogue's tcgen05.ld/st compete with the MMAs for the TMEM port)
  static constexpr int OFF_HS = OFF_F32 + (HAS_F32 ? 2 * (int) F32TILE : 0);
  static constexpr int OFF_PART = OFF_HS + (ONE_ACC ? 0 : (K / 4) * 2 * 128 * 16); // float [8 warps][K]  wden / Nyquist partials
  static constexpr int OFF_RED = OFF_PART + 8 * K * 4;              // float [8 warps][K + 4]  sum w^2 | sum w (own half), max
  static constexpr int OFF_FIN = OFF_RED + 8 * (K + 4) * 4;         // float [8][K]: wden, nyq num, inv norm, inv hden, WN, hden, sum w^2, sum w
  static constexpr int OFF_BAR = OFF_FIN + 8 * K * 4 + 16;
  static constexpr int NBAR = 2 * NS + 2 * NSO + 4 + 8 + 1 + 2;
  static constexpr int OFF_SLOT = OFF_BAR + NBAR * 8;
  static constexpr int SMEM_BYTES = OFF_SLOT + 16;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// N consecutive TMEM columns of this thread's lane
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, uint32_t (&r)[N])
{
  if constexpr (N == 8) tmem_ld8(taddr, r);
  else tmem_ld16(taddr, r);
}
template <int N>
__device__ __forceinline__ void tmem_stn(uint32_t taddr, const uint32_t (&r)[N])
{
  if constexpr (N == 8) tmem_st8(taddr, r);
  else tmem_st16(taddr, r);
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 88;" ::: "memory"); }
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// 1-D bulk copy global -> shared, completion on an mbarrier (UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// operand element index (bf16 units): row = bin (W) or frame (H), k = component
template <int K>
__device__ __forceinline__ int op_index_w(int part, int k, int b) { return ((((b >> 3) * 2 + part) * (K / 8) + (k >> 3)) << 6) + ((k & 7) << 3) + (b & 7); }
template <int K>
__device__ __forceinline__ int op_index_h(int part, int f, int k) { return ((((f >> 3) * 2 + part) * (K / 8) + (k >> 3)) << 6) + ((f & 7) << 3) + (k & 7); }
// two-part split of a pair, packed pairwise: hi = bf16(x), mid = bf16(x - hi)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& h, uint32_t& m)
{
  h = cvt2(x0, x1);
  m = cvt2(x0 - bf16lo_to_f(h), x1 - bf16hi_to_f(h));
}

// one stage of the vector butterfly: lanes with bit O set keep the upper half of a[0, 2 CNT), the others the lower half
template <int CNT, int O>
__device__ __forceinline__ void bf_stage(float* a, int lane)
{
  if constexpr (O >= 1) {
    if constexpr (CNT >= 1) {
      const bool up = (lane & O) != 0;
#pragma unroll
      for (int i = 0; i < CNT; i++) {
        const float send = up ? a[i] : a[i + CNT];
        const float keep = up ? a[i + CNT] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
      }
      bf_stage<CNT / 2, O / 2>(a, lane);
    } else {
      a[0] += __shfl_xor_sync(0xffffffffu, a[0], O); // fewer values than lanes: plain sum over the remaining bits
      bf_stage<0, O / 2>(a, lane);
    }
  }
}

struct Params {
  const float* V;      // [batch][Fp][Bp]
  float* W;            // [batch | 1][K][Bp]   fp32 state
  float* H;            // [batch][Fp][K]
  float* hden;         // [batch | 1][K]
  __nv_bfloat16* Wop;  // [batch | 1][BT/8][2][K/8][64]
  __nv_bfloat16* Hop;  // [batch][Fp/8][2][K/8][64]
  int batch, Fp, Bp, BT;
  int iters, upd_w, upd_h, shared_w, clamp_v;
  int units;           // work units: buffers (update_w) or (buffer, tile pair) (fixed W)
  long long* dbg;      // developer timeline (FB200_TCS_DEBUG_TIMELINE builds only)
};
// developer timeline: epilogue thread 0 of CTA 0 records clock64() at the boundaries of its first 96 jobs
#ifdef FB200_TCS_DEBUG_TIMELINE
#define TCS_MARK(slot) do { if (p.dbg && blockIdx.x == 0 && et == 0 && jn < 96) p.dbg[jn * 8 + (slot)] = clock64(); } while (0)
#else
#define TCS_MARK(slot) do { } while (0)
#endif

// The job sequence of one work unit, identical for every warp role: f(phase, tile, need_st, need_c, per_tile) with phase
// 0 = H job, 1 = W job.  The need_* values tell the producer how many jobs of this CTA (counted over all units, `jn`)
// must have published their tile update before it may copy operands that those updates wrote:
//   need_st            the stationary tile (only the fixed-W mode re-reads a tile it has just updated; with W updates the
//                      stationary tiles of a half-iteration were last written a whole half-iteration earlier)
//   need_c + (per_tile ? i / 2 : 0)   the streamed chunk of step i: W chunks need the whole preceding W half-iteration
//                      (the column normalisation rescales every bin), the H chunk of frames [64 i, 64 i + 64) needs only
//                      the H job of tile i / 2 -- so a W half-iteration starts while the H half-iteration before it drains.
template <class F>
__device__ __forceinline__ void for_jobs(const Params& p, int unit, uint32_t& jn, F&& f)
{
  const int T = p.Fp / 128, MT = p.BT / 128;
  if (!p.upd_w) { // fixed W: unit = (buffer, tile pair); all iterations for the pair
    const int pairs = (T + 1) / 2;
    const int t0 = 2 * (unit % pairs);
    const int nt = (t0 + 1 < T) ? 2 : 1;
    for (int it = 0; it < p.iters; it++)
      for (int i = 0; i < nt; i++) { f(0, t0 + i, it == 0 ? 0u : jn - (uint32_t) (nt - 1), 0u, false); jn++; }
    return;
  }
  uint32_t hjob0 = 0;
  bool have_h = false; // an H half-iteration of THIS unit precedes
  for (int it = 0; it < p.iters; it++) {
    // H fixed: W half-iterations follow each other directly, the W tiles were rescaled by the one just before
    const uint32_t w_st = (p.upd_h || it == 0) ? 0u : jn;
    for (int m = 0; m < MT; m++) { f(1, m, w_st, have_h ? hjob0 + 1 : 0u, have_h); jn++; }
    if (p.upd_h) {
      const uint32_t n1 = jn;
      hjob0 = n1; have_h = true;
      for (int t = 0; t < T; t++) { f(0, t, 0u, n1, false); jn++; }
    }
  }
}
__device__ __forceinline__ int unit_buffer(const Params& p, int unit) { return p.upd_w ? unit : unit / ((p.Fp / 128 + 1) / 2); }
} // namespace tcs

using namespace tcs;

// fp32 state -> split operands (once per call; the engine keeps both in step afterwards)
template <int K>
__global__ void k_tcs_pack(const float* __restrict__ W, const float* __restrict__ H, __nv_bfloat16* __restrict__ Wop,
                           __nv_bfloat16* __restrict__ Hop, int nw, int batch, int Fp, int Bp, int BT)
{
  constexpr int KB = K / 8;
  const int64_t w_items = (int64_t) nw * (BT / 8) * K, h_items = (int64_t) batch * Fp * KB;
  for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < w_items + h_items; i += (int64_t) gridDim.x * blockDim.x) {
    float x[8];
    __nv_bfloat16* dst;
    if (i < w_items) { // 8 consecutive bins of one component
      const int k = (int) (i % K);
      const int64_t r = i / K;
      const int blk = (int) (r % (BT / 8)), buf = (int) (r / (BT / 8));
      const float4* src = reinterpret_cast<const float4*>(W + ((int64_t) buf * K + k) * Bp + 8 * blk);
      const float4 a = src[0], b = src[1];
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      dst = Wop + (int64_t) buf * (BT / 8) * 2 * KB * 64 + op_index_w<K>(0, k, 8 * blk);
    } else { // 8 consecutive components of one frame
      const int64_t j = i - w_items;
      const int kb = (int) (j % KB);
      const int64_t r = j / KB;
      const int f = (int) (r % Fp), buf = (int) (r / Fp);
      const float4* src = reinterpret_cast<const float4*>(H + ((int64_t) buf * Fp + f) * K + 8 * kb);
      const float4 a = src[0], b = src[1];
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      dst = Hop + (int64_t) buf * (Fp / 8) * 2 * KB * 64 + op_index_h<K>(0, f, 8 * kb);
    }
    uint32_t ph[4], pm[4];
#pragma unroll
    for (int q = 0; q < 4; q++) split2(x[2 * q], x[2 * q + 1], ph[q], pm[q]);
    *reinterpret_cast<uint4*>(dst) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(dst + KB * 64) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
  }
}

template <int K>
__global__ void __launch_bounds__(NTHREADS, 1)
k_nmf_tcs(Params p, const __grid_constant__ CUtensorMap tmap1, const __grid_constant__ CUtensorMap tmap2)
{
  using C = Cfg<K>;
  constexpr int KB = C::KB, KS = C::KS;
  constexpr uint32_t ROWB = C::ROWB;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* part = reinterpret_cast<float*>(smem + C::OFF_PART);
  float* red = reinterpret_cast<float*>(smem + C::OFF_RED);
  float* fin = reinterpret_cast<float*>(smem + C::OFF_FIN);
  float* f_wden = fin;             // [K] 1 / max(sum_f H, eps)
  float* f_nyq = fin + K;          // [K] Nyquist numerator
  float* f_inv = fin + 2 * K;      // [K] 1 / column norm
  float* f_ihd = fin + 3 * K;      // [K] 1 / max(hden, eps)
  float* WN = fin + 4 * K;         // [K] Nyquist row of W
  float* f_hden = fin + 5 * K;     // [K]
  float* f_s2 = fin + 6 * K;       // [K] column sums of w^2 / w of the un-normalised update
  float* f_s1 = fin + 7 * K;
  float* f_gm = fin + 8 * K;       // [1] max over the tensor bins
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* v_full = bars;                 // [NS]
  uint64_t* v_empty = v_full + NS;         // [NS]
  uint64_t* o_full = v_empty + NS;         // [NSO]
  uint64_t* o_empty = o_full + NSO;        // [NSO]  both MMAs that read the chunk have completed
  uint64_t* st_full = o_empty + NSO;       // [2]
  uint64_t* st_empty = st_full + 2;        // [2]    all first MMAs of the job have completed
  uint64_t* p_full = st_empty + 2;         // [2]
  uint64_t* r_full = p_full + 2;           // [2]
  uint64_t* b_full = r_full + 2;           // [2]
  uint64_t* p_free = b_full + 2;           // [2]
  uint64_t* acc_free = p_free + 2;         // [1]    rank 64: both warpgroups have collected the group's partial
  uint64_t* r_free = acc_free + 1;         // [2]    rank 64: the second MMA that read R[g] has completed
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + C::OFF_SLOT);
  volatile uint32_t* jobs_done = slot + 1; // jobs whose tile update is complete and visible in global memory

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Fp = p.Fp, Bp = p.Bp, BT = p.BT;
  const int C1 = BT / 64, S2 = Fp / 64;

  if (tid == 0) {
    slot[1] = 0u;
    for (int i = 0; i < NS; i++) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 4); }
    for (int i = 0; i < NSO; i++) { mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 2); }
    for (int i = 0; i < 2; i++) {
      mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], C::HAS_F32 ? 9 : 1); // first-MMA commit of the job's last step (+ the 8 epilogue warps that read the fp32 tile)
      mbar_init(&p_full[i], 1); mbar_init(&r_full[i], 4); mbar_init(&b_full[i], 1); mbar_init(&p_free[i], 4);
    }
    mbar_init(acc_free, 8);
    mbar_init(&r_free[0], 1); mbar_init(&r_free[1], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmap1);
    tma_prefetch_desc(&tmap2);
  }
  if (warp == 1) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *slot;
  const int64_t wop_stride = p.shared_w ? 0 : (int64_t) (BT / 8) * 2 * KB * 64;
  const int64_t hop_stride = (int64_t) (Fp / 8) * 2 * KB * 64;

  if (warp < 4) {
    reg_dec();
    if (warp == 0) {
      // =========================================== producer =================================================
      if (lane == 0) {
        uint32_t n = 0, jn = 0;
        for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
          const int buf = unit_buffer(p, unit);
          const __nv_bfloat16* gW = p.Wop + buf * wop_stride;
          const __nv_bfloat16* gH = p.Hop + buf * hop_stride;
          for_jobs(p, unit, jn, [&](int phase, int tile, uint32_t need_st, uint32_t need_c, bool per_tile) {
            const int ns = phase == 0 ? C1 : S2;
            const uint32_t n0 = n;
            auto issue_v = [&](int i) { // V tile of step i
              const uint32_t nn = n0 + i, st = nn % NS, k = nn / NS;
              mbar_wait(&v_empty[st], (k & 1) ^ 1);
              mbar_arrive_expect_tx(&v_full[st], STAGE);
              uint8_t* dst = smem + C::OFF_V + st * STAGE;
              if (phase == 0) { // [128 frames][64 bins] as two 32-bin boxes
                tma_load_3d(dst, &tmap1, 64 * i, 128 * tile, buf, &v_full[st]);
                tma_load_3d(dst + 16384, &tmap1, 64 * i + 32, 128 * tile, buf, &v_full[st]);
              } else { // [64 frames][128 bins] as four 32-bin boxes
#pragma unroll
                for (int w = 0; w < 4; w++) tma_load_3d(dst + w * 8192, &tmap2, 128 * tile + 32 * w, 64 * i, buf, &v_full[st]);
              }
            };
            { // stationary tile: 128 rows of H (H job) / W (W job)
              if (need_st) {
                while (*jobs_done < need_st) {}
                fence_async_all(); // the tile update was written with generic stores; the copy reads through the async proxy
              }
              const uint32_t sl = jn & 1, k = jn >> 1;
              mbar_wait(&st_empty[sl], (k & 1) ^ 1);
              mbar_arrive_expect_tx(&st_full[sl], C::TILE + (C::HAS_F32 ? C::F32TILE : 0u));
              const __nv_bfloat16* src = (phase == 0 ? gH : gW) + (int64_t) tile * 16 * 2 * KB * 64;
              bulk_g2s(smem + C::OFF_ST + sl * C::TILE, src, C::TILE, &st_full[sl]);
              if constexpr (C::HAS_F32) {
                uint8_t* dst = smem + C::OFF_F32 + sl * C::F32TILE;
                if (phase == 0) {
                  bulk_g2s(dst, p.H + ((int64_t) buf * Fp + 128 * tile) * K, C::F32TILE, &st_full[sl]);
                } else {
                  const float* wrow = p.W + (int64_t) (p.shared_w ? 0 : buf) * K * Bp + 128 * tile;
                  for (int k = 0; k < K; k++) bulk_g2s(dst + k * 512, wrow + (int64_t) k * Bp, 512, &st_full[sl]);
                }
              }
            }
            // |X| does not depend on anything: the first tiles of the job are requested before the producer blocks on the
            // chunk dependencies, so a half-iteration boundary costs one L2 round trip of a chunk, not a cold pipeline
            const int pre = ns < NS ? ns : NS;
            for (int i = 0; i < pre; i++) issue_v(i);
            uint32_t seen = 0;
            for (int i = 0; i < ns; i++, n++) {
              if (i >= pre) issue_v(i);
              { // streamed operand chunk: 64 rows of W (H job) / H (W job)
                const uint32_t need = need_c + (per_tile ? (uint32_t) (i >> 1) : 0u);
                if (seen < need) {
                  while ((seen = *jobs_done) < need) {}
                  fence_async_all();
                }
                const uint32_t so = n % NSO, k = n / NSO;
                mbar_wait(&o_empty[so], (k & 1) ^ 1);
                mbar_arrive_expect_tx(&o_full[so], C::CHUNK);
                const __nv_bfloat16* src = (phase == 0 ? gW : gH) + (int64_t) i * 8 * 2 * KB * 64;
                bulk_g2s(smem + C::OFF_O + so * C::CHUNK, src, C::CHUNK, &o_full[so]);
              }
            }
          });
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // =========================================== first-MMA issuer =========================================
      constexpr uint32_t ID_H = make_idesc_bf16(128, 64, 0, 1); // A = H tile (K-major), B = W chunk (MN-major)
      constexpr uint32_t ID_W = make_idesc_bf16(128, 64, 1, 0); // A = W tile (MN-major), B = H chunk (K-major)
      constexpr uint32_t HI_A = (ROWB >> 4) | (1u << 14);       // SBO = ROWB (block rows), descriptor version 1
      constexpr uint32_t LO_A = (128u >> 4) << 16;              // LBO = 128 (component blocks along K)
      constexpr uint32_t PSTEP = (KB * 128) >> 4;               // one split part
      const uint32_t st_a = smem_u32(smem + C::OFF_ST), o_a = smem_u32(smem + C::OFF_O);
      uint32_t n = 0, jn = 0;
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        for_jobs(p, unit, jn, [&](int phase, int, uint32_t, uint32_t, bool) {
          const uint32_t sl = jn & 1;
          mbar_wait(&st_full[sl], (jn >> 1) & 1);
          const uint32_t alo = ((st_a + sl * C::TILE) >> 4) | LO_A;
          const uint32_t idesc = phase == 0 ? ID_H : ID_W;
          const int ns = phase == 0 ? C1 : S2;
          for (int i = 0; i < ns; i++, n++) {
            const uint32_t g = n & 1, so = n % NSO;
            if (n >= 2) mbar_wait(&p_free[g], ((n - 2) >> 1) & 1);
            mbar_wait(&o_full[so], (n / NSO) & 1);
            tc_fence_after();
            const uint32_t blo = ((o_a + so * C::CHUNK) >> 4) | LO_A;
            const uint32_t dP = tbase + C::TM_P + 64 * g;
#pragma unroll
            for (int j = 0; j < KS; j++) { // 16 components per k-step = two component blocks
              const uint32_t a = alo + 16 * j, b = blo + 16 * j;
              if (j == 0) mma_ss_lohi<0>(dP, a, HI_A, b, HI_A, idesc);           // hi  hi
              else mma_ss_lohi<1>(dP, a, HI_A, b, HI_A, idesc);
              mma_ss_lohi<1>(dP, a, HI_A, b + PSTEP, HI_A, idesc);               // hi  mid
              mma_ss_lohi<1>(dP, a + PSTEP, HI_A, b, HI_A, idesc);               // mid hi   (terms below 2^-16: see kernels_nmf_tc.cu)
            }
            mma_commit_warp(&p_full[g]);
            mma_commit_warp(&o_empty[so]);
          }
          mma_commit_warp(&st_empty[sl]);
        });
      }
    } else {
      // =========================================== second-MMA issuers =========================================
      // rank < 64: one issuing warp per epilogue warpgroup (accumulator g belongs to the steps of parity g);
      // rank 64:   warp 2 issues every step into the one accumulator, after both warpgroups have collected the previous one.
      const uint32_t myg = warp - 2;
      constexpr uint32_t ID_H2 = make_idesc_bf16(128, 2 * K, 0, 0), ID_H1 = make_idesc_bf16(128, K, 0, 0); // B = W chunk, K-major
      constexpr uint32_t ID_W2 = make_idesc_bf16(128, 2 * K, 0, 1), ID_W1 = make_idesc_bf16(128, K, 0, 1); // B = H chunk, MN-major
      constexpr uint32_t HI_B = (128u >> 4) | (1u << 14);  // SBO = 128 (component blocks / parts along N)
      constexpr uint32_t LO_B = (ROWB >> 4) << 16;         // LBO = ROWB (block rows along K)
      constexpr uint32_t RSTEP = ROWB >> 4;
      const uint32_t o_a = smem_u32(smem + C::OFF_O);
      uint32_t n = 0, jn = 0, grp = 0;
      if (!C::ONE_ACC || myg == 0)
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        for_jobs(p, unit, jn, [&](int phase, int, uint32_t, uint32_t, bool) {
          const uint32_t id2 = phase == 0 ? ID_H2 : ID_W2, id1 = phase == 0 ? ID_H1 : ID_W1;
          const int ns = phase == 0 ? C1 : S2;
          for (int i = 0; i < ns; i++, n++) {
            const uint32_t g = n & 1, so = n % NSO;
            if (!C::ONE_ACC && g != myg) continue;
            // r_full FIRST: it implies o_full(n) (the first MMA of this step waited for it), so the wait below can never
            // be a phase early; it only orders this thread behind the copy
            mbar_wait(&r_full[g], (n >> 1) & 1);
            mbar_wait(&o_full[so], (n / NSO) & 1);
            const bool gstart = (i % C::GROUP) == 0, gend = (i % C::GROUP) == C::GROUP - 1 || i == ns - 1;
            if (C::ONE_ACC && gstart && grp >= 1) mbar_wait(acc_free, (grp - 1) & 1);
            tc_fence_after();
            const uint32_t blo = ((o_a + so * C::CHUNK) >> 4) | LO_B;
            const uint32_t rbase = tbase + C::TM_R + 64 * g;
            const uint32_t dacc = tbase + C::TM_ACC + (C::ONE_ACC ? 0u : C::ASTRIDE * g);
#pragma unroll
            for (int j = 0; j < 4; j++) { // 16 rows of the chunk per k-step
              const uint32_t b0 = blo + 2 * j * RSTEP;
              const uint32_t rh = rbase + 8 * j;
              if (j == 0 && gstart) mma_ts_lohi<0>(dacc, rh, b0, HI_B, id2);      // R_hi [X_hi | X_mid]
              else mma_ts_lohi<1>(dacc, rh, b0, HI_B, id2);
              if (j == 0 && gstart) mma_ts_lohi<0>(dacc + 2 * K, rh + 32, b0, HI_B, id1); // R_lo  X_hi
              else mma_ts_lohi<1>(dacc + 2 * K, rh + 32, b0, HI_B, id1);
            }
            if constexpr (C::ONE_ACC) {
              mma_commit_warp(&r_free[g]);
              if (gend) { mma_commit_warp(&b_full[0]); grp++; }
            } else {
              mma_commit_warp(&b_full[g]);
            }
            mma_commit_warp(&o_empty[so]);
          }
        });
      }
    }
  } else {
    reg_inc();
    // =========================================== epilogue warps ===============================================
    const int et = tid - 128;              // 0..255
    const int wg = (warp - 4) >> 2;        // epilogue warpgroup 0/1
    const int ew = warp - 4;               // 0..7
    const int q = warp & 3;                // TMEM lane quarter
    const int r = 32 * q + lane;           // row inside a 128-row tile
    const uint32_t lane_off = (uint32_t) (32 * q) << 16;
    const uint32_t tP = tbase + C::TM_P + 64 * wg + lane_off;
    const uint32_t tR = tbase + C::TM_R + 64 * wg + lane_off;
    constexpr int K2 = K / 2;              // components this warpgroup stores / reduces in the tile updates
    constexpr int DC = C::ONE_ACC ? K2 : K; // components a thread collects from a step's partial
    constexpr int G = K2 >= 16 ? 16 : 8;   // TMEM access granule of the tile updates
    const int k0 = wg * K2;
    // this thread's slice of the accumulator / of the running sums: rank 64: components [32 wg, 32 wg + 32) of the shared
    // accumulator; rank < 64: all components of this warpgroup's own accumulator
    const uint32_t tAcc = tbase + C::TM_ACC + (C::ONE_ACC ? (uint32_t) (K2 * wg) : C::ASTRIDE * wg) + lane_off;
    const uint32_t tSum = tbase + C::TM_SUM + (uint32_t) (K2 * wg) + lane_off; // rank 64
    const uint32_t tSumAll = tbase + C::TM_SUM + lane_off;
    float* hs = reinterpret_cast<float*>(smem + C::OFF_HS);          // rank <= 32
    float4* hp = reinterpret_cast<float4*>(hs) + (wg * 128 + r);     // [j4][wg][row]
    uint32_t n = 0, jn = 0, grp = 0;
    int out_valid = 0, out_first = 0;
    uint32_t out_par = 0;

    // Add a step's fresh partial (three column groups per component) to the running sums with round-to-nearest adds.
    // rank 64: sums in TMEM; their stores are left in flight, the next tcgen05.wait::st of this thread covers them.
    auto collect = [&](bool first) {
      if constexpr (C::ONE_ACC) {
        if (!first) tmem_wait_st(); // this thread's previous update of the same columns
      }
#pragma unroll
      for (int g16 = 0; g16 < DC / 16; g16++) {
        uint32_t a0[16], a1[16], a2[16], sacc[16];
        tmem_ld16(tAcc + 16 * g16, a0);           // R_hi X_hi
        tmem_ld16(tAcc + K + 16 * g16, a1);       // R_hi X_mid
        tmem_ld16(tAcc + 2 * K + 16 * g16, a2);   // R_lo X_hi
        if constexpr (C::ONE_ACC) {
          if (!first) tmem_ld16(tSum + 16 * g16, sacc);
        }
        tmem_wait_ld();
        if constexpr (C::ONE_ACC) {
#pragma unroll
          for (int k = 0; k < 16; k += 2) {
            float x0, x1;
            add2(x0, x1, __uint_as_float(a1[k]), __uint_as_float(a1[k + 1]), __uint_as_float(a2[k]), __uint_as_float(a2[k + 1]));
            add2(x0, x1, __uint_as_float(a0[k]), __uint_as_float(a0[k + 1]), x0, x1);
            if (!first) add2(x0, x1, __uint_as_float(sacc[k]), __uint_as_float(sacc[k + 1]), x0, x1);
            sacc[k] = __float_as_uint(x0); sacc[k + 1] = __float_as_uint(x1);
          }
          tmem_st16(tSum + 16 * g16, sacc);
        } else {
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            float x[4];
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
              const int k = 4 * j4 + i;
              add2(x[i], x[i + 1], __uint_as_float(a1[k]), __uint_as_float(a1[k + 1]), __uint_as_float(a2[k]), __uint_as_float(a2[k + 1]));
              add2(x[i], x[i + 1], __uint_as_float(a0[k]), __uint_as_float(a0[k + 1]), x[i], x[i + 1]);
            }
            float4* dst = hp + (4 * g16 + j4) * 256;
            if (!first) {
              const float4 h = *dst;
              add2(x[0], x[1], h.x, h.y, x[0], x[1]);
              add2(x[2], x[3], h.z, h.w, x[2], x[3]);
            }
            *dst = make_float4(x[0], x[1], x[2], x[3]);
          }
        }
      }
    };
    // rank < 64: the previous step of this warpgroup
    auto drain = [&]() {
      if (!out_valid) return;
      mbar_wait(&b_full[wg], out_par);
      tc_fence_after();
      collect(out_first != 0);
      out_valid = 0;
    };
    // rank 64: the group of steps that has just ended (both warpgroups, each its 32 components); hands the accumulator
    // back to the second-MMA issuer
    auto drain_shared = [&](bool first) {
      mbar_wait(&b_full[0], grp & 1);
      grp++;
      tc_fence_after();
      collect(first);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
    };
    uint32_t voff[8];
#pragma unroll
    for (int x = 0; x < 8; x++) voff[x] = (uint32_t) ((((lane >> 2) ^ x) << 4) + ((lane & 3) << 2));

    auto do_step = [&](uint32_t nn, bool ph_h) {
      const uint32_t st = nn % NS;
      // p_full FIRST, v_full afterwards (see kernels_nmf_tc.cu: the V ring is the one barrier whose consecutive phases are
      // waited for by different warpgroups)
      mbar_wait(&p_full[wg], (nn >> 1) & 1);
      mbar_wait(&v_full[st], (nn / NS) & 1);
      tc_fence_after();
      uint32_t pp[64];
      tmem_ld32(tP, *reinterpret_cast<uint32_t(*)[32]>(&pp[0]));
      tmem_ld32(tP + 32, *reinterpret_cast<uint32_t(*)[32]>(&pp[32]));
      float v[64];
      if (ph_h) { // two boxes [128 frames][32 bins]: this thread's frame row, 16-byte chunks un-swizzled
        const uint8_t* row = smem + C::OFF_V + st * STAGE + r * 128;
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int c4 = 0; c4 < 8; c4++) {
            const float4 x = *reinterpret_cast<const float4*>(row + h * 16384 + ((c4 ^ (r & 7)) << 4));
            v[32 * h + 4 * c4] = x.x; v[32 * h + 4 * c4 + 1] = x.y; v[32 * h + 4 * c4 + 2] = x.z; v[32 * h + 4 * c4 + 3] = x.w;
          }
      } else { // box q [64 frames][32 bins]: this thread's bin = lane, one value per frame
        const uint8_t* vt = smem + C::OFF_V + st * STAGE + q * 8192;
#pragma unroll
        for (int j = 0; j < 64; j++) v[j] = *reinterpret_cast<const float*>(vt + j * 128 + voff[j & 7]);
      }
      tmem_wait_ld();
      // WAR across proxies: the V stage is about to be handed back to the TMA producer, and a generic-proxy LDS that is
      // merely issued is not ordered before the async-proxy refill by the mbarrier alone (seen on a 16-epilogue-warp
      // variant of kernels_nmf_tc.cu, profiles/experiments: whole rows of H off by ~0.5 % about once per 10^5 warp-steps).
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&p_free[wg]); mbar_arrive(&v_empty[st]); }
      uint32_t ph[32], pl[32];
#pragma unroll
      for (int j = 0; j < 32; j++) {
        float v0 = v[2 * j], v1 = v[2 * j + 1];
        if (p.clamp_v) { v0 = fmaxf(v0, kEps); v1 = fmaxf(v1, kEps); } // NMF.hpp:60
        float r0, r1, l0, l1;
        mul2(r0, r1, v0, v1, rcp_fast(fmaxf(__uint_as_float(pp[2 * j]), kEps)), rcp_fast(fmaxf(__uint_as_float(pp[2 * j + 1]), kEps)));
        ph[j] = cvt2(r0, r1);
        sub2(l0, l1, r0, r1, bf16lo_to_f(ph[j]), bf16hi_to_f(ph[j]));
        pl[j] = cvt2(l0, l1);
      }
      // R[wg] is still read by the second MMA of this warpgroup's previous step: rank < 64 collects that step's partial
      // here (b_full), rank 64 waits for the MMA's own completion signal
      if constexpr (!C::ONE_ACC) drain();
      else if (nn >= 2) mbar_wait(&r_free[wg], ((nn >> 1) - 1) & 1);
      tmem_st16(tR, *reinterpret_cast<uint32_t(*)[16]>(&ph[0]));
      tmem_st16(tR + 16, *reinterpret_cast<uint32_t(*)[16]>(&ph[16]));
      tmem_st16(tR + 32, *reinterpret_cast<uint32_t(*)[16]>(&pl[0]));
      tmem_st16(tR + 48, *reinterpret_cast<uint32_t(*)[16]>(&pl[16]));
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&r_full[wg]);
    };

    // K values per thread summed over the 32 lanes with shuffles (vector butterfly); afterwards lane l holds in a[0 .. PER)
    // the totals of the values PER * (l >> SH) + j  (K = 64: two per lane, K = 32: one per lane, K = 16: one per lane pair)
    auto butterfly = [&](float (&a)[K]) { bf_stage<K / 2, 16>(a, lane); };
    constexpr int SH = K == 16 ? 1 : 0;
    constexpr int PER = K == 64 ? 2 : 1;
    const bool bf_owner = K == 16 ? (lane & 1) == 0 : true; // lanes that hold distinct totals after the butterfly
    const int bf_idx = PER * (lane >> SH);                  // value index: [0,K2) first kind, [K2,K) second kind

    // wden / Nyquist-numerator partials of one frame row (new H row `h`, all K components) -> warp-private running sums
    auto frame_partials = [&](const float (&h)[K], float vn) {
      float pn = 0.f;
#pragma unroll
      for (int k = 0; k < K; k++) pn = fmaf(h[k], WN[k], pn);
      const float rn = vn / fmaxf(pn, kEps);
      float a[K];
#pragma unroll
      for (int j = 0; j < K2; j++) {
        const float x = wg ? h[K2 + j] : h[j];
        a[j] = x;            // sum_f H        (NMF.hpp:160)
        a[K2 + j] = rn * x;  // Nyquist row of (V / WH) H^T  (:159)
      }
      butterfly(a);
      if (bf_owner) {
#pragma unroll
        for (int j = 0; j < PER; j++) part[ew * K + bf_idx + j] += a[j];
      }
    };

    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const int buf = unit_buffer(p, unit);
      const int wbuf = p.shared_w ? 0 : buf;
      const float* gV = p.V + (int64_t) buf * Fp * Bp;
      float* gW = p.W + (int64_t) wbuf * K * Bp;
      float* gH = p.H + (int64_t) buf * Fp * K;
      __nv_bfloat16* gWop = p.Wop + wbuf * wop_stride;
      __nv_bfloat16* gHop = p.Hop + buf * hop_stride;
      // ---------------- unit prologue -----------------------------------------------------------------------------
      epi_bar(); // previous unit completely finished (fin / part are reused)
      if (et < K) {
        WN[et] = gW[(int64_t) et * Bp + BT];
        const float hd = p.hden[(int64_t) wbuf * K + et];
        f_hden[et] = hd;
        f_ihd[et] = 1.0f / fmaxf(hd, kEps);
      }
      for (int e = et; e < 8 * K; e += 256) part[e] = 0.f;
      epi_bar();
      bool partials_valid = false;
      int w_tiles_done = 0;
      const int MT = BT / 128;

      for_jobs(p, unit, jn, [&](int phase, int tile, uint32_t, uint32_t, bool) {
        if (phase == 1 && w_tiles_done == 0) {
          // ---------------- start of a W half-iteration: denominators + Nyquist numerators from H --------------------
          if (!partials_valid) { // first iteration (or H fixed): sweep H once; later the H jobs provide them
            for (int f0 = 0; f0 < Fp; f0 += 128) {
              float h[K];
              const float4* src = reinterpret_cast<const float4*>(gH + (int64_t) (f0 + r) * K);
#pragma unroll
              for (int j = 0; j < K / 4; j++) { const float4 x = __ldcg(src + j); h[4 * j] = x.x; h[4 * j + 1] = x.y; h[4 * j + 2] = x.z; h[4 * j + 3] = x.w; }
              float vn = gV[(int64_t) (f0 + r) * Bp + BT];
              if (p.clamp_v) vn = fmaxf(vn, kEps);
              frame_partials(h, vn);
            }
          }
          epi_bar();
          if (et < 2 * K) { // value index i < K2: component wg*K2 + i of kind 0; the 4 warps of that warpgroup hold partials
            const int kind = et / K, k = et % K;
            const int owner = (k / K2) * 4, idx = kind * K2 + (k % K2);
            float s = 0.f;
            for (int w4 = 0; w4 < 4; w4++) s += part[(owner + w4) * K + idx];
            if (kind) f_nyq[k] = s;
            else f_wden[k] = 1.0f / fmaxf(s, kEps); // the reciprocal: the tile updates multiply
          }
          epi_bar();
          for (int e = et; e < 8 * K; e += 256) part[e] = 0.f;
          for (int e = et; e < 8 * (K + 4); e += 256) red[e] = 0.f;
          epi_bar();
          partials_valid = false;
        }
        // ---------------- steps ------------------------------------------------------------------------------------
        TCS_MARK(0);
#ifdef FB200_TCS_DEBUG_TIMELINE
        if (p.dbg && blockIdx.x == 0 && et == 0 && jn < 96) p.dbg[jn * 8 + 7] = phase * 1000 + tile;
#endif
        const int ns = phase == 0 ? C1 : S2;
        float vn = 0.f;
        if (phase == 0) vn = gV[(int64_t) (128 * tile + r) * Bp + BT]; // Nyquist magnitude of this thread's frame: requested now, first
                                                                        // touched in the tile update (a use here exposes a DRAM round trip per job)
        if constexpr (C::ONE_ACC) {
          for (int i = 0; i < ns; i++, n++) {
            if ((int) (n & 1) == wg) do_step(n, phase == 0);
            // the group that ended with the previous step (one step of lag: its last second MMA runs meanwhile)
            if (i >= 1 && ((i - 1) % C::GROUP) == C::GROUP - 1) drain_shared(i - 1 < C::GROUP);
          }
        } else {
          for (int i = 0; i < ns; i++, n++) {
            if ((int) (n & 1) != wg) continue;
            do_step(n, phase == 0);
            out_valid = 1; out_first = (i == wg); out_par = (n >> 1) & 1;
          }
        }
        // The old fp32 values of the tile: from the staged copy, or (rank 64) from L2 with ld.cg -- the rows were written by
        // other threads of this CTA, and an L1 line filled around such a store can be stale -- requested before the last
        // partial is collected so that the round trip overlaps that wait.
        float old[K]; // H job: the frame's row; W job: [0, K2) this warpgroup's components of the bin
        if constexpr (!C::HAS_F32) {
          if (phase == 0) {
            const float4* src = reinterpret_cast<const float4*>(gH + (int64_t) (128 * tile + r) * K);
#pragma unroll
            for (int j = 0; j < K / 4; j++) { const float4 x = __ldcg(src + j); old[4 * j] = x.x; old[4 * j + 1] = x.y; old[4 * j + 2] = x.z; old[4 * j + 3] = x.w; }
          } else {
#pragma unroll
            for (int j = 0; j < K2; j++) old[j] = __ldcg(gW + (int64_t) (k0 + j) * Bp + 128 * tile + r);
          }
        }
        TCS_MARK(1);
        if constexpr (C::ONE_ACC) drain_shared(ns <= C::GROUP);
        else drain();
        TCS_MARK(2);
        if constexpr (C::HAS_F32) {
          mbar_wait(&st_full[jn & 1], (jn >> 1) & 1); // long complete; makes the bulk copy's bytes visible to this thread
          const float* ft = reinterpret_cast<const float*>(smem + C::OFF_F32 + (jn & 1) * C::F32TILE);
          if (phase == 0) {
            const float4* src = reinterpret_cast<const float4*>(ft + r * K);
#pragma unroll
            for (int j = 0; j < K / 4; j++) { const float4 x = src[j]; old[4 * j] = x.x; old[4 * j + 1] = x.y; old[4 * j + 2] = x.z; old[4 * j + 3] = x.w; }
          } else {
#pragma unroll
            for (int j = 0; j < K2; j++) old[j] = ft[(k0 + j) * 128 + r];
          }
        }
        // the running sums of the tile are complete; every thread reads values written by the other warpgroup
        if constexpr (C::ONE_ACC) { tmem_wait_st(); tc_fence_before(); }
        epi_bar();
        if constexpr (C::ONE_ACC) tc_fence_after();
        TCS_MARK(3);
        if (phase == 0) {
          // ---------------- H-tile update (NMF.hpp:168-170) from the fp32 state -----------------------------------------
          const int f = 128 * tile + r;
          float (&h)[K] = old;
          if (p.clamp_v) vn = fmaxf(vn, kEps); // NMF.hpp:60
          float pn = 0.f;
#pragma unroll
          for (int k = 0; k < K; k++) pn = fmaf(h[k], WN[k], pn);
          const float rn = vn / fmaxf(pn, kEps);
#pragma unroll
          if constexpr (C::ONE_ACC) {
#pragma unroll
            for (int j16 = 0; j16 < K / 16; j16++) { // all K new values (the partials of the next W-update need the whole row)
              uint32_t s0[16];
              tmem_ld16(tSumAll + 16 * j16, s0);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; i++) {
                const int k = 16 * j16 + i;
                h[k] = h[k] * fmaf(rn, WN[k], __uint_as_float(s0[i])) * f_ihd[k];
              }
            }
          } else {
#pragma unroll
            for (int j4 = 0; j4 < K / 4; j4++) {
              const float4 a = *reinterpret_cast<const float4*>(hs + ((j4 * 2) * 128 + r) * 4);
              const float4 b = *reinterpret_cast<const float4*>(hs + ((j4 * 2 + 1) * 128 + r) * 4);
              const float num[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
#pragma unroll
              for (int i = 0; i < 4; i++) {
                const int k = 4 * j4 + i;
                h[k] = h[k] * fmaf(rn, WN[k], num[i]) * f_ihd[k];
              }
            }
          }
          // this warpgroup stores its half of the row: fp32 state + split operand
          {
            float4* dst = reinterpret_cast<float4*>(gH + (int64_t) f * K + k0);
#pragma unroll
            for (int j = 0; j < K2 / 4; j++)
              dst[j] = wg ? make_float4(h[K2 + 4 * j], h[K2 + 4 * j + 1], h[K2 + 4 * j + 2], h[K2 + 4 * j + 3])
                          : make_float4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
#pragma unroll
            for (int kb = 0; kb < K2 / 8; kb++) {
              uint32_t ph[4], pm[4];
#pragma unroll
              for (int j = 0; j < 4; j++) {
                const float x0 = wg ? h[K2 + 8 * kb + 2 * j] : h[8 * kb + 2 * j], x1 = wg ? h[K2 + 8 * kb + 2 * j + 1] : h[8 * kb + 2 * j + 1];
                split2(x0, x1, ph[j], pm[j]);
              }
              *reinterpret_cast<uint4*>(gHop + op_index_h<K>(0, f, k0 + 8 * kb)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
              *reinterpret_cast<uint4*>(gHop + op_index_h<K>(1, f, k0 + 8 * kb)) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
            }
          }
          if (p.upd_w) { frame_partials(h, vn); partials_valid = true; }
        } else {
          // ---------------- W-tile update, not yet normalised (NMF.hpp:161) ------------------------------------------
          const int b = 128 * tile + r;
          float a[K]; // [0,K2) w^2, [K2,K) w of this warpgroup's components
          float mx = 0.f;
#pragma unroll
          for (int jg = 0; jg < K2 / G; jg++) {
            float num[G];
            if constexpr (C::ONE_ACC) {
              uint32_t s0[G];
              tmem_ldn<G>(tSumAll + k0 + G * jg, s0);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < G; i++) num[i] = __uint_as_float(s0[i]);
            } else {
#pragma unroll
              for (int i4 = 0; i4 < G / 4; i4++) {
                const int jj = (k0 + G * jg) / 4 + i4;
                const float4 s0 = *reinterpret_cast<const float4*>(hs + ((jj * 2) * 128 + r) * 4);
                const float4 s1 = *reinterpret_cast<const float4*>(hs + ((jj * 2 + 1) * 128 + r) * 4);
                num[4 * i4] = s0.x + s1.x; num[4 * i4 + 1] = s0.y + s1.y; num[4 * i4 + 2] = s0.z + s1.z; num[4 * i4 + 3] = s0.w + s1.w;
              }
            }
#pragma unroll
            for (int i = 0; i < G; i++) {
              const int j = G * jg + i, k = k0 + j;
              const float w = old[j] * num[i] * f_wden[k];
              gW[(int64_t) k * Bp + b] = w;
              a[j] = w * w;
              a[K2 + j] = w;
              mx = fmaxf(mx, w);
            }
          }
          butterfly(a);
          if (bf_owner) {
#pragma unroll
            for (int j = 0; j < PER; j++) red[ew * (K + 4) + bf_idx + j] += a[j];
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          if (lane == 0) red[ew * (K + 4) + K] = fmaxf(red[ew * (K + 4) + K], mx);
          w_tiles_done++;
          if (w_tiles_done == MT) {
            // ---------------- end of the W half-iteration: column normalisation (:162), hden (:169), operand refresh ----
            w_tiles_done = 0;
            epi_bar();
            if (et < K) {
              const int owner = (et / K2) * 4, idx = et % K2;
              float s2 = 0.f, s1 = 0.f;
              for (int w4 = 0; w4 < 4; w4++) { s2 += red[(owner + w4) * (K + 4) + idx]; s1 += red[(owner + w4) * (K + 4) + K2 + idx]; }
              const float wn = WN[et] * f_nyq[et] * f_wden[et]; // Nyquist bin, carried on the SIMT side
              f_s2[et] = fmaf(wn, wn, s2);
              f_s1[et] = s1 + wn;
              WN[et] = wn;
            }
            if (et == 128) {
              float gm = 0.f;
              for (int w8 = 0; w8 < 8; w8++) gm = fmaxf(gm, red[w8 * (K + 4) + K]);
              f_gm[0] = gm;
            }
            epi_bar();
            if (et < K) {
              float gm = f_gm[0];
              for (int k = 0; k < K; k++) gm = fmaxf(gm, WN[k]);
              const float s2 = f_s2[et], s1 = f_s1[et];
              const bool norm = gm > kEps;                                   // NMF.hpp:162
              const float inv = norm ? (s2 > 0.f ? 1.0f / sqrtf(s2) : 0.f) : 1.0f;
              f_inv[et] = inv;
              const float hd = s1 * inv;
              f_hden[et] = hd;
              f_ihd[et] = 1.0f / fmaxf(hd, kEps);
              p.hden[(int64_t) wbuf * K + et] = hd;
            }
            epi_bar();
            if (et < K) WN[et] *= f_inv[et];
            epi_bar();
            // Rescale sweep over W (fp32 state + split operand).  One warp pass = one component block (8 k) x 32 bins:
            // lane = (k & 7, 8-bin group), so that the fp32 rows are touched in 128-byte runs AND the 16-byte operand units
            // of a core matrix (8 k x 8 bins, 128 bytes) are written by 8 adjacent lanes.  (An item-per-thread mapping with
            // 32 different lines per instruction made this sweep 29 % of a rank-64 iteration: developer timeline, r02y.)
            {
              const int kk = lane >> 2, bq = lane & 3;
              const int n_pass = KB * (BT / 32);
              constexpr int U = 4; // passes in flight per warp (16 made no difference: the sweep moves 1.5 MB per rank-64 buffer
                                   // against the |X| streams of all other CTAs, it is bandwidth, not latency)
              for (int w0 = ew; w0 < n_pass; w0 += 8 * U) {
                float4 x0[U], x1[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                  const int wi = w0 + 8 * u;
                  if (wi < n_pass) {
                    const int k = 8 * (wi % KB) + kk, bin0 = 32 * (wi / KB) + 8 * bq;
                    const float4* src = reinterpret_cast<const float4*>(gW + (int64_t) k * Bp + bin0);
                    x0[u] = __ldcg(src); x1[u] = __ldcg(src + 1);
                  }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                  const int wi = w0 + 8 * u;
                  if (wi < n_pass) {
                    const int k = 8 * (wi % KB) + kk, bin0 = 32 * (wi / KB) + 8 * bq;
                    const float sc = f_inv[k];
                    float4 a = x0[u], b = x1[u];
                    a.x *= sc; a.y *= sc; a.z *= sc; a.w *= sc; b.x *= sc; b.y *= sc; b.z *= sc; b.w *= sc;
                    float4* dstw = reinterpret_cast<float4*>(gW + (int64_t) k * Bp + bin0);
                    dstw[0] = a; dstw[1] = b;
                    uint32_t ph[4], pm[4];
                    split2(a.x, a.y, ph[0], pm[0]); split2(a.z, a.w, ph[1], pm[1]);
                    split2(b.x, b.y, ph[2], pm[2]); split2(b.z, b.w, ph[3], pm[3]);
                    __nv_bfloat16* dst = gWop + op_index_w<K>(0, k, bin0);
                    *reinterpret_cast<uint4*>(dst) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    *reinterpret_cast<uint4*>(dst + KB * 64) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
                  }
                }
              }
            }
            if (et < K) gW[(int64_t) et * Bp + BT] = WN[et];
          }
        }
        // ---------------- publish: the tile update is in global memory ------------------------------------------------
        // generic-proxy stores -> async-proxy reads (bulk copies issued by the producer of this CTA): the writer-side
        // proxy fence plus the CTA barrier order them; no device-scope fence is needed, nobody outside the CTA reads.
        // The barrier also separates this job's reads of the running sums from the next job's first partial.
        TCS_MARK(4);
        fence_async_all();
        epi_bar();
        TCS_MARK(5);
        if (et == 0) *jobs_done = jn + 1;
        // the staged fp32 tile has been read (generic proxy) before the fence above: hand its buffer back to the producer
        if (C::HAS_F32 && lane == 0) mbar_arrive(&st_empty[jn & 1]);
      });
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------------------------
bool tcs_eligible(const NmfDev& d)
{
  const int BT = d.B - 1;
  return (d.KP == 16 || d.KP == 32 || d.KP == 64) && BT >= 128 && (BT % 128) == 0 && d.Bp == d.B + 3 && d.Fp >= 128 && (d.Fp % 128) == 0;
}

template <int K>
static int32_t tcs_run_t(Plan* p, const NmfDev& d, int iters, bool upd_w, bool upd_h)
{
  using C = Cfg<K>;
  const int BT = d.B - 1;
  const int nw = d.shared_w ? 1 : d.batch;
  const int total = d.op_total > 0 ? d.op_total : d.batch, first = d.op_total > 0 ? d.op_first : 0;
  const size_t wop_one = (size_t) (BT / 8) * C::ROWB, hop_one = (size_t) (d.Fp / 8) * C::ROWB;
  FB_CUDA(p, p->wop_buf.ensure((d.shared_w ? 1 : (size_t) total) * wop_one));
  FB_CUDA(p, p->hop_buf.ensure((size_t) total * hop_one));
  Params q{};
  q.dbg = nullptr;
#ifdef FB200_TCS_DEBUG_TIMELINE
  if (getenv("FB200_TCS_TIMELINE")) {
    FB_CUDA(p, p->out_b.ensure(sizeof(long long) * 96 * 8));
    FB_CUDA(p, cudaMemsetAsync(p->out_b.p, 0, sizeof(long long) * 96 * 8, p->stream));
    q.dbg = p->out_b.as<long long>();
  }
#endif
  q.V = d.V; q.W = d.W; q.H = d.H; q.hden = d.hden;
  q.Wop = reinterpret_cast<__nv_bfloat16*>(p->wop_buf.as<uint8_t>() + (d.shared_w ? 0 : (size_t) first * wop_one));
  q.Hop = reinterpret_cast<__nv_bfloat16*>(p->hop_buf.as<uint8_t>() + (size_t) first * hop_one);
  q.batch = d.batch; q.Fp = d.Fp; q.Bp = d.Bp; q.BT = BT;
  q.iters = iters; q.upd_w = upd_w ? 1 : 0; q.upd_h = upd_h ? 1 : 0; q.shared_w = d.shared_w; q.clamp_v = d.clamp_v;
  const int T = d.Fp / 128;
  q.units = upd_w ? d.batch : d.batch * ((T + 1) / 2);
  alignas(64) CUtensorMap tmap1, tmap2;
  FB_TRY(make_v_tensor_map(p, &tmap1, d.V, d.Bp, d.Fp, d.batch, 128));
  FB_TRY(make_v_tensor_map(p, &tmap2, d.V, d.Bp, d.Fp, d.batch, 64));
  const uint32_t bit = K == 16 ? 0x20000u : (K == 32 ? 0x40000u : 0x80000u);
  if (!(p->attr_mask & bit)) {
    FB_CUDA(p, cudaFuncSetAttribute(k_nmf_tcs<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    p->attr_mask |= bit;
  }
  {
    const int64_t items = (int64_t) nw * (BT / 8) * K + (int64_t) d.batch * d.Fp * (K / 8);
    const int blocks = (int) std::min<int64_t>((items + 255) / 256, 148 * 16);
    k_tcs_pack<K><<<blocks, 256, 0, p->stream>>>(d.W, d.H, q.Wop, q.Hop, nw, d.batch, d.Fp, d.Bp, BT);
    p->launches++;
  }
  const int grid = std::min(q.units, p->sm_count);
  while (p->kev.size() < p->kev_used + 2) { cudaEvent_t e; FB_CUDA(p, cudaEventCreate(&e)); p->kev.push_back(e); }
  cudaEventRecord(p->kev[p->kev_used++], p->stream);
  k_nmf_tcs<K><<<grid, NTHREADS, C::SMEM_BYTES, p->stream>>>(q, tmap1, tmap2);
  cudaEventRecord(p->kev[p->kev_used++], p->stream);
  p->launches++; p->launches_nmf++;
  FB_CUDA(p, cudaGetLastError());
#ifdef FB200_TCS_DEBUG_TIMELINE
  if (q.dbg) {
    long long h[96 * 8];
    FB_CUDA(p, cudaMemcpyAsync(h, q.dbg, sizeof h, cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(p, cudaStreamSynchronize(p->stream));
    fprintf(stderr, "job kind tile | start  steps_done  collected  sums_visible  updated  published | steps  update+publish  (cycles)\n");
    for (int j = 0; j < 96 && h[j * 8]; j++) {
      const long long* e = h + j * 8;
      fprintf(stderr, "%3d %s %3lld | %9lld %9lld %9lld %9lld %9lld %9lld | %8lld %8lld\n", j, e[7] >= 1000 ? "W" : "H", e[7] % 1000, e[0] - h[0], e[1] - h[0],
              e[2] - h[0], e[3] - h[0], e[4] - h[0], e[5] - h[0], e[1] - e[0], e[5] - e[1]);
    }
  }
#endif
  return FB200_OK;
}

int32_t tcs_run(Plan* p, const NmfDev& d, int iters, bool upd_w, bool upd_h)
{
  if (d.KP == 16) return tcs_run_t<16>(p, d, iters, upd_w, upd_h);
  if (d.KP == 32) return tcs_run_t<32>(p, d, iters, upd_w, upd_h);
  return tcs_run_t<64>(p, d, iters, upd_w, upd_h);
}

} // namespace fb200
