// Rank-16 tcgen05 engine with the STATIONARY operand in tensor memory (algorithms/public/NMF.hpp:144-183, 45-89).
//
// Same job structure and arithmetic as the streamed engine (kernels_nmf_tcs.cu: job(H, t) / job(W, m), 64-column steps,
// exact 3-way bf16 split, six cross terms, fresh second-MMA accumulators summed with round-to-nearest adds), with the two
// changes the profiles asked for.  Both engines were bound by the shared-memory data pipe (r01e: 75 % busy, r02e: 72 %):
//   * the A operand of the first MMA -- the 128 stationary rows (H tile or W tile) x 16 components x 3 split parts -- was
//     re-read from shared memory by each of the six split-term MMAs of every step: 24 of the ~145 KB a step moved.  It now
//     lives in TMEM (24 columns per job, double buffered): the epilogue warps write it once per job with tcgen05.st from
//     the fp32 state, the first MMA runs in TS form (A from TMEM, B = the streamed chunk from shared memory);
//   * the running numerator sums of a job (fresh per-step partials added with RN fp32 adds) went through shared memory
//     (16 KB read + written per step); they now stay in 32 TMEM columns.
// Per step the shared-memory pipe carries the |X| tile (TMA fill + one read), the streamed 6 KB chunk (fill + 6 + 8 MMA
// reads of 2 KB / 1.5 KB) and nothing else.  No shared-memory copy of the stationary tile exists at all.
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda.h>
#include <algorithm>

namespace fb200 {
using namespace tc;

namespace tcr {
constexpr int K = 16;
constexpr int KB = 2;
constexpr int NS = 4;             // V ring stages
constexpr int STAGE = 32768;      // 128 x 64 fp32
constexpr int NSO = 4;            // streamed-operand ring stages
constexpr int NTHREADS = 384;
constexpr uint32_t ROWB = 3 * KB * 128;
constexpr uint32_t CHUNK = 8 * ROWB;  // 64 rows of the streamed operand: 6 KB
// TMEM columns
constexpr uint32_t TM_P = 0;      // + 64 g
constexpr uint32_t TM_R = 128;    // + 64 g : hi [0,32) lo [32,64)
constexpr uint32_t TM_ACC = 256;  // + 64 g : [0,16) R_hi X_hi | [16,48) R_hi [X_mid|X_lo] | [48,64) R_lo X_hi
constexpr uint32_t TM_A = 384;    // + 24 slot : hi [0,8) mid [8,16) lo [16,24), bf16 pairs (k, k+1) per column
constexpr uint32_t TM_SUM = 432;  // + 16 wg : running numerator sums of the job
// shared memory
constexpr int OFF_V = 0;
constexpr int OFF_O = OFF_V + NS * STAGE;
constexpr int OFF_STF = OFF_O + NSO * (int) CHUNK;        // float [2 slots][128 rows][16]: fp32 copy of the stationary rows
constexpr int OFF_PART = OFF_STF + 2 * 128 * 16 * 4;      // float [8 warps][16]
constexpr int OFF_RED = OFF_PART + 8 * 16 * 4;            // float [8 warps][20]
constexpr int OFF_FIN = OFF_RED + 8 * 20 * 4;             // float [8][16] + 4
constexpr int OFF_BAR = OFF_FIN + 8 * 16 * 4 + 16;
constexpr int NBAR = 2 * NS + 2 * NSO + 2 + 8;
constexpr int OFF_SLOT = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_SLOT + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 88;" ::: "memory"); }
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ int op_index_w(int part, int k, int b) { return ((((b >> 3) * 3 + part) * KB + (k >> 3)) << 6) + ((k & 7) << 3) + (b & 7); }
__device__ __forceinline__ int op_index_h(int part, int f, int k) { return ((((f >> 3) * 3 + part) * KB + (k >> 3)) << 6) + ((f & 7) << 3) + (k & 7); }

struct Params {
  const float* V;      // [batch][Fp][Bp]
  float* W;            // [batch | 1][16][Bp]   fp32 state
  float* H;            // [batch][Fp][16]
  float* hden;         // [batch | 1][16]
  __nv_bfloat16* Wop;  // [batch | 1][BT/8][3][2][64]  streamed copies (split bf16)
  __nv_bfloat16* Hop;  // [batch][Fp/8][3][2][64]
  int batch, Fp, Bp, BT;
  int iters, upd_w, upd_h, shared_w, clamp_v;
  int units;
};

// The job sequence of a CTA, identical for every warp role (see kernels_nmf_tcs.cu for the schedule and the dependency
// rules): fixed W -> units are (buffer, tile pair), all iterations per pair, jobs alternate between the two tiles;
// otherwise units are buffers and every iteration is a W half-iteration (jobs over 128-bin tiles) followed by an H
// half-iteration (jobs over 128-frame tiles).
struct JobIter {
  int unit, it, idx, nt, t0;
  uint32_t jn, hjob0, n1;
  bool have_h, valid;
  int phase, tile;        // 0 = H job, 1 = W job
  uint32_t need_c;        // finished jobs the producer must see before copying the chunk of step i: need_c + (per_tile ? i / 2 : 0)
  bool per_tile;

  // (the shape parameters stay in the kernel's constant bank: the iterator itself is live across the step loop of the
  //  register-starved epilogue warps)
  __device__ void init(const Params& p, int cta)
  {
    unit = cta; jn = 0; hjob0 = 0; n1 = 0;
    valid = unit < p.units && p.iters > 0;
    if (valid) start_unit(p);
  }
  __device__ int buffer(const Params& p) const { return p.upd_w ? unit : unit / ((p.Fp / 128 + 1) / 2); }
  __device__ bool first_of_unit() const { return it == 0 && idx == 0; }
  __device__ int steps(int C1, int S2) const { return phase == 0 ? C1 : S2; }
  __device__ void start_unit(const Params& p)
  {
    it = 0; idx = 0; have_h = false;
    if (!p.upd_w) {
      const int T = p.Fp / 128;
      t0 = 2 * (unit % ((T + 1) / 2));
      nt = (t0 + 1 < T) ? 2 : 1;
    }
    compute(p);
  }
  __device__ void compute(const Params& p)
  {
    const int MT = p.BT / 128;
    if (!p.upd_w) { phase = 0; tile = t0 + idx; need_c = 0; per_tile = false; }
    else if (idx < MT) { phase = 1; tile = idx; need_c = have_h ? hjob0 + 1 : 0u; per_tile = have_h; }
    else {
      if (idx == MT) n1 = jn;
      phase = 0; tile = idx - MT; need_c = n1; per_tile = false;
    }
  }
  __device__ void next(const Params& p, int ncta)
  {
    jn++; idx++;
    const int per_it = !p.upd_w ? nt : p.BT / 128 + (p.upd_h ? p.Fp / 128 : 0);
    if (idx == per_it) {
      if (p.upd_w && p.upd_h) { hjob0 = n1; have_h = true; }
      idx = 0; it++;
      if (it == p.iters) {
        unit += ncta;
        if (unit >= p.units) { valid = false; return; }
        start_unit(p);
        return;
      }
    }
    compute(p);
  }
};
} // namespace tcr

using namespace tcr;

// fp32 state -> split operand copies for the streamed side (once per call; the engine keeps both in step afterwards)
__global__ void k_tcr_pack(const float* __restrict__ W, const float* __restrict__ H, __nv_bfloat16* __restrict__ Wop,
                           __nv_bfloat16* __restrict__ Hop, int nw, int batch, int Fp, int Bp, int BT)
{
  const int64_t w_items = (int64_t) nw * (BT / 8) * K, h_items = (int64_t) batch * Fp * KB;
  for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < w_items + h_items; i += (int64_t) gridDim.x * blockDim.x) {
    float x[8];
    __nv_bfloat16* dst;
    if (i < w_items) {
      const int k = (int) (i % K);
      const int64_t r = i / K;
      const int blk = (int) (r % (BT / 8)), buf = (int) (r / (BT / 8));
      const float4* src = reinterpret_cast<const float4*>(W + ((int64_t) buf * K + k) * Bp + 8 * blk);
      const float4 a = src[0], b = src[1];
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      dst = Wop + (int64_t) buf * (BT / 8) * 3 * KB * 64 + op_index_w(0, k, 8 * blk);
    } else {
      const int64_t j = i - w_items;
      const int kb = (int) (j % KB);
      const int64_t r = j / KB;
      const int f = (int) (r % Fp), buf = (int) (r / Fp);
      const float4* src = reinterpret_cast<const float4*>(H + ((int64_t) buf * Fp + f) * K + 8 * kb);
      const float4 a = src[0], b = src[1];
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      dst = Hop + (int64_t) buf * (Fp / 8) * 3 * KB * 64 + op_index_h(0, f, 8 * kb);
    }
    uint32_t ph[4], pm[4], pl[4];
#pragma unroll
    for (int q = 0; q < 4; q++) split3(x[2 * q], x[2 * q + 1], ph[q], pm[q], pl[q]);
    *reinterpret_cast<uint4*>(dst) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(dst + KB * 64) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
    *reinterpret_cast<uint4*>(dst + 2 * KB * 64) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_nmf_tcr(Params p, const __grid_constant__ CUtensorMap tmap1, const __grid_constant__ CUtensorMap tmap2)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  float* stf = reinterpret_cast<float*>(smem + OFF_STF);
  float* part = reinterpret_cast<float*>(smem + OFF_PART);
  float* red = reinterpret_cast<float*>(smem + OFF_RED);
  float* fin = reinterpret_cast<float*>(smem + OFF_FIN);
  float* f_wden = fin;             // [16] sum_f H
  float* f_nyq = fin + 16;         // [16] Nyquist numerator
  float* f_inv = fin + 32;         // [16] 1 / column norm
  float* f_ihd = fin + 48;         // [16] 1 / max(hden, eps)
  float* WN = fin + 64;            // [16] Nyquist row of W
  float* f_hden = fin + 80;        // [16]
  float* f_s2 = fin + 96;          // [16]
  float* f_s1 = fin + 112;         // [16]
  float* f_gm = fin + 128;         // [1]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* v_full = bars;                 // [NS]
  uint64_t* v_empty = v_full + NS;         // [NS]
  uint64_t* o_full = v_empty + NS;         // [NSO]
  uint64_t* o_empty = o_full + NSO;        // [NSO]  both MMAs that read the chunk have completed
  uint64_t* a_ready = o_empty + NSO;       // [2]    the stationary operand of a job sits in its TMEM slot
  uint64_t* p_full = a_ready + 2;          // [2]
  uint64_t* r_full = p_full + 2;           // [2]
  uint64_t* b_full = r_full + 2;           // [2]
  uint64_t* p_free = b_full + 2;           // [2]
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + OFF_SLOT);
  volatile uint32_t* jobs_done = slot + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Fp = p.Fp, Bp = p.Bp, BT = p.BT;
  const int C1 = BT / 64, S2 = Fp / 64;

  if (tid == 0) {
    slot[1] = 0u;
    for (int i = 0; i < NS; i++) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 4); }
    for (int i = 0; i < NSO; i++) { mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 2); }
    for (int i = 0; i < 2; i++) {
      mbar_init(&a_ready[i], 8);
      mbar_init(&p_full[i], 1); mbar_init(&r_full[i], 4); mbar_init(&b_full[i], 1); mbar_init(&p_free[i], 4);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmap1);
    tma_prefetch_desc(&tmap2);
  }
  if (warp == 1) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *slot;
  const int64_t wop_stride = p.shared_w ? 0 : (int64_t) (BT / 8) * 3 * KB * 64;
  const int64_t hop_stride = (int64_t) (Fp / 8) * 3 * KB * 64;
  JobIter J;
  J.init(p, blockIdx.x);

  if (warp < 4) {
    reg_dec();
    if (warp == 0) {
      // =========================================== producer =================================================
      if (lane == 0) {
        uint32_t n = 0;
        for (; J.valid; J.next(p, gridDim.x)) {
          const int buf = J.buffer(p), phase = J.phase, tile = J.tile;
          const __nv_bfloat16* gsrc = phase == 0 ? p.Wop + (p.shared_w ? 0 : buf) * wop_stride : p.Hop + buf * hop_stride;
          const int ns = J.steps(C1, S2);
          const uint32_t n0 = n;
          auto issue_v = [&](int i) {
            const uint32_t nn = n0 + i, st = nn % NS, k = nn / NS;
            mbar_wait(&v_empty[st], (k & 1) ^ 1);
            mbar_arrive_expect_tx(&v_full[st], STAGE);
            uint8_t* dst = smem + OFF_V + st * STAGE;
            if (phase == 0) {
              tma_load_3d(dst, &tmap1, 64 * i, 128 * tile, buf, &v_full[st]);
              tma_load_3d(dst + 16384, &tmap1, 64 * i + 32, 128 * tile, buf, &v_full[st]);
            } else {
#pragma unroll
              for (int w = 0; w < 4; w++) tma_load_3d(dst + w * 8192, &tmap2, 128 * tile + 32 * w, 64 * i, buf, &v_full[st]);
            }
          };
          const int pre = ns < NS ? ns : NS; // |X| depends on nothing: requested before blocking on chunk dependencies
          for (int i = 0; i < pre; i++) issue_v(i);
          uint32_t seen = 0;
          for (int i = 0; i < ns; i++, n++) {
            if (i >= pre) issue_v(i);
            const uint32_t need = J.need_c + (J.per_tile ? (uint32_t) (i >> 1) : 0u);
            if (seen < need) {
              while ((seen = *jobs_done) < need) {}
              fence_async_all(); // the chunk was written with generic stores; the copy reads through the async proxy
            }
            const uint32_t so = n % NSO, k = n / NSO;
            mbar_wait(&o_empty[so], (k & 1) ^ 1);
            mbar_arrive_expect_tx(&o_full[so], CHUNK);
            bulk_g2s(smem + OFF_O + so * CHUNK, gsrc + (int64_t) i * 8 * 3 * KB * 64, CHUNK, &o_full[so]);
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // =========================================== first-MMA issuer (A from TMEM) ===========================
      constexpr uint32_t ID_H = make_idesc_bf16(128, 64, 0, 1); // B = W chunk, MN-major
      constexpr uint32_t ID_W = make_idesc_bf16(128, 64, 0, 0); // B = H chunk, K-major
      constexpr uint32_t HI_A = (ROWB >> 4) | (1u << 14);       // SBO = ROWB (block rows), descriptor version 1
      constexpr uint32_t LO_A = (128u >> 4) << 16;              // LBO = 128 (component blocks along K)
      constexpr uint32_t PSTEP = (KB * 128) >> 4;               // one split part of the chunk
      const uint32_t o_a = smem_u32(smem + OFF_O);
      uint32_t n = 0;
      for (; J.valid; J.next(p, gridDim.x)) {
        const uint32_t sl = J.jn & 1;
        mbar_wait(&a_ready[sl], (J.jn >> 1) & 1);
        tc_fence_after();
        const uint32_t aT = tbase + TM_A + 24 * sl; // hi, +8 mid, +16 lo
        const uint32_t idesc = J.phase == 0 ? ID_H : ID_W;
        const int ns = J.steps(C1, S2);
        for (int i = 0; i < ns; i++, n++) {
          const uint32_t g = n & 1, so = n % NSO;
          if (n >= 2) mbar_wait(&p_free[g], ((n - 2) >> 1) & 1);
          mbar_wait(&o_full[so], (n / NSO) & 1);
          tc_fence_after();
          const uint32_t blo = ((o_a + so * CHUNK) >> 4) | LO_A;
          const uint32_t dP = tbase + TM_P + 64 * g;
          mma_ts_lohi<0>(dP, aT, blo, HI_A, idesc);                      // hi  hi
          mma_ts_lohi<1>(dP, aT, blo + PSTEP, HI_A, idesc);              // hi  mid
          mma_ts_lohi<1>(dP, aT + 8, blo, HI_A, idesc);                  // mid hi
          mma_ts_lohi<1>(dP, aT, blo + 2 * PSTEP, HI_A, idesc);          // hi  lo
          mma_ts_lohi<1>(dP, aT + 16, blo, HI_A, idesc);                 // lo  hi
          mma_ts_lohi<1>(dP, aT + 8, blo + PSTEP, HI_A, idesc);          // mid mid
          mma_commit_warp(&p_full[g]);
          mma_commit_warp(&o_empty[so]);
        }
      }
    } else {
      // =========================================== second-MMA issuers (one per epilogue warpgroup) ===========
      const uint32_t myg = warp - 2;
      constexpr uint32_t ID_H3 = make_idesc_bf16(128, 48, 0, 0), ID_H1 = make_idesc_bf16(128, 16, 0, 0); // B = W chunk, K-major
      constexpr uint32_t ID_W3 = make_idesc_bf16(128, 48, 0, 1), ID_W1 = make_idesc_bf16(128, 16, 0, 1); // B = H chunk, MN-major
      constexpr uint32_t HI_B = (128u >> 4) | (1u << 14);
      constexpr uint32_t LO_B = (ROWB >> 4) << 16;
      constexpr uint32_t RSTEP = ROWB >> 4;
      const uint32_t o_a = smem_u32(smem + OFF_O);
      uint32_t n = 0;
      for (; J.valid; J.next(p, gridDim.x)) {
        const uint32_t id3 = J.phase == 0 ? ID_H3 : ID_W3, id1 = J.phase == 0 ? ID_H1 : ID_W1;
        const int ns = J.steps(C1, S2);
        for (int i = 0; i < ns; i++, n++) {
          const uint32_t g = n & 1, so = n % NSO;
          if (g != myg) continue;
          mbar_wait(&r_full[g], (n >> 1) & 1); // implies o_full(n): the first MMA of this step waited for it
          mbar_wait(&o_full[so], (n / NSO) & 1);
          tc_fence_after();
          const uint32_t blo = ((o_a + so * CHUNK) >> 4) | LO_B;
          const uint32_t rbase = tbase + TM_R + 64 * g;
          const uint32_t dacc = tbase + TM_ACC + 64 * g;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t b0 = blo + 2 * j * RSTEP;
            const uint32_t rh = rbase + 8 * j;
            if (j == 0) mma_ts_lohi<0>(dacc, rh, b0, HI_B, id3);
            else mma_ts_lohi<1>(dacc, rh, b0, HI_B, id3);
            if (j == 0) mma_ts_lohi<0>(dacc + 48, rh + 32, b0, HI_B, id1);
            else mma_ts_lohi<1>(dacc + 48, rh + 32, b0, HI_B, id1);
          }
          mma_commit_warp(&b_full[g]);
          mma_commit_warp(&o_empty[so]);
        }
      }
    }
  } else {
    reg_inc();
    // =========================================== epilogue warps ===============================================
    const int et = tid - 128;
    const int wg = (warp - 4) >> 2;
    const int ew = warp - 4;
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const uint32_t lane_off = (uint32_t) (32 * q) << 16;
    const uint32_t tP = tbase + TM_P + 64 * wg + lane_off;
    const uint32_t tR = tbase + TM_R + 64 * wg + lane_off;
    const uint32_t tAcc = tbase + TM_ACC + 64 * wg + lane_off;
    const uint32_t tSum = tbase + TM_SUM + 16 * wg + lane_off;
    const int k0 = 8 * wg; // components this warpgroup stages / stores / reduces
    uint32_t n = 0;
    int out_valid = 0, out_first = 0;
    uint32_t out_par = 0;

    auto drain = [&]() {
      if (!out_valid) return;
      mbar_wait(&b_full[wg], out_par);
      tc_fence_after();
      uint32_t a[32], a2[32], w[16];
      tmem_ld32(tAcc, a);
      tmem_ld32(tAcc + 32, a2);
      if (!out_first) tmem_ld16(tSum, w);
      tmem_wait_ld();
#pragma unroll
      for (int k = 0; k < K; k += 2) { // sum += a[k] + ((a[16+k] + a2[k]) + a2[16+k])
        float x0, x1;
        add2(x0, x1, __uint_as_float(a[16 + k]), __uint_as_float(a[17 + k]), __uint_as_float(a2[k]), __uint_as_float(a2[k + 1]));
        add2(x0, x1, x0, x1, __uint_as_float(a2[16 + k]), __uint_as_float(a2[17 + k]));
        add2(x0, x1, __uint_as_float(a[k]), __uint_as_float(a[k + 1]), x0, x1);
        if (!out_first) add2(x0, x1, __uint_as_float(w[k]), __uint_as_float(w[k + 1]), x0, x1);
        w[k] = __float_as_uint(x0); w[k + 1] = __float_as_uint(x1);
      }
      tmem_st16(tSum, w);
      tmem_wait_st();
      out_valid = 0;
    };
    uint32_t voff[8];
#pragma unroll
    for (int x = 0; x < 8; x++) voff[x] = (uint32_t) ((((lane >> 2) ^ x) << 4) + ((lane & 3) << 2));

    auto do_step = [&](uint32_t nn, bool ph_h) {
      const uint32_t st = nn % NS;
      mbar_wait(&p_full[wg], (nn >> 1) & 1); // p_full FIRST, v_full afterwards (kernels_nmf_tc.cu, rule 1)
      mbar_wait(&v_full[st], (nn / NS) & 1);
      tc_fence_after();
      uint32_t pp[64];
      tmem_ld32(tP, *reinterpret_cast<uint32_t(*)[32]>(&pp[0]));
      tmem_ld32(tP + 32, *reinterpret_cast<uint32_t(*)[32]>(&pp[32]));
      float v[64];
      if (ph_h) {
        const uint8_t* row = smem + OFF_V + st * STAGE + r * 128;
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int c4 = 0; c4 < 8; c4++) {
            const float4 x = *reinterpret_cast<const float4*>(row + h * 16384 + ((c4 ^ (r & 7)) << 4));
            v[32 * h + 4 * c4] = x.x; v[32 * h + 4 * c4 + 1] = x.y; v[32 * h + 4 * c4 + 2] = x.z; v[32 * h + 4 * c4 + 3] = x.w;
          }
      } else {
        const uint8_t* vt = smem + OFF_V + st * STAGE + q * 8192;
#pragma unroll
        for (int j = 0; j < 64; j++) v[j] = *reinterpret_cast<const float*>(vt + j * 128 + voff[j & 7]);
      }
      tmem_wait_ld();
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
      drain();
      tmem_st16(tR, *reinterpret_cast<uint32_t(*)[16]>(&ph[0]));
      tmem_st16(tR + 16, *reinterpret_cast<uint32_t(*)[16]>(&ph[16]));
      tmem_st16(tR + 32, *reinterpret_cast<uint32_t(*)[16]>(&pl[0]));
      tmem_st16(tR + 48, *reinterpret_cast<uint32_t(*)[16]>(&pl[16]));
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&r_full[wg]);
    };

    // 16 values per thread summed over the 32 lanes with 15 + 1 shuffles; even lane l ends with the total of value l >> 1
    auto butterfly = [&](float (&a)[16]) {
#pragma unroll
      for (int o = 16, cnt = 8; cnt >= 1; o >>= 1, cnt >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (i < cnt) {
            const float send = up ? a[i] : a[i + cnt];
            const float keep = up ? a[i + cnt] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
      }
      a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
    };
    const bool bf_owner = (lane & 1) == 0;
    const int bf_idx = lane >> 1;

    auto frame_partials = [&](const float (&h)[16], float vn) {
      float pn = 0.f;
#pragma unroll
      for (int k = 0; k < K; k++) pn = fmaf(h[k], WN[k], pn);
      const float rn = vn / fmaxf(pn, kEps);
      float a[16];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float x = wg ? h[8 + j] : h[j];
        a[j] = x;            // sum_f H        (NMF.hpp:160)
        a[8 + j] = rn * x;   // Nyquist row of (V / WH) H^T  (:159)
      }
      butterfly(a);
      if (bf_owner) part[ew * 16 + bf_idx] += a[0];
    };

    // The stationary rows of a job: this thread's 8 components of row r of the job's tile, from the fp32 state.
    auto load_stationary = [&](const JobIter& N, float (&x)[8]) {
      const int buf = N.buffer(p);
      if (N.phase == 0) {
        const float4* src = reinterpret_cast<const float4*>(p.H + ((int64_t) buf * Fp + 128 * N.tile + r) * K + k0);
        const float4 a = src[0], b = src[1];
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      } else {
        const float* src = p.W + ((int64_t) (p.shared_w ? 0 : buf) * K + k0) * Bp + 128 * N.tile + r;
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = src[(int64_t) j * Bp];
      }
    };
    // ... into the job's TMEM slot (split operand) and the fp32 stash the tile update reads its old values from
    auto stage_stationary = [&](const JobIter& N, const float (&x)[8]) {
      const uint32_t sl = N.jn & 1;
      float4* st4 = reinterpret_cast<float4*>(stf + (sl * 128 + r) * 16 + k0);
      st4[0] = make_float4(x[0], x[1], x[2], x[3]);
      st4[1] = make_float4(x[4], x[5], x[6], x[7]);
      uint32_t ph[4], pm[4], pl[4];
#pragma unroll
      for (int j = 0; j < 4; j++) split3(x[2 * j], x[2 * j + 1], ph[j], pm[j], pl[j]);
      const uint32_t tA = tbase + TM_A + 24 * sl + 4 * wg + lane_off;
      tmem_st4(tA, ph[0], ph[1], ph[2], ph[3]);
      tmem_st4(tA + 8, pm[0], pm[1], pm[2], pm[3]);
      tmem_st4(tA + 16, pl[0], pl[1], pl[2], pl[3]);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[sl]);
    };

    if (J.valid) { // the first job's stationary operand
      float x[8];
      load_stationary(J, x);
      stage_stationary(J, x);
    }
    bool partials_valid = false;
    int w_tiles_done = 0;
    const int MT = BT / 128;
    for (; J.valid; J.next(p, gridDim.x)) {
      const int buf = J.buffer(p), phase = J.phase, tile = J.tile;
      const int wbuf = p.shared_w ? 0 : buf;
      const float* gV = p.V + (int64_t) buf * Fp * Bp;
      float* gW = p.W + (int64_t) wbuf * K * Bp;
      float* gH = p.H + (int64_t) buf * Fp * K;
      __nv_bfloat16* gWop = p.Wop + wbuf * wop_stride;
      __nv_bfloat16* gHop = p.Hop + buf * hop_stride;
      const uint32_t sl = J.jn & 1;
      if (J.first_of_unit()) {
        // ---------------- unit prologue ---------------------------------------------------------------------------
        epi_bar();
        if (et < K) {
          WN[et] = gW[(int64_t) et * Bp + BT];
          const float hd = p.hden[(int64_t) wbuf * K + et];
          f_hden[et] = hd;
          f_ihd[et] = 1.0f / fmaxf(hd, kEps);
        }
        for (int e = et; e < 8 * 16; e += 256) part[e] = 0.f;
        epi_bar();
        partials_valid = false;
        w_tiles_done = 0;
      }
      if (phase == 1 && w_tiles_done == 0) {
        // ---------------- start of a W half-iteration: denominators + Nyquist numerators from H ----------------------
        if (!partials_valid) {
          for (int f0 = 0; f0 < Fp; f0 += 128) {
            float h[16];
            const float4* src = reinterpret_cast<const float4*>(gH + (int64_t) (f0 + r) * K);
#pragma unroll
            for (int j = 0; j < 4; j++) { const float4 x = src[j]; h[4 * j] = x.x; h[4 * j + 1] = x.y; h[4 * j + 2] = x.z; h[4 * j + 3] = x.w; }
            float vn = gV[(int64_t) (f0 + r) * Bp + BT];
            if (p.clamp_v) vn = fmaxf(vn, kEps);
            frame_partials(h, vn);
          }
        }
        epi_bar();
        if (et < 2 * K) {
          const int kind = et / K, k = et % K;
          const int owner = (k / 8) * 4, idx = kind * 8 + (k % 8);
          float s = 0.f;
          for (int w4 = 0; w4 < 4; w4++) s += part[(owner + w4) * 16 + idx];
          (kind ? f_nyq : f_wden)[k] = s;
        }
        epi_bar();
        for (int e = et; e < 8 * 16; e += 256) part[e] = 0.f;
        for (int e = et; e < 8 * 20; e += 256) red[e] = 0.f;
        epi_bar();
        partials_valid = false;
      }
      // ---------------- steps ----------------------------------------------------------------------------------------
      const int ns = J.steps(C1, S2);
      float vn = 0.f;
      if (phase == 0) {
        vn = gV[(int64_t) (128 * tile + r) * Bp + BT];
        if (p.clamp_v) vn = fmaxf(vn, kEps);
      }
      for (int i = 0; i < ns; i++, n++) {
        if ((int) (n & 1) != wg) continue;
        do_step(n, phase == 0);
        out_valid = 1; out_first = (i == wg); out_par = (n >> 1) & 1;
      }
      drain();
      // the next job's stationary rows are requested now and consumed after this job's tile update (L2 latency hidden)
      JobIter N = J;
      N.next(p, gridDim.x);
      // (the rows of the next job are final: within a half-iteration the jobs touch disjoint tiles, across a boundary
      //  the stationary side was last written a whole half-iteration earlier -- except in the fixed-W mode when a unit has
      //  a single tile, where the next job re-reads the tile this job is about to update: load after the update then)
      const bool same_tile = N.valid && N.phase == phase && N.tile == tile && N.buffer(p) == buf;
      float nx[8];
      if (N.valid && !same_tile) load_stationary(N, nx);
      tc_fence_before();
      epi_bar(); // both warpgroups' sums of the tile are in TMEM
      tc_fence_after();
      uint32_t s0[16], s1[16];
      tmem_ld16(tbase + TM_SUM + lane_off, s0);
      tmem_ld16(tbase + TM_SUM + 16 + lane_off, s1);
      tmem_wait_ld();
      if (phase == 0) {
        // ---------------- H-tile update (NMF.hpp:168-170) ------------------------------------------------------------
        const int f = 128 * tile + r;
        float h[16];
        {
          const float4* src = reinterpret_cast<const float4*>(stf + (sl * 128 + r) * 16);
#pragma unroll
          for (int j = 0; j < 4; j++) { const float4 x = src[j]; h[4 * j] = x.x; h[4 * j + 1] = x.y; h[4 * j + 2] = x.z; h[4 * j + 3] = x.w; }
        }
        float pn = 0.f;
#pragma unroll
        for (int k = 0; k < K; k++) pn = fmaf(h[k], WN[k], pn);
        const float rn = vn / fmaxf(pn, kEps);
#pragma unroll
        for (int k = 0; k < K; k++) {
          const float num = __uint_as_float(s0[k]) + __uint_as_float(s1[k]);
          h[k] = h[k] * fmaf(rn, WN[k], num) * f_ihd[k];
        }
        {
          float4* dst = reinterpret_cast<float4*>(gH + (int64_t) f * K + k0);
          dst[0] = wg ? make_float4(h[8], h[9], h[10], h[11]) : make_float4(h[0], h[1], h[2], h[3]);
          dst[1] = wg ? make_float4(h[12], h[13], h[14], h[15]) : make_float4(h[4], h[5], h[6], h[7]);
          uint32_t ph[4], pm[4], pl[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float x0 = wg ? h[8 + 2 * j] : h[2 * j], x1 = wg ? h[9 + 2 * j] : h[2 * j + 1];
            split3(x0, x1, ph[j], pm[j], pl[j]);
          }
          *reinterpret_cast<uint4*>(gHop + op_index_h(0, f, k0)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4*>(gHop + op_index_h(1, f, k0)) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
          *reinterpret_cast<uint4*>(gHop + op_index_h(2, f, k0)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
        if (p.upd_w) { frame_partials(h, vn); partials_valid = true; }
        if (same_tile) { // single-tile unit of the fixed-W mode: the next job runs on the rows just computed
#pragma unroll
          for (int j = 0; j < 8; j++) nx[j] = wg ? h[8 + j] : h[j];
        }
      } else {
        // ---------------- W-tile update, not yet normalised (NMF.hpp:161) --------------------------------------------
        const int b = 128 * tile + r;
        float wold[8];
        {
          const float4* src = reinterpret_cast<const float4*>(stf + (sl * 128 + r) * 16 + k0);
          const float4 x = src[0], y = src[1];
          wold[0] = x.x; wold[1] = x.y; wold[2] = x.z; wold[3] = x.w; wold[4] = y.x; wold[5] = y.y; wold[6] = y.z; wold[7] = y.w;
        }
        float a[16];
        float mx = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int k = k0 + j;
          const float num = __uint_as_float(s0[k]) + __uint_as_float(s1[k]);
          const float w = wold[j] * num / fmaxf(f_wden[k], kEps);
          gW[(int64_t) k * Bp + b] = w;
          a[j] = w * w;
          a[8 + j] = w;
          mx = fmaxf(mx, w);
        }
        butterfly(a);
        if (bf_owner) red[ew * 20 + bf_idx] += a[0];
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[ew * 20 + 16] = fmaxf(red[ew * 20 + 16], mx);
        w_tiles_done++;
        if (w_tiles_done == MT) {
          // ---------------- end of the W half-iteration: column normalisation (:162), hden (:169), operand refresh ------
          w_tiles_done = 0;
          epi_bar();
          if (et < K) {
            const int owner = (et / 8) * 4, idx = et % 8;
            float s2 = 0.f, s1v = 0.f;
            for (int w4 = 0; w4 < 4; w4++) { s2 += red[(owner + w4) * 20 + idx]; s1v += red[(owner + w4) * 20 + 8 + idx]; }
            const float wn = WN[et] * f_nyq[et] / fmaxf(f_wden[et], kEps);
            f_s2[et] = fmaf(wn, wn, s2);
            f_s1[et] = s1v + wn;
            WN[et] = wn;
          }
          if (et == 64) {
            float gm = 0.f;
            for (int w8 = 0; w8 < 8; w8++) gm = fmaxf(gm, red[w8 * 20 + 16]);
            f_gm[0] = gm;
          }
          epi_bar();
          if (et < K) {
            float gm = f_gm[0];
            for (int k = 0; k < K; k++) gm = fmaxf(gm, WN[k]);
            const float s2 = f_s2[et], s1v = f_s1[et];
            const bool norm = gm > kEps;
            const float inv = norm ? (s2 > 0.f ? 1.0f / sqrtf(s2) : 0.f) : 1.0f;
            f_inv[et] = inv;
            const float hd = s1v * inv;
            f_hden[et] = hd;
            f_ihd[et] = 1.0f / fmaxf(hd, kEps);
            p.hden[(int64_t) wbuf * K + et] = hd;
          }
          epi_bar();
          if (et < K) WN[et] *= f_inv[et];
          const int n_items = (BT / 8) * K;
          for (int base = et; base < n_items; base += 4 * 256) {
            float4 x0[4], x1[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int it8 = base + 256 * u;
              if (it8 < n_items) {
                const float4* src = reinterpret_cast<const float4*>(gW + (int64_t) (it8 % K) * Bp + 8 * (it8 / K));
                x0[u] = src[0]; x1[u] = src[1];
              }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int it8 = base + 256 * u;
              if (it8 < n_items) {
                const int k = it8 % K, blk = it8 / K;
                const float s = f_inv[k];
                float4 a4 = x0[u], b4 = x1[u];
                a4.x *= s; a4.y *= s; a4.z *= s; a4.w *= s; b4.x *= s; b4.y *= s; b4.z *= s; b4.w *= s;
                float4* dstw = reinterpret_cast<float4*>(gW + (int64_t) k * Bp + 8 * blk);
                dstw[0] = a4; dstw[1] = b4;
                uint32_t ph[4], pm[4], pl[4];
                split3(a4.x, a4.y, ph[0], pm[0], pl[0]); split3(a4.z, a4.w, ph[1], pm[1], pl[1]);
                split3(b4.x, b4.y, ph[2], pm[2], pl[2]); split3(b4.z, b4.w, ph[3], pm[3], pl[3]);
                __nv_bfloat16* dst = gWop + op_index_w(0, k, 8 * blk);
                *reinterpret_cast<uint4*>(dst) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<uint4*>(dst + KB * 64) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
                *reinterpret_cast<uint4*>(dst + 2 * KB * 64) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
              }
            }
          }
          if (et < K) gW[(int64_t) et * Bp + BT] = WN[et];
          // the first W tile of the NEXT iteration is the stationary operand of a job that may already have been loaded
          // above (when this was also the last job before an H-less W half-iteration): only the W-only mode re-reads W
          // tiles this soon, and there the load must see the rescaled rows
          if (N.valid && N.phase == 1 && N.buffer(p) == buf) {
            epi_bar(); // the rescale sweep of all threads is complete
            load_stationary(N, nx);
          }
        }
      }
      // ---------------- stage the next job's stationary operand, publish this job ---------------------------------------
      if (N.valid) stage_stationary(N, nx);
      fence_async_all();
      epi_bar();
      if (et == 0) *jobs_done = J.jn + 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

// ---------------------------------------------------------------------------------------------------------------
bool tcr_eligible(const NmfDev& d)
{
  const int BT = d.B - 1;
  return d.KP == 16 && BT >= 128 && (BT % 128) == 0 && d.Bp == d.B + 3 && d.Fp >= 128 && (d.Fp % 128) == 0;
}

int32_t tcr_run(Plan* p, const NmfDev& d, int iters, bool upd_w, bool upd_h)
{
  const int BT = d.B - 1;
  const int nw = d.shared_w ? 1 : d.batch;
  FB_CUDA(p, p->wop_buf.ensure((size_t) nw * (BT / 8) * ROWB));
  FB_CUDA(p, p->hop_buf.ensure((size_t) d.batch * (d.Fp / 8) * ROWB));
  Params q{};
  q.V = d.V; q.W = d.W; q.H = d.H; q.hden = d.hden;
  q.Wop = p->wop_buf.as<__nv_bfloat16>(); q.Hop = p->hop_buf.as<__nv_bfloat16>();
  q.batch = d.batch; q.Fp = d.Fp; q.Bp = d.Bp; q.BT = BT;
  q.iters = iters; q.upd_w = upd_w ? 1 : 0; q.upd_h = upd_h ? 1 : 0; q.shared_w = d.shared_w; q.clamp_v = d.clamp_v;
  const int T = d.Fp / 128;
  q.units = upd_w ? d.batch : d.batch * ((T + 1) / 2);
  alignas(64) CUtensorMap tmap1, tmap2;
  FB_TRY(make_v_tensor_map(p, &tmap1, d.V, d.Bp, d.Fp, d.batch, 128));
  FB_TRY(make_v_tensor_map(p, &tmap2, d.V, d.Bp, d.Fp, d.batch, 64));
  FB_CUDA(p, cudaFuncSetAttribute(k_nmf_tcr, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  {
    const int64_t items = (int64_t) nw * (BT / 8) * K + (int64_t) d.batch * d.Fp * KB;
    const int blocks = (int) std::min<int64_t>((items + 255) / 256, 148 * 16);
    k_tcr_pack<<<blocks, 256, 0, p->stream>>>(d.W, d.H, q.Wop, q.Hop, nw, d.batch, d.Fp, d.Bp, BT);
    p->launches++;
  }
  const int grid = std::min(q.units, p->sm_count);
  while (p->kev.size() < p->kev_used + 2) { cudaEvent_t e; FB_CUDA(p, cudaEventCreate(&e)); p->kev.push_back(e); }
  cudaEventRecord(p->kev[p->kev_used++], p->stream);
  k_nmf_tcr<<<grid, NTHREADS, SMEM_BYTES, p->stream>>>(q, tmap1, tmap2);
  cudaEventRecord(p->kev[p->kev_used++], p->stream);
  p->launches++; p->launches_nmf++;
  FB_CUDA(p, cudaGetLastError());
  return FB200_OK;
}

} // namespace fb200
