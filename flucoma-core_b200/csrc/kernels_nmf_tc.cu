// tcgen05 / TMEM / TMA engine for the NMF multiplicative updates (algorithms/public/NMF.hpp:144-183), rank 16.
//
// One persistent CTA owns one buffer for ALL iterations.  W and H live in shared memory as exact three-way bf16
// splits (x = hi + mid + lo reproduces the fp32 value), stored directly in UMMA core-matrix layout, so the operand
// copies ARE the state.  V = |X| is streamed by TMA (128B swizzle) once per phase, WH is accumulated in TMEM, the ratio
// V / max(WH, eps) is computed by the epilogue warps straight out of TMEM and written back to TMEM (as a 2-way bf16
// split) as the A operand of the second MMA -- neither WH nor the ratio ever touch shared or global memory, and there
// is no inter-CTA communication.
//
//   phase 1 (H-update of a 128-frame tile t, NMF.hpp:165-170), per 64-bin chunk c:
//       P[f][b]    = H_t W_c           tcgen05.mma SS   A = H blocks (K-major)   B = W blocks (MN-major)   M128 N64 K16
//       R[f][b]    = V / max(P, eps)   epilogue: tcgen05.ld, swizzled LDS of the TMA tile, rcp, 2-way split, tcgen05.st
//       hnum[f][k] += R W_c^T          tcgen05.mma TS   A = R (TMEM)             B = W blocks (K-major)    M128 N32+N16 K64
//     then H <- H * hnum / max(hden, eps) for the tile ("tile prep").
//   phase 2 (this tile's share of the next W-update, NMF.hpp:158-160), per 128-bin tile m and 64-frame half s:
//       P[b][f]    = W_m^T H_ts^T      SS   A = W blocks (MN-major)  B = H blocks (K-major)    M128 N64 K16
//       R[b][f]    = V / max(P, eps)
//       wnum[b][k] += R H_ts           TS   A = R (TMEM)             B = H blocks (MN-major)   M128 N32+N16 K64
//   after the last tile: W <- W * wnum / max(wden, eps), conditional column normalisation (:161-162), hden = sum_b W.
//
// Precision: W and H are carried as exact three-way splits (the STATE must not be rounded: 16-bit state measured 1.3e-4
// against the fp64 CPU restatement after 200 iterations -- the NMF dynamics amplify a per-iteration rounding about 100x).
// The MMAs use what their consumers can resolve: W.H keeps the terms down to 2^-16 (hh, hm, mh), because the ratio made
// from it is itself a two-part operand, and the second MMA multiplies R_hi by [X_hi | X_mid] (one instruction, N = 32,
// the parts sit next to each other in shared memory) and R_lo by X_hi (N = 16, its own accumulator columns).  The
// epilogue sums the three 16-column groups of every step's FRESH accumulator with round-to-nearest adds (the tensor
// core's accumulate truncates; accumulated over a whole job in TMEM that bias alone measured 1.3e-4).
// Errors against the fp64 CPU restatement after 200 iterations: see tools/tc_margin.py and profiles/r02k_experiments.txt.
//
// The Nyquist bin (B = 2^m + 1) does not fit the 128-wide tiles; its column is carried on the SIMT side of the epilogue
// (a 16-term dot product per frame), so the tensor tiles cover bins 0 .. B-2 exactly.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = first-stage MMA issuer (+ TMEM owner), warps 2/3 = second-
// stage MMA issuers (one per epilogue warpgroup), warps 4-11 = two epilogue warpgroups that alternate steps (ping-pong on
// two P/R/accumulator TMEM buffers).  All reductions are fixed-order: results are bitwise repeatable.  The rules the
// mbarrier protocol relies on are listed in DESIGN.md ("Synchronisation rules"); each has a regression test.
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda.h>
#include <cstdio>
#include <cstdlib>

namespace fb200 {
using namespace tc;

namespace tcn {
constexpr int K = 16;
constexpr int KB = 2;            // 8-component blocks
constexpr int NS = 3;            // V ring stages of 32 KB
constexpr int STAGE = 32768;
constexpr int BT_MAX = 512;      // tensor bins (B - 1)
constexpr int FP_MAX = 512;      // padded frames
constexpr int NTHREADS = 384;    // warpgroup 0: producer, issuer, 2 idle warps; warpgroups 1, 2: epilogue
#ifdef FB200_TC_ROWPAD
constexpr uint32_t ROWB = 3 * KB * 128 + 16; // bytes per 8-bin (W) / 8-frame (H) block row: 3 parts x 2 component blocks, + 16 so that per-row accesses of consecutive block rows fall into different banks
#else
constexpr uint32_t ROWB = 3 * KB * 128; // bytes per 8-bin (W) / 8-frame (H) block row: 3 parts x 2 component blocks
#endif

// TMEM columns
constexpr uint32_t TM_P = 0;      // + 64 g
constexpr int RP = 2;             // parts of the ratio operand R written to TMEM (3 = exact fp32, 2 = hi + lo)
constexpr uint32_t RCOLS = 32 * RP;
constexpr uint32_t TM_R = 128;    // + RCOLS g : hi [0,32) mid [32,64) lo [64,96)
// per-step partial of the second MMA: [0,16) leading term R_hi X_hi, [16,32) R_hi X_mid, [32,48) R_lo X_hi ([48,64) unused:
// R_hi X_lo is of the order of the dropped R_lo X_mid and made no measurable difference).
// Every accumulation chain owns its columns: only MMAs of the same shape on the same accumulator address are
// pipelined in issue order; chains of different shape must not touch each other's destination columns.
constexpr uint32_t ACOLS = 64;
constexpr uint32_t TM_ACC = 256;  // + ACOLS g
constexpr uint32_t TM_WSUM = 384; // + 64 wg + 16 m : fp32 running sums of the W numerator, added by the epilogue (RN)

// shared memory map (bytes)
constexpr int OFF_V = 0;
constexpr int OFF_WOP = OFF_V + NS * STAGE;                 // bf16 [BT/8][3][KB][8][8]
constexpr int OFF_HOP = OFF_WOP + (BT_MAX / 8) * (int) ROWB; // bf16 [Fp/8][3][KB][8][8]
constexpr int OFF_VN = OFF_HOP + (FP_MAX / 8) * (int) ROWB;  // float [FP_MAX]  Nyquist column of V
constexpr int OFF_WN = OFF_VN + FP_MAX * 4;                 // float [K]       Nyquist row of W
constexpr int OFF_HDEN = OFF_WN + K * 4;                    // float [K]
constexpr int OFF_PART = OFF_HDEN + K * 4;                  // float [4 tiles][8 warps][32]
constexpr int OFF_RED = OFF_PART + 4 * 8 * 32 * 4;          // float [8 warps][36]
constexpr int OFF_FIN = OFF_RED + 8 * 36 * 4;               // float [64]
constexpr int OFF_HS = OFF_FIN + 64 * 4;                    // float [4 j4][2 wg][128 rows][4] per-warpgroup partial H numerators (16-byte lane stride)
constexpr int OFF_BAR = OFF_HS + 2 * 128 * 16 * 4;          // mbarriers
constexpr int NBAR = 2 * NS + 2 + 2 + 2 + 2 + 3;
constexpr int OFF_SLOT = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_SLOT + 16;

// element index (in bf16 units) of W[k][b] / H[f][k], split part `part`
__device__ __forceinline__ int wop_index(int part, int k, int b) { return (b >> 3) * (int) (ROWB / 2) + ((part * KB + (k >> 3)) << 6) + ((k & 7) << 3) + (b & 7); }
__device__ __forceinline__ int hop_index(int part, int f, int k) { return (f >> 3) * (int) (ROWB / 2) + ((part * KB + (k >> 3)) << 6) + ((f & 7) << 3) + (k & 7); }

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// warp q of epilogue warpgroup 0 with warp q of warpgroup 1 (named barriers 2..5, 64 threads)
__device__ __forceinline__ void pair_bar(int q) { asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory"); }
// register re-distribution between the control warpgroup and the epilogue warpgroups (whole warpgroup executes it)
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 88;" ::: "memory"); }
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory"); }

// Block schedule of one pass, shared by every warp role.
enum BlockKind { BLK_P1 = 0, BLK_PREP = 1, BLK_P2 = 2, BLK_W = 3 };
__device__ __forceinline__ int block_count(bool p1, bool p2, int T) { return (p1 && p2) ? 3 * T + 1 : (p1 ? 2 * T : 2 * T + 1); }
__device__ __forceinline__ void block_at(bool p1, bool p2, int T, int i, int& kind, int& t)
{
  if (p1 && p2) { // [P1(t) PREP(t) P2(t)] W
    // (Running phase 1 of tile t+1 ahead of phase 2 of tile t would take the prep's barrier round trip off the critical
    // path, but measured 9 % slower: phase 2 then re-reads a V tile that is 512 KB "old" per CTA instead of 256 KB, and
    // with 148 CTAs streaming that falls out of L2.)
    if (i == 3 * T) { kind = BLK_W; t = 0; }
    else { const int rr = i % 3; t = i / 3; kind = rr == 0 ? BLK_P1 : (rr == 1 ? BLK_PREP : BLK_P2); }
  } else if (p1) { // [P1(t) PREP(t)]
    kind = (i & 1) ? BLK_PREP : BLK_P1;
    t = i >> 1;
  } else { // [PREP(t) P2(t)] W
    if (i == 2 * T) { kind = BLK_W; t = 0; }
    else { kind = (i & 1) ? BLK_P2 : BLK_PREP; t = i >> 1; }
  }
}
// one call site for the body (a callback invoked from several places would be inlined several times)
// `rev`: walk the tiles in descending order.  Passes alternate direction so that the V tiles touched last in one pass
// are the first ones needed by the next pass: with 148 CTAs streaming 157 MB of V through a 126 MB L2 a fixed cyclic
// order never hits, a boustrophedon order re-uses whatever the L2 still holds.  (Any tile order is a valid schedule:
// tiles only meet in the fixed-order sums of the W-update.)
template <class F>
__device__ __forceinline__ void for_blocks(bool p1, bool p2, int T, bool rev, F&& f)
{
  const int nb = block_count(p1, p2, T);
#pragma unroll 1
  for (int i = 0; i < nb; i++) {
    int kind, t;
    block_at(p1, p2, T, i, kind, t);
    f(kind, (rev && kind != BLK_W) ? T - 1 - t : t);
  }
}

struct Sched {
  int npass, both, upd_w, upd_h, iters;
  __device__ bool p1(int pass) const { return both ? pass > 0 : upd_h != 0; }
  __device__ bool p2(int pass) const { return both ? pass < iters : upd_w != 0; }
};

// Asynchronous progress / cancel (fb200_nmf_args.progress_stride == FB200_PROGRESS_ASYNC; NMFClient.hpp:261-274 polls a
// FluidTask every iteration).  `ctrl` is device memory: ctrl[0] is the cancel request (the host raises a word in
// host-mapped memory, CTA 0 relays it),
// ctrl[1 + cta] counts the (buffer, pass) units this CTA has finished.  The TMA producer is the role that runs furthest
// ahead, so it samples the cancel word once per pass (the load is issued one pass early: no L2 round trip on its
// critical path) and publishes the first cancelled global pass index in shared memory; every role evaluates the same
// pure function of (pass, global pass index, published index), so all of them leave the loops at the same point.
// A cancelled buffer finishes the iteration it is in (fused schedule: W is one update ahead of H between passes, so one
// H-only pass follows); buffers not yet started keep their initial state.
enum PassMode { PASS_RUN = 0, PASS_FINAL = 1, PASS_STOP = 2 };
struct Ctl {
  volatile uint32_t* cs; // shared: cs[1] = number of passes decided, cs[2] = first cancelled global pass index
  bool on;
  __device__ __forceinline__ int mode(const Sched& sc, int pass, uint32_t gp) const
  {
    if (!on) return PASS_RUN;
    while (cs[1] <= gp) {}
    __threadfence_block();
    const uint32_t sg = cs[2];
    if (gp < sg) return PASS_RUN;
    return (sc.both && pass > 0 && gp == sg) ? PASS_FINAL : PASS_STOP;
  }
};
} // namespace tcn

using namespace tcn;

// developer timeline: CTA 0 records clock64() at a few points of steps [DBG_N0, DBG_N0 + 32) when a debug buffer is given
#define DBG_N0 192
// (compiled in only with -DFB200_TC_DEBUG_TIMELINE: the marks cost issue slots and registers in the per-step path)
#ifdef FB200_TC_DEBUG_TIMELINE
#define DBG_MARK(slot, nn) do { if (dbg && blockIdx.x == 0 && (nn) >= DBG_N0 && (nn) < DBG_N0 + 32 && lane == 0) dbg[((nn) - DBG_N0) * 16 + (slot)] = clock64(); } while (0)
#else
#define DBG_MARK(slot, nn) do { (void) dbg; } while (0)
#endif

__global__ void __launch_bounds__(NTHREADS, 1)
k_nmf_tc(NmfDev d, const __grid_constant__ CUtensorMap tmap1, const __grid_constant__ CUtensorMap tmap2, int iters, int upd_w, int upd_h,
         long long* dbg, unsigned int* ctrl, const unsigned int* host_cancel)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  __nv_bfloat16* wop = reinterpret_cast<__nv_bfloat16*>(smem + OFF_WOP);
  __nv_bfloat16* hop = reinterpret_cast<__nv_bfloat16*>(smem + OFF_HOP);
  float* VN = reinterpret_cast<float*>(smem + OFF_VN);
  float* WN = reinterpret_cast<float*>(smem + OFF_WN);
  float* hden = reinterpret_cast<float*>(smem + OFF_HDEN);
  float* part = reinterpret_cast<float*>(smem + OFF_PART);
  float* red = reinterpret_cast<float*>(smem + OFF_RED);
  float* fin = reinterpret_cast<float*>(smem + OFF_FIN);
  float* hs = reinterpret_cast<float*>(smem + OFF_HS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* v_full = bars;            // [NS]
  uint64_t* v_empty = bars + NS;      // [NS]
  uint64_t* p_full = bars + 2 * NS;   // [2]
  uint64_t* r_full = p_full + 2;      // [2]
  uint64_t* b_full = r_full + 2;      // [2] second MMA of a step retired: its partial sits in TM_ACC + 32 g
  // three separate "operands ready" barriers so that two completions can never pile up unobserved on one of them
  uint64_t* p_free = b_full + 2;      // [2] the epilogue holds P of its current step in registers: buffer reusable
  uint64_t* buf_ready = p_free + 2;     // buffer prologue done
  uint64_t* prep_ready = buf_ready + 1; // tile prep (H-update) done
  uint64_t* w_ready = prep_ready + 1;   // W-update done
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + OFF_SLOT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Fp = d.Fp, Bp = d.Bp, BT = d.B - 1;
  const int T = Fp / 128, MT = BT / 128, C1 = BT / 64;
  Sched sc;
  sc.both = upd_w && upd_h; sc.upd_w = upd_w; sc.upd_h = upd_h; sc.iters = iters;
  sc.npass = sc.both ? iters + 1 : iters;

  Ctl ctl;
  ctl.cs = slot; ctl.on = ctrl != nullptr;
  if (tid == 0) {
    slot[1] = 0u; slot[2] = 0xFFFFFFFFu;
    for (int i = 0; i < NS; i++) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 4); }
    for (int i = 0; i < 2; i++) { mbar_init(&p_full[i], 1); mbar_init(&r_full[i], 4); mbar_init(&b_full[i], 1); mbar_init(&p_free[i], 4); }
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

  if (warp < 4) {
  reg_dec(); // control warpgroup gives registers back; the epilogue warpgroups take them
  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      uint32_t n = 0, gp = 0, cancel_pre = 0, relay_pre = 0;
      bool stop = false;
      for (int buf = blockIdx.x; buf < d.batch && !stop; buf += gridDim.x) {
        for (int pass = 0; pass < sc.npass; pass++, gp++) {
          if (ctl.on) { // decide this pass with the cancel word sampled one pass ago, publish, sample for the next
            if (cancel_pre && slot[2] == 0xFFFFFFFFu) slot[2] = gp;
            __threadfence_block();
            slot[1] = gp + 1;
            cancel_pre = *reinterpret_cast<volatile unsigned int*>(ctrl);
            // CTA 0 relays the host's request (a word in host-mapped memory, one PCIe read per pass of ONE CTA) into the
            // device word everybody samples; the value read now is used one pass later, like cancel_pre
            if (blockIdx.x == 0) {
              if (relay_pre) *reinterpret_cast<volatile unsigned int*>(ctrl) = 1u;
              relay_pre = *reinterpret_cast<const volatile unsigned int*>(host_cancel);
            }
          }
          const int mode = ctl.mode(sc, pass, gp);
          if (mode == PASS_STOP) { stop = true; break; }
          for_blocks(sc.p1(pass), sc.p2(pass) && mode == PASS_RUN, T, (pass & 1) != 0, [&](int kind, int t) {
            if (kind == BLK_P1)
              for (int c = 0; c < C1; c++, n++) {
                const uint32_t st = n % NS, k = n / NS;
                mbar_wait(&v_empty[st], (k & 1) ^ 1);
                mbar_arrive_expect_tx(&v_full[st], STAGE);
                uint8_t* dst = smem + OFF_V + st * STAGE;
                tma_load_3d(dst, &tmap1, 64 * c, 128 * t, buf, &v_full[st]);
                tma_load_3d(dst + 16384, &tmap1, 64 * c + 32, 128 * t, buf, &v_full[st]);
              }
            else if (kind == BLK_P2)
              for (int m = 0; m < MT; m++)
                for (int s = 0; s < 2; s++, n++) {
                  const uint32_t st = n % NS, k = n / NS;
                  mbar_wait(&v_empty[st], (k & 1) ^ 1);
                  mbar_arrive_expect_tx(&v_full[st], STAGE);
                  uint8_t* dst = smem + OFF_V + st * STAGE;
#pragma unroll
                  for (int w = 0; w < 4; w++) tma_load_3d(dst + w * 8192, &tmap2, 128 * m + 32 * w, 128 * t + 64 * s, buf, &v_full[st]);
                }
          });
          if (mode == PASS_FINAL) { gp++; break; }
        }
      }
    }
    __syncwarp(); // lanes 1..31 wait here, not inside the CTA-wide barrier at the end of the kernel
  } else if (warp == 1) {
    // =========================================== MMA issuer =============================================
    { // the whole warp runs the issue loop converged; the MMA / commit wrappers elect one lane
      // The issuer is ONE thread and every step needs 22 MMAs, so its instruction count per MMA is what paces the whole
      // pipeline (measured: with naive descriptor construction the kernel ran at the same speed with the epilogue
      // math removed).  All descriptors are therefore pre-split into a constant high word and a low word that only
      // needs one integer add per MMA (the 14-bit start-address field never carries into the LBO field).
      const uint32_t wop_a = smem_u32(wop), hop_a = smem_u32(hop);
      constexpr uint32_t ID_P1A = make_idesc_bf16(128, 64, 0, 1);
      constexpr uint32_t ID_P1B32 = make_idesc_bf16(128, 32, 0, 0), ID_P1B16 = make_idesc_bf16(128, 16, 0, 0);
      constexpr uint32_t ID_P2A = make_idesc_bf16(128, 64, 1, 0);
      constexpr uint32_t ID_P2B32 = make_idesc_bf16(128, 32, 0, 1), ID_P2B16 = make_idesc_bf16(128, 16, 0, 1);
      constexpr uint32_t HI_A = (ROWB >> 4) | (1u << 14);  // first MMA operands: SBO = ROWB (block rows), version 1
      constexpr uint32_t HI_B = (128u >> 4) | (1u << 14);  // second MMA B operand: SBO = 128 (parts / component blocks along N)
      constexpr uint32_t LO_A = (128u >> 4) << 16;         // LBO = 128 (component blocks along K)
      constexpr uint32_t LO_B = (ROWB >> 4) << 16;         // LBO = ROWB (block rows along K)
      constexpr uint32_t RSTEP = ROWB >> 4;                // one block row, in descriptor address units
      constexpr uint32_t PSTEP = (KB * 128) >> 4;          // one split part
      const uint32_t wlo_a = (wop_a >> 4) | LO_A, hlo_a = (hop_a >> 4) | LO_A; // first MMA, part 0, block row 0
      const uint32_t wlo_b = (wop_a >> 4) | LO_B, hlo_b = (hop_a >> 4) | LO_B; // second MMA B operand
      uint32_t n = 0, buf_cnt = 0, prep_cnt = 0, w_cnt = 0;
      // Two issuing warps with one in-order event stream each: this warp issues the first MMA of every step ("A(n)"), as
      // soon as its P buffer is free [p_free(n-2): the epilogue signals it early in its step, once P(n-2) sits in
      // registers]; warp 2 issues the second MMAs ("B(n)") as soon as the ratio is in TMEM [r_full(n), end of the step].
      // Neither stream ever queues behind the other's barrier, so each warpgroup finds its next P tile ready.
      auto issue_a = [&](uint32_t alo, uint32_t blo, uint32_t idesc) {
        const uint32_t g = n & 1;
        DBG_MARK(0, n);
        if (n >= 2) mbar_wait(&p_free[g], ((n - 2) >> 1) & 1);
        tc_fence_after();
        DBG_MARK(1, n);
        const uint32_t dP = tbase + TM_P + 64 * g;
        // the terms of (A_hi + A_mid + A_lo)(B_hi + B_mid + B_lo) down to 2^-16: the ratio that is made from P is itself
        // carried as a two-part (16-bit) operand, so the 2^-16 .. 2^-24 terms (hi lo, lo hi, mid mid) bought nothing
        // measurable against the fp64 CPU restatement and cost half of the first MMA's operand traffic (profiles/r02k_experiments.txt)
        mma_ss_lohi<0>(dP, alo, HI_A, blo, HI_A, idesc);                          // hi  hi
        mma_ss_lohi<1>(dP, alo, HI_A, blo + PSTEP, HI_A, idesc);                  // hi  mid
        mma_ss_lohi<1>(dP, alo + PSTEP, HI_A, blo, HI_A, idesc);                  // mid hi
        mma_commit_warp(&p_full[g]);
        DBG_MARK(2, n);
        n++;
      };
      uint32_t gp = 0;
      bool stop = false;
      for (int buf = blockIdx.x; buf < d.batch && !stop; buf += gridDim.x) {
        if (ctl.mode(sc, 0, gp) == PASS_STOP) break; // cancelled before this buffer started: it keeps its initial state
        mbar_wait(buf_ready, buf_cnt & 1); buf_cnt++; // operands of this buffer are in shared memory
        tc_fence_after();
        for (int pass = 0; pass < sc.npass; pass++, gp++) {
          const int mode = ctl.mode(sc, pass, gp);
          if (mode == PASS_STOP) { stop = true; break; }
          const bool p1 = sc.p1(pass), p2 = sc.p2(pass) && mode == PASS_RUN;
          int prep_owed = 0; // tile preps whose completion this warp has not consumed yet (each must be observed
                             // before the next one can complete: see the placement of the waits below)
          for_blocks(p1, p2, T, (pass & 1) != 0, [&](int kind, int t) {
            if (kind == BLK_P1) {
              for (int c = 0; c < C1; c++) issue_a(hlo_a + 16 * t * RSTEP, wlo_a + 8 * c * RSTEP, ID_P1A); // A = H rows of tile t, B = W rows of chunk c
              if (!p2 && prep_owed) { mbar_wait(prep_ready, prep_cnt & 1); prep_cnt++; prep_owed--; }       // prep(t-1), long done
            } else if (kind == BLK_PREP) {
              prep_owed++;
              if (!p1) { mbar_wait(prep_ready, prep_cnt & 1); prep_cnt++; prep_owed--; tc_fence_after(); }  // phase 2 follows directly
            } else if (kind == BLK_P2) {
              if (prep_owed) { mbar_wait(prep_ready, prep_cnt & 1); prep_cnt++; prep_owed--; }              // H_op(t) updated
              tc_fence_after();
              for (int m = 0; m < MT; m++)
                for (int s = 0; s < 2; s++) issue_a(wlo_a + 16 * m * RSTEP, hlo_a + (16 * t + 8 * s) * RSTEP, ID_P2A); // A = W rows of tile m, B = H rows of half s
            } else {
              mbar_wait(w_ready, w_cnt & 1); w_cnt++; // W-update done: W_op rewritten
              tc_fence_after();
            }
          });
          while (prep_owed) { mbar_wait(prep_ready, prep_cnt & 1); prep_cnt++; prep_owed--; }
          tc_fence_after();
          if (mode == PASS_FINAL) { gp++; break; }
        }
      }
    }
  } else {
    // =========================================== MMA issuers, second stage ==============================
    // One issuing warp PER epilogue warpgroup (warp 2: even steps, warp 3: odd steps).  With a single in-order stream
    // the second MMA of step n queued behind that of step n-1, which belongs to the other warpgroup: whenever the two
    // warpgroups were out of phase, B(n) was issued ~1000 cycles after its ratio was ready and the owner stalled in
    // its next step collecting it (developer timeline).
    {
      const uint32_t myg = warp - 2;
      const uint32_t wop_a = smem_u32(wop), hop_a = smem_u32(hop);
      constexpr uint32_t ID_P1B32 = make_idesc_bf16(128, 32, 0, 0), ID_P1B16 = make_idesc_bf16(128, 16, 0, 0);
      constexpr uint32_t ID_P2B32 = make_idesc_bf16(128, 32, 0, 1), ID_P2B16 = make_idesc_bf16(128, 16, 0, 1);
      constexpr uint32_t HI_B = (128u >> 4) | (1u << 14);  // B operand: SBO = 128 (parts / component blocks along N)
      constexpr uint32_t LO_B = (ROWB >> 4) << 16;         // LBO = ROWB (block rows along K)
      constexpr uint32_t RSTEP = ROWB >> 4;
      const uint32_t wlo_b = (wop_a >> 4) | LO_B, hlo_b = (hop_a >> 4) | LO_B;
      uint32_t n = 0;
      // The operands read here are rewritten only after the epilogue has collected every partial (b_full) of the phase
      // that used them, so this stream needs no barrier besides r_full.
      auto issue_b = [&](uint32_t blo, uint32_t id32, uint32_t id16) {
        const uint32_t g = n & 1;
        if (g != myg) { n++; return; }
        mbar_wait(&r_full[g], (n >> 1) & 1);
        tc_fence_after();
        DBG_MARK(4, n);
        const uint32_t rbase = tbase + TM_R + RCOLS * g;
        const uint32_t dacc = tbase + TM_ACC + ACOLS * g;
        // The tensor core adds into the fp32 accumulator with truncation, so every accumulate costs up to one ulp of
        // the accumulator, always in the same direction.  Columns [0,16) therefore receive ONLY the leading term
        // R_hi X_hi (one add per K-step); all correction terms go to the small-magnitude columns [16,48).
        // The two accumulation chains own disjoint columns: MMAs of different shape / accumulator address are not
        // ordered against each other by the tensor pipe, only same-shape same-accumulator chains are.
#pragma unroll
        for (int j = 0; j < 4; j++) { // K-step j = 16 bins (phase 1) / 16 frames (phase 2) = block rows 2j, 2j+1 of the step
          const uint32_t b0 = blo + 2 * j * RSTEP; // parts hi, mid, lo side by side along N
          const uint32_t rh = rbase + 8 * j;
          if (j == 0) mma_ts_lohi<0>(dacc, rh, b0, HI_B, id32);                  // R_hi [X_hi | X_mid] -> cols [0,32)
          else mma_ts_lohi<1>(dacc, rh, b0, HI_B, id32);
          if (j == 0) mma_ts_lohi<0>(dacc + 32, rh + 32, b0, HI_B, id16);        // R_lo  X_hi         -> cols [32,48)
          else mma_ts_lohi<1>(dacc + 32, rh + 32, b0, HI_B, id16);
        }
        mma_commit_warp(&b_full[g]); // the epilogue adds this partial to its fp32 running sums
        DBG_MARK(5, n);
        n++;
      };
      uint32_t gp = 0;
      bool stop = false;
      for (int buf = blockIdx.x; buf < d.batch && !stop; buf += gridDim.x)
        for (int pass = 0; pass < sc.npass; pass++, gp++) {
          const int mode = ctl.mode(sc, pass, gp);
          if (mode == PASS_STOP) { stop = true; break; }
          for_blocks(sc.p1(pass), sc.p2(pass) && mode == PASS_RUN, T, (pass & 1) != 0, [&](int kind, int t) {
            if (kind == BLK_P1)
              for (int c = 0; c < C1; c++) issue_b(wlo_b + 8 * c * RSTEP, ID_P1B32, ID_P1B16);           // B = W rows of chunk c
            else if (kind == BLK_P2)
              for (int m = 0; m < MT; m++)
                for (int s = 0; s < 2; s++) issue_b(hlo_b + (16 * t + 8 * s) * RSTEP, ID_P2B32, ID_P2B16); // B = H rows of half s
          });
          if (mode == PASS_FINAL) { gp++; break; }
        }
    }
  }
  } else {
    reg_inc();
    // =========================================== epilogue warps =========================================
    const int et = tid - 128;              // 0..255
    const int wg = (warp - 4) >> 2;        // epilogue warpgroup 0/1
    const int ew = warp - 4;               // 0..7
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int r = 32 * q + lane;           // row (frame or bin) inside a 128-row tile
    const uint32_t lane_off = (uint32_t) (32 * q) << 16;
    const uint32_t tP = tbase + TM_P + 64 * wg + lane_off;
    const uint32_t tR = tbase + TM_R + RCOLS * wg + lane_off;
    const uint32_t tAcc = tbase + TM_ACC + ACOLS * wg + lane_off;
    const uint32_t tWsum = tbase + TM_WSUM + 64 * wg + lane_off; // + 16 m
    uint32_t n = 0;
    // The tensor core truncates when it adds into its fp32 accumulator, so long in-TMEM accumulation chains drift in
    // one direction; the NMF problem has nearly flat directions (almost identical components) along which such a
    // persistent bias piles up over the iterations.  Each step's second MMA therefore starts a fresh partial (4
    // K-steps), and the partials are summed here with round-to-nearest fp32 adds: into this thread's row of `hs` (shared
    // memory, conflict-free 16-byte lane stride; registers are the scarce resource of these warps) for phase 1, into the
    // TM_WSUM columns for phase 2.
    int out_valid = 0, out_phase = 0, out_m = 0, out_first = 0; // this warpgroup's step whose partial is not yet collected
    uint32_t out_par = 0;
    auto drain = [&]() {
      if (!out_valid) return;
      mbar_wait(&b_full[wg], out_par);
      tc_fence_after();
      uint32_t a[32], a2[16];
      tmem_ld32(tAcc, a);
      tmem_ld16(tAcc + 32, a2);
      if (out_phase == 1) {
        tmem_wait_ld();
        float4* hp = reinterpret_cast<float4*>(hs) + (wg * 128 + r); // [j4][wg][row]
#pragma unroll
        for (int j4 = 0; j4 < 4; j4++) { // hs += a[k] + (a[16+k] + a2[k]), two components per instruction
          float x[4];
#pragma unroll
          for (int i = 0; i < 4; i += 2) {
            const int k = 4 * j4 + i;
            add2(x[i], x[i + 1], __uint_as_float(a[16 + k]), __uint_as_float(a[17 + k]), __uint_as_float(a2[k]), __uint_as_float(a2[k + 1]));
            add2(x[i], x[i + 1], __uint_as_float(a[k]), __uint_as_float(a[k + 1]), x[i], x[i + 1]);
          }
          if (!out_first) {
            const float4 h = hp[j4 * 256];
            add2(x[0], x[1], h.x, h.y, x[0], x[1]);
            add2(x[2], x[3], h.z, h.w, x[2], x[3]);
          }
          hp[j4 * 256] = make_float4(x[0], x[1], x[2], x[3]);
        }
      } else {
        uint32_t w[16];
        if (!out_first) tmem_ld16(tWsum + 16 * out_m, w);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < K; k += 2) {
          float x0, x1;
          add2(x0, x1, __uint_as_float(a[16 + k]), __uint_as_float(a[17 + k]), __uint_as_float(a2[k]), __uint_as_float(a2[k + 1]));
          add2(x0, x1, __uint_as_float(a[k]), __uint_as_float(a[k + 1]), x0, x1);
          if (!out_first) add2(x0, x1, __uint_as_float(w[k]), __uint_as_float(w[k + 1]), x0, x1);
          w[k] = __float_as_uint(x0); w[k + 1] = __float_as_uint(x1);
        }
        tmem_st16(tWsum + 16 * out_m, w);
        tmem_wait_st();
      }
      out_valid = 0;
    };
    // phase-2 V addressing: byte offset inside a 128-byte tile row for (frame & 7) = x
    uint32_t voff[8];
#pragma unroll
    for (int x = 0; x < 8; x++) voff[x] = (uint32_t) ((((lane >> 2) ^ x) << 4) + ((lane & 3) << 2));

    // One step of either phase for this thread's row: 64 columns of P against 64 values of V -> split ratio in TMEM.
    // All loads are issued up front and the only mid-step synchronisation (b_full of this warpgroup's previous step,
    // whose second MMA still reads R[g]) is taken after the ratios already sit in registers.
    auto do_step = [&](uint32_t nn, uint32_t st, bool ph1) {
      // p_full FIRST, v_full only afterwards.  A parity wait is only unambiguous if the waiter cannot be a whole phase
      // ahead of the barrier.  v_full[st] is the one barrier whose consecutive phases are waited for by DIFFERENT
      // warpgroups (stage = n % 3, warpgroup = n % 2): use k-1 of this stage belonged to the other warpgroup, so this
      // warpgroup has no wait of its own that orders it behind phase k-1.  p_full(n) does: the issuer releases A(n-1), and
      // hence A(n), only after the other warpgroup has loaded P(n-3), which it does after ITS v_full wait on this stage.
      // (Sampling v_full before p_full was confirmed let a parity test pass on the still incomplete previous phase when
      // TMA was slow - cold start: stale V for the boxes not yet landed, a second expect_tx arrive inside one phase.)
      mbar_wait(&p_full[wg], (nn >> 1) & 1);
      if (q == 0) DBG_MARK(7, nn);
      mbar_wait(&v_full[st], (nn / NS) & 1);
      tc_fence_after();
      if (q == 0) DBG_MARK(8, nn);
      uint32_t p[64];
      tmem_ld32(tP, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
      tmem_ld32(tP + 32, *reinterpret_cast<uint32_t(*)[32]>(&p[32]));
      float v[64];
      if (ph1) { // two boxes [128 frames][32 bins]: this thread's row, 16-byte chunks un-swizzled
        const uint8_t* row = smem + OFF_V + st * STAGE + r * 128;
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int c4 = 0; c4 < 8; c4++) {
            const float4 x = *reinterpret_cast<const float4*>(row + h * 16384 + ((c4 ^ (r & 7)) << 4));
            v[32 * h + 4 * c4] = x.x; v[32 * h + 4 * c4 + 1] = x.y; v[32 * h + 4 * c4 + 2] = x.z; v[32 * h + 4 * c4 + 3] = x.w;
          }
      } else { // box q [64 frames][32 bins]: this thread's bin = lane, one value per frame
        const uint8_t* vt = smem + OFF_V + st * STAGE + q * 8192;
#pragma unroll
        for (int j = 0; j < 64; j++) v[j] = *reinterpret_cast<const float*>(vt + j * 128 + voff[j & 7]);
      }
      tmem_wait_ld();
      // WAR across proxies: the V stage is about to be handed back to the TMA producer, and a generic-proxy LDS that is
      // merely issued is not ordered before the async-proxy refill by the mbarrier alone (seen on a 16-epilogue-warp
      // variant of this kernel, profiles/experiments: whole rows of H off by ~0.5 % about once per 10^5 warp-steps).
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&p_free[wg]); mbar_arrive(&v_empty[st]); } // P buffer and V stage are in registers now
      uint32_t ph[32], pl[32];
#pragma unroll
      for (int j = 0; j < 32; j++) {
#ifdef FB200_TC_EXPERIMENT_NOMATH
        ph[j] = p[2 * j] ^ __float_as_uint(v[2 * j]); pl[j] = p[2 * j + 1] ^ __float_as_uint(v[2 * j + 1]);
        continue;
#endif
        float r0, r1, l0, l1;
        mul2(r0, r1, v[2 * j], v[2 * j + 1], rcp_fast(fmaxf(__uint_as_float(p[2 * j]), kEps)),
             rcp_fast(fmaxf(__uint_as_float(p[2 * j + 1]), kEps)));
        ph[j] = cvt2(r0, r1);
        sub2(l0, l1, r0, r1, bf16lo_to_f(ph[j]), bf16hi_to_f(ph[j]));
        pl[j] = cvt2(l0, l1);
      }
      if (q == 0) DBG_MARK(9, nn);
      drain();
      if (q == 0) DBG_MARK(10, nn);
      tmem_st16(tR, *reinterpret_cast<uint32_t(*)[16]>(&ph[0]));
      tmem_st16(tR + 16, *reinterpret_cast<uint32_t(*)[16]>(&ph[16]));
      tmem_st16(tR + 32, *reinterpret_cast<uint32_t(*)[16]>(&pl[0]));
      tmem_st16(tR + 48, *reinterpret_cast<uint32_t(*)[16]>(&pl[16]));
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&r_full[wg]);
      if (q == 0) DBG_MARK(11, nn);
    };
    static_assert(RP == 2, "the step epilogue writes a two-part ratio");

    uint32_t gp = 0, units_done = 0;
    bool stop = false;
    for (int buf = blockIdx.x; buf < d.batch && !stop; buf += gridDim.x) {
      if (ctl.mode(sc, 0, gp) == PASS_STOP) break; // cancelled before this buffer started: it keeps its initial state
      // ---------------- buffer prologue: state -> 3-way split operands --------------------------------------------
      const float* gW = d.W + (int64_t) buf * K * Bp;
      const float* gH = d.H + (int64_t) buf * Fp * K;
      const float* gV = d.V + (int64_t) buf * Fp * Bp;
      for (int e = et; e < K * Bp; e += 256) {
        int k = e / Bp, b = e - k * Bp;
        float w = gW[e];
        if (b < BT) {
          __nv_bfloat16 hi = __float2bfloat16_rn(w);
          float r1 = w - __bfloat162float(hi);
          __nv_bfloat16 mi = __float2bfloat16_rn(r1);
          __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mi));
          wop[wop_index(0, k, b)] = hi; wop[wop_index(1, k, b)] = mi; wop[wop_index(2, k, b)] = lo;
        } else if (b == BT) {
          WN[k] = w;
        }
      }
      for (int e = et; e < Fp * K; e += 256) {
        float h = gH[e];
        int f = e >> 4, k = e & 15;
        __nv_bfloat16 hi = __float2bfloat16_rn(h);
        float r1 = h - __bfloat162float(hi);
        __nv_bfloat16 mi = __float2bfloat16_rn(r1);
        __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mi));
        hop[hop_index(0, f, k)] = hi; hop[hop_index(1, f, k)] = mi; hop[hop_index(2, f, k)] = lo;
      }
      for (int f = et; f < Fp; f += 256) VN[f] = gV[(int64_t) f * Bp + BT];
      if (et < K) {
        hden[et] = d.hden[(int64_t) buf * K + et];
        fin[48 + et] = 1.0f / fmaxf(hden[et], kEps);
      }
      fence_proxy_async();
      epi_bar();
      if (lane == 0) mbar_arrive(buf_ready);

      for (int pass = 0; pass < sc.npass; pass++, gp++) {
        const int mode = ctl.mode(sc, pass, gp);
        if (mode == PASS_STOP) { stop = true; break; }
        const bool p1 = sc.p1(pass), p2 = sc.p2(pass) && mode == PASS_RUN;
        for_blocks(p1, p2, T, (pass & 1) != 0, [&](int kind, int t) {
          // ---------------- phase 1 steps ------------------------------------------------------------------------
          if (kind == BLK_P1) {
            for (int c = 0; c < C1; c++, n++) {
              if ((int) (n & 1) != wg) continue;
              if (q == 0) DBG_MARK(6, n);
              do_step(n, n % NS, true);
              out_valid = 1; out_phase = 1; out_first = (c == wg); out_par = (n >> 1) & 1;
            }
            drain(); // H numerator of this warpgroup's chunks complete in `hs` (all MMAs reading H_op(t) have retired)
          }
          // ---------------- tile prep: H-update (if p1), W-denominator / Nyquist partials (if p2) ----------------
          else if (kind == BLK_PREP) {
            epi_bar(); // both warpgroups' H-numerator partials of tile t are in `hs`
            const int f = 128 * t + r;
            const int k0 = 8 * wg; // this warpgroup splits / stores / reduces components [k0, k0 + 8) of every row
            float h[16];
#pragma unroll
            for (int kb = 0; kb < KB; kb++) { // H[f][8kb..8kb+7] = hi + mid + lo, one 16-byte row per core matrix
              float acc8[8];
#pragma unroll
              for (int j = 0; j < 8; j++) acc8[j] = 0.f;
#pragma unroll
              for (int pt = 2; pt >= 0; pt--) { // small parts first
                const uint4 u = *reinterpret_cast<const uint4*>(hop + hop_index(pt, f, 8 * kb));
                acc8[0] += bf16lo_to_f(u.x); acc8[1] += bf16hi_to_f(u.x); acc8[2] += bf16lo_to_f(u.y); acc8[3] += bf16hi_to_f(u.y);
                acc8[4] += bf16lo_to_f(u.z); acc8[5] += bf16hi_to_f(u.z); acc8[6] += bf16lo_to_f(u.w); acc8[7] += bf16hi_to_f(u.w);
              }
#pragma unroll
              for (int j = 0; j < 8; j++) h[8 * kb + j] = acc8[j];
            }
            const float vn = VN[f];
            // The warp with the same q in the other warpgroup works on the SAME 128 rows (it owns the other 8 components)
            // and also needs all 16 old values: nobody may store before both have loaded.
            if (p1) pair_bar(q);
            if (p1) {
              float pn = 0.f;
#pragma unroll
              for (int k = 0; k < K; k++) pn = fmaf(h[k], WN[k], pn);
              const float rn = vn / fmaxf(pn, kEps);
#pragma unroll
              for (int j4 = 0; j4 < 4; j4++) { // all 16 new values (the Nyquist term of phase 2 needs the whole new row)
                const float4 a = *reinterpret_cast<const float4*>(hs + ((j4 * 2) * 128 + r) * 4);     // even chunks
                const float4 b = *reinterpret_cast<const float4*>(hs + ((j4 * 2 + 1) * 128 + r) * 4); // odd chunks
                const float num[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                  const int k = 4 * j4 + i;
                  h[k] = h[k] * fmaf(rn, WN[k], num[i]) * fin[48 + k];   // NMF.hpp:170, fin[48+k] = 1 / max(hden[k], eps)
                }
              }
              uint32_t ph[4], pm[4], pl[4];
#pragma unroll
              for (int j = 0; j < 4; j++) {
                const float x0 = wg ? h[8 + 2 * j] : h[2 * j], x1 = wg ? h[9 + 2 * j] : h[2 * j + 1];
                split3(x0, x1, ph[j], pm[j], pl[j]);
              }
              *reinterpret_cast<uint4*>(hop + hop_index(0, f, k0)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
              *reinterpret_cast<uint4*>(hop + hop_index(1, f, k0)) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
              *reinterpret_cast<uint4*>(hop + hop_index(2, f, k0)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
            if (p2) {
              float pn = 0.f;
#pragma unroll
              for (int k = 0; k < K; k++) pn = fmaf(h[k], WN[k], pn);
              const float rn = vn / fmaxf(pn, kEps);
              float a[16]; // [0,8) W denominator terms, [8,16) Nyquist numerator terms of this warpgroup's components
#pragma unroll
              for (int j = 0; j < 8; j++) {
                const float x = wg ? h[8 + j] : h[j];
                a[j] = x;
                a[8 + j] = rn * x;
              }
              // vector butterfly: 16 values x 32 lanes reduced with 16 shuffles; lane l ends with the total of
              // index 8*b16 + 4*b8 + 2*b4 + b2 (bX = bit X of l)
#pragma unroll
              for (int o = 16, cnt = 8; o >= 2; o >>= 1, cnt >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int i = 0; i < cnt; i++) {
                  const float send = up ? a[i] : a[i + cnt];
                  const float keep = up ? a[i + cnt] : a[i];
                  a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
              }
              a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
              if ((lane & 1) == 0) {
                const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                part[(t * 8 + ew) * 32 + (idx < 8 ? k0 + idx : 16 + k0 + idx - 8)] = a[0];
              }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(prep_ready);
            epi_bar(); // `hs` may be overwritten by the next phase-1 block only after every warp has read it
          }
          // ---------------- phase 2 steps ------------------------------------------------------------------------
          else if (kind == BLK_P2) {
            for (int m = 0; m < MT; m++)
              for (int s = 0; s < 2; s++, n++) {
                if ((int) (n & 1) != wg) continue;
                do_step(n, n % NS, false);
                out_valid = 1; out_phase = 2; out_m = m; out_first = (t == ((pass & 1) ? T - 1 : 0)); out_par = (n >> 1) & 1;
              }
          }
        // ---------------- W-update (NMF.hpp:161-162) + hden refresh (:169) --------------------------------------
          else {
          drain(); // last outstanding W-numerator partial of this warpgroup
          tc_fence_before();
          epi_bar(); // all partials of all tiles written, both warpgroups' running sums complete
          tc_fence_after();
          if (et < 32) { // component k = et & 15 was reduced by the 4 warps of warpgroup k >> 3
            const int own = ((et & 15) >> 3) * 4;
            float s = 0.f;
            for (int tt = 0; tt < T; tt++)
              for (int w4 = 0; w4 < 4; w4++) s += part[(tt * 8 + own + w4) * 32 + et];
            fin[et] = et < 16 ? 1.0f / fmaxf(s, kEps) : s; // [0,16) 1 / max(wden, eps) (one reciprocal per component instead of a
                                                           // division per element of the update), [16,32) Nyquist wnum
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
              uint32_t wn[32]; // running sums of the two warpgroups (each saw one 64-frame half of every tile)
              tmem_ld16(tbase + TM_WSUM + 16 * m + lane_off, *reinterpret_cast<uint32_t(*)[16]>(&wn[0]));
              tmem_ld16(tbase + TM_WSUM + 64 + 16 * m + lane_off, *reinterpret_cast<uint32_t(*)[16]>(&wn[16]));
              tmem_wait_ld();
              const int b = 128 * m + r;
              const unsigned short* wp = reinterpret_cast<const unsigned short*>(wop);
#pragma unroll
              for (int k = 0; k < K; k++) {
                float wold = bf16_bits_to_f(wp[wop_index(2, k, b)]) + bf16_bits_to_f(wp[wop_index(1, k, b)]) + bf16_bits_to_f(wp[wop_index(0, k, b)]);
                float num = __uint_as_float(wn[k]) + __uint_as_float(wn[16 + k]);
                float w = wold * num * fin[k];
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
              float w = WN[k] * fin[16 + k] * fin[k];
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
            fin[48 + et] = 1.0f / fmaxf(s1 * inv, kEps);
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
                __nv_bfloat16 hi = __float2bfloat16_rn(w);
                float r1 = w - __bfloat162float(hi);
                __nv_bfloat16 mi = __float2bfloat16_rn(r1);
                __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mi));
                wop[wop_index(0, k, b)] = hi; wop[wop_index(1, k, b)] = mi; wop[wop_index(2, k, b)] = lo;
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
        });
        if (ctl.on && et == 0) *reinterpret_cast<volatile unsigned int*>(ctrl + 1 + blockIdx.x) = ++units_done;
        if (mode == PASS_FINAL) { gp++; break; }
      }
      // ---------------- buffer epilogue: state back to global ----------------------------------------------------
      epi_bar();
      float* oW = d.W + (int64_t) buf * K * Bp;
      float* oH = d.H + (int64_t) buf * Fp * K;
      const unsigned short* wp = reinterpret_cast<const unsigned short*>(wop);
      const unsigned short* hp = reinterpret_cast<const unsigned short*>(hop);
      for (int e = et; e < K * Bp; e += 256) {
        int k = e / Bp, b = e - k * Bp;
        float w = 0.f;
        if (b < BT) w = bf16_bits_to_f(wp[wop_index(2, k, b)]) + bf16_bits_to_f(wp[wop_index(1, k, b)]) + bf16_bits_to_f(wp[wop_index(0, k, b)]);
        else if (b == BT) w = WN[k];
        oW[e] = w;
      }
      for (int e = et; e < Fp * K; e += 256) {
        int f = e >> 4, k = e & 15;
        oH[e] = bf16_bits_to_f(hp[hop_index(2, f, k)]) + bf16_bits_to_f(hp[hop_index(1, f, k)]) + bf16_bits_to_f(hp[hop_index(0, f, k)]);
      }
      if (et < K) d.hden[(int64_t) buf * K + et] = hden[et]; // sum_b W of the W just written: a later launch resumes from it
      epi_bar(); // operands are overwritten by the next buffer's prologue
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

int tc_grid(const Plan* p, const NmfDev& d) { return std::min(d.batch, p->sm_count); }

int32_t tc_run(Plan* p, const NmfDev& d, int iters, bool upd_w, bool upd_h, unsigned int* ctrl, const unsigned int* host_cancel)
{
  alignas(64) CUtensorMap tmap1, tmap2;
  FB_TRY(make_v_tensor_map(p, &tmap1, d.V, d.Bp, d.Fp, d.batch, 128));
  FB_TRY(make_v_tensor_map(p, &tmap2, d.V, d.Bp, d.Fp, d.batch, 64));
  if (!(p->attr_mask & 0x10000u)) {
    FB_CUDA(p, cudaFuncSetAttribute(k_nmf_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    p->attr_mask |= 0x10000u;
  }
  const int grid = tc_grid(p, d);
  while (p->kev.size() < p->kev_used + 2) { cudaEvent_t e; FB_CUDA(p, cudaEventCreate(&e)); p->kev.push_back(e); }
  cudaEventRecord(p->kev[p->kev_used++], p->stream);
  long long* dbg = nullptr;
#ifdef FB200_TC_DEBUG_TIMELINE
  if (getenv("FB200_TC_TIMELINE")) { // developer aid: per-step timeline of CTA 0, dumped to stderr after the launch
    FB_CUDA(p, p->out_b.ensure(sizeof(long long) * 32 * 16));
    FB_CUDA(p, cudaMemsetAsync(p->out_b.p, 0, sizeof(long long) * 32 * 16, p->stream));
    dbg = p->out_b.as<long long>();
  }
#endif
  k_nmf_tc<<<grid, NTHREADS, SMEM_BYTES, p->stream>>>(d, tmap1, tmap2, iters, upd_w ? 1 : 0, upd_h ? 1 : 0, dbg, ctrl, host_cancel);
  if (dbg) {
    std::vector<long long> h(32 * 16);
    FB_CUDA(p, cudaMemcpyAsync(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, p->stream));
    FB_CUDA(p, cudaStreamSynchronize(p->stream));
    long long t0 = 0;
    for (auto v : h) if (v && (!t0 || v < t0)) t0 = v;
    fprintf(stderr, "step | issuer: A_start A_pfree A_done B_start B_rfull B_done | epi: start p_full v_full half0 drained done\n");
    for (int i = 0; i < 32; i++) {
      fprintf(stderr, "%4d |", DBG_N0 + i);
      for (int j = 0; j < 12; j++) fprintf(stderr, " %7lld", h[i * 16 + j] ? h[i * 16 + j] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  cudaEventRecord(p->kev[p->kev_used++], p->stream);
  p->launches++; p->launches_nmf++;
  FB_CUDA(p, cudaGetLastError());
  return FB200_OK;
}

} // namespace fb200
