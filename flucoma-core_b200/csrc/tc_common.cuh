// sm_100a building blocks shared by the tcgen05 NMF engine and its self-test: mbarrier, TMA, TMEM alloc/ld/st,
// tcgen05.mma (SS and TS forms), UMMA shared-memory / instruction descriptors.  Inline PTX only (no CUTLASS).
//
// Descriptor conventions used here (SWIZZLE_NONE "interleaved" canonical layouts; cf. CUTLASS cute/arch/mma_sm100_desc.hpp
// and cute/atom/mma_traits_sm100.hpp): an operand lives in shared memory as 8x8 "core matrices" of 16-bit elements,
// each 128 contiguous bytes (8 rows of 16 bytes).
//   K-major  operand: core matrix = 8 MN-rows x 8 K-elements
//   MN-major operand: core matrix = 8 K-rows  x 8 MN-elements
// In both cases SBO = byte stride between consecutive core matrices along MN, LBO = byte stride along K.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb200 {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t ok;
#ifdef FB200_TRYWAIT_HINT_NS
  // optional suspend-time hint (measured: a large hint delays wake-ups and costs far more than the spin it avoids)
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t) FB200_TRYWAIT_HINT_NS)
               : "memory");
#else
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
#endif
  return ok != 0;
}
// non-blocking probe
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  while (!mbar_try_wait(bar, parity)) {}
}

// ---- fences ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int x, int y, int z, uint64_t* bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(smem_dst)), "l"((uint64_t) tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap)
{
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t) tmap) : "memory");
}
// byte offset of float j (0..31) of row r inside a [rows][32 floats] tile written by TMA with CU_TENSOR_MAP_SWIZZLE_128B
__device__ __forceinline__ uint32_t swz128_off(int r, int j) { return (uint32_t) (r * 128 + ((((j >> 2) ^ (r & 7))) << 4) + ((j & 3) << 2)); }

// ---- TMEM --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols)
{ // whole warp; ncols power of two >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 bit, N consecutive columns: thread i of the warp <-> TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
               "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}

// ---- descriptors -------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  uint64_t d = 0;
  d |= (uint64_t) ((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t) ((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t) ((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t) 1 << 46;
  return d;
}
// instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate.  a_mn / b_mn: 1 = MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn)
{
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) a_mn << 15) | ((uint32_t) b_mn << 16) | ((uint32_t) (N >> 3) << 17) |
         ((uint32_t) (M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
               "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
// Same instructions with the 64-bit descriptors passed as (low, high) 32-bit halves and a compile-time accumulate flag:
// the issuer then needs one integer add per MMA to step through shared memory.  These are WARP-LEVEL calls: all 32 lanes
// must execute them converged, one elected lane issues.  (Issuing from inside an `if (lane == 0)` branch makes the
// compiler wrap every uniform-datapath UTCHMMA in a vote/elect/branch loop: measured ~50 cycles per MMA.)
template <int ACC>
__device__ __forceinline__ void mma_ss_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc)
{
  asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
               "setp.ne.b32 p, %6, 0;\n\t@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
               "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "n"(ACC)
               : "memory");
}
template <int ACC>
__device__ __forceinline__ void mma_ts_lohi(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc)
{
  asm volatile("{\n\t.reg .b64 db;\n\t.reg .pred p, pe;\n\telect.sync _|pe, 0xffffffff;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
               "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
               "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "n"(ACC)
               : "memory");
}
// warp-level commit: one elected lane arrives on `bar` when all MMAs issued so far by this warp have completed
__device__ __forceinline__ void mma_commit_warp(uint64_t* bar)
{
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
               "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
               : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- split-bf16 helpers ------------------------------------------------------------------------------------------
// x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits, enough for the 1e-4 parity bar
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo)
{
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// pack two floats as bf16x2 (a in the low half)
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b)
{
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16lo_to_f(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16hi_to_f(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }

// ---- packed fp32 / bf16 arithmetic of the update epilogues -------------------------------------------------------
// packed convert: low half <- a, high half <- b (F2FP.BF16.F32.PACK_AB, full-rate; the C++ intrinsic compiled to two
// scalar F2F on the XU pipe)
__device__ __forceinline__ uint32_t cvt2(float a, float b)
{
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// packed fp32 pairs (FMUL2 / FADD2: one issue slot for two IEEE-rounded operations, results identical to the scalar forms)
__device__ __forceinline__ void mul2(float& o0, float& o1, float a0, float a1, float b0, float b1)
{
  uint64_t a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(c));
}
__device__ __forceinline__ void add2(float& o0, float& o1, float a0, float a1, float b0, float b1)
{
  uint64_t a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(c));
}
__device__ __forceinline__ void sub2(float& o0, float& o1, float a0, float a1, float b0, float b1)
{
  uint64_t a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(c));
}
__device__ __forceinline__ float rcp_fast(float x)
{
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// exact 3-way split of a pair: x = hi + mid + lo (each bf16), packed pairwise
__device__ __forceinline__ void split3(float x0, float x1, uint32_t& h, uint32_t& m, uint32_t& l)
{
  h = cvt2(x0, x1);
  x0 -= bf16lo_to_f(h); x1 -= bf16hi_to_f(h);
  m = cvt2(x0, x1);
  x0 -= bf16lo_to_f(m); x1 -= bf16hi_to_f(m);
  l = cvt2(x0, x1);
}
__device__ __forceinline__ float bf16_bits_to_f(unsigned short u) { return __uint_as_float((uint32_t) u << 16); }

} // namespace tc
} // namespace fb200
