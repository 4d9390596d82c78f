// Self-test of the tcgen05 / TMEM / TMA building blocks in exactly the four operand forms the NMF engine uses
// (fb200_selftest_tcgen05).  Each sub-test writes a small matrix that tests/test_tcgen05_blocks.py compares with numpy:
//   1  D[128 f x 64 b]  = H[128x16] W[16x64]          SS, A K-major (H blocks), B MN-major (W blocks)
//   2  D[128 f x 16 k]  = R[128x64] W[16x64]^T        TS, A = bf16 pairs in TMEM, B K-major (same W blocks)
//   3  D[128 b x 64 f]  = W2[16x128]^T H2[64x16]^T    SS, A MN-major (W blocks), B K-major (H blocks)
//   4  D[128 b x 16 k]  = R2[128x64] H2[64x16]        TS, B MN-major (same H blocks)
//   5  TMA 3-D tiled load with 128B swizzle, read back through swz128_off()
#include "common.cuh"
#include "tc_common.cuh"

#include <cuda.h>

namespace fb200 {
using namespace tc;

// operand blocks: W_op[(bb*KB + kb)*64 + i*8 + j] = W[8kb+i][8bb+j]; H_op[(fb*KB + kb)*64 + i*8 + j] = H[8fb+i][8kb+j]
__device__ __forceinline__ int wop_index(int k, int b, int KB) { return ((b >> 3) * KB + (k >> 3)) * 64 + (k & 7) * 8 + (b & 7); }
__device__ __forceinline__ int hop_index(int f, int k, int KB) { return ((f >> 3) * KB + (k >> 3)) * 64 + (f & 7) * 8 + (k & 7); }

struct SelfTestArgs {
  const float* H1; const float* W1; const float* R1; const float* W2; const float* H2; const float* R2;
  float* out1; float* out2; float* out3; float* out4; float* out5;
};

__global__ void __launch_bounds__(128) k_tc_selftest(SelfTestArgs a, const __grid_constant__ CUtensorMap tmap)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  float* vtile = reinterpret_cast<float*>(smem);                                   // 16 KB, 1024-aligned
  __nv_bfloat16* hop1 = reinterpret_cast<__nv_bfloat16*>(smem + 16384);            // 128x16
  __nv_bfloat16* wop1 = hop1 + 128 * 16;                                           // 16x64
  __nv_bfloat16* wop2 = wop1 + 16 * 64;                                            // 16x128
  __nv_bfloat16* hop2 = wop2 + 16 * 128;                                           // 64x16
  uint64_t* bar = reinterpret_cast<uint64_t*>(hop2 + 64 * 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  for (int e = tid; e < 128 * 16; e += 128) hop1[hop_index(e / 16, e % 16, 2)] = __float2bfloat16_rn(a.H1[e]);
  for (int e = tid; e < 16 * 64; e += 128) wop1[wop_index(e / 64, e % 64, 2)] = __float2bfloat16_rn(a.W1[e]);
  for (int e = tid; e < 16 * 128; e += 128) wop2[wop_index(e / 128, e % 128, 2)] = __float2bfloat16_rn(a.W2[e]);
  for (int e = tid; e < 64 * 16; e += 128) hop2[hop_index(e / 16, e % 16, 2)] = __float2bfloat16_rn(a.H2[e]);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const uint32_t lane_off = (uint32_t) (warp * 32) << 16;
  const uint32_t tD = tbase, tR = tbase + 64, tAcc = tbase + 128;
  uint32_t parity = 0;

  // ---- 1: SS, A = H blocks K-major (LBO 128 = comp blocks, SBO 256 = frame blocks), B = W blocks MN-major (LBO 128, SBO 256)
  if (tid == 0) {
    mma_ss(tD, make_smem_desc(smem_u32(hop1), 128, 256), make_smem_desc(smem_u32(wop1), 128, 256), make_idesc_bf16(128, 64, 0, 1), 0);
    mma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], parity); parity ^= 1;
  tc_fence_after();
  for (int c = 0; c < 64; c += 16) {
    uint32_t r[16];
    tmem_ld16(tD + lane_off + c, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; j++) a.out1[tid * 64 + c + j] = __uint_as_float(r[j]);
  }
  // ---- 2: TS, A = R1 row of this thread as bf16 pairs (column c holds elements 2c, 2c+1), B = W blocks K-major
  {
    uint32_t p[16];
    for (int h = 0; h < 2; h++) {
      for (int j = 0; j < 16; j++) p[j] = pack_bf16x2(a.R1[tid * 64 + h * 32 + 2 * j], a.R1[tid * 64 + h * 32 + 2 * j + 1]);
      tmem_st16(tR + lane_off + h * 16, p);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    for (int j = 0; j < 4; j++) // K-step j = bins 16j..16j+15 = bin blocks 2j, 2j+1
      mma_ts(tAcc, tR + 8 * j, make_smem_desc(smem_u32(wop1) + (2 * j) * 256, 256, 128), make_idesc_bf16(128, 16, 0, 0), j > 0);
    mma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], parity); parity ^= 1;
  tc_fence_after();
  {
    uint32_t r[16];
    tmem_ld16(tAcc + lane_off, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; j++) a.out2[tid * 16 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  // ---- 3: SS, A = W2 blocks MN-major (M = 128 bins: SBO 256 = bin blocks, LBO 128 = comp blocks), B = H2 blocks K-major
  if (tid == 0) {
    tc_fence_after();
    mma_ss(tD, make_smem_desc(smem_u32(wop2), 128, 256), make_smem_desc(smem_u32(hop2), 128, 256), make_idesc_bf16(128, 64, 1, 0), 0);
    mma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], parity); parity ^= 1;
  tc_fence_after();
  for (int c = 0; c < 64; c += 16) {
    uint32_t r[16];
    tmem_ld16(tD + lane_off + c, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; j++) a.out3[tid * 64 + c + j] = __uint_as_float(r[j]);
  }
  // ---- 4: TS, A = R2 row (lane = bin), B = H2 blocks MN-major (N = comps: SBO 128, K = frames: LBO 256)
  {
    uint32_t p[16];
    for (int h = 0; h < 2; h++) {
      for (int j = 0; j < 16; j++) p[j] = pack_bf16x2(a.R2[tid * 64 + h * 32 + 2 * j], a.R2[tid * 64 + h * 32 + 2 * j + 1]);
      tmem_st16(tR + lane_off + h * 16, p);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    for (int j = 0; j < 4; j++) // K-step j = frames 16j..16j+15 = frame blocks 2j, 2j+1
      mma_ts(tAcc, tR + 8 * j, make_smem_desc(smem_u32(hop2) + (2 * j) * 256, 256, 128), make_idesc_bf16(128, 16, 0, 1), j > 0);
    mma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], parity); parity ^= 1;
  tc_fence_after();
  {
    uint32_t r[16];
    tmem_ld16(tAcc + lane_off, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; j++) a.out4[tid * 16 + j] = __uint_as_float(r[j]);
  }
  // ---- 5: TMA box (32 bins x 128 frames) at bin 32 of buffer 0, 128B swizzle
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar[1], 128 * 32 * 4);
    tma_load_3d(vtile, &tmap, 32, 0, 0, &bar[1]);
  }
  mbar_wait(&bar[1], 0);
  for (int j = 0; j < 32; j++) a.out5[tid * 32 + j] = *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(vtile) + swz128_off(tid, j));

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

// 3-D tensor map over V[batch][Fp][Bp] (fp32): box = 32 bins x box_rows frames, 128B swizzle, zero OOB fill
int32_t make_v_tensor_map(Plan* p, void* tmap_out, const float* V, int64_t Bp, int64_t Fp, int64_t batch, int box_rows)
{
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      p->err = "cuTensorMapEncodeTiled not available";
      return FB200_ERR_CUDA;
    }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  cuuint64_t dims[3] = {(cuuint64_t) Bp, (cuuint64_t) Fp, (cuuint64_t) batch};
  cuuint64_t strides[2] = {(cuuint64_t) Bp * 4, (cuuint64_t) Bp * Fp * 4};
  cuuint32_t box[3] = {32, (cuuint32_t) box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(reinterpret_cast<CUtensorMap*>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(V), dims,
                      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    p->err = "cuTensorMapEncodeTiled failed: " + std::to_string((int) r);
    return FB200_ERR_CUDA;
  }
  return FB200_OK;
}

// inputs (device, float): H1[128x16] W1[16x64] R1[128x64] W2[16x128] H2[64x16] R2[128x64] V[128][68]
// outputs (device, float): out1[128x64] out2[128x16] out3[128x64] out4[128x16] out5[128x32]
int32_t run_tc_selftest(Plan* p, const float* in, float* out)
{
  SelfTestArgs a;
  a.H1 = in; a.W1 = a.H1 + 128 * 16; a.R1 = a.W1 + 16 * 64; a.W2 = a.R1 + 128 * 64; a.H2 = a.W2 + 16 * 128; a.R2 = a.H2 + 64 * 16;
  const float* V = a.R2 + 128 * 64;
  a.out1 = out; a.out2 = a.out1 + 128 * 64; a.out3 = a.out2 + 128 * 16; a.out4 = a.out3 + 128 * 64; a.out5 = a.out4 + 128 * 16;
  alignas(64) CUtensorMap tmap;
  FB_TRY(make_v_tensor_map(p, &tmap, V, 68, 128, 1, 128));
  size_t smem = 16384 + 2 * (128 * 16 + 16 * 64 + 16 * 128 + 64 * 16) + 64;
  FB_CUDA(p, cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  k_tc_selftest<<<1, 128, smem, p->stream>>>(a, tmap);
  p->launches++;
  FB_CUDA(p, cudaGetLastError());
  return FB200_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Micro-benchmark of tcgen05.mma issue/execute cost for the small shapes the engine uses: `reps` back-to-back MMAs of
// one form, timed with clock64 from first issue to commit completion.  out[v] = cycles per MMA (x1000) for variant v.
__global__ void __launch_bounds__(128) k_tc_mma_timing(long long* out, int reps)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u; // small bf16 values
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  uint32_t parity = 0;
  const uint32_t sa = smem_u32(smem);
  const uint32_t lo = (sa >> 4) | (8u << 16), hi = 48u | (1u << 14);
  const uint32_t lob = (sa >> 4) | (48u << 16), hib = 8u | (1u << 14);
  for (int v = 0; v < 10; v++) {
    __syncthreads();
    if (warp == 0) {
      long long t0 = clock64();
#define FB_LOOP(stmt) for (int i = 0; i < reps; i++) { stmt; }
      switch (v) {
      case 0: FB_LOOP(mma_ss_lohi<1>(tb, lo, hi, lo, hi, make_idesc_bf16(128, 64, 0, 1))) break;                 // SS N=64 same acc
      case 1: FB_LOOP(mma_ss_lohi<1>(tb + 64 * (i & 1), lo, hi, lo, hi, make_idesc_bf16(128, 64, 0, 1))) break;  // SS N=64 two accs
      case 2: FB_LOOP(mma_ss_lohi<1>(tb, lo, hi, lo, hi, make_idesc_bf16(128, 256, 0, 1))) break;                // SS N=256
      case 3: FB_LOOP(mma_ts_lohi<1>(tb + 256, tb + 128, lob, hib, make_idesc_bf16(128, 16, 0, 0))) break;       // TS N=16 same acc
      case 4: FB_LOOP(mma_ts_lohi<1>(tb + 256 + 16 * (i & 3), tb + 128, lob, hib, make_idesc_bf16(128, 16, 0, 0))) break; // TS N=16 4 accs
      case 5: FB_LOOP(mma_ts_lohi<1>(tb + 256, tb + 128, lob, hib, make_idesc_bf16(128, 32, 0, 0))) break;       // TS N=32
      case 6: FB_LOOP(mma_ts_lohi<1>(tb + 256, tb + 128, lob, hib, make_idesc_bf16(128, 128, 0, 0))) break;      // TS N=128
      case 7: FB_LOOP(mma_ss_lohi<1>(tb + 256, lo, hi, lob, hib, make_idesc_bf16(128, 16, 0, 0))) break;         // SS N=16 same acc
      case 8: FB_LOOP(mma_ts_lohi<1>(tb + 256, tb + 128 + 8 * (i & 3), lob, hib, make_idesc_bf16(128, 16, 0, 1))) break; // TS N=16 MN-major B
      case 9: FB_LOOP(mma_ss_lohi<1>(tb + 64 * (i & 1), lo, hi, lo, hi, make_idesc_bf16(128, 64, 1, 0))) break;  // SS N=64, A MN-major
      }
#undef FB_LOOP
      long long t1 = clock64();
      mma_commit_warp(&bar[0]);
      mbar_wait(&bar[0], parity);
      long long t2 = clock64();
      if (tid == 0) {
        out[2 * v] = (t1 - t0) * 1000 / reps;      // issue cost per MMA (x1000)
        out[2 * v + 1] = (t2 - t0) * 1000 / reps;  // issue + execute per MMA (x1000)
      }
    }
    parity ^= 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int32_t run_tc_mma_timing(Plan* p, long long* d_out, int reps)
{
  size_t smem = 65536 + 64;
  FB_CUDA(p, cudaFuncSetAttribute(k_tc_mma_timing, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  k_tc_mma_timing<<<1, 128, smem, p->stream>>>(d_out, reps);
  FB_CUDA(p, cudaGetLastError());
  return FB200_OK;
}

} // namespace fb200
