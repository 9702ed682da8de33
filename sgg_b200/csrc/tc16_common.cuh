// Shared pieces of the 3xFP16 tcgen05 engine (tc16_gemm.cu) and the fused message-passing kernels (mp_fused.cu):
// tile constants, UMMA descriptors, the fp16 [hi | lo*2^11] operand split, the GRUCell pointwise math and the
// TMA tensor-map encoder.
//
// MMA pattern of a K16 step.  With x = hi + 2^-11 lo the three products are a_hi b_hi (MAIN) and a_lo b_hi + a_hi b_lo
// (CORR, carries the 2^11 scale).  Where a tile has NC <= 128 columns the kernels issue TWO instructions, not three:
// every stage stores the B_lo tile right behind the B_hi tile and every accumulator pair stores CORR right behind MAIN, so
// ONE N = 2 NC instruction computes [a_hi b_hi | a_hi b_lo] and a second, N = NC, adds a_lo b_hi into CORR ("fused pair").
// Same flops, one instruction and one A_hi fetch fewer per step; the issuing thread spends ~100 cycles per tcgen05.mma
// when operands change per instruction (tools/ubench/umma_rate.cu), so narrow tiles are bound by instruction count.
// The issuing loops run warp-uniformly and ONE ELECTED lane issues (elect.sync): under `if (lane == 0)` the compiler
// wraps each UTCHMMA / UTMALDG in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop to move its operands to uniform registers.
#pragma once
#include <cuda_fp16.h>
#include "tc_gemm.cuh"
#include "kernels.h"

namespace sgg {
namespace tc16 {

constexpr int BM = 128;
constexpr int BK = 64;                       // fp16 elements per k-block = one 128-byte swizzle span
constexpr int NTHR = 320;
constexpr int A_BYTES = BM * BK * 4;         // raw fp32 tile == hi tile + lo tile
constexpr int A_HALF = A_BYTES / 2;
constexpr int SMEM_BUDGET = 230000;
constexpr float LO_SCALE = 2048.0f, LO_INV = 1.0f / 2048.0f;


__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  // c_format = F32 (1) [4,6); a_format = b_format = F16 (0); K-major A and B; N>>3 [17,23); M>>4 [24,29)
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// K-major SWIZZLE_128B tile with 128-byte rows: SBO = 1024 B (8 rows), LBO unused, version 1, layout 2
__device__ __forceinline__ uint64_t make_sdesc128(const void *smem_tile) {
  const uint64_t addr = (uint64_t)((tc::smem_u32(smem_tile) & 0x3FFFF) >> 4);
  return addr | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32-byte global store (sm_100: STG.256), address 32-byte aligned: one full sector per thread and instruction
__device__ __forceinline__ void stg256(void *p, const uint4 &a, const uint4 &b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
               "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, const uint4 &v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// (x0, x1) -> packed fp16 hi pair and packed fp16 scaled-lo pair (x0 in the low half); cvt.rn.f16x2.f32 packs two
// conversions into one instruction
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((x0 - hf.x) * LO_SCALE, (x1 - hf.y) * LO_SCALE);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

// fp16 range guard: a packed pair of fp16 `hi` words has an all-ones exponent (inf / nan: |x| >= 65520 or a non-finite
// input) iff bit 15 / 31 of ((h & 0x7fff7fff) + 0x04000400) is set.  Kernels OR this into a register and raise the
// sticky device flag once per thread at most (sgg_tc16_overflow reads it).
__device__ __forceinline__ uint32_t f16x2_nonfinite(uint32_t h) { return ((h & 0x7fff7fffu) + 0x04000400u) & 0x80008000u; }

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// Fast, fp32-grade activations for the epilogue (it is instruction-bound: 3 transcendentals per hidden unit).
// ex2.approx / rcp.approx carry ~2^-22 relative error => |error| < 3e-7 on sigmoid / tanh values, far inside the
// 1e-4 parity bar; the accurate expf / tanhf versions cost ~25 instructions each and doubled the epilogue time.
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(-2.0f * fminf(fmaxf(x, -15.0f), 15.0f));      // clamp: tanh(+-15) == +-1 in fp32, no inf/inf
  return __fdividef(1.0f - e, 1.0f + e);
}
// torch.nn.GRUCell pointwise on 4 hidden units; gi_* / gh_* include the biases
struct Gru4 { float4 out, r, z, n; };
__device__ __forceinline__ Gru4 gru4(const float4 &gir, const float4 &ghr, const float4 &giz, const float4 &ghz,
                                     const float4 &gin, const float4 &ghn, const float4 &h) {
  Gru4 o;
  o.r.x = fast_sigmoid(gir.x + ghr.x); o.r.y = fast_sigmoid(gir.y + ghr.y);
  o.r.z = fast_sigmoid(gir.z + ghr.z); o.r.w = fast_sigmoid(gir.w + ghr.w);
  o.z.x = fast_sigmoid(giz.x + ghz.x); o.z.y = fast_sigmoid(giz.y + ghz.y);
  o.z.z = fast_sigmoid(giz.z + ghz.z); o.z.w = fast_sigmoid(giz.w + ghz.w);
  o.n.x = fast_tanh(gin.x + o.r.x * ghn.x); o.n.y = fast_tanh(gin.y + o.r.y * ghn.y);
  o.n.z = fast_tanh(gin.z + o.r.z * ghn.z); o.n.w = fast_tanh(gin.w + o.r.w * ghn.w);
  o.out.x = (1.0f - o.z.x) * o.n.x + o.z.x * h.x; o.out.y = (1.0f - o.z.y) * o.n.y + o.z.y * h.y;
  o.out.z = (1.0f - o.z.z) * o.n.z + o.z.z * h.z; o.out.w = (1.0f - o.z.w) * o.n.w + o.z.w * h.w;
  return o;
}
__device__ __forceinline__ float4 add4(const float4 &a, const float4 &b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// ---- host: TMA tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

// row-major [rows, K] of fp32 (elem 4) or fp16 (elem 2) -> boxes of box_rows x 128 bytes, SWIZZLE_128B, zero OOB fill
static inline int make_tmap(CUtensorMap *m, const void *base, int rows, int K, int box_rows, int elem) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t gstr[1] = {(cuuint64_t)K * (cuuint64_t)elem};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base,
                   gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d elem=%d", (int)r, rows, K, elem);
  return 0;
}

}  // namespace tc16
}  // namespace sgg
