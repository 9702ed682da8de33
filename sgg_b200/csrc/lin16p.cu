// LINEAR on PRE-SPLIT operands: y = act(x w^T + b) where x arrives as fp16 [hi | lo * 2^11] planes written by the kernel
// that produced it (RoIAlign, a previous layer's epilogue) and w as the fp16 planes of sgg_tc_split_weights.
// rel_model_stanford.py:100-101 (roi_fmap: fc6 on the 25088-wide RoIAlign rows, then fc7) is the consumer: the fp32-input
// engine (tc16_gemm.cu) spends a third of its shared-memory traffic and eight warps of ALU work per k-block converting the
// A tile in place; here both operands go from TMA straight to the tensor core, as in the conv and P/Q kernels.
//   tile 128 rows x 128 columns, 64-wide k-blocks, 3 stages of [A_hi | A_lo | B_hi | B_lo] (64 KB);
//   two MMAs per K16 step: [main | a_hi b_lo] as one N = 256 instruction, then corr += a_lo b_hi (tc16_common.cuh);
//   K folded into TMEM 256 at a time, two ping-pong accumulator pairs drained into fp32 registers by warps 2..9;
//   epilogue: bias, ReLU, fp32 rows and (optionally) the fp16 planes of y for the next layer.
// Ragged M / Nout / K are handled by TMA's out-of-bounds zero fill and masked stores (Nout % 4 == 0, K % 8 == 0).
#include <stdlib.h>
#include "tc16_common.cuh"

namespace sgg {
namespace lin16p {
using namespace tc16;

constexpr int NC = 128;
constexpr int A_PLANE = BM * BK * 2, B_PLANE = NC * BK * 2;     // 16 KB each
constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;                // 64 KB
constexpr int STAGES = 3;
constexpr int RING = STAGES * STAGE;
constexpr int SMEM = RING + 1024 + 256;
constexpr int KCB = 256 / BK;
__device__ unsigned int g_lin16p_overflow = 0;   // sticky: an emitted output plane left the fp16 range (reported by sgg_tc16_overflow, bit 1)

struct LinParams {
  int M, K, Nout, relu;
  const float *bias;          // nullable
  const float *out_scale;     // nullable device scalar multiplied into x w^T before bias / ReLU (scaled backward GEMMs)
  float *out;                 // [M, Nout]
  __half *out_hi, *out_lo;    // nullable: planes of the output
};

__device__ __forceinline__ void mbar_wait_b(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(done)
        : "r"(tc::smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

__global__ void __launch_bounds__(NTHR, 1)
k_lin16p(const LinParams p, const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
         const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + RING);
  uint64_t *full = bars, *empty = bars + STAGES, *tmem_full = bars + 2 * STAGES, *tmem_empty = bars + 2 * STAGES + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * NC, m0 = blockIdx.y * BM;
  const int kblocks = (p.K + BK - 1) / BK;
  const int nchunks = (kblocks + KCB - 1) / KCB;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAh); prefetch_tmap(&tmAl); prefetch_tmap(&tmBh); prefetch_tmap(&tmBl);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1); mbar_init(tmem_full + 1, 1);
    mbar_init(tmem_empty, 8); mbar_init(tmem_empty + 1, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    for (int it = 0; it < kblocks; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait_b(empty + s, ph ^ 1);
      uint8_t *st = smem + (size_t)s * STAGE;
      const int k0 = it * BK;
      if (elect_one()) {
        mbar_arrive_expect_tx(full + s, STAGE);
        tma_load_2d(st, &tmAh, full + s, k0, m0);
        tma_load_2d(st + A_PLANE, &tmAl, full + s, k0, m0);
        tma_load_2d(st + 2 * A_PLANE, &tmBh, full + s, k0, n0);
        tma_load_2d(st + 2 * A_PLANE + B_PLANE, &tmBl, full + s, k0, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_f16(BM, NC), idesc2 = make_idesc_f16(BM, 2 * NC);
    for (int it = 0; it < kblocks; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait_b(full + s, ph);
      fence_after_sync();
      uint8_t *st = smem + (size_t)s * STAGE;
      const uint64_t ah = make_sdesc128(st), al = make_sdesc128(st + A_PLANE), bh = make_sdesc128(st + 2 * A_PLANE);
      const int chunk = it / KCB, kc = it - chunk * KCB;
      if (kc == 0) {                         // the buffer pair must have been drained (two chunks ago)
        mbar_wait_b(tmem_empty + (chunk & 1), ((chunk >> 1) & 1) ^ 1);
        fence_after_sync();
      }
      const uint32_t dm = tmem_base + (uint32_t)((chunk & 1) * 2 * NC), dc = dm + (uint32_t)NC;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);
          const uint32_t acc = (kc == 0 && kk == 0) ? 0u : 1u;
          mma_f16_ss(dm, ah + o, bh + o, idesc2, acc);       // [main | a_hi b_lo]  (B_lo follows B_hi in the stage)
          mma_f16_ss(dc, al + o, bh + o, idesc, 1u);         // corr += a_lo b_hi
        }
        mma_commit(empty + s);
        if (kc == KCB - 1 || it == kblocks - 1) mma_commit(tmem_full + (chunk & 1));
      }
      __syncwarp();
    }
  } else {
    // ===================== warps 2..9: thread <-> (row of the tile, half of its columns) =====================
    constexpr int HC = NC / 2;
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HC);
    float acc[HC];
#pragma unroll
    for (int c = 0; c < HC; ++c) acc[c] = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) {
      const int b = ch & 1;
      mbar_wait_b(tmem_full + b, (ch >> 1) & 1);
      fence_after_sync();
      __syncwarp();
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 16) {
        float v[16], w[16];
        tmem_ld16(taddr + (uint32_t)(b * 2 * NC + c0), v);
        tmem_ld16(taddr + (uint32_t)(b * 2 * NC + NC + c0), w);
        tmem_wait_ld();
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) acc[c0 + cc] += fmaf(w[cc], LO_INV, v[cc]);
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + b);
    }
    const int m = m0 + row, cbase = n0 + half * HC;
    if (m < p.M) {
      uint32_t ovf = 0;
      const float osc = p.out_scale != nullptr ? __ldg(p.out_scale) : 1.0f;
      float *orow = p.out + (size_t)m * p.Nout;
#pragma unroll
      for (int c = 0; c < HC; c += 4) {
        const int col = cbase + c;
        if (col < p.Nout) {                  // Nout % 4 == 0: a group of four is inside or outside
          float4 v = make_float4(acc[c] * osc, acc[c + 1] * osc, acc[c + 2] * osc, acc[c + 3] * osc);
          if (p.bias != nullptr) {
            const float4 bv = ldg4(p.bias + col);
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          }
          if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          *reinterpret_cast<float4 *>(orow + col) = v;
          if (p.out_hi != nullptr) {
            uint2 hi, lo;
            split2(v.x, v.y, hi.x, lo.x); split2(v.z, v.w, hi.y, lo.y);
            ovf |= f16x2_nonfinite(hi.x) | f16x2_nonfinite(hi.y);
            *reinterpret_cast<uint2 *>(p.out_hi + (size_t)m * p.Nout + col) = hi;
            *reinterpret_cast<uint2 *>(p.out_lo + (size_t)m * p.Nout + col) = lo;
          }
        }
      }
      if (ovf) atomicOr(&g_lin16p_overflow, 2u);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

int overflow_flag(int reset, unsigned int *out) {
  unsigned int v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&v, g_lin16p_overflow, sizeof(v)) != cudaSuccess) return -1;
  if (reset) {
    const unsigned int z = 0;
    if (cudaMemcpyToSymbol(g_lin16p_overflow, &z, sizeof(z)) != cudaSuccess) return -1;
  }
  *out = v;
  return 0;
}

}  // namespace lin16p
}  // namespace sgg

/* y [M,Nout] = act(x w^T + b) on tcgen05 with BOTH operands pre-split: x_planes = fp16 [hi | lo * 2^11] planes of x [M,K]
 * (2 * M * K halves: sgg_node_edge_features_planes, or y_planes of a previous call), w_split16 = planes of w [Nout,K]
 * (sgg_tc_split_weights, 3xFP16 layout).  y_planes nullable: also emit the planes of y (2 * M * Nout halves).
 * out_scale nullable (device scalar multiplied into x w^T before bias / ReLU).
 * K % 8 == 0, Nout % 4 == 0; fp16 range of x is the producer's responsibility (sticky flag: sgg_tc16_overflow). */
extern "C" int sgg_tc16_linear_pre(const void *x_planes, const void *w_split16, const float *bias, float *y, void *y_planes,
                                   int M, int Nout, int K, int relu, const float *out_scale, void *stream) {
  using namespace sgg::lin16p;
  if (M <= 0 || Nout <= 0) return 0;
  if (!x_planes || !w_split16 || !y) return sgg_set_err(SGG_E_BADARG, "tc16_linear_pre: null pointer");
  if (K <= 0 || (K & 7) || (Nout & 3)) return sgg_set_err(SGG_E_BADARG, "tc16_linear_pre: K %% 8 == 0 and Nout %% 4 == 0 required");
  if ((reinterpret_cast<uintptr_t>(x_planes) & 15) || (reinterpret_cast<uintptr_t>(w_split16) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return sgg_set_err(SGG_E_BADARG, "tc16_linear_pre: 16-byte alignment required");
  static bool attr = false;
  if (!attr) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_lin16p, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr = true;
  }
  const __half *xh = (const __half *)x_planes, *wh = (const __half *)w_split16;
  CUtensorMap tm[4];
  int rc;
  if ((rc = sgg::tc16::make_tmap(&tm[0], xh, M, K, sgg::tc16::BM, 2))) return rc;
  if ((rc = sgg::tc16::make_tmap(&tm[1], xh + (size_t)M * K, M, K, sgg::tc16::BM, 2))) return rc;
  if ((rc = sgg::tc16::make_tmap(&tm[2], wh, Nout, K, NC, 2))) return rc;
  if ((rc = sgg::tc16::make_tmap(&tm[3], wh + (size_t)Nout * K, Nout, K, NC, 2))) return rc;
  LinParams p{};
  p.M = M; p.K = K; p.Nout = Nout; p.relu = relu; p.bias = bias; p.out = y; p.out_scale = out_scale;
  p.out_hi = (__half *)y_planes; p.out_lo = y_planes ? (__half *)y_planes + (size_t)M * Nout : nullptr;
  // blockIdx.x = column tile: CTAs resident together share a row block's A tiles and walk k in lockstep
  dim3 grid((Nout + NC - 1) / NC, (M + sgg::tc16::BM - 1) / sgg::tc16::BM);
  if (grid.y > 65535) return sgg_set_err(SGG_E_BADARG, "tc16_linear_pre: too many rows");
  k_lin16p<<<grid, sgg::tc16::NTHR, SMEM, (cudaStream_t)stream>>>(p, tm[0], tm[1], tm[2], tm[3]);
  SGG_RETURN_IF_LAUNCH_FAILED("k_lin16p");
  return 0;
}
