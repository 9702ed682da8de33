// VGG16 conv stack of the frozen detector backbone (rel_model_base.py:184, :310-321: torchvision vgg16 `features`
// without the last max-pool) as tcgen05 implicit GEMMs — SURVEY.md §8f rank 2.  The reference runs it through cuDNN in
// fp32 (226 GFLOP per 608 x 608 image: 89 % of its forward); here every 3x3 / pad 1 / stride 1 convolution is
//
//   out[pixel, cout] = sum_{tap, cin} in[pixel + tap, cin] * w[cout, tap, cin]
//
// with activations kept as NHWC fp16 planes [hi | lo * 2^11] (tc16_common.cuh: x = hi + 2^-11 lo, fp32-grade results), so
// the im2col operand never exists and the hardware's out-of-bounds zero fill of the TMA boxes is the padding.
// K = 9 * Cin is folded into TMEM 256 at a time (two ping-pong accumulator pairs, the tensor core's fp32 accumulate
// truncates: DESIGN.md section 4); eight warps drain the chunks into registers.  Epilogue: bias + ReLU, optional fused
// 2x2 max-pool, output again as fp16 planes (NHWC) or, for the last layer, fp32 NCHW (the layout RelModel.fmap has in the
// reference).  Two kernels, chosen per layer:
//   k_conv3x3     (v1): one 8 x 16 pixel tile per CTA; per tap and 64-channel block the A tile is ONE 4-D TMA box
//                       {64 c, 16 w, 8 h, 1 n} at (w0 + dx - 1, h0 + dy - 1).  Used from 512 input channels.
//   k_conv3x3_v2  (v2): persistent CTAs over 16 x 8 tiles; three column-shifted 18 x 8 slabs per channel block serve all
//                       nine taps through shifted shared-memory descriptors.  Used up to 256 input channels.
// Both have a cta_group::2 variant (CTA pairs, M = 256).  The first layer (Cin = 3, K = 27) is a SIMT kernel that also
// converts NCHW fp32 images to planes.
#include <stdlib.h>
#include <stdio.h>
#include "tc16_common.cuh"

namespace sgg {
namespace conv {
using namespace tc16;

constexpr int TH = 8, TW = 16;               // output tile: 8 rows x 16 columns = 128 pixels = UMMA M
__device__ unsigned int g_conv_overflow = 0; // sticky: an emitted activation left the fp16 range (sgg_conv_overflow)

struct ConvParams {
  int B, H, W, Cin, Cout;
  int tiles_w, tiles_h;
  int relu, pool;
  int kcb;                                   // v2: k-blocks folded into TMEM before a drain
  int dbg;                                   // v2: record g_conv_dbg
  int dry;                                   // v2 experiment (SGG_CONV_DRY): 1 = MMAs without operand loads, 2 = loads without MMAs,
                                             // 3 = weight loads only, 4 = activation loads only (both without MMAs)
  const float *bias;
  __half *out_hi, *out_lo;                   // NHWC planes [B, Ho, Wo, Cout] (nullable when out_f32 is set)
  float *out_f32;                            // nullable: fp32 NCHW [B, Cout, Ho, Wo]
};

__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- cta_group::2 (one 256-row UMMA tile per CTA pair; each CTA stages its own 128 pixel rows and HALF of the weight slab) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// pair loads: the destination is this CTA's shared memory, the transaction bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_4d_pair(void *smem_dst, const CUtensorMap *m, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread complete -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc::smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t ncols) {   // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

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

template <int NC, int CG>
struct CCfg {
  static constexpr int A_PLANE = BM * BK * 2;              // 16 KB
  static constexpr int B_PLANE = (NC / CG) * BK * 2;       // CG = 2: this CTA's half of the weight slab
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;  // NC = 128: 64 KB (pair: 48 KB); NC = 64: 48 KB (pair: 40 KB)
  static constexpr int STAGES = STAGE > 48 * 1024 ? 3 : 4;
  static constexpr int RING = STAGES * STAGE;
  static constexpr int SMEM = RING + 1024 + 256;
  static constexpr int TS = NC + 4;                        // fp32 staging tile row stride (pool / NCHW epilogues)
  static_assert(BM * TS * 4 <= RING, "staging tile must fit in the ring");
};

// CG = 1: one CTA per 128-pixel tile.  CG = 2: launched as clusters of two CTAs (adjacent tiles of the same channel
// block); the even CTA issues cta_group::2 MMAs (M = 256) that read both CTAs' pixel rows and the two halves of the
// weight slab, so each SM stages and re-reads only NC / 2 weight rows per k-block.
template <int NC, int CG>
__global__ void __launch_bounds__(NTHR, 1)
k_conv3x3(const ConvParams p, const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
          const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl) {
  using namespace tc;
  using C = CCfg<NC, CG>;
  constexpr int STAGES = C::STAGES, STAGE = C::STAGE, KCB = 256 / BK;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  constexpr uint32_t TMEM_COLS = 4 * NC <= 256 ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::RING);
  uint64_t *full = bars, *empty = bars + STAGES, *tmem_full = bars + 2 * STAGES, *tmem_empty = bars + 2 * STAGES + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int tw = t % p.tiles_w; t /= p.tiles_w;
  const int th = t % p.tiles_h;
  const int n = t / p.tiles_h;
  const int h0 = th * TH, w0 = tw * TW;
  const int cout0 = blockIdx.y * NC;
  const int cblocks = p.Cin / BK;
  const int kblocks = 9 * cblocks;
  const int nchunks = (kblocks + KCB - 1) / KCB;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAh); prefetch_tmap(&tmAl); prefetch_tmap(&tmBh); prefetch_tmap(&tmBl);
    // pair: `full` and `tmem_empty` are used on the leader only (one arrival per producer / per draining thread of both CTAs)
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, CG); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1); mbar_init(tmem_full + 1, 1);
    mbar_init(tmem_empty, CG == 2 ? 16 : 256); mbar_init(tmem_empty + 1, CG == 2 ? 16 : 256);   // pair: one arrival per warp
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_pair(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  fence_before_sync();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: k-block it = (channel block cb, tap) =====================
    {                                          // whole warp runs the loop, one elected lane issues
      for (int it = 0; it < kblocks; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait_b(empty + s, ph ^ 1);
        const int cb = it / 9, tap = it - cb * 9;
        const int dy = tap / 3, dx = tap - dy * 3;
        uint8_t *st = smem + (size_t)s * STAGE;
        if (elect_one()) {
        if constexpr (CG == 2) {
          const uint32_t lbar = mapa_u32(smem_u32(full + s), 0);
          if (rank == 0) mbar_arrive_expect_tx(full + s, 2 * STAGE); else mbar_arrive_cluster(lbar);
          const int brow = cout0 + (int)rank * (NC / 2);
          tma_load_4d_pair(st, &tmAh, lbar, cb * BK, w0 + dx - 1, h0 + dy - 1, n);
          tma_load_4d_pair(st + C::A_PLANE, &tmAl, lbar, cb * BK, w0 + dx - 1, h0 + dy - 1, n);
          tma_load_2d_pair(st + 2 * C::A_PLANE, &tmBh, lbar, tap * p.Cin + cb * BK, brow);
          tma_load_2d_pair(st + 2 * C::A_PLANE + C::B_PLANE, &tmBl, lbar, tap * p.Cin + cb * BK, brow);
        } else {
          mbar_arrive_expect_tx(full + s, STAGE);
          tma_load_4d(st, &tmAh, full + s, cb * BK, w0 + dx - 1, h0 + dy - 1, n);
          tma_load_4d(st + C::A_PLANE, &tmAl, full + s, cb * BK, w0 + dx - 1, h0 + dy - 1, n);
          tma_load_2d(st + 2 * C::A_PLANE, &tmBh, full + s, tap * p.Cin + cb * BK, cout0);
          tma_load_2d(st + 2 * C::A_PLANE + C::B_PLANE, &tmBl, full + s, tap * p.Cin + cb * BK, cout0);
        }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (rank == 0) {                          // whole warp runs the loop, one elected lane issues (see k_conv3x3_v2)
      constexpr uint32_t idesc = make_idesc_f16(BM * CG, NC);
      constexpr uint32_t idesc2 = make_idesc_f16(BM, 2 * NC);   // CG = 1 only
      (void)idesc2;
      for (int it = 0; it < kblocks; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait_b(full + s, ph);
        fence_after_sync();
        uint8_t *st = smem + (size_t)s * STAGE;
        const uint64_t ah = make_sdesc128(st), al = make_sdesc128(st + C::A_PLANE);
        const uint64_t bh = make_sdesc128(st + 2 * C::A_PLANE), bl = make_sdesc128(st + 2 * C::A_PLANE + C::B_PLANE);
        const int chunk = it / KCB, kc = it - chunk * KCB;
        if (kc == 0) {                       // the buffer pair must have been drained (chunk - 2)
          mbar_wait_b(tmem_empty + (chunk & 1), ((chunk >> 1) & 1) ^ 1);
          fence_after_sync();
        }
        const uint32_t dm = tmem_base + (uint32_t)((chunk & 1) * 2 * NC), dc = dm + (uint32_t)NC;
        if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);
          const uint32_t acc = (kc == 0 && kk == 0) ? 0u : 1u;
          if constexpr (CG == 2) {
            mma_f16_ss_pair(dc, al + o, bh + o, idesc, acc);
            mma_f16_ss_pair(dc, ah + o, bl + o, idesc, 1u);
            mma_f16_ss_pair(dm, ah + o, bh + o, idesc, acc);
          } else {
            // two instructions instead of three (tc16_common.cuh, "fused pair"): the B_lo tile follows the B_hi
            // tile in shared memory, so ONE N = 2 NC MMA forms [main | a_hi b_lo] in the adjacent column ranges
            mma_f16_ss(dm, ah + o, bh + o, idesc2, acc);
            mma_f16_ss(dc, al + o, bh + o, idesc, 1u);
          }
        }
        if constexpr (CG == 2) {
          mma_commit_pair(empty + s);
          if (kc == KCB - 1 || it == kblocks - 1) mma_commit_pair(tmem_full + (chunk & 1));
        } else {
          mma_commit(empty + s);
          if (kc == KCB - 1 || it == kblocks - 1) mma_commit(tmem_full + (chunk & 1));
        }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== warps 2..9: thread <-> (pixel row of the tile, half of the tile's channels) =====================
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
      if constexpr (CG == 2) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty + b), 0));
      } else {
        mbar_arrive(tmem_empty + b);
      }
    }
    // bias + ReLU
    const int cbase = cout0 + half * HC;
#pragma unroll
    for (int c = 0; c < HC; c += 4) {
      const float4 bv = p.bias != nullptr ? ldg4(p.bias + cbase + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[c] += bv.x; acc[c + 1] += bv.y; acc[c + 2] += bv.z; acc[c + 3] += bv.w;
      if (p.relu) {
        acc[c] = fmaxf(acc[c], 0.f); acc[c + 1] = fmaxf(acc[c + 1], 0.f);
        acc[c + 2] = fmaxf(acc[c + 2], 0.f); acc[c + 3] = fmaxf(acc[c + 3], 0.f);
      }
    }
    const int ty = row / TW, tx = row - ty * TW;
    const int y = h0 + ty, x = w0 + tx;
    if (!p.pool && p.out_f32 == nullptr) {
      // planes, NHWC: this thread's pixel, HC consecutive channels (HC * 2 bytes contiguous per plane)
      if (n < p.B && y < p.H && x < p.W) {
        const size_t off = (((size_t)n * p.H + y) * p.W + x) * p.Cout + cbase;
        uint32_t ovf = 0;
#pragma unroll
        for (int c = 0; c < HC; c += 8) {
          uint4 hi, lo;
          split2(acc[c], acc[c + 1], hi.x, lo.x); split2(acc[c + 2], acc[c + 3], hi.y, lo.y);
          split2(acc[c + 4], acc[c + 5], hi.z, lo.z); split2(acc[c + 6], acc[c + 7], hi.w, lo.w);
          ovf |= f16x2_nonfinite(hi.x) | f16x2_nonfinite(hi.y) | f16x2_nonfinite(hi.z) | f16x2_nonfinite(hi.w);
          *reinterpret_cast<uint4 *>(p.out_hi + off + c) = hi;
          *reinterpret_cast<uint4 *>(p.out_lo + off + c) = lo;
        }
        if (ovf) atomicOr(&g_conv_overflow, 1u);
      }
    } else {
      // stage the tile in shared memory (the ring is idle: every MMA has retired), then pool and / or transpose
      constexpr int TS = C::TS;
      float *tile = reinterpret_cast<float *>(smem);
#pragma unroll
      for (int c = 0; c < HC; c += 4)
        *reinterpret_cast<float4 *>(tile + (size_t)row * TS + half * HC + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
      named_bar_sync(1, 256);
      const int tid = (int)threadIdx.x - 64;             // 0..255
      const int Ho = p.pool ? p.H / 2 : p.H, Wo = p.pool ? p.W / 2 : p.W;
      const int npix = p.pool ? (TH / 2) * (TW / 2) : BM;      // output pixels of this tile
      const int pw = p.pool ? TW / 2 : TW;
      if (p.out_f32 != nullptr) {
        // fp32 NCHW: thread <-> (channel, pixel run): consecutive threads take consecutive pixels of one channel
        for (int i = tid; i < NC * npix; i += 256) {
          const int c = i / npix, pp = i - c * npix;
          const int oy = pp / pw, ox = pp - oy * pw;
          float v;
          if (p.pool) {
            const int r0 = (2 * oy) * TW + 2 * ox;
            v = fmaxf(fmaxf(tile[(size_t)r0 * TS + c], tile[(size_t)(r0 + 1) * TS + c]),
                      fmaxf(tile[(size_t)(r0 + TW) * TS + c], tile[(size_t)(r0 + TW + 1) * TS + c]));
          } else {
            v = tile[(size_t)pp * TS + c];
          }
          const int gy = (p.pool ? h0 / 2 : h0) + oy, gx = (p.pool ? w0 / 2 : w0) + ox;
          if (n < p.B && gy < Ho && gx < Wo) p.out_f32[(((size_t)n * p.Cout + cout0 + c) * Ho + gy) * Wo + gx] = v;
        }
      } else {
        // pooled planes, NHWC: thread <-> (output pixel, 8 consecutive channels)
        constexpr int CG = NC / 8;
        uint32_t ovf = 0;
        for (int i = tid; i < npix * CG; i += 256) {
          const int pp = i / CG, c = (i - pp * CG) * 8;
          const int oy = pp / pw, ox = pp - oy * pw;
          const int r0 = (2 * oy) * TW + 2 * ox;
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; k += 4) {
            const float4 a = *reinterpret_cast<const float4 *>(tile + (size_t)r0 * TS + c + k);
            const float4 b = *reinterpret_cast<const float4 *>(tile + (size_t)(r0 + 1) * TS + c + k);
            const float4 d = *reinterpret_cast<const float4 *>(tile + (size_t)(r0 + TW) * TS + c + k);
            const float4 e = *reinterpret_cast<const float4 *>(tile + (size_t)(r0 + TW + 1) * TS + c + k);
            v[k] = fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)); v[k + 1] = fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y));
            v[k + 2] = fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)); v[k + 3] = fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w));
          }
          const int gy = h0 / 2 + oy, gx = w0 / 2 + ox;
          if (n < p.B && gy < Ho && gx < Wo) {
            uint4 hi, lo;
            split2(v[0], v[1], hi.x, lo.x); split2(v[2], v[3], hi.y, lo.y);
            split2(v[4], v[5], hi.z, lo.z); split2(v[6], v[7], hi.w, lo.w);
            ovf |= f16x2_nonfinite(hi.x) | f16x2_nonfinite(hi.y) | f16x2_nonfinite(hi.z) | f16x2_nonfinite(hi.w);
            const size_t off = (((size_t)n * Ho + gy) * Wo + gx) * p.Cout + cout0 + c;
            *reinterpret_cast<uint4 *>(p.out_hi + off) = hi;
            *reinterpret_cast<uint4 *>(p.out_lo + off) = lo;
          }
        }
        if (ovf) atomicOr(&g_conv_overflow, 1u);
      }
    }
  }
  tc::fence_before_sync();
  if constexpr (CG == 2) {
    cluster_sync_all();                      // no remote arrival may still be in flight towards a CTA that has exited
    if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// v2: column-shifted slabs + persistent CTAs.
// The v1 kernel above fetches the A operand of every tap separately (nine shifted 128-pixel boxes per 64-channel block) and
// runs one tile per CTA; ncu (profiles/r02_ncu_conv.md) shows its deep layers at the L2 -> SM throughput cap (11.9 TB/s,
// 6300 B/clk) and the Cin = 64 layers bound by per-tile prologue / epilogue (tensor pipe 18 %).
// Here the output tile is 16 rows x 8 columns (GEMM row m = ty * 8 + tx, so an 8-row UMMA group is one tile line) and the
// A operand of a 64-channel block is THREE slabs, one per horizontal tap offset dx: the 18 x 8 pixel box at column
// w0 + dx - 1.  The three vertical taps of a slab are the same bytes read through a shared-memory descriptor that starts
// dy lines (dy * 1024 B) into the slab — the 128-byte swizzle is a function of the shared-memory ADDRESS, so a descriptor
// may start at any row of a TMA-written box (tools/ubench/umma_shift.cu).  A bytes per tile and channel block: 288 KB ->
// 108 KB, every 8-row group still a whole 1024-byte swizzle atom.  (A single 18 x 10 halo read at row offsets dy * 10 + dx
// is 46 KB and also exact, but its 8-row groups straddle two atoms and the tensor core then needs ~170 instead of ~102
// cycles to fetch A: measured 340 cycles per K16 step instead of 258.)
// CTAs are persistent: each loops over (channel tile, spatial tile) work items, the TMEM ping-pong and both shared-memory
// rings run across item boundaries, and the drain warps' epilogue (bias, ReLU, 2x2 max-pool by warp shuffles, stores)
// overlaps the next item's main loop.  k-block order inside a channel block: dx-major (tap = dy * 3 + dx).
constexpr int VH = 16, VW = 8;               // output tile (rows x columns) = 128 pixels
constexpr int SLAB_PLANE = (VH + 2) * VW * 128;   // 18432 B: 18 lines x 8 pixels x 64 channels (fp16), 1024-aligned

// SGG_CONV_DBG=1: clock64 stamps of CTA 0 for its first 16 items: [item][0] MMA issue start, [1] MMA issue end,
// [2] last chunk drained, [3] epilogue done
__device__ long long g_conv_dbg[16 * 8];
__device__ long long g_conv_cta[256 * 4];     // per CTA: items, cycles, start ns, end ns (SGG_CONV_DBG=1)

// K-major SWIZZLE_128B descriptor (SBO = 1024 B) from a shared-memory address
__device__ __forceinline__ uint64_t make_sdesc128_addr(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int NC, int CG>
struct VCfg {
  static constexpr int A_STAGE = 2 * SLAB_PLANE;            // hi + lo slab of one (channel block, dx)
  static constexpr int A_STAGES = 3;
  static constexpr int B_PLANE = (NC / CG) * BK * 2;        // one tap, one plane (this CTA's share of the channel tile)
  static constexpr int B_STAGE = 2 * B_PLANE;
  static constexpr int B_FIT = (225 * 1024 - A_STAGES * A_STAGE) / B_STAGE;
  static constexpr int B_STAGES = B_FIT > 8 ? 8 : B_FIT;
  static constexpr int RING = A_STAGES * A_STAGE + B_STAGES * B_STAGE;
  static constexpr int SMEM = RING + 1024 + 512;
  static_assert(B_STAGES >= 3, "weight-tap ring too short");
};

template <int NC, int CG>
__global__ void __launch_bounds__(NTHR, 1)
k_conv3x3_v2(const ConvParams p, const int n_items, const int n_sp, const __grid_constant__ CUtensorMap tmAh,
             const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl) {
  using namespace tc;
  using C = VCfg<NC, CG>;
  constexpr int AS = C::A_STAGES, BS = C::B_STAGES;
  const int KCB = p.kcb;                      // k-blocks per accumulation chunk (256 / BK unless overridden: SGG_CONV_KCB)
  constexpr uint32_t TMEM_COLS = 4 * NC <= 256 ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *ring_a = smem, *ring_b = smem + AS * C::A_STAGE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::RING);
  uint64_t *a_full = bars, *a_empty = bars + AS, *b_full = bars + 2 * AS, *b_empty = bars + 2 * AS + BS;
  uint64_t *tmem_full = bars + 2 * AS + 2 * BS, *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int unit = (int)blockIdx.x / CG, n_units = (int)gridDim.x / CG;   // a unit = one CTA or one CTA pair
  const int cblocks = p.Cin / BK;
  const int kblocks = 9 * cblocks;
  const int nchunks = (kblocks + KCB - 1) / KCB;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAh); prefetch_tmap(&tmAl); prefetch_tmap(&tmBh); prefetch_tmap(&tmBl);
    // pair: the `full` and `tmem_empty` barriers live on the leader (one arrival per producer / per draining warp of both CTAs)
    for (int s = 0; s < AS; ++s) { mbar_init(a_full + s, CG); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < BS; ++s) { mbar_init(b_full + s, CG); mbar_init(b_empty + s, 1); }
    mbar_init(tmem_full, 1); mbar_init(tmem_full + 1, 1);
    mbar_init(tmem_empty, 8 * CG); mbar_init(tmem_empty + 1, 8 * CG);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_pair(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  fence_before_sync();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  long long dbg_c0 = 0, dbg_n0 = 0;
  if (p.dbg && threadIdx.x == 64) { dbg_c0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_n0)); }

  // work item -> (channel tile, image, tile row, tile column) of THIS CTA; spatial index n_sp.. is padding (n = B: loads are
  // zero-filled by TMA, stores are masked)
  auto decode = [&](int item, int &cout0, int &n, int &h0, int &w0) {
    const int ct = item / n_sp;
    int t = (item - ct * n_sp) * CG + (int)rank;
    cout0 = ct * NC;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h;
    n = t / p.tiles_h; h0 = th * VH; w0 = tw * VW;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    // Slab sequence k = 0, 1, ... over (item, channel block, dx); slab k + 2 is issued after the three weight taps of slab
    // k, so an activation slab always has two slab times (six taps) to arrive.
    if (p.dry != 1) {                          // whole warp runs the loops, one elected lane issues
      uint32_t ga = 0, gb = 0;                 // slabs / taps issued so far (ring position and phase)
      int s_item = unit, s_cb = 0, s_dx = 0;   // cursor of the next slab to issue
      auto issue_slab = [&]() {
        if (s_item >= n_items) return;
        int cout0, n, h0, w0;
        decode(s_item, cout0, n, h0, w0);
        const uint32_t s = ga % AS, ph = (ga / AS) & 1; ++ga;
        mbar_wait_b(a_empty + s, ph ^ 1);
        uint8_t *st = ring_a + (size_t)s * C::A_STAGE;
        if (elect_one()) {
        if constexpr (CG == 2) {
          const uint32_t lbar = mapa_u32(smem_u32(a_full + s), 0);
          if (rank == 0) mbar_arrive_expect_tx(a_full + s, 2 * C::A_STAGE); else mbar_arrive_cluster(lbar);
          tma_load_4d_pair(st, &tmAh, lbar, s_cb * BK, w0 + s_dx - 1, h0 - 1, n);
          tma_load_4d_pair(st + SLAB_PLANE, &tmAl, lbar, s_cb * BK, w0 + s_dx - 1, h0 - 1, n);
        } else if (p.dry == 3) {                 // experiment: no activation loads
          mbar_arrive(a_full + s);
        } else {
          mbar_arrive_expect_tx(a_full + s, C::A_STAGE);
          tma_load_4d(st, &tmAh, a_full + s, s_cb * BK, w0 + s_dx - 1, h0 - 1, n);
          tma_load_4d(st + SLAB_PLANE, &tmAl, a_full + s, s_cb * BK, w0 + s_dx - 1, h0 - 1, n);
        }
        }
        __syncwarp();
        if (++s_dx == 3) { s_dx = 0; if (++s_cb == cblocks) { s_cb = 0; s_item += n_units; } }
      };
      issue_slab(); issue_slab();
      for (int item = unit; item < n_items; item += n_units) {
        int cout0, n, h0, w0;
        decode(item, cout0, n, h0, w0);
        const int brow = cout0 + (int)rank * (NC / CG);
        for (int cb = 0; cb < cblocks; ++cb)
          for (int dx = 0; dx < 3; ++dx) {
            for (int dy = 0; dy < 3; ++dy) {
              const int kcol = (dy * 3 + dx) * p.Cin + cb * BK;
              const uint32_t s = gb % BS, ph = (gb / BS) & 1; ++gb;
              mbar_wait_b(b_empty + s, ph ^ 1);
              uint8_t *st = ring_b + (size_t)s * C::B_STAGE;
              if (elect_one()) {
              if constexpr (CG == 2) {
                const uint32_t lbar = mapa_u32(smem_u32(b_full + s), 0);
                if (rank == 0) mbar_arrive_expect_tx(b_full + s, 2 * C::B_STAGE); else mbar_arrive_cluster(lbar);
                tma_load_2d_pair(st, &tmBh, lbar, kcol, brow);
                tma_load_2d_pair(st + C::B_PLANE, &tmBl, lbar, kcol, brow);
              } else if (p.dry == 4) {           // experiment: no weight loads
                mbar_arrive(b_full + s);
              } else {
                mbar_arrive_expect_tx(b_full + s, C::B_STAGE);
                tma_load_2d(st, &tmBh, b_full + s, kcol, brow);
                tma_load_2d(st + C::B_PLANE, &tmBl, b_full + s, kcol, brow);
              }
              }
              __syncwarp();
            }
            issue_slab();
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair: the leader only) =====================
    // The whole warp runs the loop (warp-uniform control flow and operands); one elected lane issues.  With the loop under
    // `if (lane == 0)` the compiler cannot prove the descriptors uniform and wraps every UTCHMMA in an
    // ELECT / R2UR.BROADCAST / BRA.U.ANY loop, which made the issue thread — not the tensor pipe — the bound.
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM * CG, NC);
      constexpr uint32_t idesc2 = make_idesc_f16(BM, 2 * NC);   // CG = 1 only
      (void)idesc2;
      uint32_t ga = 0, gb = 0, gch = 0;
      int ord = 0;
      for (int item = unit; item < n_items; item += n_units, ++ord) {
        if (p.dbg && blockIdx.x == 0 && ord < 16 && lane == 0) g_conv_dbg[ord * 8 + 0] = clock64();
        int it = 0;
        for (int cb = 0; cb < cblocks; ++cb)
          for (int dx = 0; dx < 3; ++dx) {
            const uint32_t as = ga % AS, aph = (ga / AS) & 1; ++ga;
            if (p.dry != 1) mbar_wait_b(a_full + as, aph);
            const uint32_t a_hi = smem_u32(ring_a + (size_t)as * C::A_STAGE), a_lo = a_hi + SLAB_PLANE;
            for (int dy = 0; dy < 3; ++dy, ++it) {
              const uint32_t bs = gb % BS, bph = (gb / BS) & 1; ++gb;
              if (p.dry != 1) mbar_wait_b(b_full + bs, bph);
              fence_after_sync();
              const int chunk = it / KCB, kc = it - chunk * KCB;
              const uint32_t gc = gch + (uint32_t)chunk;
              if (kc == 0) {                     // the buffer pair must have been drained (two chunks ago)
                mbar_wait_b(tmem_empty + (gc & 1), ((gc >> 1) & 1) ^ 1);
                fence_after_sync();
              }
              const uint32_t shift = (uint32_t)(dy * VW * 128);          // dy lines down: whole 1024-byte swizzle atoms
              const uint64_t ah = make_sdesc128_addr(a_hi + shift), al = make_sdesc128_addr(a_lo + shift);
              uint8_t *st = ring_b + (size_t)bs * C::B_STAGE;
              const uint64_t bh = make_sdesc128(st), bl = make_sdesc128(st + C::B_PLANE);
              const uint32_t dm = tmem_base + (uint32_t)((gc & 1) * 2 * NC), dc = dm + (uint32_t)NC;
              const bool chunk_end = kc == KCB - 1 || it == kblocks - 1;
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                  if (p.dry >= 2) break;
                  const uint64_t o = (uint64_t)(kk * 2);
                  const uint32_t acc = (kc == 0 && kk == 0) ? 0u : 1u;
                  if constexpr (CG == 2) {
                    mma_f16_ss_pair(dc, al + o, bh + o, idesc, acc);
                    mma_f16_ss_pair(dc, ah + o, bl + o, idesc, 1u);
                    mma_f16_ss_pair(dm, ah + o, bh + o, idesc, acc);
                  } else {
                    mma_f16_ss(dm, ah + o, bh + o, idesc2, acc);     // [main | a_hi b_lo]  (N = 2 NC: B_lo follows B_hi)
                    mma_f16_ss(dc, al + o, bh + o, idesc, 1u);       // corr += a_lo b_hi
                  }
                }
                if constexpr (CG == 2) {
                  mma_commit_pair(b_empty + bs);
                  if (dy == 2) mma_commit_pair(a_empty + as);
                  if (chunk_end) mma_commit_pair(tmem_full + (gc & 1));
                } else {
                  mma_commit(b_empty + bs);
                  if (dy == 2) mma_commit(a_empty + as);
                  if (chunk_end) mma_commit(tmem_full + (gc & 1));
                }
              }
              __syncwarp();
            }
          }
        gch += (uint32_t)nchunks;
        if (p.dbg && blockIdx.x == 0 && ord < 16 && lane == 0) g_conv_dbg[ord * 8 + 1] = clock64();
      }
    }
  } else {
    // ===================== warps 2..9: thread <-> (pixel of the tile, half of the tile's channels) =====================
    constexpr int HC = NC / 2;
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ty = row >> 3, tx = row & 7;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HC);
    uint32_t gch = 0, ovf = 0;
    int ord = 0;
    for (int item = unit; item < n_items; item += n_units, ++ord) {
      int cout0, n, h0, w0;
      decode(item, cout0, n, h0, w0);
      float acc[HC];
#pragma unroll
      for (int c = 0; c < HC; ++c) acc[c] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch, ++gch) {
        const uint32_t b = gch & 1;
        mbar_wait_b(tmem_full + b, (gch >> 1) & 1);
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
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty + b), 0)); else mbar_arrive(tmem_empty + b);
        }
      }
      if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64 && ord < 16) g_conv_dbg[ord * 8 + 2] = clock64();
      // ---- epilogue of this item (the MMA warp is already filling the ping-pong buffers with the next item) ----
      const int cbase = cout0 + half * HC;
#pragma unroll
      for (int c = 0; c < HC; c += 4) {
        const float4 bv = p.bias != nullptr ? ldg4(p.bias + cbase + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[c] += bv.x; acc[c + 1] += bv.y; acc[c + 2] += bv.z; acc[c + 3] += bv.w;
        if (p.relu) {
          acc[c] = fmaxf(acc[c], 0.f); acc[c + 1] = fmaxf(acc[c + 1], 0.f);
          acc[c + 2] = fmaxf(acc[c + 2], 0.f); acc[c + 3] = fmaxf(acc[c + 3], 0.f);
        }
      }
      int y = h0 + ty, x = w0 + tx, Ho = p.H, Wo = p.W;
      bool writer = true;
      if (p.pool) {
        // 2x2 windows live inside one warp: partner pixels are lanes ^ 1 (tx) and ^ 8 (ty)
#pragma unroll
        for (int c = 0; c < HC; ++c) {
          float v = acc[c];
          v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
          v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
          acc[c] = v;
        }
        writer = !(tx & 1) && !(ty & 1);
        y >>= 1; x >>= 1; Ho >>= 1; Wo >>= 1;
      }
      if (writer && n < p.B && y < Ho && x < Wo) {
        if (p.out_f32 != nullptr) {
          float *dst = p.out_f32 + (((size_t)n * p.Cout + cbase) * Ho + y) * Wo + x;
          const size_t cs = (size_t)Ho * Wo;
#pragma unroll
          for (int c = 0; c < HC; ++c) dst[(size_t)c * cs] = acc[c];
        } else {
          const size_t off = (((size_t)n * Ho + y) * Wo + x) * p.Cout + cbase;
#pragma unroll
          for (int c = 0; c < HC; c += 16) {
            uint4 hi[2], lo[2];
#pragma unroll
            for (int k = 0; k < 16; k += 8) {
              uint4 &h = hi[k >> 3], &l = lo[k >> 3];
              split2(acc[c + k], acc[c + k + 1], h.x, l.x); split2(acc[c + k + 2], acc[c + k + 3], h.y, l.y);
              split2(acc[c + k + 4], acc[c + k + 5], h.z, l.z); split2(acc[c + k + 6], acc[c + k + 7], h.w, l.w);
              ovf |= f16x2_nonfinite(h.x) | f16x2_nonfinite(h.y) | f16x2_nonfinite(h.z) | f16x2_nonfinite(h.w);
            }
            stg256(p.out_hi + off + c, hi[0], hi[1]);
            stg256(p.out_lo + off + c, lo[0], lo[1]);
          }
        }
      }
      if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64 && ord < 16) g_conv_dbg[ord * 8 + 3] = clock64();
    }
    if (ovf) atomicOr(&g_conv_overflow, 1u);
    if (p.dbg && threadIdx.x == 64) {
      long long n1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
      if (blockIdx.x == 0) { g_conv_dbg[15 * 8 + 5] = ord; g_conv_dbg[15 * 8 + 6] = clock64() - dbg_c0; g_conv_dbg[15 * 8 + 7] = n1 - dbg_n0; }
      if (blockIdx.x < 256) {
        g_conv_cta[blockIdx.x * 4] = ord; g_conv_cta[blockIdx.x * 4 + 1] = clock64() - dbg_c0;
        g_conv_cta[blockIdx.x * 4 + 2] = dbg_n0; g_conv_cta[blockIdx.x * 4 + 3] = n1;
      }
    }
  }
  tc::fence_before_sync();
  if constexpr (CG == 2) {
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// First VGG layer: 3 -> Cout (64) channels, fp32 NCHW image in, relu(conv) as NHWC fp16 planes out.
// thread <-> one pixel, ALL output channels: every lane of a warp reads the same weight, so the [27][COUT] weight table in
// shared memory is read with broadcast LDS.128 (four channels per load, no bank conflicts) and the 27 input loads of a warp
// are 128-byte coalesced rows of the NCHW image.  (The first version mapped a thread to 16 channels of a pixel: four
// different weight rows per warp load, 2-way bank conflicts, one LDS per FMA — 5.25 ms at 32 x 608 x 608, LSU-bound;
// the 3 GB of output planes are 0.55 ms of HBM time.)
template <int COUT>
__global__ void __launch_bounds__(256) k_conv_first(const float *__restrict__ img, const float *__restrict__ w,
                                                    const float *__restrict__ bias, int B, int H, int W,
                                                    __half *__restrict__ out_hi, __half *__restrict__ out_lo) {
  __shared__ __align__(16) float sw[27 * COUT + COUT];     // [j = c * 9 + dy * 3 + dx][cout], then bias
  for (int i = threadIdx.x; i < COUT * 27; i += blockDim.x) {
    const int co = i / 27, j = i - co * 27;
    sw[j * COUT + co] = w[i];
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sw[27 * COUT + i] = bias[i];
  __syncthreads();
  const size_t total = (size_t)B * H * W;
  uint32_t ovf = 0;
  for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(pix % W);
    const size_t r = pix / W;
    const int y = (int)(r % H);
    const int n = (int)(r / H);
    float in[27];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int yy = y + dy - 1, xx = x + dx - 1;
          in[c * 9 + dy * 3 + dx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (((size_t)n * 3 + c) * H + yy) * W + xx) : 0.f;
        }
    const size_t off = pix * COUT;
#pragma unroll
    for (int g = 0; g < COUT; g += 16) {                    // 16 channels at a time: 16 accumulators live
      float o[16];
#pragma unroll
      for (int k = 0; k < 16; k += 4) {
        const float4 bv = *reinterpret_cast<const float4 *>(sw + 27 * COUT + g + k);
        o[k] = bv.x; o[k + 1] = bv.y; o[k + 2] = bv.z; o[k + 3] = bv.w;
      }
#pragma unroll
      for (int j = 0; j < 27; ++j) {
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
          const float4 wv = *reinterpret_cast<const float4 *>(sw + j * COUT + g + k);   // same address in every lane: broadcast
          o[k] = fmaf(in[j], wv.x, o[k]); o[k + 1] = fmaf(in[j], wv.y, o[k + 1]);
          o[k + 2] = fmaf(in[j], wv.z, o[k + 2]); o[k + 3] = fmaf(in[j], wv.w, o[k + 3]);
        }
      }
      uint4 hi[2], lo[2];
#pragma unroll
      for (int k = 0; k < 16; k += 8) {
        uint4 &h = hi[k >> 3], &l = lo[k >> 3];
        split2(fmaxf(o[k], 0.f), fmaxf(o[k + 1], 0.f), h.x, l.x); split2(fmaxf(o[k + 2], 0.f), fmaxf(o[k + 3], 0.f), h.y, l.y);
        split2(fmaxf(o[k + 4], 0.f), fmaxf(o[k + 5], 0.f), h.z, l.z); split2(fmaxf(o[k + 6], 0.f), fmaxf(o[k + 7], 0.f), h.w, l.w);
        ovf |= f16x2_nonfinite(h.x) | f16x2_nonfinite(h.y) | f16x2_nonfinite(h.z) | f16x2_nonfinite(h.w);
      }
      stg256(out_hi + off + g, hi[0], hi[1]);                // 16 channels = one 32-byte sector per plane
      stg256(out_lo + off + g, lo[0], lo[1]);
    }
  }
  if (ovf) atomicOr(&g_conv_overflow, 1u);
}

// w [Cout, Cin, 3, 3] fp32 -> planes [Cout][tap][Cin] fp16 (hi, lo * 2^11): the K order of the implicit GEMM
__global__ void k_conv_weight_planes(const float *__restrict__ w, int Cout, int Cin, __half *__restrict__ hi, __half *__restrict__ lo) {
  const size_t total = (size_t)Cout * 9 * Cin;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int tap = (int)((i / Cin) % 9);
    const int co = (int)(i / ((size_t)9 * Cin));
    const float v = w[((size_t)co * Cin + ci) * 9 + tap];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn((v - __half2float(h)) * LO_SCALE);
  }
}

static int make_tmap_4d(CUtensorMap *m, const void *base, int B, int H, int W, int Cc, int box_w = TW, int box_h = TH) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[4] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void *)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled (4d) failed (%d) B=%d H=%d W=%d C=%d", (int)r, B, H, W, Cc);
  return 0;
}

// SGG_CONV_CG = 1 | 2 (default 1): CTAs per UMMA tile.  The cta_group::2 variants are exact and halve the weight bytes each SM
// stages, but measured no faster (profiles/r02_conv_experiments.md): these kernels are bound by the tensor pipe's
// power cap and by instruction count, not by shared-memory or L2 bandwidth, and the pair cannot use the fused two-instruction
// pattern (its 2 NC columns would interleave the two CTAs' halves).
static int conv_cta_group() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SGG_CONV_CG");
    v = (e && atoi(e) == 2) ? 2 : 1;
  }
  return v;
}

template <int NC, int CG>
static int launch_conv(const ConvParams &p, const __half *in_hi, const __half *in_lo, const __half *w_hi, const __half *w_lo,
                       cudaStream_t st) {
  using C = CCfg<NC, CG>;
  static bool attr = false;
  if (!attr) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_conv3x3<NC, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  CUtensorMap tm[4];
  int rc;
  if ((rc = make_tmap_4d(&tm[0], in_hi, p.B, p.H, p.W, p.Cin))) return rc;
  if ((rc = make_tmap_4d(&tm[1], in_lo, p.B, p.H, p.W, p.Cin))) return rc;
  if ((rc = make_tmap(&tm[2], w_hi, p.Cout, 9 * p.Cin, NC / CG, 2))) return rc;
  if ((rc = make_tmap(&tm[3], w_lo, p.Cout, 9 * p.Cin, NC / CG, 2))) return rc;
  const long long tiles = (long long)p.tiles_w * p.tiles_h * p.B;
  if (tiles > 0x7ffffff0LL) return sgg_set_err(SGG_E_BADARG, "conv3x3: too many tiles");
  if (CG == 1) {
    dim3 grid((unsigned)tiles, p.Cout / NC);
    k_conv3x3<NC, CG><<<grid, NTHR, C::SMEM, st>>>(p, tm[0], tm[1], tm[2], tm[3]);
  } else {
    // pairs of adjacent tiles; an odd tile count is padded with a CTA whose image index is B: its loads are
    // zero-filled (out of bounds) and its stores are masked
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((tiles + 1) & ~1LL), p.Cout / NC);
    cfg.blockDim = dim3(NTHR);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SGG_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_conv3x3<NC, CG>, p, tm[0], tm[1], tm[2], tm[3]));
  }
  SGG_RETURN_IF_LAUNCH_FAILED("k_conv3x3");
  return 0;
}

// SGG_CONV_V = 1 | 2: force a kernel version; default 0 = per layer: v2 (slabs + persistent CTAs) up to 256 input channels,
// v1 (per-tap boxes, one tile per CTA) from 512 — measured inside the stack at 32 x 608^2 (profiles/r02_conv_experiments.md
// section 8): v2 wins 4.58 -> 3.87, 1.96 -> 1.27, 1.40 -> 1.19 ms on the K = 576 / 1152 layers, v1 wins 2.40 -> 2.2 and
// 0.72 -> 0.65 ms on the 512-channel layers (72 k-blocks per tile amortise its prologue; less control per k-block).
static int conv_version(int Cin) {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SGG_CONV_V");
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 2) v = 0;
  }
  return v ? v : (Cin >= 512 ? 1 : 2);
}

template <int NC, int CG>
static int launch_conv_v2(ConvParams p, const __half *in_hi, const __half *in_lo, const __half *w_hi, const __half *w_lo,
                          cudaStream_t st) {
  using C = VCfg<NC, CG>;
  static bool attr = false;
  if (!attr) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_conv3x3_v2<NC, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  p.tiles_w = (p.W + VW - 1) / VW; p.tiles_h = (p.H + VH - 1) / VH;
  static const int kcb_env = getenv("SGG_CONV_KCB") ? atoi(getenv("SGG_CONV_KCB")) : 0;   // experiment knob
  p.kcb = kcb_env > 0 ? kcb_env : 256 / BK;
  static const int dry_env = getenv("SGG_CONV_DRY") ? atoi(getenv("SGG_CONV_DRY")) : 0;
  p.dry = dry_env;
  static const int dbg_env = getenv("SGG_CONV_DBG") ? atoi(getenv("SGG_CONV_DBG")) : 0;
  p.dbg = dbg_env;
  CUtensorMap tm[4];
  int rc;
  if ((rc = make_tmap_4d(&tm[0], in_hi, p.B, p.H, p.W, p.Cin, VW, VH + 2))) return rc;
  if ((rc = make_tmap_4d(&tm[1], in_lo, p.B, p.H, p.W, p.Cin, VW, VH + 2))) return rc;
  if ((rc = make_tmap(&tm[2], w_hi, p.Cout, 9 * p.Cin, NC / CG, 2))) return rc;
  if ((rc = make_tmap(&tm[3], w_lo, p.Cout, 9 * p.Cin, NC / CG, 2))) return rc;
  const long long tiles = (long long)p.tiles_w * p.tiles_h * p.B;
  const long long n_sp = (tiles + CG - 1) / CG;                 // spatial work units (tiles, or pairs of adjacent tiles)
  const long long items = n_sp * (p.Cout / NC);
  if (items > 0x7ffffff0LL) return sgg_set_err(SGG_E_BADARG, "conv3x3: too many tiles");
  const int sms = sgg_num_sms();
  const long long units = items < sms / CG ? items : sms / CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(units * CG));
  cfg.blockDim = dim3(NTHR);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  SGG_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_conv3x3_v2<NC, CG>, p, (int)items, (int)n_sp, tm[0], tm[1], tm[2], tm[3]));
  SGG_RETURN_IF_LAUNCH_FAILED("k_conv3x3_v2");
  if (p.dbg) {                                  // debugging aid: print CTA 0's phase stamps relative to its first one
    long long h[16 * 8];
    if (cudaDeviceSynchronize() == cudaSuccess && cudaMemcpyFromSymbol(h, g_conv_dbg, sizeof(h)) == cudaSuccess) {
      fprintf(stderr, "[conv dbg] Cin %d Cout %d H %d: item: mma-start mma-end | drained epilogue-done (cycles since first stamp)\n", p.Cin, p.Cout, p.H);
      const long long t0 = h[0];
      for (int i = 0; i < 10; ++i)
        fprintf(stderr, "[conv dbg] %2d: %7lld %7lld | %7lld %7lld\n", i, h[i * 8] - t0, h[i * 8 + 1] - t0, h[i * 8 + 2] - t0, h[i * 8 + 3] - t0);
      long long c[256 * 4];
      if (cudaMemcpyFromSymbol(c, g_conv_cta, sizeof(c)) == cudaSuccess) {
        const int nb = (int)(units * CG) < 256 ? (int)(units * CG) : 256;
        long long tmin = c[2], tmax = c[3], cmin = c[1], cmax = c[1], first_end = c[3];
        for (int b = 0; b < nb; ++b) {
          if (c[b * 4 + 2] < tmin) tmin = c[b * 4 + 2];
          if (c[b * 4 + 3] > tmax) tmax = c[b * 4 + 3];
          if (c[b * 4 + 3] < first_end) first_end = c[b * 4 + 3];
          if (c[b * 4 + 1] < cmin) cmin = c[b * 4 + 1];
          if (c[b * 4 + 1] > cmax) cmax = c[b * 4 + 1];
        }
        long long smax = 0;
        for (int b = 0; b < nb; ++b) if (c[b * 4 + 2] - tmin > smax) smax = c[b * 4 + 2] - tmin;
        fprintf(stderr, "[conv dbg] %d CTAs: span %lld ns (first start -> last end), last CTA start +%lld ns, first end +%lld ns, cycles min %lld max %lld\n", nb,
                tmax - tmin, smax, first_end - tmin, cmin, cmax);
      }
      fprintf(stderr, "[conv dbg] CTA 0: %lld items in %lld cycles = %lld ns (%.2f GHz): %.0f cycles / item\n", h[15 * 8 + 5], h[15 * 8 + 6], h[15 * 8 + 7],
              (double)h[15 * 8 + 6] / (double)h[15 * 8 + 7], (double)h[15 * 8 + 6] / (double)(h[15 * 8 + 5] ? h[15 * 8 + 5] : 1));
    }
  }
  return 0;
}

}  // namespace conv
}  // namespace sgg

// ---- C-ABI ------------------------------------------------------------------------------------------------------
/* w [Cout, Cin, 3, 3] fp32 -> [hi | lo] fp16 planes in implicit-GEMM order [Cout][tap][Cin] (2 * Cout * 9 * Cin halves) */
extern "C" int sgg_conv_weight_planes(const float *w, int Cout, int Cin, void *planes, void *stream) {
  if (!w || !planes || Cout <= 0 || Cin <= 0) return sgg_set_err(SGG_E_BADARG, "conv_weight_planes: bad argument");
  const size_t n = (size_t)Cout * 9 * Cin;
  __half *hi = (__half *)planes;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  sgg::conv::k_conv_weight_planes<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, hi, hi + n);
  SGG_RETURN_IF_LAUNCH_FAILED("k_conv_weight_planes");
  return 0;
}

/* first VGG layer: img [B,3,H,W] fp32 NCHW, w [64,3,3,3], bias [64] -> relu(conv) as NHWC planes [hi | lo] (2*B*H*W*64 halves) */
extern "C" int sgg_conv3x3_first(const float *img, const float *w, const float *bias, int B, int H, int W, int Cout,
                                 void *out_planes, void *stream) {
  if (!img || !w || !bias || !out_planes || B <= 0 || H <= 0 || W <= 0) return sgg_set_err(SGG_E_BADARG, "conv3x3_first: bad argument");
  if (Cout != 64) return sgg_set_err(SGG_E_BADARG, "conv3x3_first: Cout must be 64");
  const size_t n = (size_t)B * H * W * 64;
  __half *hi = (__half *)out_planes;
  const size_t threads = (size_t)B * H * W;
  const int blocks = (int)((threads + 255) / 256 < 148 * 16 ? (threads + 255) / 256 : 148 * 16);
  sgg::conv::k_conv_first<64><<<blocks, 256, 0, (cudaStream_t)stream>>>(img, w, bias, B, H, W, hi, hi + n);
  SGG_RETURN_IF_LAUNCH_FAILED("k_conv_first");
  return 0;
}

/* 3x3 / pad 1 / stride 1 convolution (+ bias, optional ReLU, optional fused 2x2 max-pool) on tcgen05.
 * in_planes: NHWC [hi | lo] planes of [B,H,W,Cin]; w_planes from sgg_conv_weight_planes.  Cin % 64 == 0, Cout % 64 == 0.
 * Output: out_planes (NHWC planes of [B,Ho,Wo,Cout]) or, if out_f32_nchw != NULL, fp32 NCHW [B,Cout,Ho,Wo]. */
extern "C" int sgg_conv3x3_tc(const void *in_planes, const void *w_planes, const float *bias, int B, int H, int W, int Cin,
                              int Cout, int relu, int pool, void *out_planes, float *out_f32_nchw, void *stream) {
  using namespace sgg::conv;
  if (!in_planes || !w_planes || (!out_planes && !out_f32_nchw)) return sgg_set_err(SGG_E_BADARG, "conv3x3_tc: null pointer");
  if (B <= 0 || H <= 0 || W <= 0 || (Cin % 64) || (Cout % 64)) return sgg_set_err(SGG_E_BADARG, "conv3x3_tc: bad shape");
  if (pool && ((H & 1) || (W & 1))) return sgg_set_err(SGG_E_BADARG, "conv3x3_tc: fused pooling needs even H, W");
  ConvParams p{};
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.relu = relu; p.pool = pool; p.bias = bias;
  p.tiles_w = (W + TW - 1) / TW; p.tiles_h = (H + TH - 1) / TH;
  const size_t n_in = (size_t)B * H * W * Cin;
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const size_t n_out = (size_t)B * Ho * Wo * Cout;
  const __half *ih = (const __half *)in_planes, *wh = (const __half *)w_planes;
  p.out_hi = out_f32_nchw ? nullptr : (__half *)out_planes;
  p.out_lo = out_f32_nchw ? nullptr : (__half *)out_planes + n_out;
  p.out_f32 = out_f32_nchw;
  const size_t wn = (size_t)Cout * 9 * Cin;
  const bool pair = conv_cta_group() == 2;
  if (conv_version(Cin) == 2) {
    if (Cout % 128 == 0)
      return pair ? launch_conv_v2<128, 2>(p, ih, ih + n_in, wh, wh + wn, (cudaStream_t)stream)
                  : launch_conv_v2<128, 1>(p, ih, ih + n_in, wh, wh + wn, (cudaStream_t)stream);
    return pair ? launch_conv_v2<64, 2>(p, ih, ih + n_in, wh, wh + wn, (cudaStream_t)stream)
                : launch_conv_v2<64, 1>(p, ih, ih + n_in, wh, wh + wn, (cudaStream_t)stream);
  }
  if (Cout % 128 == 0)
    return pair ? launch_conv<128, 2>(p, ih, ih + n_in, wh, wh + wn, (cudaStream_t)stream)
                : launch_conv<128, 1>(p, ih, ih + n_in, wh, wh + wn, (cudaStream_t)stream);
  return pair ? launch_conv<64, 2>(p, ih, ih + n_in, wh, wh + wn, (cudaStream_t)stream)
              : launch_conv<64, 1>(p, ih, ih + n_in, wh, wh + wn, (cudaStream_t)stream);
}

extern "C" int sgg_conv_overflow(int reset) {
  unsigned int v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&v, sgg::conv::g_conv_overflow, sizeof(v)) != cudaSuccess) return -1;
  if (reset) {
    const unsigned int z = 0;
    if (cudaMemcpyToSymbol(sgg::conv::g_conv_overflow, &z, sizeof(z)) != cudaSuccess) return -1;
  }
  return (int)v;
}
