// VGG16 conv stack of the frozen detector backbone (rel_model_base.py:184, :310-321: torchvision vgg16 `features`
// without the last max-pool) as tcgen05 implicit GEMMs — SURVEY.md §8f rank 2.  The reference runs it through cuDNN in
// fp32 (226 GFLOP per 608 x 608 image: 89 % of its forward); here every 3x3 / pad 1 / stride 1 convolution is
//
//   out[pixel, cout] = sum_{tap, cin} in[pixel + tap, cin] * w[cout, tap, cin]
//
// with activations kept as NHWC fp16 planes [hi | lo * 2^11] (tc16_common.cuh: x = hi + 2^-11 lo, three MMA passes give
// fp32-grade results), so the im2col operand never exists: for tap (dy, dx) and a block of 64 input channels, the A tile
// of an 8 x 16 pixel output tile is ONE 4-D TMA box {64 c, 16 w, 8 h, 1 n} at (w0 + dx - 1, h0 + dy - 1) — the hardware's
// out-of-bounds zero fill is the padding.  K = 9 * Cin is folded into TMEM 256 at a time (two ping-pong accumulator
// pairs, the tensor core's fp32 accumulate truncates: DESIGN.md section 4); eight warps drain the chunks into registers.
// Epilogue: bias + ReLU, optional fused 2x2 max-pool (the tile holds whole windows), output again as fp16 planes (NHWC)
// or, for the last layer, fp32 NCHW (the layout RelModel.fmap has in the reference).
// The first layer (Cin = 3, K = 27) is a small SIMT kernel that also converts NCHW fp32 images to planes.
#include <stdlib.h>
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

template <int NC>
struct CCfg {
  static constexpr int A_PLANE = BM * BK * 2;              // 16 KB
  static constexpr int B_PLANE = NC * BK * 2;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;  // NC = 128: 64 KB; NC = 64: 48 KB
  static constexpr int STAGES = NC == 128 ? 3 : 4;
  static constexpr int RING = STAGES * STAGE;
  static constexpr int SMEM = RING + 1024 + 256;
  static constexpr int TS = NC + 4;                        // fp32 staging tile row stride (pool / NCHW epilogues)
  static_assert(BM * TS * 4 <= RING, "staging tile must fit in the ring");
};

template <int NC>
__global__ void __launch_bounds__(NTHR, 1)
k_conv3x3(const ConvParams p, const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
          const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl) {
  using namespace tc;
  using C = CCfg<NC>;
  constexpr int STAGES = C::STAGES, STAGE = C::STAGE, KCB = 256 / BK;
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
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1); mbar_init(tmem_full + 1, 1);
    mbar_init(tmem_empty, 256); mbar_init(tmem_empty + 1, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: k-block it = (channel block cb, tap) =====================
    if (lane == 0) {
      for (int it = 0; it < kblocks; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait_b(empty + s, ph ^ 1);
        const int cb = it / 9, tap = it - cb * 9;
        const int dy = tap / 3, dx = tap - dy * 3;
        uint8_t *st = smem + (size_t)s * STAGE;
        mbar_arrive_expect_tx(full + s, STAGE);
        tma_load_4d(st, &tmAh, full + s, cb * BK, w0 + dx - 1, h0 + dy - 1, n);
        tma_load_4d(st + C::A_PLANE, &tmAl, full + s, cb * BK, w0 + dx - 1, h0 + dy - 1, n);
        tma_load_2d(st + 2 * C::A_PLANE, &tmBh, full + s, tap * p.Cin + cb * BK, cout0);
        tma_load_2d(st + 2 * C::A_PLANE + C::B_PLANE, &tmBl, full + s, tap * p.Cin + cb * BK, cout0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, NC);
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
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);
          const uint32_t acc = (kc == 0 && kk == 0) ? 0u : 1u;
          mma_f16_ss(dc, al + o, bh + o, idesc, acc);
          mma_f16_ss(dc, ah + o, bl + o, idesc, 1u);
          mma_f16_ss(dm, ah + o, bh + o, idesc, acc);
        }
        mma_commit(empty + s);
        if (kc == KCB - 1 || it == kblocks - 1) mma_commit(tmem_full + (chunk & 1));
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
      mbar_arrive(tmem_empty + b);
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
      if (y < p.H && x < p.W) {
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
          if (gy < Ho && gx < Wo) p.out_f32[(((size_t)n * p.Cout + cout0 + c) * Ho + gy) * Wo + gx] = v;
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
          if (gy < Ho && gx < Wo) {
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
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// First VGG layer: 3 -> Cout (64) channels, fp32 NCHW image in, relu(conv) as NHWC fp16 planes out.
// thread <-> (pixel, 16 output channels); weights [Cout][27] + bias staged in shared memory.
template <int COUT>
__global__ void __launch_bounds__(256) k_conv_first(const float *__restrict__ img, const float *__restrict__ w,
                                                    const float *__restrict__ bias, int B, int H, int W,
                                                    __half *__restrict__ out_hi, __half *__restrict__ out_lo) {
  __shared__ float sw[COUT * 27 + COUT];
  for (int i = threadIdx.x; i < COUT * 27; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sw[COUT * 27 + i] = bias[i];
  __syncthreads();
  constexpr int GROUPS = COUT / 16;
  const size_t total = (size_t)B * H * W * GROUPS;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int gq = (int)(i % GROUPS);
    size_t pix = i / GROUPS;
    const int x = (int)(pix % W); pix /= W;
    const int y = (int)(pix % H);
    const int n = (int)(pix / H);
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
    float o[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float *wk = sw + (gq * 16 + k) * 27;
      float s = sw[COUT * 27 + gq * 16 + k];
#pragma unroll
      for (int j = 0; j < 27; ++j) s = fmaf(in[j], wk[j], s);
      o[k] = fmaxf(s, 0.f);
    }
    const size_t off = (((size_t)n * H + y) * W + x) * COUT + gq * 16;
#pragma unroll
    for (int k = 0; k < 16; k += 8) {
      uint4 hi, lo;
      split2(o[k], o[k + 1], hi.x, lo.x); split2(o[k + 2], o[k + 3], hi.y, lo.y);
      split2(o[k + 4], o[k + 5], hi.z, lo.z); split2(o[k + 6], o[k + 7], hi.w, lo.w);
      if (f16x2_nonfinite(hi.x) | f16x2_nonfinite(hi.y) | f16x2_nonfinite(hi.z) | f16x2_nonfinite(hi.w))
        atomicOr(&g_conv_overflow, 1u);
      *reinterpret_cast<uint4 *>(out_hi + off + k) = hi;
      *reinterpret_cast<uint4 *>(out_lo + off + k) = lo;
    }
  }
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

static int make_tmap_4d(CUtensorMap *m, const void *base, int B, int H, int W, int Cc) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[4] = {(cuuint64_t)Cc, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)Cc * 2, (cuuint64_t)W * Cc * 2, (cuuint64_t)H * W * Cc * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void *)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled (4d) failed (%d) B=%d H=%d W=%d C=%d", (int)r, B, H, W, Cc);
  return 0;
}

template <int NC>
static int launch_conv(const ConvParams &p, const __half *in_hi, const __half *in_lo, const __half *w_hi, const __half *w_lo,
                       cudaStream_t st) {
  using C = CCfg<NC>;
  static bool attr = false;
  if (!attr) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_conv3x3<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  CUtensorMap tm[4];
  int rc;
  if ((rc = make_tmap_4d(&tm[0], in_hi, p.B, p.H, p.W, p.Cin))) return rc;
  if ((rc = make_tmap_4d(&tm[1], in_lo, p.B, p.H, p.W, p.Cin))) return rc;
  if ((rc = make_tmap(&tm[2], w_hi, p.Cout, 9 * p.Cin, NC, 2))) return rc;
  if ((rc = make_tmap(&tm[3], w_lo, p.Cout, 9 * p.Cin, NC, 2))) return rc;
  const long long tiles = (long long)p.tiles_w * p.tiles_h * p.B;
  if (tiles > 0x7fffffffLL) return sgg_set_err(SGG_E_BADARG, "conv3x3: too many tiles");
  dim3 grid((unsigned)tiles, p.Cout / NC);
  k_conv3x3<NC><<<grid, NTHR, C::SMEM, st>>>(p, tm[0], tm[1], tm[2], tm[3]);
  SGG_RETURN_IF_LAUNCH_FAILED("k_conv3x3");
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
  const size_t threads = (size_t)B * H * W * 4;
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
  if (Cout % 128 == 0) return launch_conv<128>(p, ih, ih + n_in, wh, wh + wn, (cudaStream_t)stream);
  return launch_conv<64>(p, ih, ih + n_in, wh, wh + wn, (cudaStream_t)stream);
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
