// Tensor-core (tcgen05 + TMEM + TMA) GEMM engine with fused epilogues, fp32-accurate via 3xTF32.
//
//   D[128 x N] (TMEM, fp32) = A_hi*B_hi  (+)  A_lo*B_hi + A_hi*B_lo        (kind::tf32, K = 8 / instruction)
//
// with x_hi = x & 0xffffe000 (exactly representable in tf32) and x_lo = x - x_hi (exact in fp32; the
// tensor core keeps its top 11 bits) => ~2^-21 relative error per product, i.e. fp32-grade results
// (the reference's 1e-4 parity bar leaves no room for single-pass TF32 on K = 4096).
//
// Accumulation accuracy.  The tensor core's fp32 accumulator truncates on every accumulate step (measured
// on B200: the error grows linearly with the number of MMAs folded into one accumulator, 1.9e-4 at K=4096).
// Two counter-measures: (1) the large products A_hi*B_hi go to a MAIN accumulator and the two small
// correction products to a separate CORR accumulator (main + corr is formed in fp32 registers by the
// epilogue), which cuts the truncating accumulations of the large term 3x; (2) the LINEAR kernels fold
// only KC = 256 of K into TMEM at a time: two TMEM buffer pairs ping-pong and the epilogue warps drain each
// finished chunk into fp32 REGISTER accumulators (round-to-nearest) while the next chunk is multiplied.
//
// One CTA = one 128-row x (NBLK x NBR)-column output tile (and one K range when split-K), 10 warps:
//   warp 0   : TMA producer  - per k-block (BKF floats = one swizzle span) loads the raw fp32 A tile and
//              the pre-split B_hi / B_lo tiles (weights are split once per weight version by
//              sgg_tc_split_weights) into a multi-stage ring (SWIZZLE_128B or _64B, K-major).
//   warp 1   : allocates TMEM; one thread issues 3 tcgen05.mma per 8 of K and commits to the stage's
//              "empty" barrier; chunk / final commits signal the epilogue warps.
//   warps 2-5: split A in shared memory (hi in place, lo to a second buffer), fence.proxy.async, arrive.
//   warps 6-9: accumulate / epilogue warps, one thread per accumulator row (TMEM lane).
// B column blocks are NBR weight rows at arbitrary row bases, so a tile can cover hidden units
// [j0, j0+NBR) of the r, z, n gates of a GRUCell (rows j0, H+j0, 2H+j0) and apply the GRU
// non-linearity - and for edges the gated gather of P = V W_ih^T rows - straight out of TMEM.
#include "tc_gemm.cuh"
#include "kernels.h"

namespace sgg {
namespace tc {

constexpr int BM = 128;
constexpr int NTHR = 320;
constexpr int SMEM_BUDGET = 224 * 1024;

enum { EPI_LINEAR = 0, EPI_GRU_INIT = 1, EPI_GRU_NODE = 2, EPI_GRU_EDGE = 3 };

struct Params {
  int M, K, H, Nout, relu;
  int kb_per_split;           // LINEAR split-K: k-blocks per blockIdx.z (== all k-blocks when gridDim.z == 1)
  const float *bias;          // LINEAR
  float *out;                 // LINEAR: [M,Nout] (or partials [splits,M,Nout]); GRU: [M,H]
  const float *b_ih, *b_hh;   // GRU
  const float *h;             // GRU_NODE / GRU_EDGE: previous state [M,H]
  const float *P;             // GRU_EDGE: [N,3H]
  const float *gates;         // GRU_EDGE: [M,4]
  const int *subj, *obj;      // GRU_EDGE
  float *cache;               // GRU: nullable [M,4,H] (r, z, n, gh_n) for the backward pass
};

template <int NBLK, int NBR, int BKF>
struct Cfg {
  static constexpr int A_BYTES = BM * BKF * 4;
  static constexpr int BB_BYTES = NBR * BKF * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * NBLK * BB_BYTES;
  // EVEN stage count: the two converter groups alternate k-blocks, so with an odd ring a group would revisit a
  // stage only every second phase and its parity wait could be satisfied by the phase it skipped (mbarrier
  // phases alias modulo 2) -> stale data / deadlock.  With an even ring group g only ever sees stages of parity g.
  static constexpr int STAGES_RAW = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int STAGES = STAGES_RAW >= 2 ? (STAGES_RAW & ~1) : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
  static constexpr int KCB = 256 / BKF;       // k-blocks per accumulation chunk (LINEAR)
};

__device__ __forceinline__ float gru_point(float gi_r, float gh_r, float gi_z, float gh_z, float gi_n, float gh_n,
                                           float h, float *cache, int H) {
  const float r = sgg_sigmoid(gi_r + gh_r);
  const float z = sgg_sigmoid(gi_z + gh_z);
  const float n = tanhf(gi_n + r * gh_n);
  if (cache != nullptr) { cache[0] = r; cache[H] = z; cache[2 * H] = n; cache[3 * H] = gh_n; }
  return (1.0f - z) * n + z * h;
}

template <int NBLK, int NBR, int NSEG, int EPI, int BKF>
__global__ void __launch_bounds__(NTHR, 1)
k_tc_gemm(Params p, const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
          const __grid_constant__ CUtensorMap tmBh0, const __grid_constant__ CUtensorMap tmBl0,
          const __grid_constant__ CUtensorMap tmBh1, const __grid_constant__ CUtensorMap tmBl1) {
  using C = Cfg<NBLK, NBR, BKF>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, A_BYTES = C::A_BYTES, BB_BYTES = C::BB_BYTES;
  constexpr int KCB = C::KCB;
  constexpr int NCOL = NBLK * NBR;                 // MMA N
  constexpr bool CHUNKED = (EPI == EPI_LINEAR);
  static_assert(BKF == 32 || BKF == 16, "k-block = one 128B or 64B swizzle span");
  static_assert(!CHUNKED || (NSEG == 1 && NCOL <= 128), "chunked LINEAR keeps NCOL register accumulators per thread");
  static_assert(NCOL % 16 == 0 && NCOL <= 256, "UMMA N");
  // TMEM columns: GRU: [main seg0..|corr seg0..]; LINEAR: buffer b = [main_b | corr_b]
  constexpr int CORR = CHUNKED ? NCOL : NSEG * NCOL;          // offset of the correction accumulator
  constexpr int ACC_COLS = CHUNKED ? 4 * NCOL : 2 * NSEG * NCOL;
  constexpr uint32_t TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static_assert(ACC_COLS <= 512, "tile too wide for TMEM");
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
  uint64_t *full = bars, *ready = bars + STAGES, *empty = bars + 2 * STAGES, *tmem_full = bars + 3 * STAGES;
  uint64_t *tmem_empty = bars + 3 * STAGES + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM;
  const int j0 = blockIdx.x * (EPI == EPI_LINEAR ? NCOL : NBR);
  const int kblocks_all = (p.K + BKF - 1) / BKF;
  // split-K (LINEAR only): this CTA reduces k-blocks [kb_lo, kb_lo + kblocks)
  const int kb_lo = CHUNKED ? (int)blockIdx.z * p.kb_per_split : 0;
  const int kblocks = CHUNKED ? min(p.kb_per_split, kblocks_all - kb_lo) : kblocks_all;
  const int total = NSEG * kblocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0); prefetch_tmap(&tmBh0); prefetch_tmap(&tmBl0);
    if (NSEG > 1) { prefetch_tmap(&tmA1); prefetch_tmap(&tmBh1); prefetch_tmap(&tmBl1); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(ready + s, 64); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1); mbar_init(tmem_full + 1, 1);
    mbar_init(tmem_empty, 128); mbar_init(tmem_empty + 1, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto stage_ptr = [&](int s) { return smem + (size_t)s * STAGE_BYTES; };   // [A_hi | A_lo | B_hi | B_lo]

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);
        const int seg = it / kblocks, k0 = (kb_lo + it - seg * kblocks) * BKF;
        uint8_t *st = stage_ptr(s);
        mbar_arrive_expect_tx(full + s, A_BYTES + 2 * NBLK * BB_BYTES);
        tma_load_2d(st, seg == 0 ? &tmA0 : &tmA1, full + s, k0, m0);
#pragma unroll
        for (int b = 0; b < NBLK; ++b) {
          const int row = (EPI == EPI_LINEAR) ? (j0 + b * NBR) : (b * p.H + j0);
          tma_load_2d(st + 2 * A_BYTES + b * BB_BYTES, seg == 0 ? &tmBh0 : &tmBh1, full + s, k0, row);
          tma_load_2d(st + 2 * A_BYTES + (NBLK + b) * BB_BYTES, seg == 0 ? &tmBl0 : &tmBl1, full + s, k0, row);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, NCOL);
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(full + s, ph);
        mbar_wait(ready + s, ph);
        fence_after_sync();
        const int seg = it / kblocks, kb = it - seg * kblocks;
        uint8_t *st = stage_ptr(s);
        const uint64_t ah = make_sdesc<BKF>(st), al = make_sdesc<BKF>(st + A_BYTES);
        const uint64_t bh = make_sdesc<BKF>(st + 2 * A_BYTES), bl = make_sdesc<BKF>(st + 2 * A_BYTES + NBLK * BB_BYTES);
        uint32_t dm, first;
        const int chunk = it / KCB, kc = it - chunk * KCB;
        if (CHUNKED) {
          if (kc == 0) {                       // buffer pair must have been drained by the epilogue warps
            mbar_wait(tmem_empty + (chunk & 1), ((chunk >> 1) & 1) ^ 1);
            fence_after_sync();
          }
          dm = tmem_base + (uint32_t)((chunk & 1) * 2 * NCOL);
          first = (kc == 0) ? 1u : 0u;
        } else {
          dm = tmem_base + (uint32_t)(seg * NCOL);
          first = (kb == 0) ? 1u : 0u;
        }
        const uint32_t dc = dm + (uint32_t)CORR;
#pragma unroll
        for (int kk = 0; kk < BKF / 8; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);     // 32 bytes >> 4
          const uint32_t acc = (first && kk == 0) ? 0u : 1u;
          mma_tf32_ss(dc, al + o, bh + o, idesc, acc);     // corrections -> CORR
          mma_tf32_ss(dc, ah + o, bl + o, idesc, 1u);
          mma_tf32_ss(dm, ah + o, bh + o, idesc, acc);     // large term  -> MAIN
        }
        mma_commit(empty + s);
        if (CHUNKED && (kc == KCB - 1 || it == total - 1)) mma_commit(tmem_full + (chunk & 1));
      }
      if (!CHUNKED) mma_commit(tmem_full);
    }
  } else if (warp < 6) {
    // ===================== split A (hi / lo) =====================
    // Two groups of two warps alternate k-blocks, so one group's fence.proxy.async (MEMBAR) and barrier
    // round-trip overlap the other group's loads/stores.
    const int grp = (warp - 2) >> 1;                 // 0: warps 2,3   1: warps 4,5
    const int t = (threadIdx.x - 64) & 63;           // 0..63 inside the group
    for (int it = grp; it < total; it += 2) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(full + s, ph);
      const uint32_t a_addr = smem_u32(stage_ptr(s)), lo_addr = a_addr + A_BYTES;
      constexpr int PER = A_BYTES / 16 / 64;         // float4 per thread
#pragma unroll
      for (int i0 = 0; i0 < PER; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = lds128(a_addr + (uint32_t)((t + (i0 + i) * 64) * 16));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 hi, l;
          hi.x = __uint_as_float(__float_as_uint(v[i].x) & 0xffffe000u); l.x = v[i].x - hi.x;
          hi.y = __uint_as_float(__float_as_uint(v[i].y) & 0xffffe000u); l.y = v[i].y - hi.y;
          hi.z = __uint_as_float(__float_as_uint(v[i].z) & 0xffffe000u); l.z = v[i].z - hi.z;
          hi.w = __uint_as_float(__float_as_uint(v[i].w) & 0xffffe000u); l.w = v[i].w - hi.w;
          const uint32_t off = (uint32_t)((t + (i0 + i) * 64) * 16);
          sts128(a_addr + off, hi);
          sts128(lo_addr + off, l);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(ready + s);
    }
  } else {
    // ===================== accumulate / epilogue: thread <-> accumulator row (TMEM lane) =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool live = m < p.M;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (EPI == EPI_LINEAR) {
      float acc[NCOL];
#pragma unroll
      for (int c = 0; c < NCOL; ++c) acc[c] = 0.f;
      const int nchunks = (total + KCB - 1) / KCB;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int b = ch & 1;
        mbar_wait(tmem_full + b, (ch >> 1) & 1);
        fence_after_sync();
        __syncwarp();
#pragma unroll
        for (int c0 = 0; c0 < NCOL; c0 += 16) {
          float v[16], w[16];
          tmem_ld16(taddr + (uint32_t)(b * 2 * NCOL + c0), v);
          tmem_ld16(taddr + (uint32_t)(b * 2 * NCOL + NCOL + c0), w);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c0 + c] += v[c] + w[c];
        }
        fence_before_sync();
        mbar_arrive(tmem_empty + b);
      }
      if (live) {
        const bool partial = gridDim.z > 1;        // split-K: raw partial sums, bias/ReLU applied by the reducer
        const bool vec = (p.Nout & 3) == 0;
        float *yrow = p.out + ((size_t)blockIdx.z * p.M + m) * p.Nout;
#pragma unroll
        for (int c0 = 0; c0 < NCOL; c0 += 4) {
          const int j = j0 + c0;
          if (j < p.Nout) {
            float v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              v[c] = acc[c0 + c];
              if (!partial) {
                if (p.bias != nullptr && j + c < p.Nout) v[c] += __ldg(p.bias + j + c);
                if (p.relu) v[c] = fmaxf(v[c], 0.f);
              }
            }
            if (vec && j + 4 <= p.Nout) {
              *reinterpret_cast<float4 *>(yrow + j) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) if (j + c < p.Nout) yrow[j + c] = v[c];
            }
          }
        }
      }
    } else {
      mbar_wait(tmem_full, 0);
      fence_after_sync();
      const int H = p.H;
      int s_id = 0, o_id = 0; float gs = 0.f, go = 0.f;
      if (EPI == EPI_GRU_EDGE && live) {
        s_id = p.subj[m]; o_id = p.obj[m];
        const float4 g4 = *reinterpret_cast<const float4 *>(p.gates + (size_t)m * 4);
        gs = g4.x; go = g4.y;
      }
      for (int c0 = 0; c0 < NBR; c0 += 8) {
        float ar[8], az[8], an[8], br[8], bz[8], bn[8], t8[8];
        __syncwarp();                              // tcgen05.ld is .sync.aligned: whole warp, converged
        // seg-0 accumulators: main + corr
        tmem_ld8(taddr + (uint32_t)(0 * NBR + c0), ar); tmem_ld8(taddr + (uint32_t)(CORR + 0 * NBR + c0), t8); tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; ++c) ar[c] += t8[c];
        tmem_ld8(taddr + (uint32_t)(1 * NBR + c0), az); tmem_ld8(taddr + (uint32_t)(CORR + 1 * NBR + c0), t8); tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; ++c) az[c] += t8[c];
        tmem_ld8(taddr + (uint32_t)(2 * NBR + c0), an); tmem_ld8(taddr + (uint32_t)(CORR + 2 * NBR + c0), t8); tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; ++c) an[c] += t8[c];
        if (EPI == EPI_GRU_NODE) {
          tmem_ld8(taddr + (uint32_t)(NCOL + 0 * NBR + c0), br); tmem_ld8(taddr + (uint32_t)(CORR + NCOL + 0 * NBR + c0), t8); tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c) br[c] += t8[c];
          tmem_ld8(taddr + (uint32_t)(NCOL + 1 * NBR + c0), bz); tmem_ld8(taddr + (uint32_t)(CORR + NCOL + 1 * NBR + c0), t8); tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c) bz[c] += t8[c];
          tmem_ld8(taddr + (uint32_t)(NCOL + 2 * NBR + c0), bn); tmem_ld8(taddr + (uint32_t)(CORR + NCOL + 2 * NBR + c0), t8); tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c) bn[c] += t8[c];
        }
        if (live) {
          const int j = j0 + c0;
          float hv[8], o8[8];
          if (EPI != EPI_GRU_INIT) {
            const float4 h0 = *reinterpret_cast<const float4 *>(p.h + (size_t)m * H + j);
            const float4 h1 = *reinterpret_cast<const float4 *>(p.h + (size_t)m * H + j + 4);
            hv[0] = h0.x; hv[1] = h0.y; hv[2] = h0.z; hv[3] = h0.w; hv[4] = h1.x; hv[5] = h1.y; hv[6] = h1.z; hv[7] = h1.w;
          }
          float psr[8], psz[8], psn[8], por[8], poz[8], pon[8];
          if (EPI == EPI_GRU_EDGE) {
            const float *ps = p.P + (size_t)s_id * 3 * H + j, *po = p.P + (size_t)o_id * 3 * H + j;
#pragma unroll
            for (int c = 0; c < 8; c += 4) {
              const float4 a0 = *reinterpret_cast<const float4 *>(ps + c), a1 = *reinterpret_cast<const float4 *>(ps + H + c),
                           a2 = *reinterpret_cast<const float4 *>(ps + 2 * H + c);
              const float4 b0 = *reinterpret_cast<const float4 *>(po + c), b1 = *reinterpret_cast<const float4 *>(po + H + c),
                           b2 = *reinterpret_cast<const float4 *>(po + 2 * H + c);
              psr[c] = a0.x; psr[c + 1] = a0.y; psr[c + 2] = a0.z; psr[c + 3] = a0.w;
              psz[c] = a1.x; psz[c + 1] = a1.y; psz[c + 2] = a1.z; psz[c + 3] = a1.w;
              psn[c] = a2.x; psn[c + 1] = a2.y; psn[c + 2] = a2.z; psn[c + 3] = a2.w;
              por[c] = b0.x; por[c + 1] = b0.y; por[c + 2] = b0.z; por[c + 3] = b0.w;
              poz[c] = b1.x; poz[c + 1] = b1.y; poz[c + 2] = b1.z; poz[c + 3] = b1.w;
              pon[c] = b2.x; pon[c + 1] = b2.y; pon[c + 2] = b2.z; pon[c + 3] = b2.w;
            }
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float bir = __ldg(p.b_ih + j + c), biz = __ldg(p.b_ih + H + j + c), bin = __ldg(p.b_ih + 2 * H + j + c);
            const float bhr = __ldg(p.b_hh + j + c), bhz = __ldg(p.b_hh + H + j + c), bhn = __ldg(p.b_hh + 2 * H + j + c);
            float *cp = p.cache ? p.cache + (size_t)m * 4 * H + j + c : nullptr;
            if (EPI == EPI_GRU_INIT) {          // acc = x W_ih^T ; h = 0 => gh = b_hh
              o8[c] = gru_point(ar[c] + bir, bhr, az[c] + biz, bhz, an[c] + bin, bhn, 0.f, cp, H);
            } else if (EPI == EPI_GRU_NODE) {   // seg 0 = ctx W_ih^T, seg 1 = V W_hh^T
              o8[c] = gru_point(ar[c] + bir, br[c] + bhr, az[c] + biz, bz[c] + bhz, an[c] + bin, bn[c] + bhn, hv[c], cp, H);
            } else {                            // EDGE: acc = Eh W_hh^T ; gi = g_s P[s] + g_o P[o] + b_ih
              const float gir = fmaf(gs, psr[c], go * por[c]) + bir;
              const float giz = fmaf(gs, psz[c], go * poz[c]) + biz;
              const float gin = fmaf(gs, psn[c], go * pon[c]) + bin;
              o8[c] = gru_point(gir, ar[c] + bhr, giz, az[c] + bhz, gin, an[c] + bhn, hv[c], cp, H);
            }
          }
          float *op = p.out + (size_t)m * H + j;
          *reinterpret_cast<float4 *>(op) = make_float4(o8[0], o8[1], o8[2], o8[3]);
          *reinterpret_cast<float4 *>(op + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// x -> (hi, lo) for weights, once per weight version.
__global__ void k_tc_split(const float *__restrict__ w, size_t n, float *__restrict__ hi, float *__restrict__ lo) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = w[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[i] = h; lo[i] = v - h;
  }
}

// split-K reducer: y = act(sum_z part[z] + bias), fixed summation order
__global__ void k_tc_splitk_reduce(const float *__restrict__ part, int splits, size_t mn, int Nout,
                                   const float *__restrict__ bias, int relu, float *__restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < mn; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(size_t)z * mn + i];
    if (bias != nullptr) s += __ldg(bias + (int)(i % Nout));
    if (relu) s = fmaxf(s, 0.f);
    y[i] = s;
  }
}

// ------------------------------- host side -------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
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

// row-major fp32 [rows, K] -> tiles of box_rows x bkf floats, swizzle span = bkf*4 bytes, zero OOB fill
static int make_tmap(CUtensorMap *m, const float *base, int rows, int K, int box_rows, int bkf) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t gstr[1] = {(cuuint64_t)K * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)bkf, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bkf == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d", (int)r, rows, K);
  return 0;
}

struct Seg { const float *A; const float *Bhi; const float *Blo; int brows; };

template <int NBLK, int NBR, int NSEG, int EPI, int BKF>
static int launch(const Params &p, const Seg *segs, int col_tiles, int splits, cudaStream_t st) {
  using C = Cfg<NBLK, NBR, BKF>;
  static bool attr = false;
  if (!attr) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_tc_gemm<NBLK, NBR, NSEG, EPI, BKF>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  CUtensorMap tm[6];
  int rc;
  for (int s = 0; s < 2; ++s) {
    const Seg &g = segs[s < NSEG ? s : 0];
    if ((rc = make_tmap(&tm[3 * s + 0], g.A, p.M, p.K, BM, BKF))) return rc;
    if ((rc = make_tmap(&tm[3 * s + 1], g.Bhi, g.brows, p.K, NBR, BKF))) return rc;
    if ((rc = make_tmap(&tm[3 * s + 2], g.Blo, g.brows, p.K, NBR, BKF))) return rc;
  }
  dim3 grid(col_tiles, (p.M + BM - 1) / BM, splits);
  k_tc_gemm<NBLK, NBR, NSEG, EPI, BKF><<<grid, NTHR, C::SMEM, st>>>(p, tm[0], tm[3], tm[1], tm[2], tm[4], tm[5]);
  SGG_RETURN_IF_LAUNCH_FAILED("k_tc_gemm");
  return 0;
}

static bool ok16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int env_flag(const char *name, int dflt) {
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
}

// split-K plan for LINEAR: enough CTAs to fill the machine, each split a whole number of 256-wide chunks
static int plan_splits(int M, int Nout, int K, int ncol, int bkf) {
  const long tiles = (long)((Nout + ncol - 1) / ncol) * ((M + BM - 1) / BM);
  const int chunks = (K + 255) / 256;
  const int sms = sgg_num_sms();
  if (tiles <= 0 || tiles >= sms || chunks < 2) return 1;
  int want = (int)((sms + tiles - 1) / tiles);
  if (want > chunks) want = chunks;
  if (want > 32) want = 32;
  (void)bkf;
  return want < 1 ? 1 : want;
}

}  // namespace tc

size_t tc32_linear_workspace_floats(int M, int Nout, int K) {
  const int ncol = Nout <= 64 ? 64 : 128;
  const int splits = tc::plan_splits(M, Nout, K, ncol, 32);
  return splits > 1 ? (size_t)splits * M * Nout : 0;
}

// y = act(x W^T + b) on tensor cores.  w_split = [hi | lo], each [Nout, K].  ws: tc_linear_workspace_floats floats
// (may be null when that is 0; if it is null although split-K would help, the kernel simply runs unsplit).
int tc32_linear(const float *x, const float *w_split, const float *b, float *y, int M, int Nout, int K, int relu,
                float *ws, cudaStream_t st) {
  if (M <= 0 || Nout <= 0) return 0;
  if ((K & 3) || !tc::ok16(x) || !tc::ok16(w_split)) return sgg_set_err(SGG_E_BADARG, "tc_linear: K %% 4 / alignment");
  static const int bk16 = tc::env_flag("SGG_TC_BK16", 0);
  const int bkf = bk16 ? 16 : 32;
  const int ncol = Nout <= 64 ? 64 : 128;
  int splits = ws ? tc::plan_splits(M, Nout, K, ncol, bkf) : 1;
  const int kblocks = (K + bkf - 1) / bkf, kcb = 256 / bkf;
  int kb_per = kblocks;
  if (splits > 1) {
    const int chunks = (kblocks + kcb - 1) / kcb;
    const int ch_per = (chunks + splits - 1) / splits;
    kb_per = ch_per * kcb;
    splits = (kblocks + kb_per - 1) / kb_per;
  }
  tc::Params p{}; p.M = M; p.K = K; p.Nout = Nout; p.relu = relu; p.bias = b; p.kb_per_split = kb_per;
  p.out = splits > 1 ? ws : y;
  tc::Seg sg[2] = {{x, w_split, w_split + (size_t)Nout * K, Nout}, {}};
  int rc;
  if (Nout <= 64) rc = bk16 ? tc::launch<1, 64, 1, tc::EPI_LINEAR, 16>(p, sg, 1, splits, st)
                            : tc::launch<1, 64, 1, tc::EPI_LINEAR, 32>(p, sg, 1, splits, st);
  else rc = bk16 ? tc::launch<2, 64, 1, tc::EPI_LINEAR, 16>(p, sg, (Nout + 127) / 128, splits, st)
                 : tc::launch<2, 64, 1, tc::EPI_LINEAR, 32>(p, sg, (Nout + 127) / 128, splits, st);
  if (rc) return rc;
  if (splits > 1) {
    const size_t mn = (size_t)M * Nout;
    int blocks = (int)((mn + 255) / 256 < 2048 ? (mn + 255) / 256 : 2048);
    tc::k_tc_splitk_reduce<<<blocks, 256, 0, st>>>(ws, splits, mn, Nout, b, relu, y);
    SGG_RETURN_IF_LAUNCH_FAILED("k_tc_splitk_reduce");
  }
  return 0;
}

// mode 0 = INIT (h = 0), 1 = NODE (x = ctx, h = V), 2 = EDGE (h = Eh, gathered gi)
int tc32_gru(int mode, const float *x, const float *h, const float *w_ih_split, const float *w_hh_split,
             const float *b_ih, const float *b_hh, const float *P, const float *gates, const int *subj, const int *obj,
           float *out, float *cache, int M, int H, cudaStream_t st) {
  if (M <= 0) return 0;
  if (H % 64) return sgg_set_err(SGG_E_BADARG, "tc_gru: H %% 64");
  static const int bk16 = tc::env_flag("SGG_TC_BK16", 0);
  tc::Params p{}; p.M = M; p.K = H; p.H = H; p.b_ih = b_ih; p.b_hh = b_hh; p.h = h; p.P = P; p.gates = gates;
  p.subj = subj; p.obj = obj; p.out = out; p.cache = cache;
  const size_t wn = (size_t)3 * H * H;
  // narrow (32 hidden units) tiles when the row count alone cannot fill the machine, and always for NODE
  // (two segments x (main + corr) accumulators must fit the 512 TMEM columns)
  const bool narrow = ((long)(H / 64) * ((M + 127) / 128) < sgg_num_sms());
  if (mode == 0) {
    tc::Seg sg[2] = {{x, w_ih_split, w_ih_split + wn, 3 * H}, {}};
    if (narrow) return bk16 ? tc::launch<3, 32, 1, tc::EPI_GRU_INIT, 16>(p, sg, H / 32, 1, st)
                            : tc::launch<3, 32, 1, tc::EPI_GRU_INIT, 32>(p, sg, H / 32, 1, st);
    return bk16 ? tc::launch<3, 64, 1, tc::EPI_GRU_INIT, 16>(p, sg, H / 64, 1, st)
                : tc::launch<3, 64, 1, tc::EPI_GRU_INIT, 32>(p, sg, H / 64, 1, st);
  } else if (mode == 1) {
    tc::Seg sg[2] = {{x, w_ih_split, w_ih_split + wn, 3 * H}, {h, w_hh_split, w_hh_split + wn, 3 * H}};
    return bk16 ? tc::launch<3, 32, 2, tc::EPI_GRU_NODE, 16>(p, sg, H / 32, 1, st)
                : tc::launch<3, 32, 2, tc::EPI_GRU_NODE, 32>(p, sg, H / 32, 1, st);
  }
  tc::Seg sg[2] = {{h, w_hh_split, w_hh_split + wn, 3 * H}, {}};
  if (narrow) return bk16 ? tc::launch<3, 32, 1, tc::EPI_GRU_EDGE, 16>(p, sg, H / 32, 1, st)
                          : tc::launch<3, 32, 1, tc::EPI_GRU_EDGE, 32>(p, sg, H / 32, 1, st);
  return bk16 ? tc::launch<3, 64, 1, tc::EPI_GRU_EDGE, 16>(p, sg, H / 64, 1, st)
              : tc::launch<3, 64, 1, tc::EPI_GRU_EDGE, 32>(p, sg, H / 64, 1, st);
}

}  // namespace sgg

namespace sgg {
int tc32_split_weights(const float *w, size_t n, float *split, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
  tc::k_tc_split<<<blocks, 256, 0, st>>>(w, n, split, split + n);
  SGG_RETURN_IF_LAUNCH_FAILED("k_tc_split");
  return 0;
}
}  // namespace sgg
