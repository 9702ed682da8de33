// Tensor-core GEMM engine v2: tcgen05 kind::f16 with a 3-pass fp16 split ("3xFP16"), fp32-grade results.
//
//   x = hi + lo * 2^-11,   hi = fp16_rn(x),   lo = fp16_rn((x - hi) * 2^11)         (x - hi is exact in fp32)
//   D_main[128 x N] = A_hi B_hi^T                      (TMEM, fp32 accumulate, K = 16 per instruction)
//   D_corr[128 x N] = A_lo B_hi^T + A_hi B_lo^T        (separate accumulator: it carries the 2^11 scale)
//   result          = D_main + 2^-11 D_corr            (fp32 registers, epilogue)
//
// fp16 x fp16 products are exact in fp32; the dropped lo*lo term is 2^-24 relative, i.e. below fp32 rounding.
// Versus the 3xTF32 engine (tc_gemm.cu) every pass runs at twice the MMA rate and every operand element costs
// 4 bytes (hi + lo) instead of 8 on the L2 -> shared-memory path, which is what bounds these kernels
// (profiles/r01_ncu_edge_gru_tc.md).  Range: |x| must stay below 65504 (fp16); smaller magnitudes lose nothing
// (hi may be subnormal, lo picks up the remainder: absolute error <= 2^-35).
//
// One CTA = one 128-row x NCOL-column output tile, 10 warps:
//   warp 0   : TMA producer.  Per 64-wide k-block: two [128 x 32] fp32 boxes of A (raw) and the pre-split fp16
//              B_hi / B_lo tiles (sgg_tc_split_weights), all SWIZZLE_128B, K-major, multi-stage mbarrier ring.
//   warps 2-5: convert the raw A tile IN PLACE into the fp16 [hi | lo] tiles the tensor core reads (the 32 KB of
//              fp32 become 16 KB hi + 16 KB lo; each 8-row swizzle atom is converted by one warp with a
//              load-all / __syncwarp / store-all sequence, so no extra staging buffer is needed).
//   warp 1   : one thread issues the tcgen05.mma triplets and commits stage / accumulator barriers.
//   warps 6-9: TMEM readers (one accumulator row per thread).
// Epilogues:
//   LINEAR  : K is folded into TMEM 256 at a time (the tensor core's accumulator truncates on every add, see
//             tc_gemm.cu); chunks ping-pong between two TMEM buffer pairs and are drained into fp32 registers.
//   GRU_*   : accumulators are dumped to shared memory (the pipeline stages are free by then) and ALL 8 worker
//             warps run the GRUCell pointwise math with a (row, 4 hidden units)-per-thread mapping, so the gathers
//             of P = V W_ih^T rows, the state read and the state write are fully coalesced.
// GRU tiles cover hidden units [j0, j0 + NBR) of the r, z, n gates (weight rows j0, H + j0, 2H + j0); NBR is
// picked per launch from {32, 64, 80} by a wave-quantisation cost model (a ragged last column tile is masked).
#include <stdlib.h>
#include "tc16_common.cuh"

namespace sgg {
namespace tc16 {

__device__ unsigned int g_tc16_overflow = 0;

enum { EPI_LINEAR = 0, EPI_GRU_INIT = 1, EPI_GRU_NODE = 2, EPI_GRU_EDGE = 3 };

struct Params {
  int M, K, H, Nout, relu;
  int kb_per_split;           // LINEAR split-K: k-blocks per blockIdx.z
  // LINEAR stream-K (sk_W > 0): the (tile, 256-wide K chunk) space of sk_W chunks is cut into gridDim.x equal
  // contiguous ranges; a CTA writes one [BM x NCOL] partial per tile it touches into sk_part (slot = cta * sk_maxseg
  // + segment) and k_tc16_streamk_reduce sums them in CTA order.  sk_C = chunks per tile, sk_ct = column tiles.
  int sk_W, sk_C, sk_ct, sk_maxseg;
  float *sk_part;
  int pf;                     // LINEAR: k-blocks of L2 prefetch distance for the streamed A operand (0 = off)
  const float *a_direct;      // LINEAR, DIRECT: the fp32 A operand [M,K] (read by the converter warps, not by TMA)
  const float *bias;          // LINEAR
  float *out;                 // LINEAR: [M,Nout] (or partials [splits,M,Nout]); GRU: [M,H]
  const float *b_ih, *b_hh;   // GRU
  const float *h;             // GRU_NODE / GRU_EDGE: previous state [M,H]
  const float *P;             // GRU_EDGE: [N,3H]
  const float *gates;         // GRU_EDGE: [M,4]
  const int *subj, *obj;      // GRU_EDGE
  float *cache;               // GRU: nullable [M,4,H] (r, z, n, gh_n) for the backward pass
  const float *out_scale;     // LINEAR: nullable device scalar multiplied into the result before bias / ReLU (backward GEMMs)
  const float *in_scale;      // LINEAR: nullable device scalar multiplied into the A operand before the fp16 split (power of two: exact)
  __half *out_hi, *out_lo;    // LINEAR: nullable fp16 planes [hi | lo * 2^11] of the final output (consumed via TMA by mp_fused.cu)
  long long *dbg;             // nullable: per-CTA phase timestamps (SGG_TC_TIMING=1, tools/tc16_phases.py)
};

constexpr int DBG_SLOTS = 8, DBG_MAX_CTAS = 1024;
__device__ long long g_dbg[DBG_SLOTS * DBG_MAX_CTAS];
#define SGG_DBG(slot)                                                                                   \
  do {                                                                                                  \
    if (p.dbg != nullptr) {                                                                             \
      const int cta_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);                  \
      if (cta_ < DBG_MAX_CTAS) p.dbg[cta_ * DBG_SLOTS + (slot)] = clock64();                            \
    }                                                                                                   \
  } while (0)

template <int NCOL>
struct Cfg {
  static constexpr int B_BYTES = NCOL * BK * 2;                 // one of hi / lo
  static constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 6 ? 6 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
  static constexpr int KCB = 256 / BK;                          // k-blocks per accumulation chunk (LINEAR)
};

// NBLK gate blocks of NBR weight rows each (LINEAR: NBLK = 1, NBR = tile width); NSEG K-segments (NODE: 2)
// DIRECT (LINEAR only): the converter warps load the fp32 A tile straight from global memory into registers (two k-blocks
// ahead), split it and write the fp16 [hi | lo] tiles — A never exists as fp32 in shared memory.  The main loop is bound by
// shared-memory bandwidth; this removes the 32 KB TMA write and the 32 KB read-back of the raw tile
// from every k-block (224 -> 160 KB at 128 columns).
template <int NBLK, int NBR, int NSEG, int EPI, int DIRECT = 0>
__global__ void __launch_bounds__(NTHR, 1)
k_tc16(Params p, const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
       const __grid_constant__ CUtensorMap tmBh0, const __grid_constant__ CUtensorMap tmBl0,
       const __grid_constant__ CUtensorMap tmBh1, const __grid_constant__ CUtensorMap tmBl1) {
  using namespace tc;
  constexpr int NCOL = NBLK * NBR;                 // MMA N
  using C = Cfg<NCOL>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, B_BYTES = C::B_BYTES, KCB = C::KCB;
  constexpr bool CHUNKED = (EPI == EPI_LINEAR);
  constexpr int NCONV = 8;                         // converter warps = all worker warps (2..9)
  static_assert(!CHUNKED || (NSEG == 1 && NBLK == 1 && NCOL <= 128 && NCOL % 32 == 0), "chunked LINEAR: NCOL/2 register accumulators per thread");
  static_assert(NCOL % 16 == 0 && NCOL <= 256 && NBR % 16 == 0, "UMMA N / TMEM load granularity");
  static_assert(STAGES >= 2, "pipeline needs two stages");
  // TMEM columns: GRU: [main seg0, main seg1 | corr seg0, corr seg1]; LINEAR: buffer b = [main_b | corr_b]
  constexpr int CORR = CHUNKED ? NCOL : NSEG * NCOL;
  constexpr int ACC_COLS = CHUNKED ? 4 * NCOL : 2 * NSEG * NCOL;
  constexpr uint32_t TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static_assert(ACC_COLS <= 512, "tile too wide for TMEM");
  // GRU epilogue staging: NV value blocks of NBR floats per row, row stride RS floats (RS/4 odd: conflict-free)
  constexpr int NV = (EPI == EPI_GRU_NODE) ? 4 : 3;
  constexpr int RS = NV * NBR + 4;
  static_assert(CHUNKED || (BM * RS * 4 <= STAGES * STAGE_BYTES), "epilogue staging must fit in the stage ring");
  static_assert(CHUNKED || ((RS / 4) & 1) == 1, "staging row stride");
  constexpr int META_OFF = BM * RS * 4;            // [BM] int4 (subj, obj, g_sub, g_obj) after the staging tile
  static_assert(CHUNKED || (META_OFF + BM * 16 <= STAGES * STAGE_BYTES), "row metadata must fit in the stage ring");

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
  uint64_t *full = bars, *ready = bars + STAGES, *empty = bars + 2 * STAGES, *tmem_full = bars + 3 * STAGES;
  uint64_t *tmem_empty = bars + 3 * STAGES + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool SK = CHUNKED && p.sk_W > 0;          // stream-K LINEAR
  const int m0 = blockIdx.y * BM;
  const int j0 = blockIdx.x * (CHUNKED ? NCOL : NBR);
  const int kblocks_all = (p.K + BK - 1) / BK;
  const int kb_lo = CHUNKED ? (int)blockIdx.z * p.kb_per_split : 0;
  const int kblocks = CHUNKED ? min(p.kb_per_split, kblocks_all - kb_lo) : kblocks_all;
  // stream-K: this CTA's range of global chunks [gc_lo, gc_hi); k-block `it` of the CTA is global k-block gc_lo*KCB + it
  const int gc_lo = SK ? (int)((long long)blockIdx.x * p.sk_W / gridDim.x) : 0;
  const int gc_hi = SK ? (int)((long long)(blockIdx.x + 1) * p.sk_W / gridDim.x) : 0;
  const int total = SK ? (gc_hi - gc_lo) * KCB : NSEG * kblocks;

  if (warp == 0 && lane == 0) {
    SGG_DBG(0);
    prefetch_tmap(&tmA0); prefetch_tmap(&tmBh0); prefetch_tmap(&tmBl0);
    if (NSEG > 1) { prefetch_tmap(&tmA1); prefetch_tmap(&tmBh1); prefetch_tmap(&tmBl1); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(ready + s, NCONV * 32); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1); mbar_init(tmem_full + 1, 1);
    mbar_init(tmem_empty, 256); mbar_init(tmem_empty + 1, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) SGG_DBG(1);

  auto stage_ptr = [&](int s) { return smem + (size_t)s * STAGE_BYTES; };   // [A (raw -> hi|lo) | B_hi | B_lo]

  if (warp == 0) {
    // ===================== TMA producer =====================
    {                                          // whole warp runs the loop, one elected lane issues (uniform operands)
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(empty + s, ph ^ 1);
        const int seg = SK ? 0 : it / kblocks;
        int k0 = (kb_lo + it - seg * kblocks) * BK, am0 = m0, bj0 = j0;
        if (SK) {                                  // global k-block -> (tile, k-block inside the tile)
          const int gk = gc_lo * KCB + it, kbt = p.sk_C * KCB;
          const int t = gk / kbt;
          k0 = (gk - t * kbt) * BK; am0 = (t / p.sk_ct) * BM; bj0 = (t % p.sk_ct) * NCOL;
        }
        uint8_t *st = stage_ptr(s);
        if (elect_one()) {
        mbar_arrive_expect_tx(full + s, (DIRECT ? 0 : A_BYTES) + 2 * B_BYTES);
        const CUtensorMap *ta = (NSEG > 1 && seg == 1) ? &tmA1 : &tmA0;
        const CUtensorMap *tbh = (NSEG > 1 && seg == 1) ? &tmBh1 : &tmBh0;
        const CUtensorMap *tbl = (NSEG > 1 && seg == 1) ? &tmBl1 : &tmBl0;
        if (!DIRECT) {
          tma_load_2d(st, ta, full + s, k0, am0);                      // fp32 k0 .. k0+31
          tma_load_2d(st + A_HALF, ta, full + s, k0 + 32, am0);        // fp32 k0+32 .. k0+63
          if (CHUNKED && !SK && p.pf > 0 && it + p.pf < total) {       // A comes from HBM (16 KB / row): pull it into L2 early
            tma_prefetch_2d(ta, k0 + p.pf * BK, am0);
            tma_prefetch_2d(ta, k0 + p.pf * BK + 32, am0);
          }
        }
#pragma unroll
        for (int b = 0; b < NBLK; ++b) {
          const int row = CHUNKED ? bj0 : (b * p.H + j0);
          tma_load_2d(st + A_BYTES + b * (NBR * BK * 2), tbh, full + s, k0, row);
          tma_load_2d(st + A_BYTES + B_BYTES + b * (NBR * BK * 2), tbl, full + s, k0, row);
        }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // whole warp runs the loop, one ELECTED lane issues: keeps the descriptors in uniform registers (under `if (lane == 0)`
    // every UTCHMMA is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop)
    {
      constexpr uint32_t idesc = make_idesc_f16(BM, NCOL);
      constexpr uint32_t idesc2 = make_idesc_f16(BM, CHUNKED ? 2 * NCOL : NCOL);
      (void)idesc2;
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(full + s, ph);
        mbar_wait(ready + s, ph);
        fence_after_sync();
        if (it == 0 && lane == 0) SGG_DBG(2);
        const int seg = CHUNKED ? 0 : it / kblocks, kb = it - seg * kblocks;
        uint8_t *st = stage_ptr(s);
        const uint64_t ah = make_sdesc128(st), al = make_sdesc128(st + A_HALF);
        const uint64_t bh = make_sdesc128(st + A_BYTES), bl = make_sdesc128(st + A_BYTES + B_BYTES);
        uint32_t dm, first;
        const int chunk = it / KCB, kc = it - chunk * KCB;
        if (CHUNKED) {
          if (kc == 0) {                       // buffer pair must have been drained by the TMEM readers
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
        if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);     // 16 fp16 = 32 bytes >> 4
          const uint32_t acc = (first && kk == 0) ? 0u : 1u;
          if constexpr (CHUNKED) {
            // LINEAR: the B_lo tile follows the B_hi tile and CORR follows MAIN, so ONE N = 2 NCOL instruction forms
            // [main | a_hi b_lo]; an SS MMA costs >= ~102 cycles whatever its N (tools/ubench/umma_rate.cu), so two
            // instructions per K16 step instead of three is 1.2x (N = 128) to 1.6x (N = 96) fewer tensor-pipe cycles
            mma_f16_ss(dm, ah + o, bh + o, idesc2, acc);
            mma_f16_ss(dc, al + o, bh + o, idesc, 1u);
          } else {
            mma_f16_ss(dc, al + o, bh + o, idesc, acc);      // corrections -> CORR
            mma_f16_ss(dc, ah + o, bl + o, idesc, 1u);
            mma_f16_ss(dm, ah + o, bh + o, idesc, acc);      // large term  -> MAIN
          }
        }
        mma_commit(empty + s);
        if (CHUNKED && (kc == KCB - 1 || it == total - 1)) mma_commit(tmem_full + (chunk & 1));
        if (!CHUNKED && it == total - 1) mma_commit(tmem_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== warps 2..9: convert A (fp32 -> fp16 [hi | lo], in place); LINEAR: + chunk drains =====================
    // Swizzle atom g = rows 8g..8g+7: its raw floats live at [g*1024, +1024) of both 16 KB boxes, exactly where its
    // hi (first box) and lo (second box) fp16 rows go.  Lane <-> (row r, 16-byte output chunk c): k = 8c .. 8c+7.
    // The conversion is instruction-bound and every k-block passes through it, so all eight worker warps take part
    // (two atoms each) and use packed f16x2 conversions.
    constexpr int APW = 16 / NCONV;                        // swizzle atoms per warp and k-block
    const int wc = warp - 2;
    const int c = lane & 7, b = c >> 2;                    // output chunk; source box
    const int ca = 2 * (c & 3);                            // first raw chunk (logical) inside the box
    uint32_t ovf = 0;                                      // fp16 range guard (sticky flag raised after the main loop)
    const float isc = (CHUNKED && p.in_scale != nullptr) ? __ldg(p.in_scale) : 1.0f;
    auto convert = [&](int it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(full + s, ph);
      const uint32_t a_addr = smem_u32(stage_ptr(s));
      uint4 v[4 * APW];
#pragma unroll
      for (int i = 0; i < 2 * APW; ++i) {                  // i = 2*atom + row half
        const int g = APW * wc + (i >> 1), r = (lane >> 3) + 4 * (i & 1);
        const uint32_t rowb = a_addr + (uint32_t)(b * A_HALF + g * 1024 + r * 128);
        // bank spread: lanes reading box 1 fetch the odd chunk first
        v[2 * i] = lds128u(rowb + (uint32_t)((((ca + b) ^ r) & 7) << 4));
        v[2 * i + 1] = lds128u(rowb + (uint32_t)((((ca + 1 - b) ^ r) & 7) << 4));
      }
      __syncwarp();                                        // every lane's reads of these atoms are done
#pragma unroll
      for (int i = 0; i < 2 * APW; ++i) {
        const int g = APW * wc + (i >> 1), r = (lane >> 3) + 4 * (i & 1);
        const uint4 f0 = b ? v[2 * i + 1] : v[2 * i], f1 = b ? v[2 * i] : v[2 * i + 1];   // floats 0-3, 4-7 of the chunk
        uint4 hi, lo;
        split2(__uint_as_float(f0.x) * isc, __uint_as_float(f0.y) * isc, hi.x, lo.x);
        split2(__uint_as_float(f0.z) * isc, __uint_as_float(f0.w) * isc, hi.y, lo.y);
        split2(__uint_as_float(f1.x) * isc, __uint_as_float(f1.y) * isc, hi.z, lo.z);
        split2(__uint_as_float(f1.z) * isc, __uint_as_float(f1.w) * isc, hi.w, lo.w);
        ovf |= f16x2_nonfinite(hi.x) | f16x2_nonfinite(hi.y) | f16x2_nonfinite(hi.z) | f16x2_nonfinite(hi.w);
        const uint32_t dst = a_addr + (uint32_t)(g * 1024 + r * 128 + (((c ^ r) & 7) << 4));
        sts128u(dst, hi);
        sts128u(dst + A_HALF, lo);
      }
      fence_proxy_async_smem();
      mbar_arrive(ready + s);
    };
    // ---- DIRECT: global -> registers -> fp16 tiles.  Lane <-> (row r of an 8-row swizzle atom, 16-byte output chunk c);
    // a warp's load instruction covers 4 rows x 256 contiguous bytes.  PRE[j] holds the 8 floats of item j.
    auto a_coords = [&](int it, int &k0, int &am0) {
      k0 = (kb_lo + it) * BK; am0 = m0;
      if (SK) {
        const int gk = gc_lo * KCB + it, kbt = p.sk_C * KCB;
        const int t = gk / kbt;
        k0 = (gk - t * kbt) * BK; am0 = (t / p.sk_ct) * BM;
      }
    };
    auto load_direct = [&](int it, float4 (&pre)[4 * APW]) {
      int k0, am0;
      a_coords(it, k0, am0);
      const int k = k0 + 8 * c;
      const bool kin = k + 8 <= p.K;                       // K % 8 == 0: a chunk is entirely inside or outside
#pragma unroll
      for (int i = 0; i < 2 * APW; ++i) {
        const int g = APW * wc + (i >> 1), r = (lane >> 3) + 4 * (i & 1);
        const int m = am0 + 8 * g + r;
        if (kin && m < p.M) {
          const float4 *src = reinterpret_cast<const float4 *>(p.a_direct + (size_t)m * p.K + k);
          pre[2 * i] = __ldg(src); pre[2 * i + 1] = __ldg(src + 1);
        } else {
          pre[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f); pre[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    auto store_direct = [&](int it, const float4 (&pre)[4 * APW]) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(empty + s, ph ^ 1);                        // the MMAs that read this stage's previous tile have retired
      const uint32_t a_addr = smem_u32(stage_ptr(s));
#pragma unroll
      for (int i = 0; i < 2 * APW; ++i) {
        const int g = APW * wc + (i >> 1), r = (lane >> 3) + 4 * (i & 1);
        const float4 f0 = pre[2 * i], f1 = pre[2 * i + 1];
        uint4 hi, lo;
        split2(f0.x * isc, f0.y * isc, hi.x, lo.x); split2(f0.z * isc, f0.w * isc, hi.y, lo.y);
        split2(f1.x * isc, f1.y * isc, hi.z, lo.z); split2(f1.z * isc, f1.w * isc, hi.w, lo.w);
        ovf |= f16x2_nonfinite(hi.x) | f16x2_nonfinite(hi.y) | f16x2_nonfinite(hi.z) | f16x2_nonfinite(hi.w);
        const uint32_t dst = a_addr + (uint32_t)(g * 1024 + r * 128 + (((c ^ r) & 7) << 4));
        sts128u(dst, hi);
        sts128u(dst + A_HALF, lo);
      }
      fence_proxy_async_smem();
      mbar_arrive(ready + s);
    };
   if (CHUNKED) {
    // ===================== LINEAR: thread <-> (output row, half of the tile's columns) =====================
    // Chunk ch (256 of K) is drained from TMEM into fp32 registers one k-block AFTER its last k-block was converted,
    // when its MMAs have (nearly) finished, so the drains do not stall the conversion stream; the MMA issuer needs
    // the buffer back only a whole chunk later.
    constexpr int HC = NCOL / 2;
    const int half = warp >= 6 ? 1 : 0;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HC);
    float acc[HC];
#pragma unroll
    for (int cc = 0; cc < HC; ++cc) acc[cc] = 0.f;
    auto drain = [&](int ch) {
      const int bsel = ch & 1;
      mbar_wait(tmem_full + bsel, (ch >> 1) & 1);
      fence_after_sync();
      __syncwarp();
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 16) {
        float v[16], w[16];
        tmem_ld16(taddr + (uint32_t)(bsel * 2 * NCOL + c0), v);
        tmem_ld16(taddr + (uint32_t)(bsel * 2 * NCOL + NCOL + c0), w);
        tmem_wait_ld();
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) acc[c0 + cc] += fmaf(w[cc], LO_INV, v[cc]);
      }
      fence_before_sync();
      mbar_arrive(tmem_empty + bsel);
    };
    const int nchunks = (total + KCB - 1) / KCB;
    int sk_seg = 0;
    // stream-K: after draining local chunk ch, flush the partial tile if that chunk closed its tile (or the CTA's range)
    auto sk_flush = [&](int ch) {
      if (!SK) return;
      if (((gc_lo + ch + 1) % p.sk_C) != 0 && ch != nchunks - 1) return;
      float *prow = p.sk_part + ((size_t)((size_t)blockIdx.x * p.sk_maxseg + sk_seg) * BM + row) * NCOL + half * HC;
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 4) {
        *reinterpret_cast<float4 *>(prow + c0) = make_float4(acc[c0], acc[c0 + 1], acc[c0 + 2], acc[c0 + 3]);
        acc[c0] = 0.f; acc[c0 + 1] = 0.f; acc[c0 + 2] = 0.f; acc[c0 + 3] = 0.f;
      }
      ++sk_seg;
    };
    // Chunk ch is drained while the LAST k-block of chunk ch+1 is being consumed (not right after its own last k-block:
    // the converters run up to STAGES k-blocks ahead of the MMA issuer, so an early drain made them wait for the tensor
    // pipe and then starved it — the ring emptied once per chunk).  The MMA issuer needs the buffer back one k-block later.
    int next_drain = 0;
    if (DIRECT) {
      float4 pre0[4 * APW], pre1[4 * APW];               // two k-blocks of A in flight per thread
      if (total > 0) load_direct(0, pre0);
      if (total > 1) load_direct(1, pre1);
      for (int it = 0; it < total; it += 2) {
        store_direct(it, pre0);
        if (it + 2 < total) load_direct(it + 2, pre0);
        if ((it % KCB) == KCB - 1 && it / KCB >= 1) { drain(next_drain); sk_flush(next_drain); ++next_drain; }
        if (it + 1 < total) {
          store_direct(it + 1, pre1);
          if (it + 3 < total) load_direct(it + 3, pre1);
          if (((it + 1) % KCB) == KCB - 1 && (it + 1) / KCB >= 1) { drain(next_drain); sk_flush(next_drain); ++next_drain; }
        }
      }
    } else {
      for (int it = 0; it < total; ++it) {
        convert(it);
        if ((it % KCB) == KCB - 1 && it / KCB >= 1) { drain(next_drain); sk_flush(next_drain); ++next_drain; }
      }
    }
    for (; next_drain < nchunks; ++next_drain) { drain(next_drain); sk_flush(next_drain); }
    if (!SK && m < p.M) {
      const bool partial = gridDim.z > 1;        // split-K: raw partial sums, bias/ReLU applied by the reducer
      const bool vec = (p.Nout & 3) == 0;
      float *yrow = p.out + ((size_t)blockIdx.z * p.M + m) * p.Nout;
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 4) {
        const int j = j0 + half * HC + c0;
        if (j < p.Nout) {
          float v[4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            v[cc] = acc[c0 + cc];
            if (!partial) {
              if (p.out_scale != nullptr) v[cc] *= __ldg(p.out_scale);
              if (p.bias != nullptr && j + cc < p.Nout) v[cc] += __ldg(p.bias + j + cc);
              if (p.relu) v[cc] = fmaxf(v[cc], 0.f);
            }
          }
          if (vec && j + 4 <= p.Nout) {
            *reinterpret_cast<float4 *>(yrow + j) = make_float4(v[0], v[1], v[2], v[3]);
            if (!partial && p.out_hi != nullptr) {
              uint2 hi, lo;
              split2(v[0], v[1], hi.x, lo.x);
              split2(v[2], v[3], hi.y, lo.y);
              if (f16x2_nonfinite(hi.x) | f16x2_nonfinite(hi.y)) atomicOr(&g_tc16_overflow, 2u);
              *reinterpret_cast<uint2 *>(p.out_hi + (size_t)m * p.Nout + j) = hi;
              *reinterpret_cast<uint2 *>(p.out_lo + (size_t)m * p.Nout + j) = lo;
            }
          } else {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) if (j + cc < p.Nout) yrow[j + cc] = v[cc];
          }
        }
      }
    }
   } else {
    for (int it = 0; it < total; ++it) convert(it);
   }
   if (ovf) atomicOr(&g_tc16_overflow, 1u);
   if (!CHUNKED && warp >= 6) {
    // ===================== GRU phase 1: TMEM -> shared staging, thread <-> accumulator row =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float *stg = reinterpret_cast<float *>(smem) + (size_t)row * RS;
    // EDGE: this row's endpoints and scalar gates, fetched while the main loop runs, parked next to the staging tile
    int4 meta = make_int4(0, 0, 0, 0);
    if (EPI == EPI_GRU_EDGE && m0 + row < p.M) {
      const float4 g4 = ld4(p.gates + (size_t)(m0 + row) * 4);
      meta = make_int4(__ldg(p.subj + m0 + row), __ldg(p.obj + m0 + row), __float_as_int(g4.x), __float_as_int(g4.y));
    }
    mbar_wait(tmem_full, 0);                      // all MMAs done => all stages consumed, ring is free
    fence_after_sync();
    __syncwarp();
    if (EPI == EPI_GRU_EDGE) reinterpret_cast<int4 *>(smem + META_OFF)[row] = meta;
    if (threadIdx.x == 192) SGG_DBG(3);
#pragma unroll
    for (int blk = 0; blk < NV; ++blk) {
      // source accumulator columns of this value block
      const int src = (EPI == EPI_GRU_NODE && blk == 3) ? (NCOL + 2 * NBR) : blk * NBR;   // NODE blk 3 = gh_n (seg 1)
      const bool add_seg1 = (EPI == EPI_GRU_NODE && blk < 2);                             // r, z: gi + gh
#pragma unroll
      for (int c0 = 0; c0 < NBR; c0 += 16) {
        float v[16], w[16];
        tmem_ld16(taddr + (uint32_t)(src + c0), v);
        tmem_ld16(taddr + (uint32_t)(CORR + src + c0), w);
        tmem_wait_ld();
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) v[cc] = fmaf(w[cc], LO_INV, v[cc]);
        if (add_seg1) {
          float v1[16], w1[16];
          tmem_ld16(taddr + (uint32_t)(NCOL + src + c0), v1);
          tmem_ld16(taddr + (uint32_t)(CORR + NCOL + src + c0), w1);
          tmem_wait_ld();
#pragma unroll
          for (int cc = 0; cc < 16; ++cc) v[cc] += fmaf(w1[cc], LO_INV, v1[cc]);
        }
#pragma unroll
        for (int cc = 0; cc < 16; cc += 4)
          *reinterpret_cast<float4 *>(stg + blk * NBR + c0 + cc) = make_float4(v[cc], v[cc + 1], v[cc + 2], v[cc + 3]);
      }
    }
    if (threadIdx.x == 192) SGG_DBG(4);
   }
  }

  if (!CHUNKED) {
    // ===================== GRU phase 2: pointwise, thread <-> (4 hidden units, every RPI-th row), coalesced =====================
    // All 10 warps take part (the TMA / MMA warps are done by now).
    // A thread keeps its hidden-unit quad for all its rows (biases live in registers); rows are processed U at a
    // time with every load of the batch issued before the first use (branch-free: out-of-range rows are clamped
    // for the loads and only their stores are predicated), so the P-row gathers overlap instead of chaining.
    __syncthreads();                              // staging complete (warps 6-9); every other warp is idle by now
    if (threadIdx.x == 64) SGG_DBG(5);
    const int H = p.H;
    constexpr int QPR = NBR / 4;                  // float4 groups per row
    constexpr int RPI = NTHR / QPR;               // rows per sweep of the CTA
    constexpr int U = 2;
    const int t2 = (int)threadIdx.x;
    const int qd = t2 % QPR, r0 = t2 / QPR;
    const int j = j0 + 4 * qd;
    const float *stg0 = reinterpret_cast<const float *>(smem);
    const int4 *meta = reinterpret_cast<const int4 *>(smem + META_OFF);
    if (r0 < RPI && j < H) {
      const float4 bir = ldg4(p.b_ih + j), biz = ldg4(p.b_ih + H + j), bin = ldg4(p.b_ih + 2 * H + j);
      const float4 bhr = ldg4(p.b_hh + j), bhz = ldg4(p.b_hh + H + j), bhn = ldg4(p.b_hh + 2 * H + j);
      const int mlast = p.M - 1;
      for (int rb = r0; rb < BM; rb += U * RPI) {
        float4 a0[U], a1[U], a2[U], a3[U], hv[U], sr[U], sz[U], sn[U], orr[U], oz[U], on[U];
        float gs[U], go[U];
        int mm[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {             // ---- issue every load of the batch
          const int row = min(rb + u * RPI, BM - 1);
          mm[u] = m0 + row;
          const int mc = min(mm[u], mlast);       // clamped row for the loads
          const float *sp = stg0 + (size_t)row * RS + 4 * qd;
          a0[u] = ld4(sp); a1[u] = ld4(sp + NBR); a2[u] = ld4(sp + 2 * NBR);
          if (EPI == EPI_GRU_NODE) a3[u] = ld4(sp + 3 * NBR);
          if (EPI != EPI_GRU_INIT) hv[u] = ld4(p.h + (size_t)mc * H + j);
          else hv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (EPI == EPI_GRU_EDGE) {
            const int4 mt = meta[row];
            gs[u] = __int_as_float(mt.z); go[u] = __int_as_float(mt.w);
            const float *ps = p.P + (size_t)mt.x * 3 * H + j, *po = p.P + (size_t)mt.y * 3 * H + j;
            sr[u] = ld4(ps); sz[u] = ld4(ps + H); sn[u] = ld4(ps + 2 * H);
            orr[u] = ld4(po); oz[u] = ld4(po + H); on[u] = ld4(po + 2 * H);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {             // ---- GRUCell pointwise + stores
          Gru4 o;
          float4 ghn;                             // gh_n incl. bias (saved for the backward pass)
          if (EPI == EPI_GRU_INIT) {              // acc = x W_ih^T ; h = 0 => gh = b_hh
            ghn = bhn;
            o = gru4(add4(a0[u], bir), bhr, add4(a1[u], biz), bhz, add4(a2[u], bin), ghn, hv[u]);
          } else if (EPI == EPI_GRU_NODE) {       // a0 = gi_r + gh_r, a1 = gi_z + gh_z, a2 = gi_n, a3 = gh_n
            ghn = add4(a3[u], bhn);
            o = gru4(add4(a0[u], bir), bhr, add4(a1[u], biz), bhz, add4(a2[u], bin), ghn, hv[u]);
          } else {                                // EDGE: acc = Eh W_hh^T ; gi = g_s P[s] + g_o P[o] + b_ih
            float4 gir, giz, gin;
            gir.x = fmaf(gs[u], sr[u].x, go[u] * orr[u].x) + bir.x; gir.y = fmaf(gs[u], sr[u].y, go[u] * orr[u].y) + bir.y;
            gir.z = fmaf(gs[u], sr[u].z, go[u] * orr[u].z) + bir.z; gir.w = fmaf(gs[u], sr[u].w, go[u] * orr[u].w) + bir.w;
            giz.x = fmaf(gs[u], sz[u].x, go[u] * oz[u].x) + biz.x; giz.y = fmaf(gs[u], sz[u].y, go[u] * oz[u].y) + biz.y;
            giz.z = fmaf(gs[u], sz[u].z, go[u] * oz[u].z) + biz.z; giz.w = fmaf(gs[u], sz[u].w, go[u] * oz[u].w) + biz.w;
            gin.x = fmaf(gs[u], sn[u].x, go[u] * on[u].x) + bin.x; gin.y = fmaf(gs[u], sn[u].y, go[u] * on[u].y) + bin.y;
            gin.z = fmaf(gs[u], sn[u].z, go[u] * on[u].z) + bin.z; gin.w = fmaf(gs[u], sn[u].w, go[u] * on[u].w) + bin.w;
            ghn = add4(a2[u], bhn);
            o = gru4(gir, add4(a0[u], bhr), giz, add4(a1[u], bhz), gin, ghn, hv[u]);
          }
          if (rb + u * RPI < BM && mm[u] < p.M) {
            *reinterpret_cast<float4 *>(p.out + (size_t)mm[u] * H + j) = o.out;
            if (p.cache != nullptr) {
              float *cp = p.cache + (size_t)mm[u] * 4 * H + j;
              *reinterpret_cast<float4 *>(cp) = o.r;
              *reinterpret_cast<float4 *>(cp + H) = o.z;
              *reinterpret_cast<float4 *>(cp + 2 * H) = o.n;
              *reinterpret_cast<float4 *>(cp + 3 * H) = ghn;
            }
          }
        }
      }
    }
  }
  if (!CHUNKED && threadIdx.x == 64) SGG_DBG(6);
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) SGG_DBG(7);
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// w -> fp16 (hi, scaled lo), once per weight version.  Streaming: 8 floats per thread and trip (two 16-byte loads,
// one 16-byte store per half-plane); n % 8 == 0 is checked at the ABI.
__global__ void __launch_bounds__(256) k_tc16_split(const float *__restrict__ w, size_t n, __half *__restrict__ hi,
                                                     __half *__restrict__ lo) {
  const size_t n8 = n >> 3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4 *>(w) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4 *>(w) + 2 * i + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half h0 = __float2half_rn(v[2 * k]), h1 = __float2half_rn(v[2 * k + 1]);
      const __half l0 = __float2half_rn((v[2 * k] - __half2float(h0)) * LO_SCALE);
      const __half l1 = __float2half_rn((v[2 * k + 1] - __half2float(h1)) * LO_SCALE);
      ph[k] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      pl[k] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    if (f16x2_nonfinite(ph[0]) | f16x2_nonfinite(ph[1]) | f16x2_nonfinite(ph[2]) | f16x2_nonfinite(ph[3]))
      atomicOr(&g_tc16_overflow, 4u);
    reinterpret_cast<uint4 *>(hi)[i] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    reinterpret_cast<uint4 *>(lo)[i] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// split-K reducer: y = act(sum_z part[z] + bias), fixed summation order
__global__ void k_tc16_splitk_reduce(const float *__restrict__ part, int splits, size_t mn, int Nout,
                                     const float *__restrict__ bias, int relu, float *__restrict__ y,
                                     __half *__restrict__ y_hi, __half *__restrict__ y_lo, const float *__restrict__ out_scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < mn; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(size_t)z * mn + i];
    if (out_scale != nullptr) s *= __ldg(out_scale);
    if (bias != nullptr) s += __ldg(bias + (int)(i % Nout));
    if (relu) s = fmaxf(s, 0.f);
    y[i] = s;
    if (y_hi != nullptr) {
      const __half h = __float2half_rn(s);
      y_hi[i] = h;
      y_lo[i] = __float2half_rn((s - __half2float(h)) * LO_SCALE);
    }
  }
}

// stream-K reducer: grid (tiles, BM / rows-per-CTA), one float4 of one output row per thread; sums the partial tiles
// of the CTAs whose chunk ranges overlap the tile in ascending CTA order (deterministic), then bias / ReLU.
// Partition arithmetic mirrors k_tc16 (lo_c = c*W/G).
__global__ void k_tc16_streamk_reduce(const float *__restrict__ part, int W, int C, int G, int maxseg, int col_tiles,
                                      int ncol, int M, int Nout, const float *__restrict__ bias, int relu,
                                      float *__restrict__ y, __half *__restrict__ y_hi, __half *__restrict__ y_lo,
                                      const float *__restrict__ out_scale) {
  const int t = blockIdx.x;
  const long long a = (long long)t * C, b = a + C;
  const int c_first = (int)(((a + 1) * G + W - 1) / W) - 1;
  const int c_last = (int)((b * G + W - 1) / W) - 1;
  const int q4 = ncol / 4;
  const int row = blockIdx.y * (blockDim.x / q4) + threadIdx.x / q4, c4 = (threadIdx.x % q4) * 4;
  const int m = (t / col_tiles) * BM + row, j = (t % col_tiles) * ncol + c4;
  if (row >= BM || m >= M || j >= Nout) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = c_first; c <= c_last; ++c) {
    const int lo = (int)((long long)c * W / G);
    const int seg = t - lo / C;
    const float4 v = *reinterpret_cast<const float4 *>(part + ((size_t)((size_t)c * maxseg + seg) * BM + row) * ncol + c4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float o[4] = {acc.x, acc.y, acc.z, acc.w};
  const float osc = out_scale != nullptr ? __ldg(out_scale) : 1.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[k] *= osc;
    if (bias != nullptr && j + k < Nout) o[k] += __ldg(bias + j + k);
    if (relu) o[k] = fmaxf(o[k], 0.f);
  }
  float *yp = y + (size_t)m * Nout + j;
  if ((Nout & 3) == 0 && j + 4 <= Nout) {
    *reinterpret_cast<float4 *>(yp) = make_float4(o[0], o[1], o[2], o[3]);
    if (y_hi != nullptr) {
      uint2 hi, lo;
      split2(o[0], o[1], hi.x, lo.x);
      split2(o[2], o[3], hi.y, lo.y);
      *reinterpret_cast<uint2 *>(y_hi + (size_t)m * Nout + j) = hi;
      *reinterpret_cast<uint2 *>(y_lo + (size_t)m * Nout + j) = lo;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) if (j + k < Nout) yp[k] = o[k];
  }
}

// ------------------------------- host side -------------------------------------------------------
struct Seg { const float *A; const __half *Bhi; const __half *Blo; int brows; };

template <int NBLK, int NBR, int NSEG, int EPI, int DIRECT = 0>
static int launch(const Params &p, const Seg *segs, int col_tiles, int splits, cudaStream_t st) {
  using C = Cfg<NBLK * NBR>;
  static bool attr = false;
  if (!attr) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_tc16<NBLK, NBR, NSEG, EPI, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  CUtensorMap tm[6];
  int rc;
  for (int s = 0; s < 2; ++s) {
    const Seg &g = segs[s < NSEG ? s : 0];
    if ((rc = make_tmap(&tm[3 * s + 0], g.A, p.M, p.K, BM, 4))) return rc;
    if ((rc = make_tmap(&tm[3 * s + 1], g.Bhi, g.brows, p.K, NBR, 2))) return rc;
    if ((rc = make_tmap(&tm[3 * s + 2], g.Blo, g.brows, p.K, NBR, 2))) return rc;
  }
  dim3 grid(col_tiles, (p.M + BM - 1) / BM, splits);
  if (p.sk_W > 0) grid = dim3(splits, 1, 1);        // stream-K: `splits` carries the CTA count
  k_tc16<NBLK, NBR, NSEG, EPI, DIRECT><<<grid, NTHR, C::SMEM, st>>>(p, tm[0], tm[3], tm[1], tm[2], tm[4], tm[5]);
  SGG_RETURN_IF_LAUNCH_FAILED("k_tc16");
  return 0;
}

static bool ok16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static long long *dbg_ptr() {
  static long long *ptr = nullptr;
  static bool init = false;
  if (!init) {
    init = true;
    const char *v = getenv("SGG_TC_TIMING");
    if (v && atoi(v) != 0) {
      void *sym = nullptr;
      if (cudaGetSymbolAddress(&sym, g_dbg) == cudaSuccess) ptr = (long long *)sym;
    }
  }
  return ptr;
}
int debug_timing(long long *host_out, int n_ctas) {
  if (n_ctas > DBG_MAX_CTAS) n_ctas = DBG_MAX_CTAS;
  SGG_CUDA_TRY(cudaMemcpyFromSymbol(host_out, g_dbg, sizeof(long long) * DBG_SLOTS * n_ctas));
  return 0;
}

// LINEAR plan.  Three shapes of work distribution, chosen by a cost model in units of "128-wide chunk times":
//   plain    : one CTA per output tile, whole K;
//   split-K  : gridDim.z equal K ranges per tile + fixed-order reducer (every split a whole number of 256-wide chunks);
//   stream-K : the (tile, chunk) space cut into one equal contiguous range per SM + reducer — removes partial waves
//              of multi-wave GEMMs (E = 9600 edges: 300 tiles = 2.03 waves; fc6: 2400 tiles).
// The reducer variants pay the partial-sum round trip, the reducer kernel and one more dependent launch.
struct LinPlan { int ncol, splits, sk_ctas, sk_maxseg; };
static size_t l2_bytes() {
  static size_t v = 0;
  if (v == 0) {
    int dev = 0, b = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&b, cudaDevAttrL2CacheSize, dev) == cudaSuccess && b > 0) v = (size_t)b;
    else v = (size_t)126 << 20;
  }
  return v;
}
static LinPlan plan_linear(int M, int Nout, int K, bool allow_split) {
  const int sms = sgg_num_sms();
  static const double split_fixed = getenv("SGG_TC16_SPLIT_COST") ? atof(getenv("SGG_TC16_SPLIT_COST")) : 3.0;   // tuning knobs
  static const int allow_sk = getenv("SGG_TC16_STREAMK") ? atoi(getenv("SGG_TC16_STREAMK")) : 1;
  const int chunks = (K + 255) / 256;
  const int rows = (M + BM - 1) / BM;
  LinPlan best{Nout <= 64 ? 64 : 128, 1, 0, 0};
  double best_cost = 1e30;
  // 96-wide tiles (plain tiling only): an option when 128-wide tiles leave half of the SMs idle, e.g. the cfg2 edge-unary
  // GEMM (19 row blocks): 4 x 19 = 76 CTAs at 128, 6 x 19 = 114 at 96.  The main loop is bound by shared-memory bandwidth
  // (3 MMA passes re-read the operand tiles), so time per k-block scales with the bytes a CTA moves, ~(128 + ncol).
  static const int allow96 = getenv("SGG_TC16_W96") ? atoi(getenv("SGG_TC16_W96")) : 1;
  const int widths[3] = {128, 96, 64};
  for (int wi = 0; wi < 3; ++wi) {
    const int ncol = widths[wi];
    if (ncol == 128 && Nout <= 64) continue;
    if (ncol == 96 && (!allow96 || Nout <= 128)) continue;
    const long tiles = (long)((Nout + ncol - 1) / ncol) * rows;
    const double unit = (128.0 + ncol) / 256.0;        // time of one chunk ~ bytes staged per k-block
    const int max_s = (allow_split && ncol != 96) ? (chunks < 32 ? chunks : 32) : 1;
    for (int s = 1; s <= max_s; ++s) {
      const int ch_per = (chunks + s - 1) / s;
      const int s_eff = (chunks + ch_per - 1) / ch_per;
      if (s_eff != s) continue;
      const long waves = (tiles * s + sms - 1) / sms;
      // fixed prologue/epilogue ~ 1.5 chunks of a 128-wide tile (measured: ~7k of ~1k-cycle k-blocks)
      const double per = ch_per * unit + 1.5;
      const double cost = waves * per + (s > 1 ? 0.3 * s + split_fixed : 0.0);
      if (cost < best_cost - 1e-9) { best_cost = cost; best = LinPlan{ncol, s, 0, 0}; }
    }
    // stream-K only when plain tiling already needs more than one wave: a GEMM with fewer tiles than SMs usually runs
    // next to other branches of the step (measured at cfg2: giving its 76-tile edge-unary GEMM all 148 SMs made the
    // GEMM 17 % faster and the step 6 % slower, because the object branch no longer overlapped)
    // ... and only while the weight operand stays L2-resident: stream-K hands every CTA a contiguous (tile, chunk) range
    // that starts at an arbitrary k offset, so concurrently resident CTAs share no operand tiles through L2.  fc6
    // (9600 x 25088 -> 4096: 411 MB of fp16 [hi|lo] weights) ran 9.32 ms stream-K vs 6.66 ms with plain launch-order
    // tiling (profiles/r02_fc6_tile_order.md), where a wave walks k in lockstep and reads A and B once per wave.
    const bool b_resident = 4.0 * (double)Nout * (double)K <= 0.5 * (double)l2_bytes();
    if (allow_split && allow_sk && ncol != 96 && b_resident && (K % 256) == 0 && tiles > sms) {
      const long W = tiles * chunks;
      const int G = (int)(W < sms ? W : sms);
      const int q = (int)((W + G - 1) / G);
      const int maxseg = (q + chunks - 2) / chunks + 1;
      const double cost = q * unit + 1.5 + split_fixed + 0.3 * maxseg;
      if (cost < best_cost - 1e-9) { best_cost = cost; best = LinPlan{ncol, 1, G, maxseg}; }
    }
  }
  return best;
}

size_t linear_workspace_floats(int M, int Nout, int K) {
  if (M <= 0 || Nout <= 0) return 0;
  const LinPlan pl = plan_linear(M, Nout, K, true);
  if (pl.sk_ctas > 0) return (size_t)pl.sk_ctas * pl.sk_maxseg * BM * pl.ncol;
  return pl.splits > 1 ? (size_t)pl.splits * M * Nout : 0;
}

int linear(const float *x, const float *w_split, const float *b, float *y, int M, int Nout, int K, int relu, float *ws,
           cudaStream_t st) {
  return linear_planes(x, w_split, b, y, nullptr, nullptr, M, Nout, K, relu, ws, st);
}
int linear_scaled(const float *x, const float *w_split, float *y, int M, int Nout, int K, const float *out_scale, float *ws,
                  cudaStream_t st, const float *in_scale) {
  return linear_planes(x, w_split, nullptr, y, nullptr, nullptr, M, Nout, K, 0, ws, st, out_scale, in_scale);
}

// same, and the epilogue (or the split-K / stream-K reducer) also writes the fp16 [hi | lo * 2^11] planes of y
// (Nout % 4 == 0 required when planes are requested)
int linear_planes(const float *x, const float *w_split, const float *b, float *y, __half *y_hi, __half *y_lo, int M,
                  int Nout, int K, int relu, float *ws, cudaStream_t st, const float *out_scale, const float *in_scale) {
  if (M <= 0 || Nout <= 0) return 0;
  if (y_hi != nullptr && (Nout & 3)) return sgg_set_err(SGG_E_BADARG, "tc16 linear: planes need Nout %% 4 == 0");
  if ((K & 7) || !ok16(x) || !ok16(w_split)) return sgg_set_err(SGG_E_BADARG, "tc16 linear: K %% 8 / alignment");
  const LinPlan pl = plan_linear(M, Nout, K, ws != nullptr);
  const int kblocks = (K + BK - 1) / BK, kcb = 256 / BK;
  const __half *wh = reinterpret_cast<const __half *>(w_split);
  Seg sg[2] = {{x, wh, wh + (size_t)Nout * K, Nout}, {}};
  Params p{}; p.M = M; p.K = K; p.Nout = Nout; p.relu = relu; p.bias = b; p.dbg = dbg_ptr();
  p.out_hi = y_hi; p.out_lo = y_lo; p.out_scale = out_scale; p.in_scale = in_scale;
  static const int pf_env = getenv("SGG_TC16_PF") ? atoi(getenv("SGG_TC16_PF")) : 0;
  p.pf = kblocks >= 32 ? pf_env : 0;
  // tried and NOT adopted (opt-in, SGG_TC16_DIRECT=1): 62 -> 65 us on the cfg2 edge-unary GEMM, 166 -> 180 us at M = 9600.
  // Global loads return through the same L1 / shared-memory data path the TMA write + LDS read-back used, so no bytes are
  // saved there, and the converters' stage-free wait shortens the effective prefetch distance.
  static const int direct_env = getenv("SGG_TC16_DIRECT") ? atoi(getenv("SGG_TC16_DIRECT")) : 0;
  const bool direct = direct_env != 0;
  p.a_direct = x;
  int rc;
  if (pl.sk_ctas > 0) {
    const int col_tiles = (Nout + pl.ncol - 1) / pl.ncol, rows = (M + BM - 1) / BM;
    p.kb_per_split = kblocks;
    p.sk_C = K / 256; p.sk_ct = col_tiles; p.sk_W = col_tiles * rows * p.sk_C; p.sk_maxseg = pl.sk_maxseg; p.sk_part = ws;
    p.out = y;
    if (direct) {
      if (pl.ncol == 64) rc = launch<1, 64, 1, EPI_LINEAR, 1>(p, sg, col_tiles, pl.sk_ctas, st);
      else rc = launch<1, 128, 1, EPI_LINEAR, 1>(p, sg, col_tiles, pl.sk_ctas, st);
    } else if (pl.ncol == 64) rc = launch<1, 64, 1, EPI_LINEAR>(p, sg, col_tiles, pl.sk_ctas, st);
    else rc = launch<1, 128, 1, EPI_LINEAR>(p, sg, col_tiles, pl.sk_ctas, st);
    if (rc) return rc;
    const int rows_per_cta = 256 / (pl.ncol / 4);
    k_tc16_streamk_reduce<<<dim3(col_tiles * rows, BM / rows_per_cta), 256, 0, st>>>(
        ws, p.sk_W, p.sk_C, pl.sk_ctas, pl.sk_maxseg, col_tiles, pl.ncol, M, Nout, b, relu, y, y_hi, y_lo, out_scale);
    SGG_RETURN_IF_LAUNCH_FAILED("k_tc16_streamk_reduce");
    return 0;
  }
  int splits = pl.splits, kb_per = kblocks;
  if (splits > 1) {
    const int chunks = (kblocks + kcb - 1) / kcb;
    const int ch_per = (chunks + splits - 1) / splits;
    kb_per = ch_per * kcb;
    splits = (kblocks + kb_per - 1) / kb_per;
  }
  p.kb_per_split = kb_per;
  p.out = splits > 1 ? ws : y;
  if (direct) {
    if (pl.ncol == 64) rc = launch<1, 64, 1, EPI_LINEAR, 1>(p, sg, (Nout + 63) / 64, splits, st);
    else if (pl.ncol == 96) rc = launch<1, 96, 1, EPI_LINEAR, 1>(p, sg, (Nout + 95) / 96, splits, st);
    else rc = launch<1, 128, 1, EPI_LINEAR, 1>(p, sg, (Nout + 127) / 128, splits, st);
  } else if (pl.ncol == 64) rc = launch<1, 64, 1, EPI_LINEAR>(p, sg, (Nout + 63) / 64, splits, st);
  else if (pl.ncol == 96) rc = launch<1, 96, 1, EPI_LINEAR>(p, sg, (Nout + 95) / 96, splits, st);
  else rc = launch<1, 128, 1, EPI_LINEAR>(p, sg, (Nout + 127) / 128, splits, st);
  if (rc) return rc;
  if (splits > 1) {
    const size_t mn = (size_t)M * Nout;
    int blocks = (int)((mn + 255) / 256 < 2048 ? (mn + 255) / 256 : 2048);
    k_tc16_splitk_reduce<<<blocks, 256, 0, st>>>(ws, splits, mn, Nout, b, relu, y, y_hi, y_lo, out_scale);
    SGG_RETURN_IF_LAUNCH_FAILED("k_tc16_splitk_reduce");
  }
  return 0;
}

// hidden-block width for INIT / EDGE tiles: minimise waves x (L2->SM bytes per tile + fixed cost)
static int plan_gru_width(int M, int H, bool edge) {
  const int sms = sgg_num_sms();
  const int rows = (M + BM - 1) / BM;
  const int cand[3] = {80, 64, 32};
  int best = 64; double best_cost = 1e30;
  for (int i = 0; i < 3; ++i) {
    const int nbr = cand[i];
    const long tiles = (long)((H + nbr - 1) / nbr) * rows;
    const long waves = (tiles + sms - 1) / sms;
    const double bytes = (128.0 + 3.0 * nbr) * H * 4.0 + 128.0 * nbr * 4.0 * (edge ? 8.0 : 2.0) + 130e3;
    const double cost = waves * bytes;
    if (cost < best_cost) { best_cost = cost; best = nbr; }
  }
  return best;
}

// mode 0 = INIT (h = 0), 1 = NODE (x = ctx, h = V), 2 = EDGE (h = Eh, gathered gi)
int gru(int mode, const float *x, const float *h, const float *w_ih_split, const float *w_hh_split, const float *b_ih,
        const float *b_hh, const float *P, const float *gates, const int *subj, const int *obj, float *out, float *cache,
        int M, int H, cudaStream_t st) {
  if (M <= 0) return 0;
  if (H % 16) return sgg_set_err(SGG_E_BADARG, "tc16 gru: H %% 16");
  Params p{}; p.M = M; p.K = H; p.H = H; p.b_ih = b_ih; p.b_hh = b_hh; p.h = h; p.P = P; p.gates = gates;
  p.subj = subj; p.obj = obj; p.out = out; p.cache = cache; p.dbg = dbg_ptr();
  const size_t wn = (size_t)3 * H * H;
  const __half *wih = reinterpret_cast<const __half *>(w_ih_split), *whh = reinterpret_cast<const __half *>(w_hh_split);
  if (mode == 1) {
    Seg sg[2] = {{x, wih, wih + wn, 3 * H}, {h, whh, whh + wn, 3 * H}};
    return launch<3, 32, 2, EPI_GRU_NODE>(p, sg, (H + 31) / 32, 1, st);
  }
  const int nbr = plan_gru_width(M, H, mode == 2);
  const int ct = (H + nbr - 1) / nbr;
  if (mode == 0) {
    Seg sg[2] = {{x, wih, wih + wn, 3 * H}, {}};
    if (nbr == 80) return launch<3, 80, 1, EPI_GRU_INIT>(p, sg, ct, 1, st);
    if (nbr == 64) return launch<3, 64, 1, EPI_GRU_INIT>(p, sg, ct, 1, st);
    return launch<3, 32, 1, EPI_GRU_INIT>(p, sg, ct, 1, st);
  }
  Seg sg[2] = {{h, whh, whh + wn, 3 * H}, {}};
  if (nbr == 80) return launch<3, 80, 1, EPI_GRU_EDGE>(p, sg, ct, 1, st);
  if (nbr == 64) return launch<3, 64, 1, EPI_GRU_EDGE>(p, sg, ct, 1, st);
  return launch<3, 32, 1, EPI_GRU_EDGE>(p, sg, ct, 1, st);
}

// sticky fp16 range flag: bit 0 = an activation operand, bit 1 = an emitted output plane, bit 2 = a weight left the fp16
// range (or was non-finite) since the last reset.  Synchronises the device.
int overflow_flag(int reset, unsigned int *out) {
  SGG_CUDA_TRY(cudaDeviceSynchronize());
  SGG_CUDA_TRY(cudaMemcpyFromSymbol(out, g_tc16_overflow, sizeof(unsigned int)));
  if (reset) {
    const unsigned int z = 0;
    SGG_CUDA_TRY(cudaMemcpyToSymbol(g_tc16_overflow, &z, sizeof(unsigned int)));
  }
  return 0;
}

int split_weights(const float *w, size_t n, void *split, cudaStream_t st) {
  const size_t want = (n / 8 + 255) / 256;
  const int cap = sgg_num_sms() * 8;
  int blocks = (int)(want < (size_t)cap ? (want > 0 ? want : 1) : (size_t)cap);
  __half *hp = reinterpret_cast<__half *>(split);
  k_tc16_split<<<blocks, 256, 0, st>>>(w, n, hp, hp + n);
  SGG_RETURN_IF_LAUNCH_FAILED("k_tc16_split");
  return 0;
}

}  // namespace tc16
}  // namespace sgg
