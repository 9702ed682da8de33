// Fused message-passing iteration (RelModelStanford.message_pass, sgg_models/rel_model_stanford.py:74-92) for the
// 3xFP16 tcgen05 engine: TWO launches per iteration instead of seven, no stand-alone gate kernels.
//
// State representation between launches (all L2-resident at the shapes of this model):
//   V, Eh            fp32 [rows, H]          (the GRU "h" operand, the ctx gather, the tape)
//   V, Eh, ctx       fp16 planes [hi | lo]   written by the PRODUCING epilogue (x = hi + 2^-11 lo, tc16_common.cuh), so
//                                            the consumer's TMA feeds tcgen05.mma directly — no in-kernel conversion
//   gate partials    pa[ct][row][4]          per column tile ct of the producer: sum_{j in tile} state[row][j] * w_k[j]
//                                            for the four scalar gate heads (:78-81, :86-89); consumers add the <= 7
//                                            partials in tile order (deterministic) and apply the sigmoid inline
//
//   launch A  k_mp_pre : [LIN tiles]  PQ = V_i [W_ih_e ; W_hh_n]^T  ->  P [N,3H] (edge-GRU input side)
//                                                                      Q [N,3H] (node-GRU hidden side, hoisted: it only
//                                                                                needs V_i, so it leaves launch B)
//                        [CTX CTAs]   ctx_i[n] = sum_out g_out Eh_i[e] + sum_in g_in Eh_i[e]   (:86-91, CSR, fixed order)
//   launch B  k_mp_gru : [EDGE tiles] Eh_{i+1} = GRU(g_s P[s] + g_o P[o] + b_ih, Eh_i W_hh_e^T + b_hh, Eh_i)   (:83)
//                        [NODE tiles] V_{i+1}  = GRU(ctx_i W_ih_n^T + b_ih, Q + b_hh, V_i)                      (:92)
//              both roles: 128 rows x (3 gates x 80 hidden units), K = H, one tcgen05 main loop + the GRUCell pointwise
//              epilogue, which also emits the fp16 planes and the gate partials of the new state.
//   the hx = 0 initial step (:68-72) is launch B with mode INIT for both roles.
// cfg2 (N = 240, E = 2400): launch B = 19 x 7 EDGE + 2 x 7 NODE = 147 CTAs = one wave of the 148 SMs.
#include <stdlib.h>
#include "tc16_common.cuh"

namespace sgg {
namespace mpf {
using namespace tc16;

constexpr int NBR = 80;                      // hidden units per column tile (x 3 gates = MMA N 240)
constexpr int NCOL = 3 * NBR;

enum { MODE_INIT = 0, MODE_NODE = 1, MODE_EDGE = 2 };

struct GruRole {
  int M, mode;
  int nct_n, nct_e;                 // EDGE: number of partial slabs of the node / edge gate dots
  int pa_n_rows, pa_e_rows;         // EDGE: rows per slab (N, E)
  const float *h;                   // previous state [M,H] (NODE, EDGE)
  const float *b_ih, *b_hh;
  const float *PQ;                  // EDGE: P [N,3H]; NODE: Q [M,3H]
  const int *subj, *obj;            // EDGE
  const float *pa_n, *pa_e;         // EDGE: gate partials of the current states
  const float *gate_b[4];           // EDGE
  float *gates_out;                 // EDGE: nullable [M,4] (tape)
  float *out;                       // [M,H]
  __half *out_hi, *out_lo;          // [M,H] planes of the new state (nullable)
  float *cache;                     // nullable [M,4,H]: (r, z, n, gh_n)
  const float *gw[4];               // gate weight vectors (H floats each) dotted with the NEW state; null: no partials
  float *pa_out;                    // [ct][M][4]
};

struct GruParams {
  int H, rb_edge, pdl;
  GruRole r[2];                     // 0 = edge rows, 1 = node rows
  long long *dbg;
};

constexpr int DBG_SLOTS = 8, DBG_MAX_CTAS = 1024;
__device__ long long g_dbg[DBG_SLOTS * DBG_MAX_CTAS];
__device__ long long g_dbg_pre[DBG_SLOTS * DBG_MAX_CTAS];
#define MPF_DBG(slot)                                                                                   \
  do {                                                                                                  \
    if (p.dbg != nullptr) {                                                                             \
      const int cta_ = blockIdx.x + gridDim.x * blockIdx.y;                                             \
      if (cta_ < DBG_MAX_CTAS) p.dbg[cta_ * DBG_SLOTS + (slot)] = clock64();                            \
    }                                                                                                   \
  } while (0)

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// bounded mbarrier wait: a protocol bug traps (clean CUDA error) instead of hanging the GPU
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

__device__ __forceinline__ float sigm(float x) { return fast_sigmoid(x); }

// K-major swizzled tile descriptor: 128-byte rows (BKT = 64, SWIZZLE_128B) or 64-byte rows (BKT = 32, SWIZZLE_64B)
template <int BKT>
__device__ __forceinline__ uint64_t make_sdesc_k(const void *smem_tile) {
  const uint64_t addr = (uint64_t)((tc::smem_u32(smem_tile) & 0x3FFFF) >> 4);
  constexpr uint64_t sbo = (uint64_t)((8 * BKT * 2) >> 4);
  constexpr uint64_t layout = BKT == 64 ? 2ull : 4ull;
  return addr | (sbo << 32) | (1ull << 46) | (layout << 61);
}

template <int BKT>
struct GCfg {
  static constexpr int A_PLANE = BM * BKT * 2;
  static constexpr int B_BLK = NBR * BKT * 2;
  static constexpr int B_PLANE = 3 * B_BLK;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int STAGES = BKT == 64 ? 2 : 4;
  static constexpr int RING = STAGES * STAGE;
  static constexpr int SMEM = RING + 1024 + 256;
  // epilogue overlay of the (by then idle) ring
  static constexpr int RS = 3 * NBR + 4;                 // staging row stride in floats (RS/4 odd: conflict-free)
  static constexpr int META_OFF = BM * RS * 4;           // [BM] int4 (subj, obj, g_sub, g_obj)
  static constexpr int PART_OFF = META_OFF + BM * 16;    // [BM][NBR/4] float4 gate-dot partials
  static constexpr int EPI_END = PART_OFF + BM * (NBR / 4) * 16;
  static_assert(EPI_END <= RING, "epilogue overlay must fit in the stage ring");
  static_assert(((RS / 4) & 1) == 1, "staging row stride");
};

// GRUCell pointwise math of one tile (torch.nn.GRUCell semantics) + fp16 planes and gate-dot partials of the new state.
// A thread keeps its hidden-unit quad for all its rows; rows are processed U at a time with every global load of the
// batch issued before the first use (out-of-range rows are clamped for the loads, only stores are predicated).
template <int BKT, int MODE>
__device__ __forceinline__ void pointwise(const GruRole &R, uint8_t *smem, int m0, int j0, int H) {
  using C = GCfg<BKT>;
  constexpr int RS = C::RS;
  constexpr int QPR = NBR / 4;                  // 20 float4 groups per row
  constexpr int RPI = NTHR / QPR;               // 16 rows per sweep
  constexpr int U = 2;
  const int t2 = (int)threadIdx.x;
  const int qd = t2 % QPR, r0 = t2 / QPR;
  const int j = j0 + 4 * qd;
  const float *stg0 = reinterpret_cast<const float *>(smem);
  const int4 *meta = reinterpret_cast<const int4 *>(smem + C::META_OFF);
  float4 *part = reinterpret_cast<float4 *>(smem + C::PART_OFF);
  const bool want_pa = R.pa_out != nullptr;
  if (j < H) {
    // r and z only ever see b_ih + b_hh; the n gate keeps them apart (n = tanh(gi_n + r * gh_n))
    const float4 brz_r = add4(ldg4(R.b_ih + j), ldg4(R.b_hh + j)), brz_z = add4(ldg4(R.b_ih + H + j), ldg4(R.b_hh + H + j));
    const float4 bin = ldg4(R.b_ih + 2 * H + j), bhn = ldg4(R.b_hh + 2 * H + j);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int mlast = R.M - 1;
    for (int rb = r0; rb < BM; rb += U * RPI) {
      float4 hv[U], x0[U], x1[U], x2[U], y0[U], y1[U], y2[U];
      float gs[U], go[U];
      int mm[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {             // ---- issue every global load of the batch (L2 latency overlaps)
        const int row = min(rb + u * RPI, BM - 1);
        mm[u] = m0 + row;
        const int mc = min(mm[u], mlast);       // clamped row for the loads
        hv[u] = zero4;
        if (MODE != MODE_INIT) hv[u] = ld4(R.h + (size_t)mc * H + j);
        if (MODE == MODE_EDGE) {
          const int4 mt = meta[row];
          gs[u] = __int_as_float(mt.z); go[u] = __int_as_float(mt.w);
          const float4 *ps = reinterpret_cast<const float4 *>(R.PQ + (size_t)mt.x * 3 * H + j);
          const float4 *po = reinterpret_cast<const float4 *>(R.PQ + (size_t)mt.y * 3 * H + j);
          x0[u] = ps[0]; x1[u] = ps[H / 4]; x2[u] = ps[H / 2];
          y0[u] = po[0]; y1[u] = po[H / 4]; y2[u] = po[H / 2];
        } else if (MODE == MODE_NODE) {
          const float4 *pq = reinterpret_cast<const float4 *>(R.PQ + (size_t)mc * 3 * H + j);
          x0[u] = pq[0]; x1[u] = pq[H / 4]; x2[u] = pq[H / 2];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {             // ---- GRUCell pointwise + stores
        const int row = rb + u * RPI;
        const float *sp = stg0 + (size_t)min(row, BM - 1) * RS + 4 * qd;
        const float4 a0v = ld4(sp), a1v = ld4(sp + NBR), a2v = ld4(sp + 2 * NBR);
        float4 pr, pz, gin, ghn;                // pre-activations of r and z (biases folded), gi_n, gh_n incl. bias
        if (MODE == MODE_INIT) {                // acc = x W_ih^T ; h = 0 => gh = b_hh
          pr = add4(a0v, brz_r); pz = add4(a1v, brz_z); gin = add4(a2v, bin); ghn = bhn;
        } else if (MODE == MODE_NODE) {         // acc = ctx W_ih^T ; gh = Q + b_hh
          pr = add4(add4(a0v, x0[u]), brz_r); pz = add4(add4(a1v, x1[u]), brz_z); gin = add4(a2v, bin); ghn = add4(x2[u], bhn);
        } else {                                // EDGE: acc = Eh W_hh^T ; gi = g_s P[s] + g_o P[o] + b_ih
          pr.x = fmaf(gs[u], x0[u].x, go[u] * y0[u].x) + a0v.x + brz_r.x; pr.y = fmaf(gs[u], x0[u].y, go[u] * y0[u].y) + a0v.y + brz_r.y;
          pr.z = fmaf(gs[u], x0[u].z, go[u] * y0[u].z) + a0v.z + brz_r.z; pr.w = fmaf(gs[u], x0[u].w, go[u] * y0[u].w) + a0v.w + brz_r.w;
          pz.x = fmaf(gs[u], x1[u].x, go[u] * y1[u].x) + a1v.x + brz_z.x; pz.y = fmaf(gs[u], x1[u].y, go[u] * y1[u].y) + a1v.y + brz_z.y;
          pz.z = fmaf(gs[u], x1[u].z, go[u] * y1[u].z) + a1v.z + brz_z.z; pz.w = fmaf(gs[u], x1[u].w, go[u] * y1[u].w) + a1v.w + brz_z.w;
          gin.x = fmaf(gs[u], x2[u].x, go[u] * y2[u].x) + bin.x; gin.y = fmaf(gs[u], x2[u].y, go[u] * y2[u].y) + bin.y;
          gin.z = fmaf(gs[u], x2[u].z, go[u] * y2[u].z) + bin.z; gin.w = fmaf(gs[u], x2[u].w, go[u] * y2[u].w) + bin.w;
          ghn = add4(a2v, bhn);
        }
        const Gru4 o = gru4(pr, zero4, pz, zero4, gin, ghn, hv[u]);
        if (row < BM) {
          if (want_pa) {                        // this thread's share of the four gate dots of the new state
            const float4 gw0 = ldg4(R.gw[0] + j), gw1 = ldg4(R.gw[1] + j), gw2 = ldg4(R.gw[2] + j), gw3 = ldg4(R.gw[3] + j);
            float4 d;
            d.x = o.out.x * gw0.x + o.out.y * gw0.y + o.out.z * gw0.z + o.out.w * gw0.w;
            d.y = o.out.x * gw1.x + o.out.y * gw1.y + o.out.z * gw1.z + o.out.w * gw1.w;
            d.z = o.out.x * gw2.x + o.out.y * gw2.y + o.out.z * gw2.z + o.out.w * gw2.w;
            d.w = o.out.x * gw3.x + o.out.y * gw3.y + o.out.z * gw3.z + o.out.w * gw3.w;
            part[row * QPR + qd] = d;
          }
          if (mm[u] < R.M) {
            const size_t off = (size_t)mm[u] * H + j;
            *reinterpret_cast<float4 *>(R.out + off) = o.out;
            if (R.out_hi != nullptr) {
              uint2 hi, lo;
              split2(o.out.x, o.out.y, hi.x, lo.x);
              split2(o.out.z, o.out.w, hi.y, lo.y);
              *reinterpret_cast<uint2 *>(R.out_hi + off) = hi;
              *reinterpret_cast<uint2 *>(R.out_lo + off) = lo;
            }
            if (R.cache != nullptr) {
              float *cp = R.cache + (size_t)mm[u] * 4 * H + j;
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
  if (want_pa) {                                // row sums of the partial dots, fixed order over the row's quads
    __syncthreads();
    if (t2 < BM && m0 + t2 < R.M) {
      const int nq = min(QPR, (H - j0 + 3) / 4);
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < nq; ++k) {
        const float4 v = part[t2 * QPR + k];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      reinterpret_cast<float4 *>(R.pa_out)[(size_t)blockIdx.x * R.M + m0 + t2] = s;
    }
  }
}

template <int BKT>
__global__ void __launch_bounds__(NTHR, 1)
k_mp_gru(const GruParams p, const __grid_constant__ CUtensorMap tmAh0, const __grid_constant__ CUtensorMap tmAl0,
         const __grid_constant__ CUtensorMap tmBh0, const __grid_constant__ CUtensorMap tmBl0,
         const __grid_constant__ CUtensorMap tmAh1, const __grid_constant__ CUtensorMap tmAl1,
         const __grid_constant__ CUtensorMap tmBh1, const __grid_constant__ CUtensorMap tmBl1) {
  using namespace tc;
  using C = GCfg<BKT>;
  constexpr int STAGES = C::STAGES, STAGE = C::STAGE, RS = C::RS;
  constexpr uint32_t TMEM_COLS = 512;                    // [main 240 | corr 240]
  constexpr int CORR = NCOL;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::RING);
  uint64_t *full = bars, *empty = bars + STAGES, *tmem_full = bars + 2 * STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = (int)blockIdx.y >= p.rb_edge ? 1 : 0;
  const GruRole &R = p.r[role];
  const int m0 = ((int)blockIdx.y - (role ? p.rb_edge : 0)) * BM;
  const int j0 = blockIdx.x * NBR;
  const int H = p.H;
  const int kblocks = (H + BKT - 1) / BKT;
  const CUtensorMap *tAh = role ? &tmAh1 : &tmAh0, *tAl = role ? &tmAl1 : &tmAl0;
  const CUtensorMap *tBh = role ? &tmBh1 : &tmBh0, *tBl = role ? &tmBl1 : &tmBl0;

  if (warp == 0 && lane == 0) {
    MPF_DBG(0);
    prefetch_tmap(tAh); prefetch_tmap(tAl); prefetch_tmap(tBh); prefetch_tmap(tBl);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) MPF_DBG(1);

  if (warp == 0) {
    // ===================== TMA producer =====================
    {                                          // whole warp runs the loop, one elected lane issues (uniform operands)
      // NODE tiles read the ctx planes written by the preceding launch A; EDGE / INIT operands are older
      if (p.pdl && R.mode == MODE_NODE) { griddep_wait(); asm volatile("fence.proxy.async;" ::: "memory"); }
      for (int it = 0; it < kblocks; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait_b(empty + s, ph ^ 1);
        uint8_t *st = smem + (size_t)s * STAGE;
        const int k0 = it * BKT;
        if (elect_one()) {
          mbar_arrive_expect_tx(full + s, STAGE);
          tma_load_2d(st, tAh, full + s, k0, m0);
          tma_load_2d(st + C::A_PLANE, tAl, full + s, k0, m0);
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            tma_load_2d(st + 2 * C::A_PLANE + b * C::B_BLK, tBh, full + s, k0, b * H + j0);
            tma_load_2d(st + 2 * C::A_PLANE + C::B_PLANE + b * C::B_BLK, tBl, full + s, k0, b * H + j0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop and ONE ELECTED lane issues: with the loop under `if (lane == 0)` the compiler cannot
    // prove the descriptors warp-uniform and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop.
    {
      constexpr uint32_t idesc = make_idesc_f16(BM, NCOL);
      const uint32_t dm = tmem_base, dc = tmem_base + (uint32_t)CORR;
      for (int it = 0; it < kblocks; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait_b(full + s, ph);
        fence_after_sync();
        if (it == 0 && lane == 0) MPF_DBG(2);
        uint8_t *st = smem + (size_t)s * STAGE;
        const uint64_t ah = make_sdesc_k<BKT>(st), al = make_sdesc_k<BKT>(st + C::A_PLANE);
        const uint64_t bh = make_sdesc_k<BKT>(st + 2 * C::A_PLANE), bl = make_sdesc_k<BKT>(st + 2 * C::A_PLANE + C::B_PLANE);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BKT / 16; ++kk) {
            const uint64_t o = (uint64_t)(kk * 2);           // 16 fp16 = 32 bytes >> 4
            const uint32_t acc = (it == 0 && kk == 0) ? 0u : 1u;
            mma_f16_ss(dc, al + o, bh + o, idesc, acc);      // corrections -> CORR
            mma_f16_ss(dc, ah + o, bl + o, idesc, 1u);
            mma_f16_ss(dm, ah + o, bh + o, idesc, acc);      // large term  -> MAIN
          }
          mma_commit(empty + s);
          if (it == kblocks - 1) mma_commit(tmem_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== warps 2..9 =====================
    // (a) EDGE: warps 6..9 fetch each row's endpoints and finish its scalar gates from the partial dots while the main
    //     loop runs; (b) all eight warps drain TMEM into the shared staging tile (two warps per TMEM lane quarter).
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    int4 meta = make_int4(0, 0, 0, 0);
    if (half == 1 && R.mode == MODE_EDGE && m0 + row < R.M) {
      const int e = m0 + row;
      const int s_ = __ldg(R.subj + e), o_ = __ldg(R.obj + e);
      float4 a = make_float4(__ldg(R.gate_b[0]), __ldg(R.gate_b[1]), __ldg(R.gate_b[2]), __ldg(R.gate_b[3]));
      for (int ct = 0; ct < R.nct_e; ++ct) {
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(R.pa_e) + (size_t)ct * R.pa_e_rows + e);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      for (int ct = 0; ct < R.nct_n; ++ct) {
        const float4 vs = __ldcg(reinterpret_cast<const float4 *>(R.pa_n) + (size_t)ct * R.pa_n_rows + s_);
        const float4 vo = __ldcg(reinterpret_cast<const float4 *>(R.pa_n) + (size_t)ct * R.pa_n_rows + o_);
        a.x += vs.x; a.y += vo.y; a.z += vs.z; a.w += vo.w;   // sub_vert / out_edge use the subject, obj_vert / in_edge the object
      }
      const float gs = sigm(a.x), go = sigm(a.y);
      if (R.gates_out != nullptr && blockIdx.x == 0)
        *reinterpret_cast<float4 *>(R.gates_out + (size_t)e * 4) = make_float4(gs, go, sigm(a.z), sigm(a.w));
      meta = make_int4(s_, o_, __float_as_int(gs), __float_as_int(go));
    }
    mbar_wait_b(tmem_full, 0);                    // all MMAs done => all stages consumed, the ring is free
    fence_after_sync();
    __syncwarp();
    if (half == 1 && R.mode == MODE_EDGE) reinterpret_cast<int4 *>(smem + C::META_OFF)[row] = meta;
    if (threadIdx.x == 64) MPF_DBG(3);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float *stg = reinterpret_cast<float *>(smem) + (size_t)row * RS;
    constexpr int NCH = NCOL / 16;                // 15 chunks of 16 columns; half 0 takes 8, half 1 takes 7
    const int c_lo = half ? 8 : 0, c_hi = half ? NCH : 8;
    for (int ch = c_lo; ch < c_hi; ++ch) {
      float v[16], w[16];
      tmem_ld16(taddr + (uint32_t)(ch * 16), v);
      tmem_ld16(taddr + (uint32_t)(CORR + ch * 16), w);
      tmem_wait_ld();
#pragma unroll
      for (int cc = 0; cc < 16; cc += 4)
        *reinterpret_cast<float4 *>(stg + ch * 16 + cc) =
            make_float4(fmaf(w[cc], LO_INV, v[cc]), fmaf(w[cc + 1], LO_INV, v[cc + 1]), fmaf(w[cc + 2], LO_INV, v[cc + 2]),
                        fmaf(w[cc + 3], LO_INV, v[cc + 3]));
    }
    if (threadIdx.x == 64) MPF_DBG(4);
  }

  // ===================== pointwise phase: thread <-> (4 hidden units, every RPI-th row), coalesced =====================
  if (p.pdl && R.mode != MODE_INIT) griddep_wait();      // P / Q come from the preceding launch A
  __syncthreads();
  if (threadIdx.x == 64) MPF_DBG(5);
  if (R.mode == MODE_EDGE) pointwise<BKT, MODE_EDGE>(R, smem, m0, j0, H);
  else if (R.mode == MODE_NODE) pointwise<BKT, MODE_NODE>(R, smem, m0, j0, H);
  else pointwise<BKT, MODE_INIT>(R, smem, m0, j0, H);
  if (threadIdx.x == 64) MPF_DBG(6);
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) MPF_DBG(7);
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// launch A: PQ = V [W_ih_e ; W_hh_n]^T (tcgen05, 128 x 128 tiles, K folded 256 at a time) + the vertex context gather
struct LinProb {                   // y[M,Nout] = A[M,H] W[Nout,H]^T (+ bias): A as fp16 planes, W pre-split
  int M, Nout, ld, ct, tiles;      // ld = row stride of y; ct = column tiles; tiles = row blocks * ct
  const float *bias;               // nullable
  float *y;
};
struct PreParams {
  int N, E, H, n_lin, nct_n, nct_e, pdl;
  LinProb lp[2];                   // launch A: P = V W_ih_e^T, Q = V W_hh_n^T; heads launch: rel_fc(Eh_T), obj_fc(V_T)
  // CTX
  const float *Eh;                 // [E,H] fp32
  const float *pa_n, *pa_e;        // gate partials of V_i / Eh_i
  const float *gate_b2, *gate_b3;  // out_edge / in_edge biases
  const int *out_ptr, *out_idx, *in_ptr, *in_idx;
  float *ctx;                      // nullable fp32 [N,H] (tape)
  __half *ctx_hi, *ctx_lo;         // [N,H]
  long long *dbg;
};

constexpr int LCOL = 128;                                  // LIN tile width
constexpr int L_STAGE = 2 * (BM * BK * 2) + 2 * (LCOL * BK * 2);     // 64 KB
constexpr int L_STAGES = 3;
constexpr int L_RING = L_STAGES * L_STAGE;
constexpr int L_SMEM = L_RING + 1024 + 256;

__global__ void __launch_bounds__(NTHR, 1)
k_mp_pre(const PreParams p, const __grid_constant__ CUtensorMap tmA0h, const __grid_constant__ CUtensorMap tmA0l,
         const __grid_constant__ CUtensorMap tmB0h, const __grid_constant__ CUtensorMap tmB0l,
         const __grid_constant__ CUtensorMap tmA1h, const __grid_constant__ CUtensorMap tmA1l,
         const __grid_constant__ CUtensorMap tmB1h, const __grid_constant__ CUtensorMap tmB1l) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H;
  if (p.pdl && threadIdx.x == 0) griddep_launch();        // the dependent launch B may start its main loops now
#define PRE_DBG(slot)                                                                                   \
  do {                                                                                                  \
    if (p.dbg != nullptr && blockIdx.x < DBG_MAX_CTAS) p.dbg[blockIdx.x * DBG_SLOTS + (slot)] = clock64(); \
  } while (0)
  if (threadIdx.x == 0) PRE_DBG(0);

  if ((int)blockIdx.x >= p.n_lin) {
    // ===================== CTX role: one warp per node, ten nodes in flight per CTA =====================
    // lane <-> list entry while the scalar gates are formed (partials summed in tile order, sigmoid), then the warp
    // walks the list in order (gate and edge id by shuffle) with lane <-> 4 float4 column groups of the edge state.
    const int n_ctx = (int)gridDim.x - p.n_lin;
    const int nslots = n_ctx * (NTHR / 32);
    const float b2 = __ldg(p.gate_b2), b3 = __ldg(p.gate_b3);
    const int nq = H / 128;                                // float4 column groups per lane (H % 128 == 0, H <= 512)
    for (int n = warp * n_ctx + ((int)blockIdx.x - p.n_lin); n < p.N; n += nslots) {     // warp-major: N / n_ctx nodes per CTA
      float an_z = 0.f, an_w = 0.f;                       // vertex halves of the out_edge / in_edge logits (:86, :88)
      for (int ct = 0; ct < p.nct_n; ++ct) {
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(p.pa_n) + (size_t)ct * p.N + n);
        an_z += v.z; an_w += v.w;
      }
      float4 acc[2][4];
#pragma unroll
      for (int l = 0; l < 2; ++l)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[l][k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int l = 0; l < 2; ++l) {                       // l = 0: out list (gate 2), l = 1: in list (gate 3)
        const int *ptr = l ? p.in_ptr : p.out_ptr, *idx = l ? p.in_idx : p.out_idx;
        const int lb = __ldg(ptr + n), le = __ldg(ptr + n + 1);
        const float base = l ? (an_w + b3) : (an_z + b2);
        for (int c0 = lb; c0 < le; c0 += 32) {
          const int cnt = min(32, le - c0);
          int e = 0;
          float g = 0.f;
          if (lane < cnt) {
            e = __ldg(idx + c0 + lane);
            float sacc = base;
            for (int ct = 0; ct < p.nct_e; ++ct) {
              const float4 v = __ldcg(reinterpret_cast<const float4 *>(p.pa_e) + (size_t)ct * p.E + e);
              sacc += l ? v.w : v.z;
            }
            g = sigm(sacc);
          }
          int k = 0;
          for (; k + 2 <= cnt; k += 2) {                  // two edges (8 x 16-byte loads per lane) in flight
            const int e0 = __shfl_sync(0xffffffffu, e, k), e1 = __shfl_sync(0xffffffffu, e, k + 1);
            const float g0 = __shfl_sync(0xffffffffu, g, k), g1 = __shfl_sync(0xffffffffu, g, k + 1);
            float4 v0[4], v1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < nq) {
                v0[q] = __ldcg(reinterpret_cast<const float4 *>(p.Eh + (size_t)e0 * H + 128 * q) + lane);
                v1[q] = __ldcg(reinterpret_cast<const float4 *>(p.Eh + (size_t)e1 * H + 128 * q) + lane);
              }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < nq) {
                float4 &a = acc[l][q];
                a.x += g0 * v0[q].x; a.y += g0 * v0[q].y; a.z += g0 * v0[q].z; a.w += g0 * v0[q].w;
                a.x += g1 * v1[q].x; a.y += g1 * v1[q].y; a.z += g1 * v1[q].z; a.w += g1 * v1[q].w;
              }
          }
          if (k < cnt) {
            const int e0 = __shfl_sync(0xffffffffu, e, k);
            const float g0 = __shfl_sync(0xffffffffu, g, k);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q < nq) {
                const float4 v0 = __ldcg(reinterpret_cast<const float4 *>(p.Eh + (size_t)e0 * H + 128 * q) + lane);
                float4 &a = acc[l][q];
                a.x += g0 * v0.x; a.y += g0 * v0.y; a.z += g0 * v0.z; a.w += g0 * v0.w;
              }
          }
        }
      }
      // ctx = out-sum + in-sum; fp16 planes for the node GRU's TMA, fp32 for the tape
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < nq) {
          const float4 a = acc[0][q], b = acc[1][q];
          const float4 c = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
          const size_t off = (size_t)n * H + 128 * q + 4 * lane;
          if (p.ctx != nullptr) *reinterpret_cast<float4 *>(p.ctx + off) = c;
          uint2 hi, lo;
          split2(c.x, c.y, hi.x, lo.x);
          split2(c.z, c.w, hi.y, lo.y);
          *reinterpret_cast<uint2 *>(p.ctx_hi + off) = hi;
          *reinterpret_cast<uint2 *>(p.ctx_lo + off) = lo;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { PRE_DBG(1); PRE_DBG(2); PRE_DBG(3); PRE_DBG(4); PRE_DBG(5); PRE_DBG(6); PRE_DBG(7); }
    return;
  }

  // ===================== LIN role: one 128 x 128 tile of [P | Q] =====================
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L_RING);
  uint64_t *full = bars, *empty = bars + L_STAGES, *tmem_full = bars + 2 * L_STAGES;       // tmem_full[2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * L_STAGES + 2);
  const int which = (int)blockIdx.x >= p.lp[0].tiles ? 1 : 0;
  const LinProb &L = p.lp[which];
  const int tl = (int)blockIdx.x - (which ? p.lp[0].tiles : 0);
  const int m0 = (tl / L.ct) * BM;
  const int j0 = (tl % L.ct) * LCOL;
  const int kblocks = (H + BK - 1) / BK;
  constexpr int KCB = 256 / BK;
  const int nchunks = (kblocks + KCB - 1) / KCB;           // <= 2 (host checks H <= 512)
  constexpr int A_PLANE = BM * BK * 2, B_PLANE = LCOL * BK * 2;
  const CUtensorMap *tAh = which ? &tmA1h : &tmA0h, *tAl = which ? &tmA1l : &tmA0l;
  const CUtensorMap *tBh = which ? &tmB1h : &tmB0h, *tBl = which ? &tmB1l : &tmB0l;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(tAh); prefetch_tmap(tAl); prefetch_tmap(tBh); prefetch_tmap(tBl);
    for (int s = 0; s < L_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1); mbar_init(tmem_full + 1, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) PRE_DBG(1);

  if (warp == 0) {
    {                                          // whole warp runs the loop, one elected lane issues
      for (int it = 0; it < kblocks; ++it) {
        const int s = it % L_STAGES, ph = (it / L_STAGES) & 1;
        mbar_wait_b(empty + s, ph ^ 1);
        uint8_t *st = smem + (size_t)s * L_STAGE;
        const int k0 = it * BK;
        if (elect_one()) {
          mbar_arrive_expect_tx(full + s, L_STAGE);
          tma_load_2d(st, tAh, full + s, k0, m0);
          tma_load_2d(st + A_PLANE, tAl, full + s, k0, m0);
          tma_load_2d(st + 2 * A_PLANE, tBh, full + s, k0, j0);
          tma_load_2d(st + 2 * A_PLANE + B_PLANE, tBl, full + s, k0, j0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {                                          // whole warp runs the loop, one elected lane issues (see k_mp_gru)
      constexpr uint32_t idesc = make_idesc_f16(BM, LCOL), idesc2 = make_idesc_f16(BM, 2 * LCOL);
      for (int it = 0; it < kblocks; ++it) {
        const int s = it % L_STAGES, ph = (it / L_STAGES) & 1;
        mbar_wait_b(full + s, ph);
        fence_after_sync();
        if (it == 0 && lane == 0) PRE_DBG(2);
        uint8_t *st = smem + (size_t)s * L_STAGE;
        const uint64_t ah = make_sdesc128(st), al = make_sdesc128(st + A_PLANE);
        const uint64_t bh = make_sdesc128(st + 2 * A_PLANE), bl = make_sdesc128(st + 2 * A_PLANE + B_PLANE);
        const int chunk = it / KCB, kc = it - chunk * KCB;
        const uint32_t dm = tmem_base + (uint32_t)(chunk * 2 * LCOL), dc = dm + (uint32_t)LCOL;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t o = (uint64_t)(kk * 2);
            const uint32_t acc = (kc == 0 && kk == 0) ? 0u : 1u;
            mma_f16_ss(dm, ah + o, bh + o, idesc2, acc);     // [main | a_hi b_lo]: one N = 2 LCOL MMA (B_lo follows B_hi in the stage)
            mma_f16_ss(dc, al + o, bh + o, idesc, 1u);       // corr += a_lo b_hi
          }
          mma_commit(empty + s);
          if (kc == KCB - 1 || it == kblocks - 1) mma_commit(tmem_full + chunk);
        }
        __syncwarp();
      }
    }
  } else {
    // warps 2..9: thread <-> (row, half of the tile's columns); chunk sums in fp32 registers
    constexpr int HC = LCOL / 2;
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HC);
    float acc[HC];
#pragma unroll
    for (int c = 0; c < HC; ++c) acc[c] = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) {
      mbar_wait_b(tmem_full + ch, 0);
      fence_after_sync();
      __syncwarp();
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 16) {
        float v[16], w[16];
        tmem_ld16(taddr + (uint32_t)(ch * 2 * LCOL + c0), v);
        tmem_ld16(taddr + (uint32_t)(ch * 2 * LCOL + LCOL + c0), w);
        tmem_wait_ld();
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) acc[c0 + cc] += fmaf(w[cc], LO_INV, v[cc]);
      }
    }
    if (threadIdx.x == 64) { PRE_DBG(3); PRE_DBG(4); }
    // registers -> shared tile (the ring is idle: every MMA has retired) -> coalesced 512-byte row stores
    constexpr int TS = LCOL + 4;                           // tile row stride in floats (TS/4 odd: conflict-free)
    float *tile = reinterpret_cast<float *>(smem);
#pragma unroll
    for (int c0 = 0; c0 < HC; c0 += 4)
      *reinterpret_cast<float4 *>(tile + (size_t)row * TS + half * HC + c0) = make_float4(acc[c0], acc[c0 + 1], acc[c0 + 2], acc[c0 + 3]);
    named_bar_sync(1, 256);
    if (threadIdx.x == 64) PRE_DBG(5);
    const int wq = warp - 2;                               // 0..7
    const int j = j0 + 4 * lane;
    if (j < L.Nout) {
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (L.bias != nullptr) {
        bv.x = __ldg(L.bias + j);
        if (j + 1 < L.Nout) bv.y = __ldg(L.bias + j + 1);
        if (j + 2 < L.Nout) bv.z = __ldg(L.bias + j + 2);
        if (j + 3 < L.Nout) bv.w = __ldg(L.bias + j + 3);
      }
      const bool vec = (L.ld & 3) == 0 && j + 4 <= L.Nout;
      for (int r = wq; r < BM; r += 8) {
        if (m0 + r >= L.M) break;
        float4 v = *reinterpret_cast<const float4 *>(tile + (size_t)r * TS + 4 * lane);
        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        float *yp = L.y + (size_t)(m0 + r) * L.ld + j;
        if (vec) {
          *reinterpret_cast<float4 *>(yp) = v;
        } else {
          yp[0] = v.x;
          if (j + 1 < L.Nout) yp[1] = v.y;
          if (j + 2 < L.Nout) yp[2] = v.z;
          if (j + 3 < L.Nout) yp[3] = v.w;
        }
      }
    }
  }
  if (threadIdx.x == 64) PRE_DBG(6);
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) PRE_DBG(7);
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
#undef PRE_DBG
}

// x fp32 -> fp16 planes [hi | lo * 2^11] (+ optional gate partial dots as ONE slab): entry of the L0 API, where the
// initial states come from the caller in fp32
__global__ void __launch_bounds__(256) k_split_planes(const float *__restrict__ x, size_t n4, __half *__restrict__ hi,
                                                      __half *__restrict__ lo) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4 *>(x)[i];
    uint2 h, l;
    split2(v.x, v.y, h.x, l.x);
    split2(v.z, v.w, h.y, l.y);
    reinterpret_cast<uint2 *>(hi)[i] = h;
    reinterpret_cast<uint2 *>(lo)[i] = l;
  }
}

// ------------------------------- host side -------------------------------------------------------
static int make_tmap_sw(CUtensorMap *m, const void *base, int rows, int K, int box_rows, int box_k) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d", (int)r, rows, K);
  return 0;
}

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
static long long *dbg_pre_ptr() {
  static long long *ptr = nullptr;
  static bool init = false;
  if (!init) {
    init = true;
    void *sym = nullptr;
    if (dbg_ptr() != nullptr && cudaGetSymbolAddress(&sym, g_dbg_pre) == cudaSuccess) ptr = (long long *)sym;
  }
  return ptr;
}
int debug_timing(long long *host_out, int n_ctas, int which) {
  if (n_ctas > DBG_MAX_CTAS) n_ctas = DBG_MAX_CTAS;
  if (which) SGG_CUDA_TRY(cudaMemcpyFromSymbol(host_out, g_dbg_pre, sizeof(long long) * DBG_SLOTS * n_ctas));
  else SGG_CUDA_TRY(cudaMemcpyFromSymbol(host_out, g_dbg, sizeof(long long) * DBG_SLOTS * n_ctas));
  return 0;
}

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
}
static int bk_choice() { static int v = env_int("SGG_MPF_BK", 32); return v == 64 ? 64 : 32; }
static int pdl_choice() { static int v = env_int("SGG_MPF_PDL", 1); return v != 0; }

struct Scratch {
  Planes V[2], Eh[2], ctx;
  float *Vf[2], *Ef[2];            // fp32 ping-pong states (inference; training writes into the tape)
  float *P, *Q, *ctxf;
  float *paN[2], *paE[2];
  int nct;
};

static size_t layout(Scratch *s, void *ws, int N, int E, int H) {
  SggArena ar(ws, (size_t)-1);
  const size_t n1 = N > 0 ? N : 1, e1 = E > 0 ? E : 1;
  const int nct = (H + NBR - 1) / NBR;
  s->nct = nct;
  for (int i = 0; i < 2; ++i) {
    s->V[i].hi = ar.take<__half>(n1 * H); s->V[i].lo = ar.take<__half>(n1 * H);
    s->Eh[i].hi = ar.take<__half>(e1 * H); s->Eh[i].lo = ar.take<__half>(e1 * H);
    s->Vf[i] = ar.take<float>(n1 * H); s->Ef[i] = ar.take<float>(e1 * H);
    s->paN[i] = ar.take<float>((size_t)nct * n1 * 4); s->paE[i] = ar.take<float>((size_t)nct * e1 * 4);
  }
  s->ctx.hi = ar.take<__half>(n1 * H); s->ctx.lo = ar.take<__half>(n1 * H);
  s->P = ar.take<float>(n1 * 3 * H); s->Q = ar.take<float>(n1 * 3 * H); s->ctxf = ar.take<float>(n1 * H);
  return ar.off;
}
size_t workspace_bytes(int N, int E, int H) {
  Scratch s;
  return layout(&s, nullptr, N, E, H);
}

bool supported(const sgg_mp_weights *w, int N, int E, int H) {
  static int enabled = env_int("SGG_MP_FUSED", 1);
  // Many-wave graphs (cfg5: 903 GRU tiles = 6.1 waves) run 7 % faster on the multi-stream schedule of mp.cu, whose edge
  // kernel, ctx gather, P GEMM and node GRU overlap across waves; up to ~4 waves (cfg3 / cfg4 shards) the fused one wins.
  if (enabled == 1) {
    const long tiles = (long)((E + BM - 1) / BM + (N + BM - 1) / BM) * ((H + NBR - 1) / NBR);
    if (tiles > 5L * sgg_num_sms()) return false;
  }
  return enabled && sgg_tc_get_mode() == 1 && N > 0 && E > 0 && H % 128 == 0 && H <= 512 && w->edge_w_ih_split && w->edge_w_hh_split &&
         w->node_w_ih_split && w->node_w_hh_split;
}

template <int BKT>
static int launch_gru_t(const GruParams &p, const CUtensorMap *tm, int col_tiles, int rows_total, bool pdl, cudaStream_t st) {
  using C = GCfg<BKT>;
  static bool attr = false;
  if (!attr) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_mp_gru<BKT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(col_tiles, rows_total, 1);
  cfg.blockDim = dim3(NTHR, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  SGG_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_mp_gru<BKT>, p, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7]));
  SGG_RETURN_IF_LAUNCH_FAILED("k_mp_gru");
  return 0;
}

struct GruOperands { const __half *a_hi, *a_lo; const __half *b_split; int M; };

// one launch B (or the INIT launch): role 0 = edge rows, role 1 = node rows
static int launch_gru(GruParams &p, const GruOperands &oe, const GruOperands &on, int H, bool pdl, cudaStream_t st) {
  const int bk = bk_choice();
  const size_t wn = (size_t)3 * H * H;
  CUtensorMap tm[8];
  const GruOperands *ops2[2] = {&oe, &on};
  int rc;
  for (int r = 0; r < 2; ++r) {
    const GruOperands &o = *ops2[r];
    if ((rc = make_tmap_sw(&tm[4 * r + 0], o.a_hi, o.M, H, BM, bk))) return rc;
    if ((rc = make_tmap_sw(&tm[4 * r + 1], o.a_lo, o.M, H, BM, bk))) return rc;
    if ((rc = make_tmap_sw(&tm[4 * r + 2], o.b_split, 3 * H, H, NBR, bk))) return rc;
    if ((rc = make_tmap_sw(&tm[4 * r + 3], o.b_split + wn, 3 * H, H, NBR, bk))) return rc;
  }
  p.H = H; p.rb_edge = (oe.M + BM - 1) / BM; p.pdl = pdl ? 1 : 0; p.dbg = dbg_ptr();
  const int rows_total = p.rb_edge + (on.M + BM - 1) / BM;
  const int ct = (H + NBR - 1) / NBR;
  return bk == 32 ? launch_gru_t<32>(p, tm, ct, rows_total, pdl, st) : launch_gru_t<64>(p, tm, ct, rows_total, pdl, st);
}

struct LinOperands { const __half *a_hi, *a_lo; const __half *b_split; int b_rows; };

// k_mp_pre launch: up to two linear problems (K = H) + optionally the ctx gather (a.N CTX nodes when n_ctx_max > 0)
static int launch_pre(PreParams &a, const LinOperands *lo, int H, int n_ctx, cudaStream_t st) {
  CUtensorMap tm[8];
  int rc;
  for (int r = 0; r < 2; ++r) {
    const LinOperands &o = lo[a.lp[r].tiles > 0 ? r : 0];
    const int rows = a.lp[a.lp[r].tiles > 0 ? r : 0].M;
    const size_t wn = (size_t)o.b_rows * H;
    if ((rc = make_tmap_sw(&tm[4 * r + 0], o.a_hi, rows, H, BM, 64))) return rc;
    if ((rc = make_tmap_sw(&tm[4 * r + 1], o.a_lo, rows, H, BM, 64))) return rc;
    if ((rc = make_tmap_sw(&tm[4 * r + 2], o.b_split, o.b_rows, H, LCOL, 64))) return rc;
    if ((rc = make_tmap_sw(&tm[4 * r + 3], o.b_split + wn, o.b_rows, H, LCOL, 64))) return rc;
  }
  static bool attr = false;
  if (!attr) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_mp_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM));
    attr = true;
  }
  a.H = H; a.n_lin = a.lp[0].tiles + a.lp[1].tiles; a.dbg = dbg_pre_ptr();
  k_mp_pre<<<a.n_lin + n_ctx, NTHR, L_SMEM, st>>>(a, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7]);
  SGG_RETURN_IF_LAUNCH_FAILED("k_mp_pre");
  return 0;
}
static LinProb lin_prob(int M, int Nout, int ld, const float *bias, float *y) {
  LinProb l{};
  l.M = M; l.Nout = Nout; l.ld = ld; l.bias = bias; l.y = y;
  l.ct = (Nout + LCOL - 1) / LCOL;
  l.tiles = M > 0 ? ((M + BM - 1) / BM) * l.ct : 0;
  return l;
}

// classifier heads (rel_model_stanford.py:107) on the final states, read through their fp16 planes: ONE launch
int heads(const Planes &Vp, const Planes &Ep, const sgg_head_weights *hw, int N, int E, int H, int n_cls, int n_rel,
          float *obj_dists, float *rel_dists, cudaStream_t st) {
  PreParams a{};
  a.lp[0] = lin_prob(E, n_rel, n_rel, hw->rel_fc_b, rel_dists);
  a.lp[1] = lin_prob(N, n_cls, n_cls, hw->obj_fc_b, obj_dists);
  LinOperands lo[2] = {{Ep.hi, Ep.lo, reinterpret_cast<const __half *>(hw->rel_fc_w_split), n_rel},
                       {Vp.hi, Vp.lo, reinterpret_cast<const __half *>(hw->obj_fc_w_split), n_cls}};
  return launch_pre(a, lo, H, 0, st);
}

// Whole message-passing loop.  Same contract as sgg::mp_forward (mp.cu): `saved` != null records the training tape.
// last_planes (nullable, [2] = {V_T, E_T}): the last launch also emits the planes of the final states (for heads()).
int forward(const float *obj_rep, const float *rel_rep, const Planes *obj_planes, const Planes *rel_planes,
            const void *graph_ws, const sgg_mp_weights *w, int N, int E, int H, int T, float *V_out, float *E_out,
            float *saved, void *ws, size_t ws_bytes, cudaStream_t st, Planes *last_planes, int only) {
  // only >= 0 (bench probe): issue a single launch of iteration 0 on a workspace a full run has initialised:
  // 0 = INIT launch, 1 = launch A, 2 = launch B
  Scratch s;
  const size_t need = layout(&s, ws, N, E, H);
  if (!ws || need > ws_bytes) return sgg_set_err(SGG_E_WORKSPACE, "mp_fused: workspace %zu < %zu", ws_bytes, need);
  SggGraphView g = sgg_graph_view(graph_ws, N, E);
  const size_t vN = (size_t)N * H, eN = (size_t)E * H;
  MpTape tape = mp_tape_view(saved, N, E, H, T);
  auto cacheV = [&](int it) -> float * { return saved ? tape.cacheV + (size_t)it * N * 4 * H : nullptr; };
  auto cacheE = [&](int it) -> float * { return saved ? tape.cacheE + (size_t)it * E * 4 * H : nullptr; };
  auto vbuf = [&](int it) -> float * {
    if (saved) return saved + (size_t)it * (vN + eN);
    return it == T ? V_out : s.Vf[it & 1];
  };
  auto ebuf = [&](int it) -> float * {
    if (saved) return saved + (size_t)it * (vN + eN) + vN;
    return it == T ? E_out : s.Ef[it & 1];
  };
  const int nct = s.nct;
  const bool pdl = pdl_choice();
  int rc;
  // planes of the inputs (the L1 entry gets them from the unary GEMM epilogues)
  Planes op = obj_planes ? *obj_planes : s.ctx, rp = rel_planes ? *rel_planes : s.Eh[1];
  if (!obj_planes && only < 0) {
    k_split_planes<<<(int)((vN / 4 + 255) / 256 < 1184 ? (vN / 4 + 255) / 256 : 1184), 256, 0, st>>>(obj_rep, vN / 4, op.hi, op.lo);
    SGG_RETURN_IF_LAUNCH_FAILED("k_split_planes");
  }
  if (!rel_planes && only < 0) {
    k_split_planes<<<(int)((eN / 4 + 255) / 256 < 1184 ? (eN / 4 + 255) / 256 : 1184), 256, 0, st>>>(rel_rep, eN / 4, rp.hi, rp.lo);
    SGG_RETURN_IF_LAUNCH_FAILED("k_split_planes");
  }
  auto fill_common = [&](GruRole &r, bool edge, int it_out) {
    r.M = edge ? E : N;
    r.b_ih = edge ? w->edge_b_ih : w->node_b_ih; r.b_hh = edge ? w->edge_b_hh : w->node_b_hh;
    r.out = edge ? ebuf(it_out) : vbuf(it_out);
    const bool last = it_out == T;
    const bool pl = !last || last_planes != nullptr;
    r.out_hi = !pl ? nullptr : (edge ? s.Eh[it_out & 1].hi : s.V[it_out & 1].hi);
    r.out_lo = !pl ? nullptr : (edge ? s.Eh[it_out & 1].lo : s.V[it_out & 1].lo);
    r.cache = edge ? cacheE(it_out) : cacheV(it_out);
    for (int k = 0; k < 4; ++k) r.gw[k] = last ? nullptr : (edge ? w->gate_w[k] + H : w->gate_w[k]);
    r.pa_out = last ? nullptr : (edge ? s.paE[it_out & 1] : s.paN[it_out & 1]);
  };
  if (only < 0 || only == 0) {  // hx = 0 initial step (:68-72)
    GruParams p{};
    p.r[0].mode = MODE_INIT; p.r[1].mode = MODE_INIT;
    fill_common(p.r[0], true, 0); fill_common(p.r[1], false, 0);
    GruOperands oe{rp.hi, rp.lo, reinterpret_cast<const __half *>(w->edge_w_ih_split), E};
    GruOperands on{op.hi, op.lo, reinterpret_cast<const __half *>(w->node_w_ih_split), N};
    if ((rc = launch_gru(p, oe, on, H, false, st))) return rc;
  }
  const size_t wn = (size_t)3 * H * H;
  for (int it = 0; it < (only >= 0 ? (T > 0 ? 1 : 0) : T); ++it) {
    const int cur = it & 1;
    float *P = saved ? tape.P + (size_t)it * N * 3 * H : s.P;
    float *ctxf = saved ? tape.ctx + (size_t)it * N * H : nullptr;
    float *gates = saved ? tape.gates + (size_t)it * E * 4 : nullptr;
    if (only < 0 || only == 1) {  // launch A
      PreParams a{};
      a.N = N; a.E = E; a.nct_n = nct; a.nct_e = nct; a.pdl = pdl ? 1 : 0;
      a.lp[0] = lin_prob(N, 3 * H, 3 * H, nullptr, P);
      a.lp[1] = lin_prob(N, 3 * H, 3 * H, nullptr, s.Q);
      a.Eh = ebuf(it); a.pa_n = s.paN[cur]; a.pa_e = s.paE[cur];
      a.gate_b2 = w->gate_b[2]; a.gate_b3 = w->gate_b[3];
      a.out_ptr = g.out_ptr; a.out_idx = g.out_idx; a.in_ptr = g.in_ptr; a.in_idx = g.in_idx;
      a.ctx = ctxf; a.ctx_hi = s.ctx.hi; a.ctx_lo = s.ctx.lo;
      LinOperands lo[2] = {{s.V[cur].hi, s.V[cur].lo, reinterpret_cast<const __half *>(w->edge_w_ih_split), 3 * H},
                           {s.V[cur].hi, s.V[cur].lo, reinterpret_cast<const __half *>(w->node_w_hh_split), 3 * H}};
      const int sms = sgg_num_sms(), n_lin = a.lp[0].tiles + a.lp[1].tiles;
      int n_ctx = sms - (n_lin % sms);                    // fill the wave the LIN tiles leave open
      if (n_ctx < 32) n_ctx += sms;
      if (n_ctx > N) n_ctx = N;
      if ((rc = launch_pre(a, lo, H, n_ctx, st))) return rc;
    }
    if (only < 0 || only == 2) {  // launch B
      GruParams p{};
      GruRole &re = p.r[0], &rn = p.r[1];
      re.mode = MODE_EDGE; rn.mode = MODE_NODE;
      fill_common(re, true, it + 1); fill_common(rn, false, it + 1);
      re.h = ebuf(it); re.PQ = P; re.subj = g.subj; re.obj = g.obj;
      re.nct_n = nct; re.nct_e = nct; re.pa_n_rows = N; re.pa_e_rows = E; re.pa_n = s.paN[cur]; re.pa_e = s.paE[cur];
      for (int k = 0; k < 4; ++k) re.gate_b[k] = w->gate_b[k];
      re.gates_out = gates;
      rn.h = vbuf(it); rn.PQ = s.Q;
      GruOperands oe{s.Eh[cur].hi, s.Eh[cur].lo, reinterpret_cast<const __half *>(w->edge_w_hh_split), E};
      GruOperands on{s.ctx.hi, s.ctx.lo, reinterpret_cast<const __half *>(w->node_w_ih_split), N};
      if ((rc = launch_gru(p, oe, on, H, pdl && only < 0, st))) return rc;
    }
  }
  if (only >= 0) return 0;
  if (saved) {   // outputs are the last saved slot
    SGG_CUDA_TRY(cudaMemcpyAsync(V_out, vbuf(T), vN * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SGG_CUDA_TRY(cudaMemcpyAsync(E_out, ebuf(T), eN * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  if (last_planes != nullptr) { last_planes[0] = s.V[T & 1]; last_planes[1] = s.Eh[T & 1]; }
  return 0;
}

}  // namespace mpf
}  // namespace sgg
