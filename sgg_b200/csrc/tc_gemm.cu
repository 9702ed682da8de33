// Tensor-core (tcgen05 + TMEM + TMA) GEMM engine with fused epilogues, fp32-accurate via 3xTF32.
//
//   D[128 x N] (TMEM, fp32) += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo      (kind::tf32, K = 8 / instruction)
//
// with x_hi = x & 0xffffe000 (exactly representable in tf32) and x_lo = x - x_hi (exact in fp32; the
// tensor core keeps its top 11 bits) => ~2^-21 relative error per product, i.e. fp32-grade results
// (the reference's 1e-4 parity bar leaves no room for single-pass TF32 on K = 4096).
//
// The tensor core's fp32 accumulator truncates on every accumulate step (measured: error grows linearly
// with K, 1.9e-4 at K = 4096), so the LINEAR kernels accumulate only KC = 256 of K per TMEM pass: two TMEM
// accumulator buffers ping-pong, and the epilogue warps drain each finished chunk into fp32 REGISTER
// accumulators (round-to-nearest adds) while the next chunk is being multiplied.
//
// One CTA = one 128-row x (NBLK x 64)-column output tile, 10 warps:
//   warp 0   : TMA producer  — per k-block (32 floats = one 128-byte swizzle span) loads the raw fp32 A
//              tile and the pre-split B_hi / B_lo tiles (weights are split once per weight version by
//              sgg_tc_split_weights) into a multi-stage ring (SWIZZLE_128B, K-major).
//   warps 2-5: split A in shared memory (hi in place, lo to a second buffer), fence.proxy.async, arrive.
//   warps 6-9: accumulate / epilogue warps, one thread per accumulator row (TMEM lane): drain chunks
//              (LINEAR) and run the fused epilogue (bias/ReLU or the GRU cell).
//   warp 1   : allocates TMEM; one thread issues 12 tcgen05.mma per k-block and commits to the
//              stage's "empty" barrier; final commit signals the epilogue.
// B column blocks are 64 weight rows at arbitrary row bases, so a tile can cover hidden units
// [j0, j0+64) of the r, z, n gates of a GRUCell (rows j0, H+j0, 2H+j0) and apply the GRU
// non-linearity — and for edges the gated gather of P = V W_ih^T rows — straight out of TMEM.
#include "tc_gemm.cuh"
#include "kernels.h"

namespace sgg {
namespace tc {

constexpr int BM = 128;
constexpr int BKF = 32;                      // floats per k-block (128 bytes)
constexpr int NBR = 64;                      // rows per B block
constexpr int A_BYTES = BM * BKF * 4;        // 16 KB
constexpr int BB_BYTES = NBR * BKF * 4;      // 8 KB
constexpr int NTHR = 320;
constexpr int KCB = 8;                       // k-blocks (of 32) per accumulation chunk => KC = 256
constexpr int SMEM_BUDGET = 200 * 1024;

enum { EPI_LINEAR = 0, EPI_GRU_INIT = 1, EPI_GRU_NODE = 2, EPI_GRU_EDGE = 3 };

struct Params {
  int M, K, H, Nout, relu;
  const float *bias;          // LINEAR
  float *out;                 // LINEAR: [M,Nout]; GRU: [M,H]
  const float *b_ih, *b_hh;   // GRU
  const float *h;             // GRU_NODE / GRU_EDGE: previous state [M,H]
  const float *P;             // GRU_EDGE: [N,3H]
  const float *gates;         // GRU_EDGE: [M,4]
  const int *subj, *obj;      // GRU_EDGE
  float *cache;               // GRU: nullable [M,4,H] (r, z, n, gh_n) for the backward pass
};

template <int NBLK>
struct Cfg {
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * NBLK * BB_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 4 ? 4 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float gru_point(float gi_r, float gh_r, float gi_z, float gh_z, float gi_n, float gh_n,
                                           float h, float *cache, int H) {
  const float r = sgg_sigmoid(gi_r + gh_r);
  const float z = sgg_sigmoid(gi_z + gh_z);
  const float n = tanhf(gi_n + r * gh_n);
  if (cache != nullptr) { cache[0] = r; cache[H] = z; cache[2 * H] = n; cache[3 * H] = gh_n; }
  return (1.0f - z) * n + z * h;
}

template <int NBLK, int NSEG, int EPI>
__global__ void __launch_bounds__(NTHR, 1)
k_tc_gemm(Params p, const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
          const __grid_constant__ CUtensorMap tmBh0, const __grid_constant__ CUtensorMap tmBl0,
          const __grid_constant__ CUtensorMap tmBh1, const __grid_constant__ CUtensorMap tmBl1) {
  constexpr int STAGES = Cfg<NBLK>::STAGES;
  constexpr int STAGE_BYTES = Cfg<NBLK>::STAGE_BYTES;
  constexpr int NCOL = NBLK * NBR;                 // MMA N
  constexpr bool CHUNKED = (EPI == EPI_LINEAR);
  static_assert(!CHUNKED || (NSEG == 1 && NCOL <= 128), "chunked LINEAR keeps NCOL register accumulators per thread");
  constexpr int ACC_COLS = CHUNKED ? 2 * NCOL : NSEG * NCOL;
  constexpr uint32_t TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static_assert(ACC_COLS <= 512 && NCOL <= 256, "tile too wide");
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
  uint64_t *full = bars, *ready = bars + STAGES, *empty = bars + 2 * STAGES, *tmem_full = bars + 3 * STAGES;
  uint64_t *tmem_empty = bars + 3 * STAGES + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM;
  const int j0 = blockIdx.x * (EPI == EPI_LINEAR ? NCOL : NBR);
  const int kblocks = (p.K + BKF - 1) / BKF;
  const int total = NSEG * kblocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0); prefetch_tmap(&tmBh0); prefetch_tmap(&tmBl0);
    if (NSEG > 1) { prefetch_tmap(&tmA1); prefetch_tmap(&tmBh1); prefetch_tmap(&tmBl1); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(ready + s, 128); mbar_init(empty + s, 1); }
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
        const int seg = it / kblocks, k0 = (it - seg * kblocks) * BKF;
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
        const uint64_t ah = make_sdesc_sw128(st), al = make_sdesc_sw128(st + A_BYTES);
        const uint64_t bh = make_sdesc_sw128(st + 2 * A_BYTES), bl = make_sdesc_sw128(st + 2 * A_BYTES + NBLK * BB_BYTES);
        uint32_t d, first;
        const int chunk = it / KCB, kc = it - chunk * KCB;
        if (CHUNKED) {
          if (kc == 0) {                       // buffer must have been drained by the epilogue warps
            mbar_wait(tmem_empty + (chunk & 1), ((chunk >> 1) & 1) ^ 1);
            fence_after_sync();
          }
          d = tmem_base + (uint32_t)((chunk & 1) * NCOL);
          first = (kc == 0) ? 1u : 0u;
        } else {
          d = tmem_base + (uint32_t)(seg * NCOL);
          first = (kb == 0) ? 1u : 0u;
        }
#pragma unroll
        for (int kk = 0; kk < BKF / 8; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);     // 32 bytes >> 4
          mma_tf32_ss(d, al + o, bh + o, idesc, (first && kk == 0) ? 0u : 1u);
          mma_tf32_ss(d, ah + o, bl + o, idesc, 1u);
          mma_tf32_ss(d, ah + o, bh + o, idesc, 1u);
        }
        mma_commit(empty + s);
        if (CHUNKED && (kc == KCB - 1 || it == total - 1)) mma_commit(tmem_full + (chunk & 1));
      }
      if (!CHUNKED) mma_commit(tmem_full);
    }
  } else if (warp < 6) {
    // ===================== split A (hi / lo) =====================
    const int t = threadIdx.x - 64;     // 0..127
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(full + s, ph);
      float4 *a = reinterpret_cast<float4 *>(stage_ptr(s));
      float4 *lo = reinterpret_cast<float4 *>(stage_ptr(s) + A_BYTES);
#pragma unroll
      for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
        const int idx = t + i * 128;
        const float4 v = a[idx];
        float4 hi, l;
        hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - hi.x;
        hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - hi.y;
        hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - hi.z;
        hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - hi.w;
        a[idx] = hi;
        lo[idx] = l;
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
          float v[16];
          tmem_ld16(taddr + (uint32_t)(b * NCOL + c0), v);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c0 + c] += v[c];
        }
        fence_before_sync();
        mbar_arrive(tmem_empty + b);
      }
      if (live) {
        const bool vec = (p.Nout & 3) == 0;
        float *yrow = p.out + (size_t)m * p.Nout;
#pragma unroll
        for (int c0 = 0; c0 < NCOL; c0 += 4) {
          const int j = j0 + c0;
          if (j < p.Nout) {
            float v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              v[c] = acc[c0 + c];
              if (p.bias != nullptr && j + c < p.Nout) v[c] += __ldg(p.bias + j + c);
              if (p.relu) v[c] = fmaxf(v[c], 0.f);
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
        float ar[8], az[8], an[8], br[8], bz[8], bn[8];
        __syncwarp();                              // tcgen05.ld is .sync.aligned: whole warp, converged
        tmem_ld8(taddr + (uint32_t)(0 * NBR + c0), ar);
        tmem_ld8(taddr + (uint32_t)(1 * NBR + c0), az);
        tmem_ld8(taddr + (uint32_t)(2 * NBR + c0), an);
        if (EPI == EPI_GRU_NODE) {
          tmem_ld8(taddr + (uint32_t)(NCOL + 0 * NBR + c0), br);
          tmem_ld8(taddr + (uint32_t)(NCOL + 1 * NBR + c0), bz);
          tmem_ld8(taddr + (uint32_t)(NCOL + 2 * NBR + c0), bn);
        }
        tmem_wait_ld();
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

// row-major fp32 [rows, K] -> tiles of box_rows x 32 floats, 128-byte swizzle, zero OOB fill
static int make_tmap(CUtensorMap *m, const float *base, int rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t gstr[1] = {(cuuint64_t)K * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BKF, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sgg_set_err(SGG_E_BADARG, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d", (int)r, rows, K);
  return 0;
}

struct Seg { const float *A; const float *Bhi; const float *Blo; int brows; };

template <int NBLK, int NSEG, int EPI>
static int launch(const Params &p, const Seg *segs, int col_tiles, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_tc_gemm<NBLK, NSEG, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<NBLK>::SMEM));
    attr = true;
  }
  CUtensorMap tm[6];
  int rc;
  for (int s = 0; s < 2; ++s) {
    const Seg &g = segs[s < NSEG ? s : 0];
    if ((rc = make_tmap(&tm[3 * s + 0], g.A, p.M, p.K, BM))) return rc;
    if ((rc = make_tmap(&tm[3 * s + 1], g.Bhi, g.brows, p.K, NBR))) return rc;
    if ((rc = make_tmap(&tm[3 * s + 2], g.Blo, g.brows, p.K, NBR))) return rc;
  }
  dim3 grid(col_tiles, (p.M + BM - 1) / BM);
  k_tc_gemm<NBLK, NSEG, EPI><<<grid, NTHR, Cfg<NBLK>::SMEM, st>>>(p, tm[0], tm[3], tm[1], tm[2], tm[4], tm[5]);
  SGG_RETURN_IF_LAUNCH_FAILED("k_tc_gemm");
  return 0;
}

static bool ok16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace tc

// y = act(x W^T + b) on tensor cores.  w_split = [hi | lo], each [Nout, K].
int tc_linear(const float *x, const float *w_split, const float *b, float *y, int M, int Nout, int K, int relu,
              cudaStream_t st) {
  if (M <= 0 || Nout <= 0) return 0;
  if ((K & 3) || !tc::ok16(x) || !tc::ok16(w_split)) return sgg_set_err(SGG_E_BADARG, "tc_linear: K %% 4 / alignment");
  tc::Params p{}; p.M = M; p.K = K; p.Nout = Nout; p.relu = relu; p.bias = b; p.out = y;
  tc::Seg sg[2] = {{x, w_split, w_split + (size_t)Nout * K, Nout}, {}};
  if (Nout <= 64) return tc::launch<1, 1, tc::EPI_LINEAR>(p, sg, 1, st);
  return tc::launch<2, 1, tc::EPI_LINEAR>(p, sg, (Nout + 127) / 128, st);
}

int tc_gru(int mode, const float *x, const float *h, const float *w_ih_split, const float *w_hh_split,
           const float *b_ih, const float *b_hh, const float *P, const float *gates, const int *subj, const int *obj,
           float *out, float *cache, int M, int H, cudaStream_t st) {
  if (M <= 0) return 0;
  if (H % 64) return sgg_set_err(SGG_E_BADARG, "tc_gru: H %% 64");
  tc::Params p{}; p.M = M; p.K = H; p.H = H; p.b_ih = b_ih; p.b_hh = b_hh; p.h = h; p.P = P; p.gates = gates;
  p.subj = subj; p.obj = obj; p.out = out; p.cache = cache;
  const size_t wn = (size_t)3 * H * H;
  if (mode == 0) {
    tc::Seg sg[2] = {{x, w_ih_split, w_ih_split + wn, 3 * H}, {}};
    return tc::launch<3, 1, tc::EPI_GRU_INIT>(p, sg, H / 64, st);
  } else if (mode == 1) {
    tc::Seg sg[2] = {{x, w_ih_split, w_ih_split + wn, 3 * H}, {h, w_hh_split, w_hh_split + wn, 3 * H}};
    return tc::launch<3, 2, tc::EPI_GRU_NODE>(p, sg, H / 64, st);
  }
  tc::Seg sg[2] = {{h, w_hh_split, w_hh_split + wn, 3 * H}, {}};
  return tc::launch<3, 1, tc::EPI_GRU_EDGE>(p, sg, H / 64, st);
}

}  // namespace sgg

extern "C" int sgg_tc_split_weights(const float *w, size_t n, float *split, void *stream) {
  if (n == 0) return 0;
  if (!w || !split) return sgg_set_err(SGG_E_BADARG, "tc_split_weights: null pointer");
  int blocks = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
  sgg::tc::k_tc_split<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, n, split, split + n);
  SGG_RETURN_IF_LAUNCH_FAILED("k_tc_split");
  return 0;
}

extern "C" int sgg_tc_linear_forward(const float *x, const float *w_split, const float *b, float *y, int M, int Nout,
                                     int K, int relu, void *stream) {
  if ((M > 0 && Nout > 0) && (!x || !w_split || !y)) return sgg_set_err(SGG_E_BADARG, "tc_linear: null pointer");
  return sgg::tc_linear(x, w_split, b, y, M, Nout, K, relu, (cudaStream_t)stream);
}
