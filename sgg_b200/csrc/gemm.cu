// Generic fp32 tile GEMM over the SIMT core, used by the backward passes:
//   C[M,N] (=|+=) A(m,k) * B(n,k),  A(m,k) = a_col ? A[k*lda+m] : A[m*lda+k],  B(n,k) = b_col ? B[k*ldb+n] : B[n*ldb+k]
//   dX = dY W      : A = dY  (row), B = W (col: reduce over W's rows)
//   dW += dY^T X   : A = dY  (col), B = X (col), accumulate
#include <cuda_fp16.h>
#include "gemm_core.cuh"
#include "kernels.h"

namespace sgg {

template <int BM, int NW, bool A_COL, bool B_COL>
__global__ void __launch_bounds__(NTHREADS, (BM == 64 ? 2 : 1))
k_gemm(const float *__restrict__ A, int lda, const float *__restrict__ B, int ldb, float *__restrict__ C, int ldc, int M,
       int N, int K, int accumulate) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TM = BM / 16;
  const int m0 = blockIdx.y * BM;
  const int j0 = blockIdx.x * (NW * BN);
  if (gridDim.z > 1) {
    // split-K: slice z reduces k in [z*per, min(K, (z+1)*per)) into its own partial matrix (C = partials, ldc = N,
    // accumulate = 0); k_splitk_sum adds the slices in a fixed order.  per is a multiple of 16 (loader granularity).
    const int per = (((K + (int)gridDim.z - 1) / (int)gridDim.z) + 15) & ~15;
    const int k_lo = blockIdx.z * per;
    A += A_COL ? (size_t)k_lo * lda : (size_t)k_lo;
    B += B_COL ? (size_t)k_lo * ldb : (size_t)k_lo;
    C += (size_t)blockIdx.z * M * ldc;
    K = K - k_lo < per ? K - k_lo : per;
  }
  float acc[TM][NW][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int b = 0; b < NW; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][b][c] = 0.f;
  ARows Ar{A, lda, M, m0};
  WBlocks<NW> Wb;
  Wb.W = B; Wb.ldw = ldb;
#pragma unroll
  for (int b = 0; b < NW; ++b) {
    int rb = j0 + b * BN, nv = N - rb;
    Wb.rowbase[b] = nv > 0 ? rb : 0;
    Wb.nvalid[b] = nv < 0 ? 0 : (nv > BN ? BN : nv);
  }
  gemm_segment<BM, NW, NW, AccMap<0, 1, 2, 3>, A_COL, B_COL>(acc, Ar, Wb, K, smem);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool vec = ((ldc & 3) == 0) && sgg_aligned16(C);
#pragma unroll
  for (int b = 0; b < NW; ++b) {
    const int j = j0 + b * BN + tx * 4;
    if (j >= N) continue;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m >= M) continue;
      float *cp = C + (size_t)m * ldc + j;
      if (vec && j + 3 < N) {
        float4 v = make_float4(acc[i][b][0], acc[i][b][1], acc[i][b][2], acc[i][b][3]);
        if (accumulate) { const float4 o = *reinterpret_cast<float4 *>(cp); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        *reinterpret_cast<float4 *>(cp) = v;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (j + c < N) cp[c] = accumulate ? cp[c] + acc[i][b][c] : acc[i][b][c];
      }
    }
  }
}

template <int BM, int NW, bool A_COL, bool B_COL>
static int launch_gemm_t(const float *A, int lda, const float *B, int ldb, float *C, int ldc, int M, int N, int K,
                         int accumulate, cudaStream_t st, int splits = 1) {
  static bool attr_done = false;
  const size_t smem = TileSmem<BM, NW>::bytes;
  if (!attr_done) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_gemm<BM, NW, A_COL, B_COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((N + NW * BN - 1) / (NW * BN), (M + BM - 1) / BM, splits);
  k_gemm<BM, NW, A_COL, B_COL><<<grid, NTHREADS, smem, st>>>(A, lda, B, ldb, C, ldc, M, N, K, accumulate);
  SGG_RETURN_IF_LAUNCH_FAILED("k_gemm");
  return 0;
}

// C (=|+=) sum over split-K slices, fixed order (deterministic)
__global__ void k_splitk_sum(const float *__restrict__ part, int splits, int M, int N, float *__restrict__ C, int ldc,
                             int accumulate) {
  const size_t mn = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < mn; i += (size_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += part[(size_t)z * mn + i];
    float *cp = C + (i / N) * (size_t)ldc + (i % N);
    *cp = accumulate ? *cp + v : v;
  }
}

template <bool A_COL, bool B_COL>
static int launch_gemm_l(const float *A, int lda, const float *B, int ldb, float *C, int ldc, int M, int N, int K,
                         int accumulate, cudaStream_t st, float *skws, size_t skws_floats) {
  const int sms = sgg_num_sms();
  int best_bm = 64, best_nw = 1, best_s = 1;
  double best = 1e30;
  const int bms[2] = {64, 128};
  // split-K (weight-gradient GEMMs: few output tiles, very long K) needs a partials workspace and slices of >= 256 k
  int max_s = 1;
  if (skws != nullptr && A_COL && B_COL) {
    max_s = K / 256;
    const size_t fit = skws_floats / ((size_t)M * N);
    if ((size_t)max_s > fit) max_s = (int)fit;
    if (max_s > 8) max_s = 8;
    if (max_s < 1) max_s = 1;
  }
  for (int bi = 0; bi < 2; ++bi)
    for (int nw = 1; nw <= 3; ++nw)
      for (int sp = 1; sp <= max_s; ++sp) {
        const int bm = bms[bi];
        long tiles = (long)((N + nw * BN - 1) / (nw * BN)) * ((M + bm - 1) / bm) * sp;
        int occ = bm == 64 ? 2 : 1;
        long waves = (tiles + (long)sms * occ - 1) / ((long)sms * occ);
        // ~microseconds: per-tile cost (FMAs + operand traffic) x slice length, calibrated on B200 (a <128,1> tile
        // over K = 9600 takes ~690 us); a split adds the reducer launch and its partials traffic
        double cost = (double)waves * occ * (bm * nw + 0.35 * (bm + nw * BN)) * ((double)K / sp) * 3.7e-4 +
                      (sp > 1 ? 4.0 + (double)M * N * (sp + 1) * 4.0 / 3.0e6 : 0.0);
        if (cost < best) { best = cost; best_bm = bm; best_nw = nw; best_s = sp; }
      }
  float *Cd = C;
  int ldd = ldc, accd = accumulate;
  if (best_s > 1) { Cd = skws; ldd = N; accd = 0; }
  int rc = -1;
#define SGG_CASE(BM_, NW_) \
  if (best_bm == BM_ && best_nw == NW_) rc = launch_gemm_t<BM_, NW_, A_COL, B_COL>(A, lda, B, ldb, Cd, ldd, M, N, K, accd, st, best_s)
  SGG_CASE(64, 1); SGG_CASE(64, 2); SGG_CASE(64, 3);
  SGG_CASE(128, 1); SGG_CASE(128, 2); SGG_CASE(128, 3);
#undef SGG_CASE
  if (rc == -1) return sgg_set_err(SGG_E_BADARG, "gemm: no tile config");
  if (rc != 0 || best_s == 1) return rc;
  const size_t mn = (size_t)M * N;
  const int blocks = (int)((mn + 255) / 256 < 1184 ? (mn + 255) / 256 : 1184);
  k_splitk_sum<<<blocks, 256, 0, st>>>(skws, best_s, M, N, C, ldc, accumulate);
  SGG_RETURN_IF_LAUNCH_FAILED("k_splitk_sum");
  return 0;
}

int launch_gemm(const float *A, int lda, bool a_col, const float *B, int ldb, bool b_col, float *C, int ldc, int M,
                int N, int K, bool accumulate, cudaStream_t st, float *skws, size_t skws_floats) {
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0) {
    if (!accumulate) SGG_CUDA_TRY(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, st));
    return 0;
  }
  if (!a_col && !b_col) return launch_gemm_l<false, false>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, st, skws, skws_floats);
  if (!a_col && b_col) return launch_gemm_l<false, true>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, st, skws, skws_floats);
  if (a_col && !b_col) return launch_gemm_l<true, false>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, st, skws, skws_floats);
  return launch_gemm_l<true, true>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, st, skws, skws_floats);
}

// column sums: out[c] (=|+=) sum_r X[r*ld + c]   — two deterministic stages (partials, then fixed-order sum)
// Column sums, deterministic, two launches: per (column block, row slice) partial sums, then a fixed-order final sum.
// (Tried: one launch, the last-arriving CTA of a column block — device-scope counter — adds the slices.  Exact and
// deterministic, but the fence + serial tail made the 23 calls of a train step 26 us each instead of 10 + 6.6 us.)
__global__ void k_colsum_partial(const float *__restrict__ X, int ld, int rows, int cols, int rows_per, float *__restrict__ part) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;          // independent chains: four loads in flight per thread
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 += X[(size_t)r * ld + c]; s1 += X[(size_t)(r + 1) * ld + c];
    s2 += X[(size_t)(r + 2) * ld + c]; s3 += X[(size_t)(r + 3) * ld + c];
  }
  for (; r < r1; ++r) s0 += X[(size_t)r * ld + c];
  part[(size_t)blockIdx.y * cols + c] = (s0 + s1) + (s2 + s3);
}
__global__ void k_colsum_final(const float *__restrict__ part, int nparts, int cols, float *__restrict__ out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  // eight independent chains (fixed order, deterministic)
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int p = 0;
  for (; p + 7 < nparts; p += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += part[(size_t)(p + j) * cols + c];
  }
  for (; p < nparts; ++p) s[0] += part[(size_t)p * cols + c];
  const float t = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  out[c] = accumulate ? out[c] + t : t;
}

static int colsum_parts(int rows) {          // row slices: >= 32 rows each, at most 96 slices (the final stage walks them serially)
  int nparts = (rows + 31) / 32; if (nparts > 96) nparts = 96; if (nparts < 1) nparts = 1;
  return nparts;
}
size_t colsum_workspace_floats(int rows, int cols) { return (size_t)colsum_parts(rows) * cols; }

int launch_colsum(const float *X, int ld, int rows, int cols, float *out, bool accumulate, float *ws, cudaStream_t st) {
  if (cols <= 0) return 0;
  if (rows <= 0) { if (!accumulate) SGG_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)cols * 4, st)); return 0; }
  int nparts = colsum_parts(rows);
  const int rows_per = (rows + nparts - 1) / nparts;
  nparts = (rows + rows_per - 1) / rows_per;
  dim3 grid((cols + 127) / 128, nparts);
  k_colsum_partial<<<grid, 128, 0, st>>>(X, ld, rows, cols, rows_per, ws);
  SGG_RETURN_IF_LAUNCH_FAILED("k_colsum_partial");
  k_colsum_final<<<(cols + 127) / 128, 128, 0, st>>>(ws, nparts, cols, out, accumulate ? 1 : 0);
  SGG_RETURN_IF_LAUNCH_FAILED("k_colsum_final");
  return 0;
}

// out[k][c] (+)= sum_r W4[r][k] * X[r][c], k = 0..3 — the scalar-gate weight gradients (a 4 x cols "GEMM" with a
// very long reduction; as a GEMM it occupied 8 CTAs).  Two deterministic stages like colsum.
__global__ void __launch_bounds__(128) k_wsum4_partial(const float *__restrict__ W4, const float *__restrict__ X, int rows,
                                                       int cols, int rows_per, float *__restrict__ part) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  for (int r = r0; r < r1; ++r) {
    const float4 w = __ldg(reinterpret_cast<const float4 *>(W4) + r);
    const float4 x = *reinterpret_cast<const float4 *>(X + (size_t)r * cols + c);
    a0.x += w.x * x.x; a0.y += w.x * x.y; a0.z += w.x * x.z; a0.w += w.x * x.w;
    a1.x += w.y * x.x; a1.y += w.y * x.y; a1.z += w.y * x.z; a1.w += w.y * x.w;
    a2.x += w.z * x.x; a2.y += w.z * x.y; a2.z += w.z * x.z; a2.w += w.z * x.w;
    a3.x += w.w * x.x; a3.y += w.w * x.y; a3.z += w.w * x.z; a3.w += w.w * x.w;
  }
  float *p = part + (size_t)blockIdx.y * 4 * cols + c;
  *reinterpret_cast<float4 *>(p) = a0;
  *reinterpret_cast<float4 *>(p + cols) = a1;
  *reinterpret_cast<float4 *>(p + 2 * (size_t)cols) = a2;
  *reinterpret_cast<float4 *>(p + 3 * (size_t)cols) = a3;
}
static int wsum4_parts(int rows) {
  int nparts = (rows + 15) / 16; if (nparts > 296) nparts = 296; if (nparts < 1) nparts = 1;
  return nparts;
}
size_t wsum4_workspace_floats(int rows, int cols) { return (size_t)wsum4_parts(rows) * 4 * cols; }

int launch_wsum4(const float *W4, const float *X, int rows, int cols, float *out, bool accumulate, float *ws,
                 cudaStream_t st) {
  if (cols <= 0) return 0;
  if (cols & 3) return sgg_set_err(SGG_E_BADARG, "wsum4: cols must be a multiple of 4");
  if (rows <= 0) { if (!accumulate) SGG_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)cols * 16, st)); return 0; }
  int nparts = wsum4_parts(rows);
  const int rows_per = (rows + nparts - 1) / nparts;
  nparts = (rows + rows_per - 1) / rows_per;
  dim3 grid((cols / 4 + 127) / 128, nparts);
  k_wsum4_partial<<<grid, 128, 0, st>>>(W4, X, rows, cols, rows_per, ws);
  SGG_RETURN_IF_LAUNCH_FAILED("k_wsum4_partial");
  k_colsum_final<<<(4 * cols + 127) / 128, 128, 0, st>>>(ws, nparts, 4 * cols, out, accumulate ? 1 : 0);
  SGG_RETURN_IF_LAUNCH_FAILED("k_colsum_final");
  return 0;
}

}  // namespace sgg

// ---- nn.Linear backward: dx = dy W ; dW += dy^T x ; db += colsum(dy).  dy already masked by ReLU if any. ----
// workspace = column-sum partials (bias gradient) + split-K partials for the weight-gradient GEMM (few output tiles,
// reduction over all M rows: e.g. rel_fc's [51,512] gradient over 9600 edges occupied 8 CTAs without it)
static size_t lb_splitk_floats(int M, int Nout, int K) {
  if (M < 512) return 0;                                   // short reductions are not split (launch_gemm_l: >= 256 per slice)
  size_t per = (size_t)Nout * K, want = per * 8, cap = (size_t)8 << 20;     // <= 8 slices, <= 32 MB
  return want < cap ? want : (cap / per) * per;
}
extern "C" size_t sgg_linear_backward_workspace_bytes(int M, int Nout, int K) {
  return sgg_align_up(sgg::colsum_workspace_floats(M, Nout) * sizeof(float)) +
         sgg_align_up(lb_splitk_floats(M, Nout, K) * sizeof(float)) + 256;
}

extern "C" int sgg_linear_backward_ex(const float *x, const float *w, const float *dy, int M, int Nout, int K, float *dx,
                                      float *dw, float *db, int accumulate, void *ws, size_t ws_bytes, void *stream);
extern "C" int sgg_linear_backward(const float *x, const float *w, const float *dy, int M, int Nout, int K, float *dx,
                                   float *dw, float *db, void *ws, size_t ws_bytes, void *stream) {
  return sgg_linear_backward_ex(x, w, dy, M, Nout, K, dx, dw, db, 1, ws, ws_bytes, stream);
}

extern "C" int sgg_linear_backward_ex(const float *x, const float *w, const float *dy, int M, int Nout, int K, float *dx,
                                      float *dw, float *db, int accumulate, void *ws, size_t ws_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool acc = accumulate != 0;
  int rc;
  if (M < 0 || Nout <= 0 || K <= 0) return sgg_set_err(SGG_E_BADARG, "linear_backward: bad shape");
  const bool ws_ok = ws != nullptr && ws_bytes >= sgg_linear_backward_workspace_bytes(M, Nout, K);
  float *cs = (float *)ws;
  float *sk = ws_ok ? (float *)((char *)ws + sgg_align_up(sgg::colsum_workspace_floats(M, Nout) * sizeof(float))) : nullptr;
  const size_t sk_floats = ws_ok ? lb_splitk_floats(M, Nout, K) : 0;
  if (dx && (rc = sgg::launch_gemm(dy, Nout, false, w, K, true, dx, K, M, K, Nout, false, st))) return rc;
  if (dw && (rc = sgg::launch_gemm(dy, Nout, true, x, K, true, dw, K, Nout, K, M, acc, st, sk_floats ? sk : nullptr,
                                   sk_floats))) return rc;
  if (db) {
    if (!ws_ok) return sgg_set_err(SGG_E_WORKSPACE, "linear_backward: workspace too small");
    if ((rc = sgg::launch_colsum(dy, Nout, M, Nout, db, acc, cs, st))) return rc;
  }
  return 0;
}

// ---- tensor-core building blocks of the nn.Linear backward (3xTF32 engine: fp32 exponent range, gradients of any
// magnitude are safe).  The engine computes y = x w^T with x raw fp32 [M,K] and w pre-split [hi | lo] [Nout,K], so
//   dX = dY W      -> x = dY,             w = split(W^T)  (sgg_bwd_transpose(W, split = 1), cacheable per weight version)
//   dW = dY^T X    -> x = dY^T [Nout,Mp], w = split(X^T) [K,Mp], Mp = M padded with zero rows to a multiple of 32
// composed by sgg_b200.ops.linear_backward (which also chunks dW by output rows so that the gradient all-reduce of one
// chunk overlaps the GEMM of the next).
extern "C" int sgg_bwd_transpose(const float *in, long long ldin, int R, int C, float *out, int Rpad, int split, void *stream) {
  if (!in || !out || R < 0 || C <= 0 || Rpad < R || ldin < C) return sgg_set_err(SGG_E_BADARG, "bwd_transpose: bad argument");
  if (ldin > 0x7fffffffLL) return sgg_set_err(SGG_E_BADARG, "bwd_transpose: ldin too large");
  return sgg::launch_transpose_ld(in, (int)ldin, R, C, out, Rpad, split != 0, (cudaStream_t)stream);
}
extern "C" size_t sgg_tc32_linear_workspace_bytes(int M, int Nout, int K) {
  return sgg::tc32_linear_workspace_floats(M, Nout, K) * sizeof(float);
}
extern "C" int sgg_tc32_linear_forward(const float *x, const float *w_split, const float *b, float *y, int M, int Nout, int K,
                                       int relu, void *ws, size_t ws_bytes, void *stream) {
  if ((M > 0 && Nout > 0) && (!x || !w_split || !y)) return sgg_set_err(SGG_E_BADARG, "tc32_linear: null pointer");
  if (K % 4) return sgg_set_err(SGG_E_BADARG, "tc32_linear: K %% 4");
  if (sgg_tc32_linear_workspace_bytes(M, Nout, K) > 0 && (!ws || ws_bytes < sgg_tc32_linear_workspace_bytes(M, Nout, K)))
    return sgg_set_err(SGG_E_WORKSPACE, "tc32_linear: workspace too small");
  return sgg::tc32_linear(x, w_split, b, y, M, Nout, K, relu, (float *)ws, (cudaStream_t)stream);
}

// ---- scaled 3xFP16 variants of the backward building blocks (twice the MMA rate of 3xTF32) -------------------------
// The 3xFP16 split (x = hi + lo / 2^11, both fp16) is fp32-grade only while |x| stays in fp16's normal range; gradients do
// not (1e-6 .. 1e-9 is typical).  Multiplying the gradient operand by s = 2^k with max |s dY| in [1024, 2048) is exact in
// fp32, brings everything within 2^-24 of the maximum into the normal range, and is undone exactly in the GEMM epilogue
// (out_scale = 1 / s).  The scale lives in device memory: nothing synchronises.
namespace sgg {
constexpr int P2_PARTS = 1184;
__global__ void __launch_bounds__(256) k_absmax_partial(const float *__restrict__ x, size_t n, float *__restrict__ part) {
  __shared__ float sh[8];
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = fabsf(x[i]);
    if (v < INFINITY) m = fmaxf(m, v);          // NaN / inf do not take part (they poison the product anyway)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = sh[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    if (threadIdx.x == 0) part[blockIdx.x] = m;
  }
}
// sc[0] = s = 2^(10 - floor(log2 amax)), sc[1] = 1 / s; s = 1 when amax == 0
// blockIdx.x selects one of several independent reductions (partials `part_stride` apart, results 4 floats apart)
__global__ void k_pow2_scale(const float *__restrict__ part, int nparts, float *__restrict__ sc, int part_stride = 0) {
  part += (size_t)blockIdx.x * part_stride; sc += 4 * blockIdx.x;
  float m = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 32) m = fmaxf(m, part[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (threadIdx.x == 0) {
    int k = 0;
    if (m > 0.f) {
      k = 10 - ilogbf(m);
      k = k > 100 ? 100 : (k < -100 ? -100 : k);
    }
    sc[0] = ldexpf(1.0f, k);
    sc[1] = ldexpf(1.0f, -k);
  }
}
// in [R,C] (row stride ldin) -> out [C,Rpad] (rows >= R zero).  MODE 0: fp32 out scaled by sc[0] (sc nullable = 1);
// MODE 1: fp16 [hi | lo * 2^11] planes (operand split of the 3xFP16 engine)
template <int MODE>
__global__ void __launch_bounds__(256) k_transpose16(const float *__restrict__ in, int ldin, int R, int C, void *__restrict__ out,
                                                     int Rpad, const float *__restrict__ sc) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float s = sc != nullptr ? sc[0] : 1.f;               // power-of-two scale of a gradient operand (exact)
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * ldin + c] * s : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < Rpad) {
      const float v = tile[threadIdx.x][i];
      if (MODE == 0) {
        reinterpret_cast<float *>(out)[(size_t)c * Rpad + r] = v;
      } else {
        __half *hi = reinterpret_cast<__half *>(out), *lo = hi + (size_t)C * Rpad;
        const __half h = __float2half_rn(v);
        hi[(size_t)c * Rpad + r] = h;
        lo[(size_t)c * Rpad + r] = __float2half_rn((v - __half2float(h)) * 2048.0f);
      }
    }
  }
}
// y = x * sc[0] elementwise (A operand of a dX-type GEMM: dY stays row-major, only the scale is applied)
__global__ void k_scale_by(const float4 *__restrict__ x, size_t n4, const float *__restrict__ sc, float4 *__restrict__ y) {
  const float s = sc[0];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    y[i] = v;
  }
}
}  // namespace sgg

/* sc[0] = 2^k with max|x| * 2^k in [1024, 2048), sc[1] = 2^-k (device floats; ws: sgg_pow2_scale_workspace_bytes()) */
namespace sgg {
// (s, 1/s) from per-block abs-max partials a producer kernel already wrote (mp_bwd.cu: k_gru_bwd)
int launch_pow2_from_parts(const float *parts, int nparts, float *sc, int count, int part_stride, cudaStream_t st) {
  k_pow2_scale<<<count, 32, 0, st>>>(parts, nparts, sc, part_stride);
  SGG_RETURN_IF_LAUNCH_FAILED("k_pow2_scale");
  return 0;
}
}  // namespace sgg
extern "C" size_t sgg_pow2_scale_workspace_bytes(void) { return sgg::P2_PARTS * sizeof(float); }
extern "C" int sgg_pow2_scale(const float *x, long long n, float *sc, void *ws, size_t ws_bytes, void *stream) {
  if (!x || !sc || !ws || n <= 0 || ws_bytes < sgg_pow2_scale_workspace_bytes()) return sgg_set_err(SGG_E_BADARG, "pow2_scale: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int parts = (int)(((size_t)n + 255) / 256 < (size_t)sgg::P2_PARTS ? ((size_t)n + 255) / 256 : (size_t)sgg::P2_PARTS);
  sgg::k_absmax_partial<<<parts, 256, 0, st>>>(x, (size_t)n, (float *)ws);
  SGG_RETURN_IF_LAUNCH_FAILED("k_absmax_partial");
  sgg::k_pow2_scale<<<1, 32, 0, st>>>((const float *)ws, parts, sc);
  SGG_RETURN_IF_LAUNCH_FAILED("k_pow2_scale");
  return 0;
}
/* in [R,C] (row stride ldin) -> out [C,Rpad], multiplied by sc[0] (sc nullable): mode 0 = fp32; mode 1 = fp16 [hi | lo] planes */
extern "C" int sgg_bwd_transpose16(const float *in, long long ldin, int R, int C, void *out, int Rpad, int mode, const float *sc,
                                   void *stream) {
  if (!in || !out || R < 0 || C <= 0 || Rpad < R || ldin < C || ldin > 0x7fffffffLL || (mode != 0 && mode != 1))
    return sgg_set_err(SGG_E_BADARG, "bwd_transpose16: bad argument");
  dim3 grid((Rpad + 31) / 32, (C + 31) / 32);
  if (grid.y > 65535) return sgg_set_err(SGG_E_BADARG, "bwd_transpose16: too many columns");
  if (mode == 0) sgg::k_transpose16<0><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, (int)ldin, R, C, out, Rpad, sc);
  else sgg::k_transpose16<1><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, (int)ldin, R, C, out, Rpad, sc);
  SGG_RETURN_IF_LAUNCH_FAILED("k_transpose16");
  return 0;
}
extern "C" int sgg_scale_by(const float *x, long long n, const float *sc, float *y, void *stream) {
  if (!x || !y || !sc || n <= 0 || (n & 3)) return sgg_set_err(SGG_E_BADARG, "scale_by: bad argument");
  const size_t n4 = (size_t)n / 4;
  const int blocks = (int)((n4 + 255) / 256 < 2368 ? (n4 + 255) / 256 : 2368);
  sgg::k_scale_by<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)x, n4, sc, (float4 *)y);
  SGG_RETURN_IF_LAUNCH_FAILED("k_scale_by");
  return 0;
}
/* y = out_scale[0] * x w^T on the 3xFP16 engine (w_split16: fp16 [hi | lo] planes [Nout,K]); K % 8 == 0 */
extern "C" size_t sgg_tc16_linear_workspace_bytes(int M, int Nout, int K) {
  return sgg::tc16::linear_workspace_floats(M, Nout, K) * sizeof(float);
}
extern "C" int sgg_tc16_linear_scaled(const float *x, const void *w_split16, float *y, int M, int Nout, int K,
                                      const float *out_scale, void *ws, size_t ws_bytes, void *stream) {
  if ((M > 0 && Nout > 0) && (!x || !w_split16 || !y)) return sgg_set_err(SGG_E_BADARG, "tc16_linear_scaled: null pointer");
  if (ws && ws_bytes < sgg_tc16_linear_workspace_bytes(M, Nout, K)) return sgg_set_err(SGG_E_WORKSPACE, "tc16_linear_scaled: workspace too small");
  return sgg::tc16::linear_scaled(x, (const float *)w_split16, y, M, Nout, K, out_scale, (float *)ws, (cudaStream_t)stream);
}
