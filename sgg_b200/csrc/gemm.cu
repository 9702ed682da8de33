// Generic fp32 tile GEMM over the SIMT core, used by the backward passes:
//   C[M,N] (=|+=) A(m,k) * B(n,k),  A(m,k) = a_col ? A[k*lda+m] : A[m*lda+k],  B(n,k) = b_col ? B[k*ldb+n] : B[n*ldb+k]
//   dX = dY W      : A = dY  (row), B = W (col: reduce over W's rows)
//   dW += dY^T X   : A = dY  (col), B = X (col), accumulate
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
                         int accumulate, cudaStream_t st) {
  static bool attr_done = false;
  const size_t smem = TileSmem<BM, NW>::bytes;
  if (!attr_done) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_gemm<BM, NW, A_COL, B_COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((N + NW * BN - 1) / (NW * BN), (M + BM - 1) / BM);
  k_gemm<BM, NW, A_COL, B_COL><<<grid, NTHREADS, smem, st>>>(A, lda, B, ldb, C, ldc, M, N, K, accumulate);
  SGG_RETURN_IF_LAUNCH_FAILED("k_gemm");
  return 0;
}

template <bool A_COL, bool B_COL>
static int launch_gemm_l(const float *A, int lda, const float *B, int ldb, float *C, int ldc, int M, int N, int K,
                         int accumulate, cudaStream_t st) {
  const int sms = sgg_num_sms();
  int best_bm = 64, best_nw = 1;
  double best = 1e30;
  const int bms[2] = {64, 128};
  for (int bi = 0; bi < 2; ++bi)
    for (int nw = 1; nw <= 3; ++nw) {
      const int bm = bms[bi];
      long tiles = (long)((N + nw * BN - 1) / (nw * BN)) * ((M + bm - 1) / bm);
      int occ = bm == 64 ? 2 : 1;
      long waves = (tiles + (long)sms * occ - 1) / ((long)sms * occ);
      double cost = (double)waves * occ * (bm * nw + 0.35 * (bm + nw * BN));
      if (cost < best) { best = cost; best_bm = bm; best_nw = nw; }
    }
#define SGG_CASE(BM_, NW_) \
  if (best_bm == BM_ && best_nw == NW_) return launch_gemm_t<BM_, NW_, A_COL, B_COL>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, st)
  SGG_CASE(64, 1); SGG_CASE(64, 2); SGG_CASE(64, 3);
  SGG_CASE(128, 1); SGG_CASE(128, 2); SGG_CASE(128, 3);
#undef SGG_CASE
  return sgg_set_err(SGG_E_BADARG, "gemm: no tile config");
}

int launch_gemm(const float *A, int lda, bool a_col, const float *B, int ldb, bool b_col, float *C, int ldc, int M,
                int N, int K, bool accumulate, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0) {
    if (!accumulate) SGG_CUDA_TRY(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, st));
    return 0;
  }
  if (!a_col && !b_col) return launch_gemm_l<false, false>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, st);
  if (!a_col && b_col) return launch_gemm_l<false, true>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, st);
  if (a_col && !b_col) return launch_gemm_l<true, false>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, st);
  return launch_gemm_l<true, true>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, st);
}

// column sums: out[c] (=|+=) sum_r X[r*ld + c]   — two deterministic stages (partials, then fixed-order sum)
__global__ void k_colsum_partial(const float *__restrict__ X, int ld, int rows, int cols, int rows_per, float *__restrict__ part) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per, r1 = min(rows, r0 + rows_per);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += X[(size_t)r * ld + c];
  part[(size_t)blockIdx.y * cols + c] = s;
}
__global__ void k_colsum_final(const float *__restrict__ part, int nparts, int cols, float *__restrict__ out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * cols + c];
  out[c] = accumulate ? out[c] + s : s;
}

size_t colsum_workspace_floats(int rows, int cols) {
  int nparts = (rows + 255) / 256; if (nparts > 128) nparts = 128; if (nparts < 1) nparts = 1;
  return (size_t)nparts * cols;
}

int launch_colsum(const float *X, int ld, int rows, int cols, float *out, bool accumulate, float *ws, cudaStream_t st) {
  if (cols <= 0) return 0;
  if (rows <= 0) { if (!accumulate) SGG_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)cols * 4, st)); return 0; }
  int nparts = (rows + 255) / 256; if (nparts > 128) nparts = 128;
  const int rows_per = (rows + nparts - 1) / nparts;
  nparts = (rows + rows_per - 1) / rows_per;
  dim3 grid((cols + 127) / 128, nparts);
  k_colsum_partial<<<grid, 128, 0, st>>>(X, ld, rows, cols, rows_per, ws);
  SGG_RETURN_IF_LAUNCH_FAILED("k_colsum_partial");
  k_colsum_final<<<(cols + 127) / 128, 128, 0, st>>>(ws, nparts, cols, out, accumulate ? 1 : 0);
  SGG_RETURN_IF_LAUNCH_FAILED("k_colsum_final");
  return 0;
}

}  // namespace sgg

// ---- nn.Linear backward: dx = dy W ; dW += dy^T x ; db += colsum(dy).  dy already masked by ReLU if any. ----
extern "C" size_t sgg_linear_backward_workspace_bytes(int M, int Nout) {
  return sgg::colsum_workspace_floats(M, Nout) * sizeof(float) + 256;
}

extern "C" int sgg_linear_backward(const float *x, const float *w, const float *dy, int M, int Nout, int K, float *dx,
                                   float *dw, float *db, void *ws, size_t ws_bytes, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (M < 0 || Nout <= 0 || K <= 0) return sgg_set_err(SGG_E_BADARG, "linear_backward: bad shape");
  if (dx && (rc = sgg::launch_gemm(dy, Nout, false, w, K, true, dx, K, M, K, Nout, false, st))) return rc;
  if (dw && (rc = sgg::launch_gemm(dy, Nout, true, x, K, true, dw, K, Nout, K, M, true, st))) return rc;
  if (db) {
    if (!ws || ws_bytes < sgg_linear_backward_workspace_bytes(M, Nout))
      return sgg_set_err(SGG_E_WORKSPACE, "linear_backward: workspace too small");
    if ((rc = sgg::launch_colsum(dy, Nout, M, Nout, db, true, (float *)ws, st))) return rc;
  }
  return 0;
}
