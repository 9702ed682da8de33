// fp32 SIMT tile-GEMM core shared by the linear, GRU and backward kernels.
//
//   acc[BM x (blocks of 64 cols)] += A[BM x K] * B[cols x K]^T
//
// One CTA = 256 threads laid out 16 (rows) x 16 (cols); a thread owns TM = BM/16
// rows x 4 consecutive columns in each column block.  A column block is 64
// consecutive B-rows starting at an arbitrary base, so one tile can cover hidden
// units [j0, j0+64) of the r, z and n gates of a GRUCell (weight rows j0, H+j0,
// 2H+j0) and run the GRU non-linearity in the GEMM epilogue.
//
// Operand layouts (template flags):
//   A_COL = false : A(m,k) = base[(m0+m)*ld + k]    (activations, K-major)
//   A_COL = true  : A(m,k) = base[k*ld + m0+m]      (transposed view, e.g. dY^T for dW = dY^T X)
//   B_COL = false : B(j,k) = W[(rowbase+j)*ld + k]  (nn.Linear / GRUCell weight [out,K])
//   B_COL = true  : B(j,k) = W[k*ld + rowbase+j]    (dX = dY W: reduce over W's rows)
// K is consumed in chunks of 16 through double-buffered shared memory laid out
// k-major (inner-product reads are float4 broadcasts / conflict-free); global loads
// of chunk c+1 are issued before the FMAs of chunk c (register staging).  All edges
// (M, cols, K) are masked with zero fill; unaligned operands fall back to scalar loads.
#pragma once
#include "common.cuh"

namespace sgg {

constexpr int BK = 16;
constexpr int BN = 64;        // columns per block
constexpr int NTHREADS = 256;

template <int BM, int NW>
struct TileSmem {
  static constexpr int LDA = BM + 4;
  static constexpr int LDB = NW * BN + 4;
  static constexpr size_t bytes = sizeof(float) * 2 * BK * (LDA + LDB);
};

template <int A, int B, int C, int D = -1>
struct AccMap {
  __device__ static constexpr int at(int i) { return i == 0 ? A : (i == 1 ? B : (i == 2 ? C : D)); }
};

struct ARows {           // A operand of one tile
  const float *base; int lda; int M; int m0;
};

template <int NW>
struct WBlocks {         // NW column blocks of the B operand
  const float *W; int ldw; int rowbase[NW]; int nvalid[NW];
};

__device__ __forceinline__ bool sgg_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BM, int NACC, int NW, class Map, bool A_COL = false, bool B_COL = false>
__device__ __forceinline__ void gemm_segment(float (&acc)[BM / 16][NACC][4], const ARows &A,
                                             const WBlocks<NW> &Wb, int K, float *smem) {
  constexpr int TM = BM / 16;
  constexpr int LDA = TileSmem<BM, NW>::LDA;
  constexpr int LDB = TileSmem<BM, NW>::LDB;
  constexpr int A_LD = BM / 64;  // float4 loads per thread for the A chunk
  float *As = smem;                    // [2][BK][LDA]
  float *Bs = smem + 2 * BK * LDA;     // [2][BK][LDB]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // ---- loader mappings -------------------------------------------------------------
  // row mode: 4 threads per 16-float k-chunk of one row  -> (row = idx>>2, kq = idx&3)
  // col mode: one float4 of 4 consecutive rows/cols for one k -> (kk = idx / (W/4), q = idx % (W/4))
  const bool a_vec = ((A.lda & 3) == 0) && sgg_aligned16(A.base) && (A_COL || (K & 3) == 0);
  const bool b_vec = ((Wb.ldw & 3) == 0) && sgg_aligned16(Wb.W) && (B_COL || (K & 3) == 0);
  bool b_vec_blk[NW];
#pragma unroll
  for (int b = 0; b < NW; ++b) b_vec_blk[b] = b_vec && (!B_COL || (Wb.rowbase[b] & 3) == 0);

  float4 ra[A_LD], rb[NW];
  auto gload = [&](int k0) {
#pragma unroll
    for (int l = 0; l < A_LD; ++l) {
      const int idx = tid + l * NTHREADS;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (!A_COL) {
        const int r = idx >> 2, k = k0 + (idx & 3) * 4;
        if (A.m0 + r < A.M) {
          const float *p = A.base + (size_t)(A.m0 + r) * A.lda + k;
          if (a_vec && k + 3 < K) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (k + c < K) v[c] = __ldg(p + c);
          }
        }
      } else {
        const int kk = idx / (BM / 4), m = (idx % (BM / 4)) * 4, k = k0 + kk;
        if (k < K) {
          const float *p = A.base + (size_t)k * A.lda + A.m0 + m;
          if (a_vec && A.m0 + m + 3 < A.M) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (A.m0 + m + c < A.M) v[c] = __ldg(p + c);
          }
        }
      }
      ra[l] = make_float4(v[0], v[1], v[2], v[3]);
    }
#pragma unroll
    for (int b = 0; b < NW; ++b) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (!B_COL) {
        const int r = tid >> 2, k = k0 + (tid & 3) * 4;
        if (r < Wb.nvalid[b]) {
          const float *p = Wb.W + (size_t)(Wb.rowbase[b] + r) * Wb.ldw + k;
          if (b_vec_blk[b] && k + 3 < K) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (k + c < K) v[c] = __ldg(p + c);
          }
        }
      } else {
        const int kk = tid >> 4, j = (tid & 15) * 4, k = k0 + kk;
        if (k < K && j < Wb.nvalid[b]) {
          const float *p = Wb.W + (size_t)k * Wb.ldw + Wb.rowbase[b] + j;
          if (b_vec_blk[b] && j + 3 < Wb.nvalid[b]) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (j + c < Wb.nvalid[b]) v[c] = __ldg(p + c);
          }
        }
      }
      rb[b] = make_float4(v[0], v[1], v[2], v[3]);
    }
  };
  auto sstore = [&](int buf) {
    float *as = As + buf * BK * LDA;
    float *bs = Bs + buf * BK * LDB;
#pragma unroll
    for (int l = 0; l < A_LD; ++l) {
      const int idx = tid + l * NTHREADS;
      if (!A_COL) {
        const int r = idx >> 2, kq = (idx & 3) * 4;
        as[(kq + 0) * LDA + r] = ra[l].x;
        as[(kq + 1) * LDA + r] = ra[l].y;
        as[(kq + 2) * LDA + r] = ra[l].z;
        as[(kq + 3) * LDA + r] = ra[l].w;
      } else {
        const int kk = idx / (BM / 4), m = (idx % (BM / 4)) * 4;
        *reinterpret_cast<float4 *>(as + kk * LDA + m) = ra[l];
      }
    }
#pragma unroll
    for (int b = 0; b < NW; ++b) {
      if (!B_COL) {
        const int c = b * BN + (tid >> 2), kq = (tid & 3) * 4;
        bs[(kq + 0) * LDB + c] = rb[b].x;
        bs[(kq + 1) * LDB + c] = rb[b].y;
        bs[(kq + 2) * LDB + c] = rb[b].z;
        bs[(kq + 3) * LDB + c] = rb[b].w;
      } else {
        const int kk = tid >> 4, j = (tid & 15) * 4;
        *reinterpret_cast<float4 *>(bs + kk * LDB + b * BN + j) = rb[b];
      }
    }
  };

  const int nchunk = (K + BK - 1) / BK;
  __syncthreads();           // previous users of smem (earlier segment) are done
  if (nchunk == 0) return;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int c = 0; c < nchunk; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunk) gload((c + 1) * BK);
    const float *as = As + buf * BK * LDA + ty * TM;
    const float *bs = Bs + buf * BK * LDB + tx * 4;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4 *>(as + kk * LDA + i);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int b = 0; b < NW; ++b) {
        float4 w = *reinterpret_cast<const float4 *>(bs + kk * LDB + b * BN);
        const int s = Map::at(b);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          acc[i][s][0] = fmaf(a[i], w.x, acc[i][s][0]);
          acc[i][s][1] = fmaf(a[i], w.y, acc[i][s][1]);
          acc[i][s][2] = fmaf(a[i], w.z, acc[i][s][2]);
          acc[i][s][3] = fmaf(a[i], w.w, acc[i][s][3]);
        }
      }
    }
    if (c + 1 < nchunk) sstore(buf ^ 1);
    __syncthreads();
  }
}

}  // namespace sgg
