// fp32 SIMT tile-GEMM core shared by the linear and GRU kernels.
//
//   acc[BM x (blocks of 64 cols)] += A[BM x K] * W[rows, K]^T
//
// Both operands are K-major (activations [M,K] row-major, nn.Linear / GRUCell
// weights [out,K] row-major), i.e. a "TN" GEMM.  One CTA = 256 threads laid out
// 16 (rows) x 16 (cols); a thread owns TM = BM/16 rows x 4 consecutive columns in
// each of the column blocks.  A column block is 64 consecutive weight rows
// starting at an arbitrary row base, so one tile can cover e.g. hidden units
// [j0, j0+64) of the r, z and n gates of a GRUCell (rows j0, H+j0, 2H+j0), which
// lets the GRU non-linearity run in the GEMM epilogue.
//
// K is consumed in chunks of 16 through double-buffered shared memory (k-major,
// so the inner product reads are float4 broadcasts / conflict-free); global loads
// for chunk c+1 are issued before the FMAs of chunk c (register staging).
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

// Rows of the A operand: direct rows m0.. of a row-major [M, lda] matrix (zero beyond M).
struct ARows {
  const float *base; int lda; int M; int m0;
};

// NW weight row-blocks; block b = rows [rowbase[b], rowbase[b] + nvalid[b]) of W (row stride ldw).
template <int NW>
struct WBlocks {
  const float *W; int ldw; int rowbase[NW]; int nvalid[NW];
};

template <int BM, int NACC, int NW, class Map>
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
  const int lrow = tid >> 2, lq = tid & 3;   // loader mapping: 4 threads per 16-float row chunk

  float4 ra[A_LD], rb[NW];
  const float *ap[A_LD];
  bool aok[A_LD];
#pragma unroll
  for (int l = 0; l < A_LD; ++l) {
    int r = lrow + l * 64;
    aok[l] = (A.m0 + r) < A.M;
    ap[l] = A.base + (size_t)(aok[l] ? A.m0 + r : 0) * A.lda + lq * 4;
  }
  const float *bp[NW];
  bool bok[NW];
#pragma unroll
  for (int b = 0; b < NW; ++b) {
    bok[b] = lrow < Wb.nvalid[b];
    bp[b] = Wb.W + (size_t)(Wb.rowbase[b] + (bok[b] ? lrow : 0)) * Wb.ldw + lq * 4;
  }
  auto gload = [&](int k0) {
#pragma unroll
    for (int l = 0; l < A_LD; ++l)
      ra[l] = aok[l] ? __ldg(reinterpret_cast<const float4 *>(ap[l] + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < NW; ++b)
      rb[b] = bok[b] ? __ldg(reinterpret_cast<const float4 *>(bp[b] + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto sstore = [&](int buf) {
    float *as = As + buf * BK * LDA;
    float *bs = Bs + buf * BK * LDB;
#pragma unroll
    for (int l = 0; l < A_LD; ++l) {
      int r = lrow + l * 64;
      as[(lq * 4 + 0) * LDA + r] = ra[l].x;
      as[(lq * 4 + 1) * LDA + r] = ra[l].y;
      as[(lq * 4 + 2) * LDA + r] = ra[l].z;
      as[(lq * 4 + 3) * LDA + r] = ra[l].w;
    }
#pragma unroll
    for (int b = 0; b < NW; ++b) {
      int c = b * BN + lrow;
      bs[(lq * 4 + 0) * LDB + c] = rb[b].x;
      bs[(lq * 4 + 1) * LDB + c] = rb[b].y;
      bs[(lq * 4 + 2) * LDB + c] = rb[b].z;
      bs[(lq * 4 + 3) * LDB + c] = rb[b].w;
    }
  };

  const int nchunk = K / BK;
  __syncthreads();           // previous users of smem (earlier segment) are done
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
        constexpr int dummy = 0; (void)dummy;
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
