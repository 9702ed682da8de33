// nn.Linear forward (+bias, +ReLU) on the fp32 SIMT tile core.
// Reference call sites: obj_unary / edge_unary / obj_fc / rel_fc
// (sgg_models/rel_model_stanford.py:29-33,103-107), roi_fmap / roi_fmap_obj
// (sgg_models/rel_model_base.py:110-111).
#include "gemm_core.cuh"
#include "kernels.h"

namespace sgg {

template <int BM, int NW>
__global__ void __launch_bounds__(NTHREADS, (BM == 64 ? 2 : 1))
k_linear(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
         float *__restrict__ y, int M, int Nout, int K, int relu) {
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

  ARows A{x, K, M, m0};
  WBlocks<NW> Wb;
  Wb.W = w; Wb.ldw = K;
#pragma unroll
  for (int b = 0; b < NW; ++b) {
    int rb = j0 + b * BN;
    int nv = Nout - rb;
    Wb.rowbase[b] = nv > 0 ? rb : 0;
    Wb.nvalid[b] = nv < 0 ? 0 : (nv > BN ? BN : nv);
  }
  gemm_segment<BM, NW, NW, AccMap<0, 1, 2, 3>>(acc, A, Wb, K, smem);

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool vec = (Nout & 3) == 0;
#pragma unroll
  for (int b = 0; b < NW; ++b) {
    const int j = j0 + b * BN + tx * 4;
    if (j >= Nout) continue;
    float bv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) bv[c] = (bias != nullptr && j + c < Nout) ? __ldg(bias + j + c) : 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m >= M) continue;
      float v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        v[c] = acc[i][b][c] + bv[c];
        if (relu) v[c] = fmaxf(v[c], 0.f);
      }
      float *yp = y + (size_t)m * Nout + j;
      if (vec) {
        *reinterpret_cast<float4 *>(yp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (j + c < Nout) yp[c] = v[c];
      }
    }
  }
}

template <int BM, int NW>
static int launch_linear_t(const float *x, const float *w, const float *b, float *y, int M, int Nout, int K,
                           int relu, cudaStream_t st) {
  static bool attr_done = false;
  const size_t smem = TileSmem<BM, NW>::bytes;
  if (!attr_done) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_linear<BM, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((Nout + NW * BN - 1) / (NW * BN), (M + BM - 1) / BM);
  k_linear<BM, NW><<<grid, NTHREADS, smem, st>>>(x, w, b, y, M, Nout, K, relu);
  SGG_RETURN_IF_LAUNCH_FAILED("k_linear");
  return 0;
}

// Pick the tile shape with the smallest estimated time: cost = waves * BM * NW.
int launch_linear(const float *x, const float *w, const float *b, float *y, int M, int Nout, int K, int relu,
                  cudaStream_t st) {
  if (M <= 0 || Nout <= 0) return 0;
  if (K <= 0) return sgg_set_err(SGG_E_BADARG, "linear: K=%d must be positive", K);
  const int sms = sgg_num_sms();
  int best_bm = 64, best_nw = 1;
  double best = 1e30;
  const int bms[2] = {64, 128};
  for (int bi = 0; bi < 2; ++bi)
    for (int nw = 1; nw <= 3; ++nw) {
      const int bm = bms[bi];
      long tiles = (long)((Nout + nw * BN - 1) / (nw * BN)) * ((M + bm - 1) / bm);
      int occ = bm == 64 ? 2 : 1;
      long waves = (tiles + (long)sms * occ - 1) / ((long)sms * occ);
      // per-tile time ~ bm*nw FMAs per k with a fixed smem/issue overhead that favours fat tiles
      double cost = (double)waves * occ * (bm * nw + 0.35 * (bm + nw * BN));
      if (cost < best) { best = cost; best_bm = bm; best_nw = nw; }
    }
#define SGG_CASE(BM_, NW_) if (best_bm == BM_ && best_nw == NW_) return launch_linear_t<BM_, NW_>(x, w, b, y, M, Nout, K, relu, st)
  SGG_CASE(64, 1); SGG_CASE(64, 2); SGG_CASE(64, 3);
  SGG_CASE(128, 1); SGG_CASE(128, 2); SGG_CASE(128, 3);
#undef SGG_CASE
  return sgg_set_err(SGG_E_BADARG, "linear: no tile config");
}

}  // namespace sgg

extern "C" int sgg_linear_forward(const float *x, const float *w, const float *b, float *y, int M, int Nout, int K,
                                  int relu, void *stream) {
  if ((M > 0 && Nout > 0) && (!x || !w || !y)) return sgg_set_err(SGG_E_BADARG, "linear: null pointer");
  return sgg::launch_linear(x, w, b, y, M, Nout, K, relu, (cudaStream_t)stream);
}
