// Internal launch helpers shared between translation units.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
namespace sgg {
cudaStream_t side_stream(cudaStream_t main, int idx = 0);
int stream_order(cudaStream_t from, cudaStream_t to);
int launch_linear(const float *x, const float *w, const float *b, float *y, int M, int Nout, int K, int relu,
                  cudaStream_t st);
// skws (nullable): split-K partials for weight-gradient GEMMs (a_col && b_col); up to skws_floats / (M*N) slices
int launch_gemm(const float *A, int lda, bool a_col, const float *B, int ldb, bool b_col, float *C, int ldc, int M,
                int N, int K, bool accumulate, cudaStream_t st, float *skws = nullptr, size_t skws_floats = 0);
size_t wsum4_workspace_floats(int rows, int cols);
int launch_wsum4(const float *W4, const float *X, int rows, int cols, float *out, bool accumulate, float *ws,
                 cudaStream_t st);
// `count` independent (s, 1/s) pairs (4 floats apart) from abs-max partials `part_stride` floats apart
int launch_pow2_from_parts(const float *parts, int nparts, float *sc, int count, int part_stride, cudaStream_t st);
size_t colsum_workspace_floats(int rows, int cols);
int launch_colsum(const float *X, int ld, int rows, int cols, float *out, bool accumulate, float *ws, cudaStream_t st);
// 3xTF32 engine (tc_gemm.cu), callable directly: fp32 exponent range, used by the backward GEMMs whatever the
// forward engine is.  w_split = [hi | lo], each [Nout, K] fp32 (k_tc_split convention: hi = x & 0xffffe000).
size_t tc32_linear_workspace_floats(int M, int Nout, int K);
int tc32_linear(const float *x, const float *w_split, const float *b, float *y, int M, int Nout, int K, int relu,
                float *ws, cudaStream_t st);
// in [R, C] (row stride ldin) -> out [C, Rpad] (rows R..Rpad-1 zero); split: out = [hi | lo] 3xTF32 planes of C*Rpad floats
int launch_transpose_ld(const float *in, int ldin, int R, int C, float *out, int Rpad, bool split, cudaStream_t st);
size_t tc_linear_workspace_floats(int M, int Nout, int K);
int tc_linear(const float *x, const float *w_split, const float *b, float *y, int M, int Nout, int K, int relu,
              float *ws, cudaStream_t st);
// mode 0 = INIT (h = 0), 1 = NODE (x = ctx, h = V), 2 = EDGE (h = Eh, gathered gi)
int tc_gru(int mode, const float *x, const float *h, const float *w_ih_split, const float *w_hh_split,
           const float *b_ih, const float *b_hh, const float *P, const float *gates, const int *subj, const int *obj,
           float *out, float *cache, int M, int H, cudaStream_t st);

namespace tc16 {
// LINEAR whose epilogue / reducer also writes the fp16 [hi | lo * 2^11] planes of y (nullable; Nout % 4 == 0)
int linear_planes(const float *x, const float *w_split, const float *b, float *y, __half *y_hi, __half *y_lo, int M,
                  int Nout, int K, int relu, float *ws, cudaStream_t st, const float *out_scale = nullptr,
                  const float *in_scale = nullptr);
// y = out_scale[0] * ((in_scale[0] * x) w^T) on the 3xFP16 engine whatever the process-wide engine is (scaled backward
// GEMMs; in_scale nullable: the A operand is multiplied by it inside the kernel's fp32 -> fp16 split)
int linear_scaled(const float *x, const float *w_split, float *y, int M, int Nout, int K, const float *out_scale, float *ws,
                  cudaStream_t st, const float *in_scale = nullptr);
size_t linear_workspace_floats(int M, int Nout, int K);
}  // namespace tc16
// Fused message-passing loop of the 3xFP16 engine (mp_fused.cu): 2 launches per iteration.
namespace mpf {
struct Planes { __half *hi, *lo; };
bool supported(const sgg_mp_weights *w, int N, int E, int H);
size_t workspace_bytes(int N, int E, int H);
int forward(const float *obj_rep, const float *rel_rep, const Planes *obj_planes, const Planes *rel_planes,
            const void *graph_ws, const sgg_mp_weights *w, int N, int E, int H, int T, float *V_out, float *E_out,
            float *saved, void *ws, size_t ws_bytes, cudaStream_t st, Planes *last_planes = nullptr, int only = -1);
int heads(const Planes &Vp, const Planes &Ep, const sgg_head_weights *hw, int N, int E, int H, int n_cls, int n_rel,
          float *obj_dists, float *rel_dists, cudaStream_t st);
int debug_timing(long long *host_out, int n_ctas, int which);
}  // namespace mpf

struct MpTape {
  int N, E, H, T;
  float *states;   // (T+1) x [V_t (N*H) | E_t (E*H)]
  float *cacheV;   // (T+1) x N x 4H   (r, z, n, gh_n) of the GRU call that produced V_t
  float *cacheE;   // (T+1) x E x 4H
  float *gates;    // T x E x 4
  float *ctx;      // T x N x H
  float *P;        // T x N x 3H
  size_t floats;
};
MpTape mp_tape_view(float *base, int N, int E, int H, int T);
}
