// Tensor-core engine selection + the C-ABI entry points of the tcgen05 GEMM path.
//   mode 0 : 3xTF32 (tc_gemm.cu)   — kind::tf32, fp32 [hi | lo] weight splits (2n floats)
//   mode 1 : 3xFP16 (tc16_gemm.cu) — kind::f16,  fp16 [hi | lo*2^11] weight splits (2n halves, first half of the buffer)
// A split-weight buffer is only valid for the mode it was made in (the Python cache is keyed by mode).
#include <stdlib.h>
#include <atomic>
#include "kernels.h"

namespace sgg {
size_t tc32_linear_workspace_floats(int M, int Nout, int K);
int tc32_linear(const float *x, const float *w_split, const float *b, float *y, int M, int Nout, int K, int relu,
                float *ws, cudaStream_t st);
int tc32_gru(int mode, const float *x, const float *h, const float *w_ih_split, const float *w_hh_split,
             const float *b_ih, const float *b_hh, const float *P, const float *gates, const int *subj, const int *obj,
             float *out, float *cache, int M, int H, cudaStream_t st);
int tc32_split_weights(const float *w, size_t n, float *split, cudaStream_t st);
namespace tc16 {
size_t linear_workspace_floats(int M, int Nout, int K);
int linear(const float *x, const float *w_split, const float *b, float *y, int M, int Nout, int K, int relu, float *ws,
           cudaStream_t st);
int gru(int mode, const float *x, const float *h, const float *w_ih_split, const float *w_hh_split, const float *b_ih,
        const float *b_hh, const float *P, const float *gates, const int *subj, const int *obj, float *out, float *cache,
        int M, int H, cudaStream_t st);
int split_weights(const float *w, size_t n, void *split, cudaStream_t st);
int debug_timing(long long *host_out, int n_ctas);
int overflow_flag(int reset, unsigned int *out);
}  // namespace tc16

static std::atomic<int> g_tc_mode{-1};
static int tc_mode() {
  int m = g_tc_mode.load(std::memory_order_relaxed);
  if (m < 0) {
    const char *v = getenv("SGG_TC_MODE");
    m = v ? (atoi(v) != 0 ? 1 : 0) : SGG_TC_DEFAULT_MODE;
    g_tc_mode.store(m);
  }
  return m;
}

// the workspace covers either engine, so a mode switch between sizing and launching stays safe
size_t tc_linear_workspace_floats(int M, int Nout, int K) {
  const size_t a = tc32_linear_workspace_floats(M, Nout, K), b = tc16::linear_workspace_floats(M, Nout, K);
  return a > b ? a : b;
}
int tc_linear(const float *x, const float *w_split, const float *b, float *y, int M, int Nout, int K, int relu,
              float *ws, cudaStream_t st) {
  return tc_mode() ? tc16::linear(x, w_split, b, y, M, Nout, K, relu, ws, st)
                   : tc32_linear(x, w_split, b, y, M, Nout, K, relu, ws, st);
}
int tc_gru(int mode, const float *x, const float *h, const float *w_ih_split, const float *w_hh_split,
           const float *b_ih, const float *b_hh, const float *P, const float *gates, const int *subj, const int *obj,
           float *out, float *cache, int M, int H, cudaStream_t st) {
  return tc_mode() ? tc16::gru(mode, x, h, w_ih_split, w_hh_split, b_ih, b_hh, P, gates, subj, obj, out, cache, M, H, st)
                   : tc32_gru(mode, x, h, w_ih_split, w_hh_split, b_ih, b_hh, P, gates, subj, obj, out, cache, M, H, st);
}
}  // namespace sgg

extern "C" int sgg_tc_set_mode(int mode) {
  if (mode != 0 && mode != 1) return sgg_set_err(SGG_E_BADARG, "tc_set_mode: mode must be 0 (3xTF32) or 1 (3xFP16)");
  sgg::g_tc_mode.store(mode);
  return 0;
}
extern "C" int sgg_tc_get_mode(void) { return sgg::tc_mode(); }

extern "C" int sgg_tc_split_weights(const float *w, size_t n, float *split, void *stream) {
  if (n == 0) return 0;
  if (!w || !split) return sgg_set_err(SGG_E_BADARG, "tc_split_weights: null pointer");
  if (sgg::tc_mode()) {
    if (n & 7) return sgg_set_err(SGG_E_BADARG, "tc_split_weights (3xFP16): element count must be a multiple of 8");
    return sgg::tc16::split_weights(w, n, split, (cudaStream_t)stream);
  }
  return sgg::tc32_split_weights(w, n, split, (cudaStream_t)stream);
}

extern "C" size_t sgg_tc_linear_workspace_bytes(int M, int Nout, int K) {
  return sgg::tc_linear_workspace_floats(M, Nout, K) * sizeof(float);
}

extern "C" int sgg_tc_linear_forward(const float *x, const float *w_split, const float *b, float *y, int M, int Nout,
                                     int K, int relu, void *ws, size_t ws_bytes, void *stream) {
  if ((M > 0 && Nout > 0) && (!x || !w_split || !y)) return sgg_set_err(SGG_E_BADARG, "tc_linear: null pointer");
  // ws == NULL is allowed (no split-K / stream-K); a workspace that is too small is an error, as the header says
  if (ws && ws_bytes < sgg_tc_linear_workspace_bytes(M, Nout, K))
    return sgg_set_err(SGG_E_WORKSPACE, "tc_linear: workspace %zu < %zu", ws_bytes, sgg_tc_linear_workspace_bytes(M, Nout, K));
  return sgg::tc_linear(x, w_split, b, y, M, Nout, K, relu, (float *)ws, (cudaStream_t)stream);
}

/* debug: per-CTA phase timestamps (8 clock64 values per CTA) of the most recent 3xFP16 kernel; only recorded when the
 * process runs with SGG_TC_TIMING=1.  Synchronises the device. */
extern "C" int sgg_tc_debug_timing(long long *host_out, int n_ctas) {
  if (!host_out || n_ctas <= 0) return sgg_set_err(SGG_E_BADARG, "tc_debug_timing: bad argument");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return sgg_set_err((int)e, "tc_debug_timing: %s", cudaGetErrorString(e));
  return sgg::tc16::debug_timing(host_out, n_ctas);
}

extern "C" int sgg_mpf_debug_timing(long long *host_out, int n_ctas, int which) {
  if (!host_out || n_ctas <= 0) return sgg_set_err(SGG_E_BADARG, "mpf_debug_timing: bad argument");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return sgg_set_err((int)e, "mpf_debug_timing: %s", cudaGetErrorString(e));
  return sgg::mpf::debug_timing(host_out, n_ctas, which);
}

namespace sgg { namespace lin16p { int overflow_flag(int reset, unsigned int *out); } int roi_overflow_flag(int reset, unsigned int *out); int bcast_overflow_flag(int reset, unsigned int *out); }
/* fp16 range guard of the 3xFP16 engine (|x| must stay below 65504): returns the sticky flag (0 = every operand was in
 * range since the last reset; bit 0 activations, bit 1 emitted planes, bit 2 weights), negative on error. */
extern "C" int sgg_tc16_overflow(int reset) {
  unsigned int v = 0, v2 = 0, v3 = 0;
  int rc = sgg::tc16::overflow_flag(reset, &v);
  if (rc == 0) rc = sgg::lin16p::overflow_flag(reset, &v2);     // pre-split LINEAR (lin16p.cu): emitted planes
  if (rc == 0) rc = sgg::roi_overflow_flag(reset, &v3);         // RoIAlign rows emitted as operand planes
  unsigned int v4 = 0;
  if (rc == 0) rc = sgg::bcast_overflow_flag(reset, &v4);       // pools + geom rows emitted as operand planes (train)
  return rc ? -1 : (int)(v | v2 | v3 | v4);
}
