// Shared helpers for the sgg_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/sgg_b200.h"

int sgg_set_err(int code, const char *fmt, ...);
void sgg_count_launch();

#define SGG_RETURN_IF_LAUNCH_FAILED(name)                                              \
  do {                                                                                 \
    cudaError_t e_ = cudaGetLastError();                                               \
    if (e_ != cudaSuccess) return sgg_set_err((int)e_, "%s: %s", name, cudaGetErrorString(e_)); \
    sgg_count_launch();                                                                \
  } while (0)

#define SGG_CUDA_TRY(expr)                                                             \
  do {                                                                                 \
    cudaError_t e_ = (expr);                                                           \
    if (e_ != cudaSuccess) return sgg_set_err((int)e_, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

static inline size_t sgg_align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (never allocates).
struct SggArena {
  char *base; size_t cap; size_t off;
  SggArena(void *p, size_t bytes) : base((char *)p), cap(bytes), off(0) {}
  template <typename T> T *take(size_t n) {
    size_t b = sgg_align_up(n * sizeof(T));
    char *p = base ? base + off : nullptr;
    off += b;
    return (T *)p;
  }
  bool ok() const { return off <= cap; }
};

int sgg_num_sms();

__device__ __forceinline__ float sgg_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float sgg_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Layout of the graph workspace written by sgg_graph_build (graph.cu).
struct SggGraphView {
  int *err;                 // [4]  err[0] != 0: an index was out of range
  int *out_ptr, *in_ptr;    // [N+1] CSR row pointers (by subject / by object)
  int *cur_s, *cur_o;       // [N]   fill cursors (scratch)
  int *subj, *obj;          // [E]   int32 copies of the endpoints
  int *out_idx, *in_idx;    // [E]   edge ids grouped by subject / object, ascending within a node
  size_t zero_bytes;        // leading bytes that must be zeroed before counting
  size_t total_bytes;
};
static inline SggGraphView sgg_graph_view(const void *ws, int N, int E) {
  SggArena a((void *)ws, (size_t)-1);
  SggGraphView g;
  const int e1 = E > 0 ? E : 1, n1 = N > 0 ? N : 1;
  g.err = a.take<int>(4);
  g.out_ptr = a.take<int>(N + 1);
  g.in_ptr = a.take<int>(N + 1);
  g.cur_s = a.take<int>(n1);
  g.cur_o = a.take<int>(n1);
  g.zero_bytes = a.off;
  g.subj = a.take<int>(e1);
  g.obj = a.take<int>(e1);
  g.out_idx = a.take<int>(e1);
  g.in_idx = a.take<int>(e1);
  g.total_bytes = a.off;
  return g;
}
