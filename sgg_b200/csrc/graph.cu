// Ragged graph index for message passing: replaces the two dense [N,E] incidence
// matrices the reference builds per forward (sgg_models/rel_model_stanford.py:58-66)
// with int32 endpoints + CSR-by-subject + CSR-by-object, O(N+E) memory.
// Per-node edge lists are sorted by edge id => every segment reduction that walks
// them has a fixed summation order (deterministic, no float atomics anywhere).
#include "common.cuh"

namespace sgg {

__global__ void k_graph_count(const int64_t *__restrict__ rel, int64_t stride, int cs, int co, int N, int E,
                              int *__restrict__ subj, int *__restrict__ obj, int *out_ptr, int *in_ptr, int *err) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = rel[e * stride + cs], o = rel[e * stride + co];
  if (s < 0 || s >= N || o < 0 || o >= N) {
    atomicExch(err, 1);
    s = s < 0 ? 0 : (s >= N ? N - 1 : s);
    o = o < 0 ? 0 : (o >= N ? N - 1 : o);
  }
  subj[e] = (int)s; obj[e] = (int)o;
  atomicAdd(out_ptr + s + 1, 1);
  atomicAdd(in_ptr + o + 1, 1);
}

// Single-CTA inclusive scan of both pointer arrays (n = N+1 entries each; entry 0 stays 0).
__global__ void k_graph_scan(int *out_ptr, int *in_ptr, int n) {
  __shared__ int wsum[32];
  __shared__ int carry;
  for (int which = 0; which < 2; ++which) {
    int *p = which ? in_ptr : out_ptr;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
      int i = base + threadIdx.x;
      int v = i < n ? p[i] : 0;
      int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
      }
      if (lane == 31) wsum[wid] = x;
      __syncthreads();
      if (wid == 0) {
        int w = lane < (blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += t;
        }
        wsum[lane] = w;
      }
      __syncthreads();
      int prefix = carry + (wid > 0 ? wsum[wid - 1] : 0);
      if (i < n) p[i] = x + prefix;
      __syncthreads();
      if (threadIdx.x == blockDim.x - 1) carry += wsum[(blockDim.x >> 5) - 1];
      __syncthreads();
    }
  }
}

__global__ void k_graph_fill(const int *__restrict__ subj, const int *__restrict__ obj, int E,
                             const int *__restrict__ out_ptr, const int *__restrict__ in_ptr, int *cur_s, int *cur_o,
                             int *out_idx, int *in_idx) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int s = subj[e], o = obj[e];
  out_idx[out_ptr[s] + atomicAdd(cur_s + s, 1)] = e;
  in_idx[in_ptr[o] + atomicAdd(cur_o + o, 1)] = e;
}

// One thread per (node, side): insertion sort of a short list (degree <= ~2*(boxes-1)).
__global__ void k_graph_sort(const int *__restrict__ out_ptr, const int *__restrict__ in_ptr, int N, int *out_idx,
                             int *in_idx) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * N) return;
  int n = t >> 1;
  const int *ptr = (t & 1) ? in_ptr : out_ptr;
  int *idx = (t & 1) ? in_idx : out_idx;
  int b = ptr[n], e = ptr[n + 1];
  for (int i = b + 1; i < e; ++i) {
    int v = idx[i], j = i - 1;
    while (j >= b && idx[j] > v) { idx[j + 1] = idx[j]; --j; }
    idx[j + 1] = v;
  }
}

}  // namespace sgg

extern "C" size_t sgg_graph_workspace_bytes(int N, int E) {
  return sgg_graph_view(nullptr, N < 0 ? 0 : N, E < 0 ? 0 : E).total_bytes;
}

extern "C" int sgg_graph_build(const int64_t *rel_inds, int64_t row_stride, int col_subj, int col_obj, int N, int E,
                               void *graph_ws, size_t graph_ws_bytes, void *stream) {
  if (N < 0 || E < 0 || !graph_ws || (E > 0 && !rel_inds)) return sgg_set_err(SGG_E_BADARG, "graph_build: bad argument");
  if (E > 0 && N == 0) return sgg_set_err(SGG_E_BADARG, "graph_build: edges without objects");
  SggGraphView g = sgg_graph_view(graph_ws, N, E);
  if (g.total_bytes > graph_ws_bytes)
    return sgg_set_err(SGG_E_WORKSPACE, "graph_build: workspace %zu < %zu", graph_ws_bytes, g.total_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  SGG_CUDA_TRY(cudaMemsetAsync(graph_ws, 0, g.zero_bytes, st));
  if (E > 0) {
    sgg::k_graph_count<<<(E + 255) / 256, 256, 0, st>>>(rel_inds, row_stride, col_subj, col_obj, N, E, g.subj, g.obj,
                                                         g.out_ptr, g.in_ptr, g.err);
    SGG_RETURN_IF_LAUNCH_FAILED("k_graph_count");
    sgg::k_graph_scan<<<1, 1024, 0, st>>>(g.out_ptr, g.in_ptr, N + 1);
    SGG_RETURN_IF_LAUNCH_FAILED("k_graph_scan");
    sgg::k_graph_fill<<<(E + 255) / 256, 256, 0, st>>>(g.subj, g.obj, E, g.out_ptr, g.in_ptr, g.cur_s, g.cur_o,
                                                        g.out_idx, g.in_idx);
    SGG_RETURN_IF_LAUNCH_FAILED("k_graph_fill");
    sgg::k_graph_sort<<<(2 * N + 127) / 128, 128, 0, st>>>(g.out_ptr, g.in_ptr, N, g.out_idx, g.in_idx);
    SGG_RETURN_IF_LAUNCH_FAILED("k_graph_sort");
  }
  return 0;
}

extern "C" int sgg_graph_check(const void *graph_ws, int N, int E, void *stream) {
  if (!graph_ws) return sgg_set_err(SGG_E_BADARG, "graph_check: null workspace");
  SggGraphView g = sgg_graph_view(graph_ws, N, E);
  int flag = 0;
  SGG_CUDA_TRY(cudaMemcpyAsync(&flag, g.err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SGG_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  if (flag) return sgg_set_err(SGG_E_INDEX, "rel_inds contain an object id outside [0, %d)", N);
  return 0;
}
