// Training tail of the path: classification losses, global-norm gradient clipping and the
// SGD-with-momentum step, each as ONE launch over all rows / all parameter tensors.
//
// Reference call sites:
//   lib/losses.py:5-70   edge_losses  ('baseline', 'dnorm', 'dnorm-fgbg' weighting of the per-edge cross entropy)
//   lib/losses.py:73-74  node_losses  (mean cross entropy)
//   lib/pytorch_misc.py:70-73,625-664  grad_clip / clip_grad_norm (one norm over all gradients, in-place scaling)
//   lib/pytorch_misc.py:130-157        get_optim: optim.SGD(weight_decay, momentum=0.9), lr/10 group for roi_fmap*
//   main.py:105-120      losses -> backward -> grad_clip -> optimizer.step
//
// All three are HBM-bound streaming work.  The reference issues ~6 elementwise kernels per parameter tensor per step
// (norm, mul_, add wd, mul/add momentum, add to param: ~240 launches over 40 tensors, ~44 B/param); here a step is
// 2 sweeps: squared-norm partials (4 B/param) and the fused clip + weight-decay + momentum + update sweep
// (12 B read + 8 B written per param, + 4 B when it also emits the fp16 [hi | lo] tensor-core operand split of the
// new weight so the next forward does not have to re-split it).  Reductions use fixed-order trees over per-chunk
// partials: results are deterministic and independent of the grid size.
#include <cuda_fp16.h>
#include "common.cuh"

namespace sgg {
namespace train {

constexpr int NT = 256;                 // threads per CTA
constexpr int CHUNK = NT * 4 * 4;       // elements per (tensor, chunk): 4 float4 per thread
constexpr int MAX_TENSORS = 1024;
static_assert(sizeof(sgg_mt_tensor) == 56, "sgg_mt_tensor layout is part of the ABI (ctypes mirror in _lib.py)");
constexpr float LO_SCALE = 2048.0f;     // must match tc16_gemm.cu (lo half is stored scaled by 2^11)

// Device-side table: [sgg_mt_tensor x n][int chunk_start x (n + 1)] (chunk_start = prefix sum of ceil(n_i / CHUNK))
__host__ __device__ inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }
struct TableView {
  const sgg_mt_tensor *t;
  const int *chunk_start;
};
__host__ __device__ inline size_t table_bytes(int n) {
  return al256((size_t)n * sizeof(sgg_mt_tensor)) + al256((size_t)(n + 1) * sizeof(int));
}
__host__ __device__ inline TableView table_view(const void *ws, int n) {
  TableView v;
  v.t = reinterpret_cast<const sgg_mt_tensor *>(ws);
  v.chunk_start = reinterpret_cast<const int *>((const char *)ws + al256((size_t)n * sizeof(sgg_mt_tensor)));
  return v;
}

// which tensor owns global chunk c (largest i with chunk_start[i] <= c); the prefix array is small and L1-resident
__device__ __forceinline__ int find_tensor(const int *__restrict__ cs, int n, int c) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(cs + mid) <= c) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ float block_sum(float v, float *sh) {
  v = sgg_warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.f;
    r = sgg_warp_sum(r);
  }
  __syncthreads();
  return r;   // valid in warp 0
}

// ---------------------------------------------------------------------------------------------------------------
// sweep 1: per-chunk sum of squares of the gradients
__global__ void __launch_bounds__(NT) k_mt_sqnorm(const void *__restrict__ table, int n_tensors, int total_chunks,
                                                  float *__restrict__ partials) {
  __shared__ float sh[NT / 32];
  const TableView tv = table_view(table, n_tensors);
  for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const int ti = find_tensor(tv.chunk_start, n_tensors, c);
    const sgg_mt_tensor T = tv.t[ti];
    const long long base = (long long)(c - __ldg(tv.chunk_start + ti)) * CHUNK;
    const long long rem = T.n - base;
    const int cnt = rem < CHUNK ? (int)rem : CHUNK;
    const float *g = T.g + base;
    float s = 0.f;
    if (T.g != nullptr) {
      if (cnt == CHUNK && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4 *>(g) + u * NT + threadIdx.x);
#pragma unroll
        for (int u = 0; u < 4; ++u) s += v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
      } else {
        for (int i = threadIdx.x; i < cnt; i += NT) { const float x = g[i]; s += x * x; }
      }
    }
    const float r = block_sum(s, sh);
    if (threadIdx.x == 0) partials[c] = r;
  }
}

// fixed-order reduction of the partials -> norm_out = {total_norm, clip_coef, scale actually applied, sum of squares}
// (lib/pytorch_misc.py:651-653: clip_coef = max_norm / (total_norm + 1e-6), applied only when < 1)
// grad_scale: the gradients in memory are grad_scale^-1 times the ones the step means (data-parallel SUM instead of
// the average): total_norm and the clip factor are those of the scaled gradients, and the applied scale includes it.
__global__ void __launch_bounds__(1024) k_mt_norm_finish(const float *__restrict__ partials, int total_chunks,
                                                          float max_norm, float grad_scale, float *__restrict__ norm_out) {
  __shared__ double shd[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < total_chunks; i += 1024) s += (double)partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double r = shd[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (threadIdx.x == 0) {
      const float tn = (float)sqrt(r) * grad_scale;
      const float coef = max_norm > 0.f ? max_norm / (tn + 1e-6f) : 1.f;
      norm_out[0] = tn;
      norm_out[1] = coef;
      norm_out[2] = ((max_norm > 0.f && coef < 1.f) ? coef : 1.f) * grad_scale;
      norm_out[3] = (float)r;
    }
  }
}

// grad_clip drop-in: g *= scale (scale read from the device; a no-op sweep is skipped when scale == 1)
__global__ void __launch_bounds__(NT) k_mt_scale_grads(const void *__restrict__ table, int n_tensors, int total_chunks,
                                                       const float *__restrict__ norm) {
  const float scale = norm[2];
  if (scale == 1.f) return;
  const TableView tv = table_view(table, n_tensors);
  for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const int ti = find_tensor(tv.chunk_start, n_tensors, c);
    const sgg_mt_tensor T = tv.t[ti];
    if (T.g == nullptr) continue;
    const long long base = (long long)(c - __ldg(tv.chunk_start + ti)) * CHUNK;
    const long long rem = T.n - base;
    const int cnt = rem < CHUNK ? (int)rem : CHUNK;
    float *g = const_cast<float *>(T.g) + base;
    if (cnt == CHUNK && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 v = reinterpret_cast<float4 *>(g)[u * NT + threadIdx.x];
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        reinterpret_cast<float4 *>(g)[u * NT + threadIdx.x] = v;
      }
    } else {
      for (int i = threadIdx.x; i < cnt; i += NT) g[i] *= scale;
    }
  }
}

// sweep 2: torch.optim.SGD semantics (dampening 0, no nesterov), gradient pre-scaled by the clip factor:
//   d = g * scale + wd * p;   m = first_step ? d : momentum * m + d;   p -= lr * m
// The arithmetic keeps torch's operation sequence (add-with-alpha = one fused multiply-add, the momentum update a
// separate multiply and add), so a step agrees with torch.optim.SGD on the same inputs to the last bit or two.
__device__ __forceinline__ float sgd_one(float p, float g, float &m, float scale, float wd, float mom, float lr,
                                         bool first, bool clip) {
  float d = clip ? __fmul_rn(g, scale) : g;
  if (wd != 0.f) d = fmaf(wd, p, d);                       // torch: grad.add(param, alpha=weight_decay)
  m = first ? d : __fadd_rn(__fmul_rn(m, mom), d);         //        buf.mul_(momentum).add_(grad)
  return fmaf(-lr, m, p);                                  //        param.add_(buf, alpha=-lr)
}

template <bool SPLIT>
__device__ __forceinline__ void sgd_chunk_vec(const sgg_mt_tensor &T, long long base, float scale, float mom, bool clip,
                                              bool write_g) {
  float4 *p4 = reinterpret_cast<float4 *>(T.p + base);
  float4 *m4 = reinterpret_cast<float4 *>(T.m + base);
  float4 *g4 = reinterpret_cast<float4 *>(const_cast<float *>(T.g) + base);
  const bool first = (T.flags & 1) != 0;
  float4 p[4], g[4], m[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = u * NT + threadIdx.x;
    p[u] = p4[i];
    g[u] = __ldcs(g4 + i);
    m[u] = first ? make_float4(0.f, 0.f, 0.f, 0.f) : m4[i];
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = u * NT + threadIdx.x;
    p[u].x = sgd_one(p[u].x, g[u].x, m[u].x, scale, T.wd, mom, T.lr, first, clip);
    p[u].y = sgd_one(p[u].y, g[u].y, m[u].y, scale, T.wd, mom, T.lr, first, clip);
    p[u].z = sgd_one(p[u].z, g[u].z, m[u].z, scale, T.wd, mom, T.lr, first, clip);
    p[u].w = sgd_one(p[u].w, g[u].w, m[u].w, scale, T.wd, mom, T.lr, first, clip);
    p4[i] = p[u];
    m4[i] = m[u];
    if (write_g && clip) {
      g[u].x *= scale; g[u].y *= scale; g[u].z *= scale; g[u].w *= scale;
      g4[i] = g[u];
    }
    if (SPLIT) {   // fp16 [hi | lo * 2^11] operand split of the NEW weight (tc16_gemm.cu: k_tc16_split)
      __half *hi = reinterpret_cast<__half *>(T.split) + base + 4 * i;
      __half *lo = hi + T.n;
      const float pv[4] = {p[u].x, p[u].y, p[u].z, p[u].w};
      __half h[4], l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        h[k] = __float2half_rn(pv[k]);
        l[k] = __float2half_rn((pv[k] - __half2float(h[k])) * LO_SCALE);
      }
      *reinterpret_cast<uint2 *>(hi) = make_uint2(
          (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16),
          (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16));
      *reinterpret_cast<uint2 *>(lo) = make_uint2(
          (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16),
          (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16));
    }
  }
}

__global__ void __launch_bounds__(NT, 3) k_mt_sgd(const void *__restrict__ table, int n_tensors, int total_chunks,
                                               const float *__restrict__ norm, float momentum, int write_g) {
  const TableView tv = table_view(table, n_tensors);
  const float scale = norm != nullptr ? norm[2] : 1.f;
  const bool clip = scale != 1.f;
  for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const int ti = find_tensor(tv.chunk_start, n_tensors, c);
    const sgg_mt_tensor T = tv.t[ti];
    if (T.g == nullptr || T.p == nullptr) continue;   // no gradient this step (torch skips it) / gradients-only row
    const long long base = (long long)(c - __ldg(tv.chunk_start + ti)) * CHUNK;
    const long long rem = T.n - base;
    const int cnt = rem < CHUNK ? (int)rem : CHUNK;
    const uintptr_t al = reinterpret_cast<uintptr_t>(T.p + base) | reinterpret_cast<uintptr_t>(T.g + base) |
                         reinterpret_cast<uintptr_t>(T.m + base);
    // the split halves need n % 8 == 0 (checked on the host), so hi/lo stay 8-byte aligned per float4 group
    if (cnt == CHUNK && (al & 15) == 0) {
      if (T.split != nullptr) sgd_chunk_vec<true>(T, base, scale, momentum, clip, write_g != 0);
      else sgd_chunk_vec<false>(T, base, scale, momentum, clip, write_g != 0);
    } else {
      const bool first = (T.flags & 1) != 0;
      float *gp = const_cast<float *>(T.g) + base;
      for (int i = threadIdx.x; i < cnt; i += NT) {
        float m = first ? 0.f : T.m[base + i];
        const float g = gp[i];
        const float pn = sgd_one(T.p[base + i], g, m, scale, T.wd, momentum, T.lr, first, clip);
        T.p[base + i] = pn;
        T.m[base + i] = m;
        if (write_g && clip) gp[i] = g * scale;
        if (T.split != nullptr) {
          __half *hi = reinterpret_cast<__half *>(T.split) + base + i;
          const __half h = __float2half_rn(pn);
          hi[0] = h;
          hi[T.n] = __float2half_rn((pn - __half2float(h)) * LO_SCALE);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Losses.  ws ints: [0] M_FG (label > 0), [1] M_BG (label == 0), [2] rows that count for the mean (label != -100),
// [3] error flag (label out of [0, C) and not -100).
__global__ void k_ce_count(const int64_t *__restrict__ labels, const int8_t *__restrict__ category, int M, int C,
                           int *__restrict__ counts) {
  int fg = 0, bg = 0, valid = 0, bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    const long long l = labels[i];
    if (category != nullptr) { fg += category[i] == 1; bg += category[i] == 2; }
    else { fg += l > 0; bg += l == 0; }
    valid += l != -100;
    bad += (l != -100) && (l < 0 || l >= C);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    fg += __shfl_xor_sync(0xffffffffu, fg, o); bg += __shfl_xor_sync(0xffffffffu, bg, o);
    valid += __shfl_xor_sync(0xffffffffu, valid, o); bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {     // integer atomics: order-independent
    if (fg) atomicAdd(counts + 0, fg);
    if (bg) atomicAdd(counts + 1, bg);
    if (valid) atomicAdd(counts + 2, valid);
    if (bad) atomicAdd(counts + 3, bad);
  }
}

// one warp per row: log-softmax, weighted negative log-likelihood, d loss / d logits
__global__ void __launch_bounds__(256) k_ce_rows(const float *__restrict__ logits, const int64_t *__restrict__ labels,
                                                 const int8_t *__restrict__ category, int M, int C, int mode,
                                                 float alpha, float beta, float gamma, const int *__restrict__ counts,
                                                 float *__restrict__ row_loss, float *__restrict__ dlogits) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int m_fg = counts[0], m_bg = counts[1], valid = counts[2];
  const long long lab = labels[row];
  const bool ignored = lab == -100 || lab < 0 || lab >= C;
  // per-row weight (lib/losses.py:38-62)
  float w;
  if (mode == SGG_LOSS_MEAN) {
    w = valid > 0 ? 1.f / (float)valid : 0.f;
  } else if (mode == SGG_LOSS_BASELINE) {
    w = gamma / (float)M;
  } else {
    const int cat = category != nullptr ? (int)category[row] : (lab > 0 ? 1 : (lab == 0 ? 2 : 0));
    w = 1.f;                                                      // edge_weights = torch.ones(M)
    if (cat == 1 && m_fg > 0) w = alpha / (float)m_fg;
    if (cat == 2) {
      if (mode == SGG_LOSS_DNORM) { if (m_bg > 0 && m_fg > 0) w = beta / (float)m_fg; }
      else if (m_bg > 0) w = beta / (float)m_bg;
    }
    w *= gamma;
  }
  const float *x = logits + (size_t)row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(x[c] - mx);
  se = sgg_warp_sum(se);
  const float lse = logf(se) + mx;
  if (lane == 0) row_loss[row] = ignored ? 0.f : w * (lse - x[lab]);
  if (dlogits != nullptr) {
    float *d = dlogits + (size_t)row * C;
    const float inv = 1.f / se;
    for (int c = lane; c < C; c += 32) {
      const float sm = expf(x[c] - mx) * inv;
      d[c] = ignored ? 0.f : w * (sm - (c == (int)lab ? 1.f : 0.f));
    }
  }
}

__global__ void __launch_bounds__(1024) k_ce_finish(const float *__restrict__ row_loss, int M, const int *__restrict__ cnt,
                                                    float *__restrict__ loss, int *__restrict__ counts_out) {
  __shared__ double shd[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < M; i += 1024) s += (double)row_loss[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double r = shd[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (threadIdx.x == 0) {
      loss[0] = (float)r;
      if (counts_out != nullptr) { counts_out[0] = cnt[0]; counts_out[1] = cnt[1]; counts_out[2] = cnt[2]; counts_out[3] = cnt[3]; }
    }
  }
}

static int grid_for(int total_chunks, int ctas_per_sm) {
  const int cap = sgg_num_sms() * ctas_per_sm;   // resident CTAs of 256 threads per SM (register-limited: k_mt_sgd 3, others 8)
  return total_chunks < cap ? (total_chunks > 0 ? total_chunks : 1) : cap;
}

}  // namespace train
}  // namespace sgg

using namespace sgg::train;

extern "C" int sgg_mt_chunk_elems(void) { return CHUNK; }

extern "C" size_t sgg_mt_table_bytes(int n_tensors) { return table_bytes(n_tensors > 0 ? n_tensors : 1); }

extern "C" long long sgg_mt_total_chunks(const sgg_mt_tensor *host, int n_tensors) {
  long long c = 0;
  for (int i = 0; i < n_tensors; ++i) c += (host[i].n + CHUNK - 1) / CHUNK;
  return c;
}

extern "C" size_t sgg_mt_workspace_bytes(long long total_chunks) {
  return sgg_align_up((size_t)(total_chunks > 0 ? total_chunks : 1) * sizeof(float));
}

extern "C" int sgg_mt_table_upload(const sgg_mt_tensor *host, int n_tensors, void *table, size_t table_bytes_,
                                   void *stream) {
  if (n_tensors <= 0 || n_tensors > MAX_TENSORS || !host || !table)
    return sgg_set_err(SGG_E_BADARG, "mt_table_upload: need 1..%d tensors and non-null pointers", MAX_TENSORS);
  if (table_bytes_ < table_bytes(n_tensors)) return sgg_set_err(SGG_E_WORKSPACE, "mt_table_upload: table too small");
  int cs[MAX_TENSORS + 1];
  long long c = 0;
  for (int i = 0; i < n_tensors; ++i) {
    // p / m may both be NULL in a gradients-only table (sgg_mt_grad_norm / sgg_mt_scale_grads); sgd skips such rows
    if (host[i].n < 0 || ((host[i].p == nullptr) != (host[i].m == nullptr)))
      return sgg_set_err(SGG_E_BADARG, "mt_table_upload: tensor %d: n < 0 or only one of p / m given", i);
    if (host[i].split && (host[i].n & 7)) return sgg_set_err(SGG_E_BADARG, "mt_table_upload: tensor %d: split needs n %% 8 == 0", i);
    cs[i] = (int)c;
    c += (host[i].n + CHUNK - 1) / CHUNK;
    if (c > 0x7fffffffLL) return sgg_set_err(SGG_E_BADARG, "mt_table_upload: too many chunks");
  }
  cs[n_tensors] = (int)c;
  const TableView tv = table_view(table, n_tensors);
  // pageable sources: cudaMemcpyAsync stages them before returning, so the stack / caller buffers may die afterwards
  SGG_CUDA_TRY(cudaMemcpyAsync((void *)tv.t, host, (size_t)n_tensors * sizeof(sgg_mt_tensor), cudaMemcpyHostToDevice,
                               (cudaStream_t)stream));
  SGG_CUDA_TRY(cudaMemcpyAsync((void *)tv.chunk_start, cs, (size_t)(n_tensors + 1) * sizeof(int), cudaMemcpyHostToDevice,
                               (cudaStream_t)stream));
  return 0;
}

extern "C" int sgg_mt_grad_norm_scaled(const void *table, int n_tensors, long long total_chunks, float max_norm,
                                       float grad_scale, float *norm_out, void *ws, size_t ws_bytes, void *stream);
extern "C" int sgg_mt_grad_norm(const void *table, int n_tensors, long long total_chunks, float max_norm,
                                float *norm_out, void *ws, size_t ws_bytes, void *stream) {
  return sgg_mt_grad_norm_scaled(table, n_tensors, total_chunks, max_norm, 1.0f, norm_out, ws, ws_bytes, stream);
}
extern "C" int sgg_mt_grad_norm_scaled(const void *table, int n_tensors, long long total_chunks, float max_norm,
                                       float grad_scale, float *norm_out, void *ws, size_t ws_bytes, void *stream) {
  if (!table || !norm_out || !ws || n_tensors <= 0 || n_tensors > MAX_TENSORS)
    return sgg_set_err(SGG_E_BADARG, "mt_grad_norm: bad argument");
  if (ws_bytes < sgg_mt_workspace_bytes(total_chunks)) return sgg_set_err(SGG_E_WORKSPACE, "mt_grad_norm: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (total_chunks > 0) {
    k_mt_sqnorm<<<grid_for((int)total_chunks, 8), NT, 0, st>>>(table, n_tensors, (int)total_chunks, (float *)ws);
    SGG_RETURN_IF_LAUNCH_FAILED("k_mt_sqnorm");
  }
  k_mt_norm_finish<<<1, 1024, 0, st>>>((const float *)ws, (int)total_chunks, max_norm, grad_scale, norm_out);
  SGG_RETURN_IF_LAUNCH_FAILED("k_mt_norm_finish");
  return 0;
}

extern "C" int sgg_mt_scale_grads(const void *table, int n_tensors, long long total_chunks, const float *norm,
                                  void *stream) {
  if (!table || !norm || n_tensors <= 0 || n_tensors > MAX_TENSORS) return sgg_set_err(SGG_E_BADARG, "mt_scale_grads: bad argument");
  if (total_chunks <= 0) return 0;
  k_mt_scale_grads<<<grid_for((int)total_chunks, 8), NT, 0, (cudaStream_t)stream>>>(table, n_tensors, (int)total_chunks, norm);
  SGG_RETURN_IF_LAUNCH_FAILED("k_mt_scale_grads");
  return 0;
}

extern "C" int sgg_mt_sgd_step(const void *table, int n_tensors, long long total_chunks, const float *norm,
                               float momentum, int write_clipped_grads, void *stream) {
  if (!table || n_tensors <= 0 || n_tensors > MAX_TENSORS) return sgg_set_err(SGG_E_BADARG, "mt_sgd_step: bad argument");
  if (total_chunks <= 0) return 0;
  k_mt_sgd<<<grid_for((int)total_chunks, 3), NT, 0, (cudaStream_t)stream>>>(table, n_tensors, (int)total_chunks, norm, momentum,
                                                                         write_clipped_grads);
  SGG_RETURN_IF_LAUNCH_FAILED("k_mt_sgd");
  return 0;
}

extern "C" size_t sgg_ce_loss_workspace_bytes(int M) {
  return sgg_align_up(4 * sizeof(int)) + sgg_align_up((size_t)(M > 0 ? M : 1) * sizeof(float));
}

extern "C" int sgg_ce_loss(const float *logits, const int64_t *labels, const int8_t *category, int M, int C, int mode,
                           float alpha, float beta, float gamma, float *loss, float *dlogits, int *counts_out,
                           void *ws, size_t ws_bytes, void *stream) {
  if (M < 0 || C <= 0 || !loss || !ws || (M > 0 && (!logits || !labels)))
    return sgg_set_err(SGG_E_BADARG, "ce_loss: bad argument");
  if (mode < SGG_LOSS_MEAN || mode > SGG_LOSS_DNORM_FGBG) return sgg_set_err(SGG_E_BADARG, "ce_loss: unknown mode %d", mode);
  if (mode == SGG_LOSS_BASELINE && !(alpha == 1.f && beta == 1.f))     // lib/losses.py:42 asserts this
    return sgg_set_err(SGG_E_BADARG, "ce_loss: 'baseline' requires alpha == beta == 1");
  if (ws_bytes < sgg_ce_loss_workspace_bytes(M)) return sgg_set_err(SGG_E_WORKSPACE, "ce_loss: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int *cnt = (int *)ws;
  float *row_loss = (float *)((char *)ws + sgg_align_up(4 * sizeof(int)));
  SGG_CUDA_TRY(cudaMemsetAsync(cnt, 0, 4 * sizeof(int), st));
  if (M > 0) {
    const int cb = (M + 255) / 256 < 592 ? (M + 255) / 256 : 592;
    k_ce_count<<<cb, 256, 0, st>>>(labels, category, M, C, cnt);
    SGG_RETURN_IF_LAUNCH_FAILED("k_ce_count");
    k_ce_rows<<<(M + 7) / 8, 256, 0, st>>>(logits, labels, category, M, C, mode, alpha, beta, gamma, cnt, row_loss, dlogits);
    SGG_RETURN_IF_LAUNCH_FAILED("k_ce_rows");
  }
  k_ce_finish<<<1, 1024, 0, st>>>(row_loss, M, cnt, loss, counts_out);
  SGG_RETURN_IF_LAUNCH_FAILED("k_ce_finish");
  return 0;
}
