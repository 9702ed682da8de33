// Evaluation tail: lib/surgery.py:17-55 filter_dets on the device.
//
//   triple score  = max_{p >= 1} softmax(rel_dists)[p] * obj_scores[subj] * obj_scores[obj]      (:42-48)
//   order         = argsort(score, descending)                                                      (:49)
//   rels, preds   = rel_inds[order], softmax(rel_dists)[order]                                      (:51-52)
//
// The reference runs softmax, two gathers, max, two multiplies, torch.sort (unstable) and two index ops as separate
// kernels and then five D2H copies per image; here: one scoring kernel (softmax fused), a bitonic sort of
// (image, score, edge id) keys — ties broken by edge id, so the order is deterministic — and one gather kernel.
// With an image column the same launch ranks every image of a batch at once (edges grouped by image, each group
// sorted by descending score), which is what makes batched evaluation possible (the reference evaluates one image
// per batch, dataloaders/visual_genome.py:730).
#include "common.cuh"

namespace sgg {
namespace rank {

constexpr int TILE = 2048;          // elements sorted inside one CTA's shared memory (1024 threads, 2 per thread)

struct Item { unsigned long long key; unsigned int idx; };

__device__ __forceinline__ bool item_less(unsigned long long ka, unsigned int ia, unsigned long long kb, unsigned int ib) {
  return ka < kb || (ka == kb && ia < ib);
}

// ascending uint order of -score: larger scores first; NaNs (never produced by a softmax) sort last
__device__ __forceinline__ unsigned int desc_bits(float s) {
  unsigned int b = __float_as_uint(s);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);     // ascending-ordered bits
  return ~b;
}

// one warp per edge: softmax (optional), triple score, sort key
__global__ void __launch_bounds__(256) k_rank_scores(const float *__restrict__ rel_dists, int apply_softmax,
                                                     const float *__restrict__ obj_scores,
                                                     const int64_t *__restrict__ rel_inds, int64_t stride, int col_img,
                                                     int col_subj, int col_obj, int N, int E, int P, int Epad,
                                                     float *__restrict__ probs, float *__restrict__ scores,
                                                     unsigned long long *__restrict__ keys, unsigned int *__restrict__ idx,
                                                     int *__restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= Epad) return;
  if (e >= E) {                                         // padding sorts to the end
    if (lane == 0) { keys[e] = ~0ull; idx[e] = 0xffffffffu; }
    return;
  }
  const float *x = rel_dists + (size_t)e * P;
  float best = -INFINITY;                               // max over p >= 1 of the probabilities
  if (apply_softmax) {
    float mx = -INFINITY;
    for (int c = lane; c < P; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int c = lane; c < P; c += 32) se += expf(x[c] - mx);
    se = sgg_warp_sum(se);
    for (int c = lane; c < P; c += 32) {
      const float pr = expf(x[c] - mx) / se;
      probs[(size_t)e * P + c] = pr;
      if (c >= 1) best = fmaxf(best, pr);
    }
  } else {
    for (int c = lane; c < P; c += 32) {
      const float pr = x[c];
      probs[(size_t)e * P + c] = pr;
      if (c >= 1) best = fmaxf(best, pr);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane == 0) {
    const long long s = rel_inds[(size_t)e * stride + col_subj], o = rel_inds[(size_t)e * stride + col_obj];
    const long long im = col_img >= 0 ? rel_inds[(size_t)e * stride + col_img] : 0;
    float sc = 0.f;
    if (s < 0 || s >= N || o < 0 || o >= N || im < 0 || im > 0x7fffffffLL) atomicExch(err, 1);
    else sc = best * obj_scores[s] * obj_scores[o];     // (max * s0) * s1, the reference's order (:48)
    scores[e] = sc;
    keys[e] = ((unsigned long long)(unsigned int)im << 32) | desc_bits(sc);
    idx[e] = (unsigned int)e;
  }
}

__device__ __forceinline__ void cmp_swap(unsigned long long &ka, unsigned int &ia, unsigned long long &kb,
                                         unsigned int &ib, bool up) {
  const bool a_less = item_less(ka, ia, kb, ib);
  if (a_less != up) {
    const unsigned long long tk = ka; ka = kb; kb = tk;
    const unsigned int ti = ia; ia = ib; ib = ti;
  }
}

// Bitonic network on n = 2^m elements.  Stage k (run length), stride j.  Element i is compared with i ^ j; the pair
// is put in ascending order iff (i & k) == 0.
// k_bitonic_tile: all (k, j) steps with k in [k_from, k_to] restricted to j < TILE, done in shared memory.
//   first call: k_from = 2, k_to = TILE (full sort of each tile); later, for k > TILE: only the tail j = TILE/2 .. 1.
__global__ void __launch_bounds__(1024) k_bitonic_tile(unsigned long long *__restrict__ keys, unsigned int *__restrict__ idx,
                                                       int n, int k_from, int k_to) {
  __shared__ unsigned long long sk[TILE];
  __shared__ unsigned int si[TILE];
  const int base = blockIdx.x * TILE;
  for (int t = threadIdx.x; t < TILE; t += blockDim.x) {
    const int i = base + t;
    sk[t] = i < n ? keys[i] : ~0ull;
    si[t] = i < n ? idx[i] : 0xffffffffu;
  }
  __syncthreads();
  for (int k = k_from; k <= k_to; k <<= 1) {
    for (int j = (k >> 1) < TILE ? (k >> 1) : (TILE >> 1); j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < TILE / 2; t += blockDim.x) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));     // index with bit j clear
        const int hi = lo | j;
        const bool up = ((base + lo) & k) == 0;
        cmp_swap(sk[lo], si[lo], sk[hi], si[hi], up);
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < TILE; t += blockDim.x) {
    const int i = base + t;
    if (i < n) { keys[i] = sk[t]; idx[i] = si[t]; }
  }
}

// one global compare-exchange pass (stride j >= TILE)
__global__ void __launch_bounds__(256) k_bitonic_global(unsigned long long *__restrict__ keys, unsigned int *__restrict__ idx,
                                                        int n, int k, int j) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n / 2) return;
  const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
  const int hi = lo | j;
  unsigned long long ka = keys[lo], kb = keys[hi];
  unsigned int ia = idx[lo], ib = idx[hi];
  const bool up = (lo & k) == 0;
  const bool a_less = item_less(ka, ia, kb, ib);
  if (a_less != up) { keys[lo] = kb; keys[hi] = ka; idx[lo] = ib; idx[hi] = ia; }
}

// one warp per output row: rels_out[i] = rel_inds[order[i]] (subject, object), pred_out[i] = probs[order[i]]
__global__ void __launch_bounds__(256) k_rank_gather(const unsigned int *__restrict__ idx, const int64_t *__restrict__ rel_inds,
                                                     int64_t stride, int col_subj, int col_obj,
                                                     const float *__restrict__ probs, const float *__restrict__ scores,
                                                     int E, int P, int64_t *__restrict__ rels_out,
                                                     float *__restrict__ pred_out, float *__restrict__ score_out,
                                                     int *__restrict__ order_out) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= E) return;
  const unsigned int src = idx[i];
  if (src >= (unsigned int)E) return;                   // cannot happen (padding sorts behind the E real rows)
  if (lane == 0) {
    rels_out[(size_t)i * 2] = rel_inds[(size_t)src * stride + col_subj];
    rels_out[(size_t)i * 2 + 1] = rel_inds[(size_t)src * stride + col_obj];
    if (score_out != nullptr) score_out[i] = scores[src];
    if (order_out != nullptr) order_out[i] = (int)src;
  }
  for (int c = lane; c < P; c += 32) pred_out[(size_t)i * P + c] = probs[(size_t)src * P + c];
}

static int pow2_at_least(int n) { int p = 2; while (p < n) p <<= 1; return p; }

struct View { int *err; float *probs, *scores; unsigned long long *keys; unsigned int *idx; size_t bytes; };
static View view(void *ws, int E, int P) {
  SggArena a(ws, (size_t)-1);
  View v;
  const size_t e1 = E > 0 ? E : 1, ep = pow2_at_least(E > 0 ? E : 1);
  v.err = a.take<int>(4);
  v.probs = a.take<float>(e1 * P);
  v.scores = a.take<float>(e1);
  v.keys = a.take<unsigned long long>(ep);
  v.idx = a.take<unsigned int>(ep);
  v.bytes = a.off;
  return v;
}

}  // namespace rank
}  // namespace sgg

extern "C" size_t sgg_rank_relations_workspace_bytes(int E, int P) {
  return sgg::rank::view(nullptr, E < 0 ? 0 : E, P < 1 ? 1 : P).bytes;
}

extern "C" int sgg_rank_relations(const float *rel_dists, int apply_softmax, const float *obj_scores,
                                  const int64_t *rel_inds, int64_t row_stride, int col_img, int col_subj, int col_obj,
                                  int N, int E, int P, int64_t *rels_out, float *pred_out, float *score_out,
                                  int *order_out, void *ws, size_t ws_bytes, void *stream) {
  using namespace sgg::rank;
  if (E < 0 || N < 0 || P < 2 || !ws) return sgg_set_err(SGG_E_BADARG, "rank_relations: bad shape / null workspace");
  if (E == 0) return 0;
  if (!rel_dists || !obj_scores || !rel_inds || !rels_out || !pred_out)
    return sgg_set_err(SGG_E_BADARG, "rank_relations: null pointer");
  if (ws_bytes < sgg_rank_relations_workspace_bytes(E, P)) return sgg_set_err(SGG_E_WORKSPACE, "rank_relations: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  View v = view(ws, E, P);
  const int n = pow2_at_least(E);
  SGG_CUDA_TRY(cudaMemsetAsync(v.err, 0, 4 * sizeof(int), st));
  k_rank_scores<<<(n + 7) / 8, 256, 0, st>>>(rel_dists, apply_softmax, obj_scores, rel_inds, row_stride, col_img, col_subj,
                                            col_obj, N, E, P, n, v.probs, v.scores, v.keys, v.idx, v.err);
  SGG_RETURN_IF_LAUNCH_FAILED("k_rank_scores");
  const int tiles = (n + TILE - 1) / TILE;
  k_bitonic_tile<<<tiles, 1024, 0, st>>>(v.keys, v.idx, n, 2, n < TILE ? n : TILE);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bitonic_tile");
  for (int k = 2 * TILE; k <= n; k <<= 1) {
    for (int j = k >> 1; j >= TILE; j >>= 1) {
      k_bitonic_global<<<(n / 2 + 255) / 256, 256, 0, st>>>(v.keys, v.idx, n, k, j);
      SGG_RETURN_IF_LAUNCH_FAILED("k_bitonic_global");
    }
    k_bitonic_tile<<<tiles, 1024, 0, st>>>(v.keys, v.idx, n, k, k);      // strides TILE/2 .. 1 of stage k
    SGG_RETURN_IF_LAUNCH_FAILED("k_bitonic_tile");
  }
  k_rank_gather<<<(E + 7) / 8, 256, 0, st>>>(v.idx, rel_inds, row_stride, col_subj, col_obj, v.probs, v.scores, E, P,
                                            rels_out, pred_out, score_out, order_out);
  SGG_RETURN_IF_LAUNCH_FAILED("k_rank_gather");
  return 0;
}

// host-synchronising check (tests / debug): SGG_E_INDEX if an endpoint was outside [0, N) in the last call
extern "C" int sgg_rank_relations_check(const void *ws, void *stream) {
  int e[4] = {0, 0, 0, 0};
  SGG_CUDA_TRY(cudaMemcpyAsync(e, ws, sizeof(e), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SGG_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return e[0] ? sgg_set_err(SGG_E_INDEX, "rank_relations: rel_inds endpoint outside [0, N)") : 0;
}
