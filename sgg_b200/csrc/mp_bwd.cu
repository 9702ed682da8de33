// Backward of the iterative message passing (BPTT through T iterations) — the autograd
// counterpart of sgg_models/rel_model_stanford.py:48-94, which the reference gets from
// torch.autograd over ~100 small kernels and two dense [N,E] incidence matmuls.
//
// Forward (mp.cu) saved on the tape, per iteration t: states V_t, E_t; GRU caches
// (r, z, n, gh_n) of the calls that produced V_{t+1}, E_{t+1}; edge gates g_t; ctx_t; P_t.
// Per iteration, in reverse:
//   k_gru_bwd          dgi, dgh, dh*z from dh' and the cache (GRUCell pointwise backward)
//   GEMMs (gemm.cu)    dW += dG^T X ; dX (+)= dG W ; db += colsum(dG)
//   k_edge_bwd         warp per edge: dg_sub/obj = <dgi_e, P[s|o]>, dg_out/in = <dctx[s|o], E[e]>,
//                      dlogit = dg * g (1-g); dE[e] += g_out dctx[s] + g_in dctx[o] + sum_k dlogit_k w_k[H:]
//   k_node_bwd         CTA per node (CSR, fixed order): dP[n] = sum_out g_sub dgi_e + sum_in g_obj dgi_e;
//                      da[n,k] = sum of edge dlogits; dV[n] += sum_k da[n,k] w_k[:H]
// All reductions are deterministic (no float atomics).
#include <string.h>
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"

namespace sgg {

// dgi [rows,3H], dgh [rows,3H], dh_prev [rows,H] (= dh' * z, overwritten)
// amax_gi / amax_gh (nullable, [gridDim.x]): per-block max |dgi|, max |dgh| — the abs-max pass of the scaled 3xFP16 GEMMs
// that consume these tensors, folded into their producer (blockDim.x = 256)
__global__ void k_gru_bwd(const float *__restrict__ dh_next, const float *__restrict__ cache,
                          const float *__restrict__ h_prev, int rows, int H, float *__restrict__ dgi,
                          float *__restrict__ dgh, float *__restrict__ dh_prev, float *__restrict__ amax_gi,
                          float *__restrict__ amax_gh) {
  const size_t total = (size_t)rows * (H / 4);
  float mi = 0.f, mh = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t m = i / (H / 4);
    const int j = (int)(i % (H / 4)) * 4;
    const float4 d4 = *reinterpret_cast<const float4 *>(dh_next + m * H + j);
    const float *cp = cache + m * 4 * H + j;
    const float4 r4 = *reinterpret_cast<const float4 *>(cp), z4 = *reinterpret_cast<const float4 *>(cp + H);
    const float4 n4 = *reinterpret_cast<const float4 *>(cp + 2 * H), g4 = *reinterpret_cast<const float4 *>(cp + 3 * H);
    float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h_prev != nullptr) h4 = *reinterpret_cast<const float4 *>(h_prev + m * H + j);
    const float d[4] = {d4.x, d4.y, d4.z, d4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w}, z[4] = {z4.x, z4.y, z4.z, z4.w};
    const float n[4] = {n4.x, n4.y, n4.z, n4.w}, gn[4] = {g4.x, g4.y, g4.z, g4.w}, h[4] = {h4.x, h4.y, h4.z, h4.w};
    float pr[4], pz[4], pn[4], pnr[4], dp[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float dn = d[c] * (1.f - z[c]);
      const float dz = d[c] * (h[c] - n[c]);
      pn[c] = dn * (1.f - n[c] * n[c]);
      pnr[c] = pn[c] * r[c];
      const float dr = pn[c] * gn[c];
      pr[c] = dr * r[c] * (1.f - r[c]);
      pz[c] = dz * z[c] * (1.f - z[c]);
      dp[c] = d[c] * z[c];
      const float mrz = fmaxf(fabsf(pr[c]), fabsf(pz[c]));
      mi = fmaxf(mi, fmaxf(mrz, fabsf(pn[c])));
      mh = fmaxf(mh, fmaxf(mrz, fabsf(pnr[c])));
    }
    float *gi = dgi + m * 3 * H + j, *gh = dgh + m * 3 * H + j;
    const float4 vr = make_float4(pr[0], pr[1], pr[2], pr[3]), vz = make_float4(pz[0], pz[1], pz[2], pz[3]);
    *reinterpret_cast<float4 *>(gi) = vr;
    *reinterpret_cast<float4 *>(gi + H) = vz;
    *reinterpret_cast<float4 *>(gi + 2 * H) = make_float4(pn[0], pn[1], pn[2], pn[3]);
    *reinterpret_cast<float4 *>(gh) = vr;
    *reinterpret_cast<float4 *>(gh + H) = vz;
    *reinterpret_cast<float4 *>(gh + 2 * H) = make_float4(pnr[0], pnr[1], pnr[2], pnr[3]);
    if (dh_prev != nullptr) *reinterpret_cast<float4 *>(dh_prev + m * H + j) = make_float4(dp[0], dp[1], dp[2], dp[3]);
  }
  if (amax_gi != nullptr) {
    __shared__ float smi[8], smh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mi = fmaxf(mi, __shfl_xor_sync(0xffffffffu, mi, o));
      mh = fmaxf(mh, __shfl_xor_sync(0xffffffffu, mh, o));
    }
    if ((threadIdx.x & 31) == 0) { smi[threadIdx.x >> 5] = mi; smh[threadIdx.x >> 5] = mh; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = smi[0], b = smh[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a = fmaxf(a, smi[w]); b = fmaxf(b, smh[w]); }
      amax_gi[blockIdx.x] = a; amax_gh[blockIdx.x] = b;
    }
  }
}

// one warp per edge
__global__ void k_edge_bwd(const float *__restrict__ dgi_e, const float *__restrict__ P, const float *__restrict__ dctx,
                           const float *__restrict__ Eh, const float *__restrict__ g, const int *__restrict__ subj,
                           const int *__restrict__ obj, int E, int H, const float *w0, const float *w1, const float *w2,
                           const float *w3, float *__restrict__ dl, float *__restrict__ dE) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= E) return;
  const int s = subj[e], o = obj[e];
  const float *gi = dgi_e + (size_t)e * 3 * H, *ps = P + (size_t)s * 3 * H, *po = P + (size_t)o * 3 * H;
  float ds = 0.f, dob = 0.f;
  for (int k = lane * 4; k < 3 * H; k += 128) {
    const float4 a = *reinterpret_cast<const float4 *>(gi + k);
    const float4 b = *reinterpret_cast<const float4 *>(ps + k), c = *reinterpret_cast<const float4 *>(po + k);
    ds += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    dob += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
  }
  const float *cs = dctx + (size_t)s * H, *co = dctx + (size_t)o * H, *er = Eh + (size_t)e * H;
  float dout = 0.f, din = 0.f;
  for (int k = lane * 4; k < H; k += 128) {
    const float4 a = *reinterpret_cast<const float4 *>(er + k);
    const float4 b = *reinterpret_cast<const float4 *>(cs + k), c = *reinterpret_cast<const float4 *>(co + k);
    dout += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    din += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
  }
  ds = sgg_warp_sum(ds); dob = sgg_warp_sum(dob); dout = sgg_warp_sum(dout); din = sgg_warp_sum(din);
  const float4 g4 = *reinterpret_cast<const float4 *>(g + (size_t)e * 4);
  const float l0 = ds * g4.x * (1.f - g4.x), l1 = dob * g4.y * (1.f - g4.y);
  const float l2 = dout * g4.z * (1.f - g4.z), l3 = din * g4.w * (1.f - g4.w);
  if (lane == 0) *reinterpret_cast<float4 *>(dl + (size_t)e * 4) = make_float4(l0, l1, l2, l3);
  float *de = dE + (size_t)e * H;
  for (int k = lane * 4; k < H; k += 128) {
    float4 acc = *reinterpret_cast<float4 *>(de + k);
    const float4 b = *reinterpret_cast<const float4 *>(cs + k), c = *reinterpret_cast<const float4 *>(co + k);
    const float4 a0 = __ldg(reinterpret_cast<const float4 *>(w0 + H + k)), a1 = __ldg(reinterpret_cast<const float4 *>(w1 + H + k));
    const float4 a2 = __ldg(reinterpret_cast<const float4 *>(w2 + H + k)), a3 = __ldg(reinterpret_cast<const float4 *>(w3 + H + k));
    acc.x += g4.z * b.x + g4.w * c.x + l0 * a0.x + l1 * a1.x + l2 * a2.x + l3 * a3.x;
    acc.y += g4.z * b.y + g4.w * c.y + l0 * a0.y + l1 * a1.y + l2 * a2.y + l3 * a3.y;
    acc.z += g4.z * b.z + g4.w * c.z + l0 * a0.z + l1 * a1.z + l2 * a2.z + l3 * a3.z;
    acc.w += g4.z * b.w + g4.w * c.w + l0 * a0.w + l1 * a1.w + l2 * a2.w + l3 * a3.w;
    *reinterpret_cast<float4 *>(de + k) = acc;
  }
}

// one CTA per node; CSR lists ascending => fixed summation order
__global__ void k_node_bwd(const float *__restrict__ dgi_e, const float *__restrict__ g, const float *__restrict__ dl,
                           const int *__restrict__ out_ptr, const int *__restrict__ out_idx,
                           const int *__restrict__ in_ptr, const int *__restrict__ in_idx, int H, const float *w0,
                           const float *w1, const float *w2, const float *w3, float *__restrict__ dP,
                           float *__restrict__ da, float *__restrict__ dV) {
  const int n = blockIdx.x;
  const int ob = out_ptr[n], oe = out_ptr[n + 1], ib = in_ptr[n], ie = in_ptr[n + 1];
  for (int j = threadIdx.x * 4; j < 3 * H; j += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = ob; k < oe; ++k) {
      const int e = out_idx[k];
      const float gs = g[(size_t)e * 4 + 0];
      const float4 v = *reinterpret_cast<const float4 *>(dgi_e + (size_t)e * 3 * H + j);
      acc.x += gs * v.x; acc.y += gs * v.y; acc.z += gs * v.z; acc.w += gs * v.w;
    }
    for (int k = ib; k < ie; ++k) {
      const int e = in_idx[k];
      const float go = g[(size_t)e * 4 + 1];
      const float4 v = *reinterpret_cast<const float4 *>(dgi_e + (size_t)e * 3 * H + j);
      acc.x += go * v.x; acc.y += go * v.y; acc.z += go * v.z; acc.w += go * v.w;
    }
    *reinterpret_cast<float4 *>(dP + (size_t)n * 3 * H + j) = acc;
  }
  // every thread redundantly reduces the 4 scalar gate-logit grads (short lists)
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int k = ob; k < oe; ++k) { const int e = out_idx[k]; a0 += dl[(size_t)e * 4 + 0]; a2 += dl[(size_t)e * 4 + 2]; }
  for (int k = ib; k < ie; ++k) { const int e = in_idx[k]; a1 += dl[(size_t)e * 4 + 1]; a3 += dl[(size_t)e * 4 + 3]; }
  if (threadIdx.x == 0) *reinterpret_cast<float4 *>(da + (size_t)n * 4) = make_float4(a0, a1, a2, a3);
  for (int j = threadIdx.x * 4; j < H; j += blockDim.x * 4) {
    float4 acc = *reinterpret_cast<float4 *>(dV + (size_t)n * H + j);
    const float4 v0 = __ldg(reinterpret_cast<const float4 *>(w0 + j)), v1 = __ldg(reinterpret_cast<const float4 *>(w1 + j));
    const float4 v2 = __ldg(reinterpret_cast<const float4 *>(w2 + j)), v3 = __ldg(reinterpret_cast<const float4 *>(w3 + j));
    acc.x += a0 * v0.x + a1 * v1.x + a2 * v2.x + a3 * v3.x;
    acc.y += a0 * v0.y + a1 * v1.y + a2 * v2.y + a3 * v3.y;
    acc.z += a0 * v0.z + a1 * v1.z + a2 * v2.z + a3 * v3.z;
    acc.w += a0 * v0.w + a1 * v1.w + a2 * v2.w + a3 * v3.w;
    *reinterpret_cast<float4 *>(dV + (size_t)n * H + j) = acc;
  }
}

// gate_w grads: dw[k][0:H] += tV[k], dw[k][H:2H] += tE[k]; db[k] += gb[k]
__global__ void k_gate_grad_finish(const float *__restrict__ tV, const float *__restrict__ tE,
                                   const float *__restrict__ gb, int H, float *d0, float *d1, float *d2, float *d3,
                                   float *b0, float *b1, float *b2, float *b3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float *dw[4] = {d0, d1, d2, d3};
  float *db[4] = {b0, b1, b2, b3};
  if (i < 4 * H) {
    const int k = i / H, j = i % H;
    if (dw[k]) { dw[k][j] += tV[i]; dw[k][H + j] += tE[i]; }
  }
  if (i < 4 && db[i]) db[i][0] += gb[i];
}

// ---- tensor-core backward GEMMs (3xTF32 engine: fp32 exponent range, so gradients of any magnitude are safe) ----
// The engine computes y = x w^T with x raw fp32 [M,K] and w pre-split [hi | lo] [Nout,K] (K contiguous), so
//   dX = dY W        -> x = dY,            w = split(W^T)                       (transpose of a weight, per call)
//   dW += dY^T X     -> x = dY^T [3H,Mp],  w = split(X^T) [H,Mp], Mp = M padded with zero rows to a multiple of 32
// Transposes go through 32x32 shared-memory tiles (coalesced both ways); rows in [R, Rpad) are written as zeros.
template <bool SPLIT>
__global__ void __launch_bounds__(256) k_transpose32(const float *__restrict__ in, int ldin, int R, int C,
                                                     float *__restrict__ hi, float *__restrict__ lo, int Rpad) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(size_t)r * ldin + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < Rpad) {
      const float v = tile[threadIdx.x][i];
      if (SPLIT) {
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);     // tc_gemm.cu: k_tc_split
        hi[(size_t)c * Rpad + r] = h;
        lo[(size_t)c * Rpad + r] = v - h;
      } else {
        hi[(size_t)c * Rpad + r] = v;
      }
    }
  }
}
// in [R,C] row-major -> out [C,Rpad]; split = true writes [hi | lo] planes of C*Rpad floats each
int launch_transpose_ld(const float *in, int ldin, int R, int C, float *out, int Rpad, bool split, cudaStream_t st) {
  dim3 grid((Rpad + 31) / 32, (C + 31) / 32);
  if (grid.y > 65535) return sgg_set_err(SGG_E_BADARG, "transpose: too many columns");
  if (split) k_transpose32<true><<<grid, dim3(32, 8), 0, st>>>(in, ldin, R, C, out, out + (size_t)C * Rpad, Rpad);
  else k_transpose32<false><<<grid, dim3(32, 8), 0, st>>>(in, ldin, R, C, out, nullptr, Rpad);
  SGG_RETURN_IF_LAUNCH_FAILED("k_transpose32");
  return 0;
}
static int launch_transpose(const float *in, int R, int C, float *out, int Rpad, bool split, cudaStream_t st) {
  return launch_transpose_ld(in, C, R, C, out, Rpad, split, st);
}
__global__ void k_add_inplace(float *__restrict__ c, const float *__restrict__ t, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4 *>(c)[i];
    const float4 b = reinterpret_cast<const float4 *>(t)[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4 *>(c)[i] = a;
  }
}
static int launch_add(float *c, const float *t, size_t n, cudaStream_t st) {      // n % 4 == 0 (H is a multiple of 64)
  const size_t n4 = n / 4;
  const int blocks = (int)((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184);
  k_add_inplace<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(c, t, n4);
  SGG_RETURN_IF_LAUNCH_FAILED("k_add_inplace");
  return 0;
}
constexpr int TC_BWD_MIN_ROWS = 256;     // dX-type GEMMs with fewer rows stay on the SIMT tiles
constexpr int TC_BWD_MIN_K = 256;        // dW-type GEMMs with a shorter reduction stay on the SIMT tiles (split-K)
static inline int pad32(int x) { return (x + 31) & ~31; }

struct BwdScratch {
  float *dV[2], *dE[2], *dgi_n, *dgh_n, *dgi_e, *dgh_e, *dctx, *dP, *dl, *da, *tV, *tE, *gb, *cs, *ws4, *sk;
  size_t sk_floats;
  // tensor-core backward: W^T splits (node_ih, node_hh, edge_ih, edge_hh), transposed operands, GEMM output, split-K partials
  float *wt[4], *xT, *bT, *tmp, *lin;
  // scaled 3xFP16 engine (default): exact power-of-two scale pair (s, 1/s) from the device-side abs-max, abs-max partials
  float *sc, *p2;
  float *amax;            // [2][4096] per-block abs-max partials of k_gru_bwd (dgi, dgh)
  float *sc_gi, *sc_gh;   // (s, 1/s) of the current dgi / dgh
};
constexpr int SPLITK_MAX = 4;      // split-K slices of the [3H,H] weight-gradient GEMMs (reduction over E or N rows)
static size_t bwd_layout(BwdScratch *s, void *ws, int N, int E, int H) {
  SggArena ar(ws, (size_t)-1);
  const size_t n1 = N > 0 ? N : 1, e1 = E > 0 ? E : 1;
  for (int i = 0; i < 2; ++i) { s->dV[i] = ar.take<float>(n1 * H); s->dE[i] = ar.take<float>(e1 * H); }
  s->dgi_n = ar.take<float>(n1 * 3 * H); s->dgh_n = ar.take<float>(n1 * 3 * H);
  s->dgi_e = ar.take<float>(e1 * 3 * H); s->dgh_e = ar.take<float>(e1 * 3 * H);
  s->dctx = ar.take<float>(n1 * H); s->dP = ar.take<float>(n1 * 3 * H);
  s->dl = ar.take<float>(e1 * 4); s->da = ar.take<float>(n1 * 4);
  s->tV = ar.take<float>((size_t)4 * H); s->tE = ar.take<float>((size_t)4 * H); s->gb = ar.take<float>(4);
  s->cs = ar.take<float>(colsum_workspace_floats((int)(e1 > n1 ? e1 : n1), 3 * H));
  s->ws4 = ar.take<float>(wsum4_workspace_floats((int)(e1 > n1 ? e1 : n1), H));
  s->sk_floats = (size_t)SPLITK_MAX * 3 * H * H;
  s->sk = ar.take<float>(s->sk_floats);
  const size_t big = e1 > n1 ? e1 : n1;
  const int mp = pad32((int)big);
  for (int i = 0; i < 4; ++i) s->wt[i] = ar.take<float>((size_t)2 * 3 * H * H);
  s->xT = ar.take<float>((size_t)3 * H * mp);
  s->bT = ar.take<float>((size_t)2 * H * mp);
  s->tmp = ar.take<float>(big * H > (size_t)3 * H * H ? big * H : (size_t)3 * H * H);
  size_t lin = tc32_linear_workspace_floats(3 * H, H, mp);                       // dW-type
  const size_t l1 = tc32_linear_workspace_floats((int)e1, H, 3 * H), l2 = tc32_linear_workspace_floats((int)n1, H, 3 * H);
  if (l1 > lin) lin = l1;
  if (l2 > lin) lin = l2;
  const size_t l3 = tc16::linear_workspace_floats(3 * H, H, mp), l4 = tc16::linear_workspace_floats((int)big, H, 3 * H);
  if (l3 > lin) lin = l3;
  if (l4 > lin) lin = l4;
  s->lin = ar.take<float>(lin > 0 ? lin : 1);
  s->sc = ar.take<float>(4);
  s->amax = ar.take<float>(2 * 4096);
  s->sc_gi = ar.take<float>(8); s->sc_gh = s->sc_gi + 4;
  s->p2 = ar.take<float>(sgg_pow2_scale_workspace_bytes() / sizeof(float));
  return ar.off;
}

}  // namespace sgg

extern "C" size_t sgg_mp_backward_workspace_bytes(int N, int E, int H, int T) {
  (void)T;
  sgg::BwdScratch s;
  return sgg::bwd_layout(&s, nullptr, N < 0 ? 0 : N, E < 0 ? 0 : E, H);
}

// obj_rep / rel_rep: the forward inputs.  dV_T / dE_T: grads of the outputs.  `grads` mirrors sgg_mp_weights and is
// ACCUMULATED into (NULL fields are skipped).  d_obj_rep / d_rel_rep are overwritten (nullable).
extern "C" int sgg_mp_backward(const float *obj_rep, const float *rel_rep, const void *graph_ws, const sgg_mp_weights *w,
                               const float *tape_base, int N, int E, int H, int T, const float *dV_T, const float *dE_T,
                               const sgg_mp_grads *grads, float *d_obj_rep, float *d_rel_rep, void *ws, size_t ws_bytes,
                               void *stream) {
  using namespace sgg;
  if (!w || !graph_ws || !tape_base || !grads || !ws) return sgg_set_err(SGG_E_BADARG, "mp_backward: null pointer");
  if (N < 0 || E < 0 || T < 0 || H <= 0 || (H % 64) != 0) return sgg_set_err(SGG_E_BADARG, "mp_backward: bad shape");
  BwdScratch s;
  const size_t need = bwd_layout(&s, ws, N, E, H);
  if (need > ws_bytes) return sgg_set_err(SGG_E_WORKSPACE, "mp_backward: workspace %zu < %zu", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  SggGraphView g = sgg_graph_view(graph_ws, N, E);
  MpTape tp = mp_tape_view(const_cast<float *>(tape_base), N, E, H, T);
  const size_t vN = (size_t)N * H, eN = (size_t)E * H;
  auto Vt = [&](int t) { return tp.states + (size_t)t * (vN + eN); };
  auto Et = [&](int t) { return tp.states + (size_t)t * (vN + eN) + vN; };
  int rc;
  SGG_CUDA_TRY(cudaMemsetAsync(s.tV, 0, sizeof(float) * 4 * (size_t)H, st));
  SGG_CUDA_TRY(cudaMemsetAsync(s.tE, 0, sizeof(float) * 4 * (size_t)H, st));
  SGG_CUDA_TRY(cudaMemsetAsync(s.gb, 0, sizeof(float) * 4, st));
  auto ew_blocks = [&](size_t n) { size_t b = (n + 255) / 256; return (int)(b < 4096 ? (b ? b : 1) : 4096); };
  auto gemm = [&](const float *A, int lda, bool ac, const float *B, int ldb, bool bc, float *C, int ldc, int M, int Nn,
                  int K, bool acc) { return launch_gemm(A, lda, ac, B, ldb, bc, C, ldc, M, Nn, K, acc, st, s.sk, s.sk_floats); };
  // tensor-core routes (enabled when the forward ran on tensor cores, i.e. the caller passed operand splits)
  static const bool tc_env = [] { const char *v = getenv("SGG_BWD_TC"); return v == nullptr || atoi(v) != 0; }();
  const bool tc_on = tc_env && w->edge_w_hh_split != nullptr && w->node_w_hh_split != nullptr;
  enum { WT_NODE_IH = 0, WT_NODE_HH = 1, WT_EDGE_IH = 2, WT_EDGE_HH = 3 };
  // engine of the tensor-core routes: scaled 3xFP16 (default; the gradient operand is multiplied by an exact power of two
  // so that max|s dY| lies in [1024, 2048) — same scheme as ops.linear_backward) or 3xTF32 (SGG_BWD_ENGINE=tc32)
  static const bool tc16_env = [] { const char *v = getenv("SGG_BWD_ENGINE"); return v == nullptr || strcmp(v, "tc32") != 0; }();
  const bool use16 = tc_on && tc16_env;
  if (tc_on) {      // W [3H,H] -> W^T [H,3H] operand planes, once per backward call
    const float *ws4[4] = {w->node_w_ih, w->node_w_hh, w->edge_w_ih, w->edge_w_hh};
    for (int i = 0; i < 4; ++i) {
      if (use16) rc = sgg_bwd_transpose16(ws4[i], H, 3 * H, H, s.wt[i], 3 * H, 1, nullptr, st);     // fp16 [hi | lo] planes
      else rc = launch_transpose(ws4[i], 3 * H, H, s.wt[i], 3 * H, true, st);                        // 3xTF32 [hi | lo]
      if (rc) return rc;
    }
  }
  // power-of-two scale pair of a gradient operand: s.sc = (s, 1/s).  One computation serves the dX and the dW GEMM of dY;
  // the scale is applied inside the consumers (the LINEAR kernel's fp32 -> fp16 split, the transpose).
  const float *scaled_of = nullptr;
  auto scale16 = [&](const float *dY, int M) -> int {
    if (scaled_of == dY) return 0;
    const int r = sgg_pow2_scale(dY, (long long)M * 3 * H, s.sc, s.p2, sgg_pow2_scale_workspace_bytes(), st);
    scaled_of = r == 0 ? dY : nullptr;
    return r;
  };
  // dX[M,H] (=|+=) dY[M,3H] W[3H,H]
  // `scp`: the (s, 1/s) pair of dY when its producer already reduced the abs-max (k_gru_bwd), else computed here
  auto gemm_dx = [&](const float *dY, int wi, const float *W, float *out, int M, bool acc, const float *scp) -> int {
    if (!tc_on || M < TC_BWD_MIN_ROWS) return gemm(dY, 3 * H, false, W, H, true, out, H, M, H, 3 * H, acc);
    int r;
    if (use16) {
      if (scp == nullptr) { if ((r = scale16(dY, M))) return r; scp = s.sc; }
      r = tc16::linear_scaled(dY, s.wt[wi], acc ? s.tmp : out, M, H, 3 * H, scp + 1, s.lin, st, scp);
    } else {
      r = tc32_linear(dY, s.wt[wi], nullptr, acc ? s.tmp : out, M, H, 3 * H, 0, s.lin, st);
    }
    if (r == 0 && acc) r = launch_add(out, s.tmp, (size_t)M * H, st);
    return r;
  };
  // dW[3H,H] += dY[M,3H]^T X[M,H]
  auto gemm_dw = [&](const float *dY, const float *X, float *dW, int M, const float *scp) -> int {
    if (!tc_on || M < TC_BWD_MIN_K) return gemm(dY, 3 * H, true, X, H, true, dW, H, 3 * H, H, M, true);
    const int mp = pad32(M);
    int r;
    if (use16) {
      if (scp == nullptr) { if ((r = scale16(dY, M))) return r; scp = s.sc; }
      // (the pre-split kernel, lin16p.cu, was tried here: its 48 tiles of a [3H, H] output without split-K leave 100 SMs idle —
      // L1 train step 3.99 -> 4.08 ms — so these GEMMs stay on the split-K LINEAR engine)
      r = sgg_bwd_transpose16(dY, 3 * H, M, 3 * H, s.xT, mp, 0, scp, st);                           // (s dY)^T  [3H, mp] fp32
      if (r == 0) r = sgg_bwd_transpose16(X, H, M, H, s.bT, mp, 1, nullptr, st);                     // X^T planes [H, mp]
      if (r == 0) r = tc16::linear_scaled(s.xT, s.bT, s.tmp, 3 * H, H, mp, scp + 1, s.lin, st);
    } else {
      r = launch_transpose(dY, M, 3 * H, s.xT, mp, false, st);
      if (r == 0) r = launch_transpose(X, M, H, s.bT, mp, true, st);
      if (r == 0) r = tc32_linear(s.xT, s.bT, nullptr, s.tmp, 3 * H, H, mp, 0, s.lin, st);
    }
    if (r == 0) r = launch_add(dW, s.tmp, (size_t)3 * H * H, st);
    return r;
  };
  auto colsum = [&](const float *X, int rows, int cols, float *out) -> int {
    return out ? launch_colsum(X, cols, rows, cols, out, true, s.cs, st) : 0;
  };

  // GRUCell backward of `rows` rows; with the 3xFP16 engine it also leaves the scale pairs of dgi / dgh in s.sc_gi / s.sc_gh
  auto gru_bwd = [&](const float *dh_next, const float *cache, const float *h_prev, int rows, float *dgi, float *dgh,
                     float *dh_prev) -> int {
    const int blocks = ew_blocks((size_t)rows * H / 4);
    const bool fuse = use16 && rows >= TC_BWD_MIN_ROWS;
    k_gru_bwd<<<blocks, 256, 0, st>>>(dh_next, cache, h_prev, rows, H, dgi, dgh, dh_prev, fuse ? s.amax : nullptr,
                                      fuse ? s.amax + 4096 : nullptr);
    SGG_RETURN_IF_LAUNCH_FAILED("k_gru_bwd");
    scaled_of = nullptr;                  // the gradient buffers were just rewritten
    return fuse ? launch_pow2_from_parts(s.amax, blocks, s.sc_gi, 2, 4096, st) : 0;
  };
  auto sc_of = [&](int rows, const float *scp) -> const float * { return (use16 && rows >= TC_BWD_MIN_ROWS) ? scp : nullptr; };

  const float *dVn = dV_T, *dEn = dE_T;   // grads w.r.t. V_{t+1}, E_{t+1}
  for (int t = T - 1; t >= 0; --t) {
    float *dV = s.dV[t & 1], *dE = s.dE[t & 1];
    const float *V = Vt(t), *Eh = Et(t);
    const float *gt = tp.gates + (size_t)t * E * 4, *ctx = tp.ctx + (size_t)t * N * H, *P = tp.P + (size_t)t * N * 3 * H;
    if (N > 0) {
      if ((rc = gru_bwd(dVn, tp.cacheV + (size_t)(t + 1) * N * 4 * H, V, N, s.dgi_n, s.dgh_n, dV))) return rc;
      const float *sgi = sc_of(N, s.sc_gi), *sgh = sc_of(N, s.sc_gh);
      if (grads->node_w_ih && (rc = gemm_dw(s.dgi_n, ctx, grads->node_w_ih, N, sgi))) return rc;
      if ((rc = gemm_dx(s.dgi_n, WT_NODE_IH, w->node_w_ih, s.dctx, N, false, sgi))) return rc;
      if (grads->node_w_hh && (rc = gemm_dw(s.dgh_n, V, grads->node_w_hh, N, sgh))) return rc;
      if ((rc = gemm_dx(s.dgh_n, WT_NODE_HH, w->node_w_hh, dV, N, true, sgh))) return rc;
      if ((rc = colsum(s.dgi_n, N, 3 * H, grads->node_b_ih))) return rc;
      if ((rc = colsum(s.dgh_n, N, 3 * H, grads->node_b_hh))) return rc;
    }
    if (E > 0) {
      if ((rc = gru_bwd(dEn, tp.cacheE + (size_t)(t + 1) * E * 4 * H, Eh, E, s.dgi_e, s.dgh_e, dE))) return rc;
      const float *sgh = sc_of(E, s.sc_gh);
      if (grads->edge_w_hh && (rc = gemm_dw(s.dgh_e, Eh, grads->edge_w_hh, E, sgh))) return rc;
      if ((rc = gemm_dx(s.dgh_e, WT_EDGE_HH, w->edge_w_hh, dE, E, true, sgh))) return rc;
      if ((rc = colsum(s.dgh_e, E, 3 * H, grads->edge_b_hh))) return rc;
      if ((rc = colsum(s.dgi_e, E, 3 * H, grads->edge_b_ih))) return rc;
      k_edge_bwd<<<(int)(((size_t)E * 32 + 255) / 256), 256, 0, st>>>(s.dgi_e, P, s.dctx, Eh, gt, g.subj, g.obj, E, H,
                                                                    w->gate_w[0], w->gate_w[1], w->gate_w[2],
                                                                    w->gate_w[3], s.dl, dE);
      SGG_RETURN_IF_LAUNCH_FAILED("k_edge_bwd");
    }
    if (N > 0) {
      if (E == 0) {
        SGG_CUDA_TRY(cudaMemsetAsync(s.dP, 0, sizeof(float) * vN * 3, st));
        SGG_CUDA_TRY(cudaMemsetAsync(s.da, 0, sizeof(float) * N * 4, st));
      } else {
        k_node_bwd<<<N, 128, 0, st>>>(s.dgi_e, gt, s.dl, g.out_ptr, g.out_idx, g.in_ptr, g.in_idx, H, w->gate_w[0],
                                      w->gate_w[1], w->gate_w[2], w->gate_w[3], s.dP, s.da, dV);
        SGG_RETURN_IF_LAUNCH_FAILED("k_node_bwd");
        scaled_of = nullptr;
        if (grads->edge_w_ih && (rc = gemm_dw(s.dP, V, grads->edge_w_ih, N, nullptr))) return rc;
        if ((rc = gemm_dx(s.dP, WT_EDGE_IH, w->edge_w_ih, dV, N, true, nullptr))) return rc;
        if ((rc = launch_wsum4(s.da, V, N, H, s.tV, true, s.ws4, st))) return rc;      // tV[k] += sum_n da[n][k] V[n]
        if ((rc = launch_wsum4(s.dl, Eh, E, H, s.tE, true, s.ws4, st))) return rc;     // tE[k] += sum_e dl[e][k] E[e]
        if ((rc = launch_colsum(s.dl, 4, E, 4, s.gb, true, s.cs, st))) return rc;
      }
    }
    dVn = dV; dEn = dE;
  }
  // initial step (h = 0): rel_model_stanford.py:68-72
  if (N > 0) {
    if ((rc = gru_bwd(dVn, tp.cacheV, nullptr, N, s.dgi_n, s.dgh_n, nullptr))) return rc;
    const float *sgi = sc_of(N, s.sc_gi);
    if (grads->node_w_ih && (rc = gemm_dw(s.dgi_n, obj_rep, grads->node_w_ih, N, sgi))) return rc;
    if (d_obj_rep && (rc = gemm_dx(s.dgi_n, WT_NODE_IH, w->node_w_ih, d_obj_rep, N, false, sgi))) return rc;
    if ((rc = colsum(s.dgi_n, N, 3 * H, grads->node_b_ih))) return rc;
    if ((rc = colsum(s.dgh_n, N, 3 * H, grads->node_b_hh))) return rc;
  }
  if (E > 0) {
    if ((rc = gru_bwd(dEn, tp.cacheE, nullptr, E, s.dgi_e, s.dgh_e, nullptr))) return rc;
    const float *sgi = sc_of(E, s.sc_gi);
    if (grads->edge_w_ih && (rc = gemm_dw(s.dgi_e, rel_rep, grads->edge_w_ih, E, sgi))) return rc;
    if (d_rel_rep && (rc = gemm_dx(s.dgi_e, WT_EDGE_IH, w->edge_w_ih, d_rel_rep, E, false, sgi))) return rc;
    if ((rc = colsum(s.dgi_e, E, 3 * H, grads->edge_b_ih))) return rc;
    if ((rc = colsum(s.dgh_e, E, 3 * H, grads->edge_b_hh))) return rc;
  }
  k_gate_grad_finish<<<(4 * H + 255) / 256, 256, 0, st>>>(s.tV, s.tE, s.gb, H, grads->gate_w[0], grads->gate_w[1],
                                                          grads->gate_w[2], grads->gate_w[3], grads->gate_b[0],
                                                          grads->gate_b[1], grads->gate_b[2], grads->gate_b[3]);
  SGG_RETURN_IF_LAUNCH_FAILED("k_gate_grad_finish");
  return 0;
}
