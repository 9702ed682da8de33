// Iterative message passing (IMP) forward — RelModelStanford.message_pass,
// sgg_models/rel_model_stanford.py:48-94 — on ragged CSR graphs.
//
// Algebra (exact up to fp32 reassociation):
//   x_e = g_s * V[s] + g_o * V[o]           (:78-81)
//   W_ih x_e = g_s * (W_ih V[s]) + g_o * (W_ih V[o])   ("linearity shortcut")
//   w . [v ; e] = w_v . v + w_e . e                    (gate logits split per half)
// so per iteration the E-row input-side GEMM of the edge GRU collapses into one
// N-row GEMM  P = V W_ih_e^T  whose rows are gathered (and scaled by the scalar
// gates) in the epilogue of the E-row hidden-side GEMM  R = Eh W_hh_e^T.
//
// Per iteration i (all reads are iteration-i states, Jacobi, :74-92):
//   k_gate_node   a[n,4]   = V[n] . w_k[:H]                       warp-shuffle dots
//   k_gate_edge   g[e,4]   = sigmoid(a[s|o,k] + Eh[e] . w_k[H:] + b_k)
//   linear        P[N,3H]  = V W_ih_e^T
//   k_ctx         ctx[n]   = sum_out g_out[e] Eh[e] + sum_in g_in[e] Eh[e]   (CSR, fixed order)
//   k_gru<EDGE>   Eh'      = GRU(gi = g_s P[s] + g_o P[o] + b_ih, gh = Eh W_hh_e^T + b_hh, h = Eh)
//   k_gru<NODE>   V'       = GRU(gi = ctx W_ih_n^T + b_ih, gh = V W_hh_n^T + b_hh, h = V)
// The GRU non-linearity runs in the GEMM epilogue: a tile covers hidden units
// [j0, j0+64) of all three gates (weight rows j0, H+j0, 2H+j0).
#include "gemm_core.cuh"
#include "kernels.h"

namespace sgg {

enum { GRU_INIT = 0, GRU_NODE = 1, GRU_EDGE = 2 };

struct GruArgs {
  // GEMM operands
  const float *x;      // INIT: input rows [M,H]; NODE: ctx [M,H]; EDGE: unused
  const float *h;      // NODE/EDGE: previous state [M,H]; INIT: null (h = 0)
  const float *w_ih, *w_hh, *b_ih, *b_hh;
  // EDGE only
  const float *P;      // [N,3H] = V W_ih^T
  const float *gates;  // [M,4]  (g_sub, g_obj, g_out, g_in)
  const int *subj, *obj;
  float *out;          // [M,H]
  float *cache;        // nullable [M,4,H]: (r, z, n, gh_n) saved for the backward pass
  int M, H;
};

template <int BM, int MODE>
__global__ void __launch_bounds__(NTHREADS, (BM == 64 ? 2 : 1))
k_gru(GruArgs p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TM = BM / 16;
  constexpr int NACC = (MODE == GRU_NODE) ? 4 : 3;   // r, z, n_i [, n_h]
  const int H = p.H;
  const int m0 = blockIdx.y * BM;
  const int j0 = blockIdx.x * BN;
  float acc[TM][NACC][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int b = 0; b < NACC; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][b][c] = 0.f;

  WBlocks<3> Wb;
  Wb.ldw = H;
#pragma unroll
  for (int g = 0; g < 3; ++g) { Wb.rowbase[g] = g * H + j0; Wb.nvalid[g] = BN; }

  if (MODE == GRU_INIT || MODE == GRU_NODE) {
    ARows A{p.x, H, p.M, m0};
    Wb.W = p.w_ih;
    gemm_segment<BM, NACC, 3, AccMap<0, 1, 2>>(acc, A, Wb, H, smem);
  }
  if (MODE == GRU_NODE) {
    ARows A{p.h, H, p.M, m0};
    Wb.W = p.w_hh;
    gemm_segment<BM, NACC, 3, AccMap<0, 1, 3>>(acc, A, Wb, H, smem);
  }
  if (MODE == GRU_EDGE) {
    ARows A{p.h, H, p.M, m0};
    Wb.W = p.w_hh;
    gemm_segment<BM, NACC, 3, AccMap<0, 1, 2>>(acc, A, Wb, H, smem);
  }

  // ---- epilogue: GRUCell pointwise (torch.nn.GRUCell semantics) ----
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int j = j0 + tx * 4;
  const float4 bir = __ldg(reinterpret_cast<const float4 *>(p.b_ih + j));
  const float4 biz = __ldg(reinterpret_cast<const float4 *>(p.b_ih + H + j));
  const float4 bin = __ldg(reinterpret_cast<const float4 *>(p.b_ih + 2 * H + j));
  const float4 bhr = __ldg(reinterpret_cast<const float4 *>(p.b_hh + j));
  const float4 bhz = __ldg(reinterpret_cast<const float4 *>(p.b_hh + H + j));
  const float4 bhn = __ldg(reinterpret_cast<const float4 *>(p.b_hh + 2 * H + j));
  const float bi_r[4] = {bir.x, bir.y, bir.z, bir.w}, bi_z[4] = {biz.x, biz.y, biz.z, biz.w},
              bi_n[4] = {bin.x, bin.y, bin.z, bin.w};
  const float bh_r[4] = {bhr.x, bhr.y, bhr.z, bhr.w}, bh_z[4] = {bhz.x, bhz.y, bhz.z, bhz.w},
              bh_n[4] = {bhn.x, bhn.y, bhn.z, bhn.w};
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
    float hv[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE != GRU_INIT) {
      float4 t = *reinterpret_cast<const float4 *>(p.h + (size_t)m * H + j);
      hv[0] = t.x; hv[1] = t.y; hv[2] = t.z; hv[3] = t.w;
    }
    float gi_r[4], gi_z[4], gi_n[4], gh_r[4], gh_z[4], gh_n[4];
    if (MODE == GRU_EDGE) {
      const int s = p.subj[m], o = p.obj[m];
      const float4 g4 = *reinterpret_cast<const float4 *>(p.gates + (size_t)m * 4);
      const float gs = g4.x, go = g4.y;
      const float *ps = p.P + (size_t)s * 3 * H + j, *po = p.P + (size_t)o * 3 * H + j;
      const float4 sr = *reinterpret_cast<const float4 *>(ps), orr = *reinterpret_cast<const float4 *>(po);
      const float4 sz = *reinterpret_cast<const float4 *>(ps + H), oz = *reinterpret_cast<const float4 *>(po + H);
      const float4 sn = *reinterpret_cast<const float4 *>(ps + 2 * H), on = *reinterpret_cast<const float4 *>(po + 2 * H);
      const float a_sr[4] = {sr.x, sr.y, sr.z, sr.w}, a_or[4] = {orr.x, orr.y, orr.z, orr.w};
      const float a_sz[4] = {sz.x, sz.y, sz.z, sz.w}, a_oz[4] = {oz.x, oz.y, oz.z, oz.w};
      const float a_sn[4] = {sn.x, sn.y, sn.z, sn.w}, a_on[4] = {on.x, on.y, on.z, on.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        gi_r[c] = fmaf(gs, a_sr[c], go * a_or[c]) + bi_r[c];
        gi_z[c] = fmaf(gs, a_sz[c], go * a_oz[c]) + bi_z[c];
        gi_n[c] = fmaf(gs, a_sn[c], go * a_on[c]) + bi_n[c];
        gh_r[c] = acc[i][0][c] + bh_r[c];
        gh_z[c] = acc[i][1][c] + bh_z[c];
        gh_n[c] = acc[i][2][c] + bh_n[c];
      }
    } else if (MODE == GRU_NODE) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {   // acc0/acc1 already hold gi+gh for r and z
        gi_r[c] = acc[i][0][c] + bi_r[c]; gh_r[c] = bh_r[c];
        gi_z[c] = acc[i][1][c] + bi_z[c]; gh_z[c] = bh_z[c];
        gi_n[c] = acc[i][2][c] + bi_n[c];
        gh_n[c] = acc[i][NACC - 1][c] + bh_n[c];
      }
    } else {  // INIT: h = 0 => gh = b_hh
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        gi_r[c] = acc[i][0][c] + bi_r[c]; gh_r[c] = bh_r[c];
        gi_z[c] = acc[i][1][c] + bi_z[c]; gh_z[c] = bh_z[c];
        gi_n[c] = acc[i][2][c] + bi_n[c]; gh_n[c] = bh_n[c];
      }
    }
    float o4[4], cr[4], cz[4], cn[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float r = sgg_sigmoid(gi_r[c] + gh_r[c]);
      const float z = sgg_sigmoid(gi_z[c] + gh_z[c]);
      const float n = tanhf(gi_n[c] + r * gh_n[c]);
      o4[c] = (1.0f - z) * n + z * hv[c];
      cr[c] = r; cz[c] = z; cn[c] = n;
    }
    *reinterpret_cast<float4 *>(p.out + (size_t)m * H + j) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    if (p.cache != nullptr) {
      float *cp = p.cache + (size_t)m * 4 * H + j;
      *reinterpret_cast<float4 *>(cp) = make_float4(cr[0], cr[1], cr[2], cr[3]);
      *reinterpret_cast<float4 *>(cp + H) = make_float4(cz[0], cz[1], cz[2], cz[3]);
      *reinterpret_cast<float4 *>(cp + 2 * H) = make_float4(cn[0], cn[1], cn[2], cn[3]);
      *reinterpret_cast<float4 *>(cp + 3 * H) = make_float4(gh_n[0], gh_n[1], gh_n[2], gh_n[3]);
    }
  }
}

// a[n,k] = V[n] . gate_w[k][:H]   — one warp per node, warp-shuffle reduction.
__global__ void k_gate_node(const float *__restrict__ V, int N, int H, const float *w0, const float *w1,
                            const float *w2, const float *w3, float *__restrict__ a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float *row = V + (size_t)warp * H;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int k = lane * 4; k < H; k += 128) {
    const float4 v = *reinterpret_cast<const float4 *>(row + k);
    const float4 a0 = __ldg(reinterpret_cast<const float4 *>(w0 + k));
    const float4 a1 = __ldg(reinterpret_cast<const float4 *>(w1 + k));
    const float4 a2 = __ldg(reinterpret_cast<const float4 *>(w2 + k));
    const float4 a3 = __ldg(reinterpret_cast<const float4 *>(w3 + k));
    s0 += v.x * a0.x + v.y * a0.y + v.z * a0.z + v.w * a0.w;
    s1 += v.x * a1.x + v.y * a1.y + v.z * a1.z + v.w * a1.w;
    s2 += v.x * a2.x + v.y * a2.y + v.z * a2.z + v.w * a2.w;
    s3 += v.x * a3.x + v.y * a3.y + v.z * a3.z + v.w * a3.w;
  }
  s0 = sgg_warp_sum(s0); s1 = sgg_warp_sum(s1); s2 = sgg_warp_sum(s2); s3 = sgg_warp_sum(s3);
  if (lane == 0) *reinterpret_cast<float4 *>(a + (size_t)warp * 4) = make_float4(s0, s1, s2, s3);
}

// g[e,k] = sigmoid(a[s or o, k] + Eh[e] . gate_w[k][H:] + b[k])   (rel_model_stanford.py:78-81,86-89)
__global__ void k_gate_edge(const float *__restrict__ Eh, int E, int H, const float *w0, const float *w1,
                            const float *w2, const float *w3, const float *b0, const float *b1, const float *b2,
                            const float *b3, const float *__restrict__ a, const int *__restrict__ subj,
                            const int *__restrict__ obj, float *__restrict__ g) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= E) return;
  const float *row = Eh + (size_t)warp * H;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int k = lane * 4; k < H; k += 128) {
    const float4 v = *reinterpret_cast<const float4 *>(row + k);
    const float4 a0 = __ldg(reinterpret_cast<const float4 *>(w0 + H + k));
    const float4 a1 = __ldg(reinterpret_cast<const float4 *>(w1 + H + k));
    const float4 a2 = __ldg(reinterpret_cast<const float4 *>(w2 + H + k));
    const float4 a3 = __ldg(reinterpret_cast<const float4 *>(w3 + H + k));
    s0 += v.x * a0.x + v.y * a0.y + v.z * a0.z + v.w * a0.w;
    s1 += v.x * a1.x + v.y * a1.y + v.z * a1.z + v.w * a1.w;
    s2 += v.x * a2.x + v.y * a2.y + v.z * a2.z + v.w * a2.w;
    s3 += v.x * a3.x + v.y * a3.y + v.z * a3.z + v.w * a3.w;
  }
  s0 = sgg_warp_sum(s0); s1 = sgg_warp_sum(s1); s2 = sgg_warp_sum(s2); s3 = sgg_warp_sum(s3);
  if (lane == 0) {
    const int s = subj[warp], o = obj[warp];
    const float4 as = *reinterpret_cast<const float4 *>(a + (size_t)s * 4);
    const float4 ao = *reinterpret_cast<const float4 *>(a + (size_t)o * 4);
    float4 r;
    r.x = sgg_sigmoid(as.x + s0 + __ldg(b0));   // sub_vert: subject vertex
    r.y = sgg_sigmoid(ao.y + s1 + __ldg(b1));   // obj_vert: object vertex
    r.z = sgg_sigmoid(as.z + s2 + __ldg(b2));   // out_edge: uses the subject vertex (:86)
    r.w = sgg_sigmoid(ao.w + s3 + __ldg(b3));   // in_edge:  uses the object vertex  (:88)
    *reinterpret_cast<float4 *>(g + (size_t)warp * 4) = r;
  }
}

// ctx[n] = sum_{e in out(n)} g_out[e] Eh[e] + sum_{e in in(n)} g_in[e] Eh[e]   (:86-91)
// One CTA per node, one float4 column group per thread; CSR lists are in ascending edge id.
__global__ void k_ctx(const float *__restrict__ Eh, const float *__restrict__ g, const int *__restrict__ out_ptr,
                      const int *__restrict__ out_idx, const int *__restrict__ in_ptr, const int *__restrict__ in_idx,
                      int H, float *__restrict__ ctx) {
  const int n = blockIdx.x;
  for (int j = threadIdx.x * 4; j < H; j += blockDim.x * 4) {
    float4 so = make_float4(0.f, 0.f, 0.f, 0.f), si = so;
    int b = out_ptr[n], e = out_ptr[n + 1];
    int k = b;
    for (; k + 4 <= e; k += 4) {
      int e0 = out_idx[k], e1 = out_idx[k + 1], e2 = out_idx[k + 2], e3 = out_idx[k + 3];
      float g0 = g[(size_t)e0 * 4 + 2], g1 = g[(size_t)e1 * 4 + 2], g2 = g[(size_t)e2 * 4 + 2], g3 = g[(size_t)e3 * 4 + 2];
      float4 v0 = *reinterpret_cast<const float4 *>(Eh + (size_t)e0 * H + j);
      float4 v1 = *reinterpret_cast<const float4 *>(Eh + (size_t)e1 * H + j);
      float4 v2 = *reinterpret_cast<const float4 *>(Eh + (size_t)e2 * H + j);
      float4 v3 = *reinterpret_cast<const float4 *>(Eh + (size_t)e3 * H + j);
      so.x += g0 * v0.x; so.y += g0 * v0.y; so.z += g0 * v0.z; so.w += g0 * v0.w;
      so.x += g1 * v1.x; so.y += g1 * v1.y; so.z += g1 * v1.z; so.w += g1 * v1.w;
      so.x += g2 * v2.x; so.y += g2 * v2.y; so.z += g2 * v2.z; so.w += g2 * v2.w;
      so.x += g3 * v3.x; so.y += g3 * v3.y; so.z += g3 * v3.z; so.w += g3 * v3.w;
    }
    for (; k < e; ++k) {
      int e0 = out_idx[k];
      float g0 = g[(size_t)e0 * 4 + 2];
      float4 v0 = *reinterpret_cast<const float4 *>(Eh + (size_t)e0 * H + j);
      so.x += g0 * v0.x; so.y += g0 * v0.y; so.z += g0 * v0.z; so.w += g0 * v0.w;
    }
    b = in_ptr[n]; e = in_ptr[n + 1];
    k = b;
    for (; k + 4 <= e; k += 4) {
      int e0 = in_idx[k], e1 = in_idx[k + 1], e2 = in_idx[k + 2], e3 = in_idx[k + 3];
      float g0 = g[(size_t)e0 * 4 + 3], g1 = g[(size_t)e1 * 4 + 3], g2 = g[(size_t)e2 * 4 + 3], g3 = g[(size_t)e3 * 4 + 3];
      float4 v0 = *reinterpret_cast<const float4 *>(Eh + (size_t)e0 * H + j);
      float4 v1 = *reinterpret_cast<const float4 *>(Eh + (size_t)e1 * H + j);
      float4 v2 = *reinterpret_cast<const float4 *>(Eh + (size_t)e2 * H + j);
      float4 v3 = *reinterpret_cast<const float4 *>(Eh + (size_t)e3 * H + j);
      si.x += g0 * v0.x; si.y += g0 * v0.y; si.z += g0 * v0.z; si.w += g0 * v0.w;
      si.x += g1 * v1.x; si.y += g1 * v1.y; si.z += g1 * v1.z; si.w += g1 * v1.w;
      si.x += g2 * v2.x; si.y += g2 * v2.y; si.z += g2 * v2.z; si.w += g2 * v2.w;
      si.x += g3 * v3.x; si.y += g3 * v3.y; si.z += g3 * v3.z; si.w += g3 * v3.w;
    }
    for (; k < e; ++k) {
      int e0 = in_idx[k];
      float g0 = g[(size_t)e0 * 4 + 3];
      float4 v0 = *reinterpret_cast<const float4 *>(Eh + (size_t)e0 * H + j);
      si.x += g0 * v0.x; si.y += g0 * v0.y; si.z += g0 * v0.z; si.w += g0 * v0.w;
    }
    *reinterpret_cast<float4 *>(ctx + (size_t)n * H + j) = make_float4(so.x + si.x, so.y + si.y, so.z + si.z, so.w + si.w);
  }
}

template <int BM, int MODE>
static int launch_gru_t(const GruArgs &a, cudaStream_t st) {
  static bool attr_done = false;
  const size_t smem = TileSmem<BM, 3>::bytes;
  if (!attr_done) {
    SGG_CUDA_TRY(cudaFuncSetAttribute(k_gru<BM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid(a.H / BN, (a.M + BM - 1) / BM);
  k_gru<BM, MODE><<<grid, NTHREADS, smem, st>>>(a);
  SGG_RETURN_IF_LAUNCH_FAILED("k_gru");
  return 0;
}

template <int MODE>
static int launch_gru(const GruArgs &a, cudaStream_t st) {
  if (a.M <= 0) return 0;
  const int sms = sgg_num_sms();
  const long t128 = (long)(a.H / BN) * ((a.M + 127) / 128), t64 = (long)(a.H / BN) * ((a.M + 63) / 64);
  const double c128 = (double)((t128 + sms - 1) / sms) * 2.0;
  const double c64 = (double)((t64 + 2L * sms - 1) / (2L * sms)) * 2.0 * 1.08;   // 2 CTAs/SM, slightly less efficient
  if (c64 < c128) return launch_gru_t<64, MODE>(a, st);
  return launch_gru_t<128, MODE>(a, st);
}

struct MpScratch {
  float *V[2], *Eh[2], *P, *a, *g, *ctx, *lin_ws;
  void *fused; size_t fused_bytes;        // mp_fused.cu scratch (3xFP16 engine)
};

static size_t mp_layout(MpScratch *s, void *ws, int N, int E, int H) {
  SggArena ar(ws, (size_t)-1);
  const size_t n1 = N > 0 ? N : 1, e1 = E > 0 ? E : 1;
  for (int i = 0; i < 2; ++i) { s->V[i] = ar.take<float>(n1 * H); s->Eh[i] = ar.take<float>(e1 * H); }
  s->P = ar.take<float>(n1 * 3 * H);
  s->a = ar.take<float>(n1 * 4);
  s->g = ar.take<float>(e1 * 4);
  s->ctx = ar.take<float>(n1 * H);
  s->lin_ws = ar.take<float>(tc_linear_workspace_floats(N, 3 * H, H) + 4);   // split-K partials of P = V W_ih^T
  s->fused_bytes = mpf::workspace_bytes(N, E, H);
  s->fused = ar.take<char>(s->fused_bytes);
  return ar.off;
}

// Tape written by the forward pass in training mode and consumed by mp_backward (mp_bwd.cu).
MpTape mp_tape_view(float *base, int N, int E, int H, int T) {
  MpTape t;
  const size_t n1 = N > 0 ? N : 1, e1 = E > 0 ? E : 1;
  size_t off = 0;
  auto take = [&](size_t n) { float *p = base ? base + off : nullptr; off += (n + 63) / 64 * 64; return p; };
  t.N = N; t.E = E; t.H = H; t.T = T;
  t.states = take((size_t)(T + 1) * ((size_t)N + E) * H);
  t.cacheV = take((size_t)(T + 1) * n1 * 4 * H);
  t.cacheE = take((size_t)(T + 1) * e1 * 4 * H);
  t.gates = take((size_t)(T > 0 ? T : 1) * e1 * 4);
  t.ctx = take((size_t)(T > 0 ? T : 1) * n1 * H);
  t.P = take((size_t)(T > 0 ? T : 1) * n1 * 3 * H);
  t.floats = off;
  return t;
}

// Two-stream schedule: the object branch (P = V W_ih^T, ctx, node GRU) runs on a side stream concurrently with the
// edge branch (gates, edge GRU) on the caller's stream; they exchange exactly what the data flow needs (V, gates, P)
// through events, which CUDA-graph capture turns into graph edges.  `obj_on_side`: obj_rep was produced on the side
// stream (L1 entry) so the node-side initial step may start there immediately.
int mp_forward(const float *obj_rep, const float *rel_rep, const void *graph_ws, const sgg_mp_weights *w, int N,
               int E, int H, int T, float *V_out, float *E_out, float *saved, void *ws, size_t ws_bytes,
               cudaStream_t st, bool obj_on_side, const mpf::Planes *obj_planes = nullptr,
               const mpf::Planes *rel_planes = nullptr, mpf::Planes *last_planes = nullptr) {
  if (N < 0 || E < 0 || T < 0 || H <= 0 || (H % BN) != 0)
    return sgg_set_err(SGG_E_BADARG, "mp_forward: N=%d E=%d H=%d T=%d (H must be a multiple of %d)", N, E, H, T, BN);
  if (!w || !graph_ws || (N > 0 && (!obj_rep || !V_out)) || (E > 0 && (!rel_rep || !E_out)))
    return sgg_set_err(SGG_E_BADARG, "mp_forward: null pointer");
  MpScratch s;
  const size_t need = mp_layout(&s, ws, N, E, H);
  if (need > ws_bytes || !ws) return sgg_set_err(SGG_E_WORKSPACE, "mp_forward: workspace %zu < %zu", ws_bytes, need);
  if (mpf::supported(w, N, E, H)) {
    // 3xFP16 engine: fused loop on ONE stream (2 launches per iteration, mp_fused.cu); join the object branch first
    if (obj_on_side) {
      cudaStream_t sb0 = side_stream(st, 0);
      int rc0;
      if (sb0 != nullptr && (rc0 = stream_order(sb0, st))) return rc0;
    }
    return mpf::forward(obj_rep, rel_rep, obj_planes, rel_planes, graph_ws, w, N, E, H, T, V_out, E_out, saved, s.fused,
                        s.fused_bytes, st, last_planes);
  }
  SggGraphView g = sgg_graph_view(graph_ws, N, E);
  const size_t vN = (size_t)N * H, eN = (size_t)E * H;
  MpTape tape = mp_tape_view(saved, N, E, H, T);
  auto cacheV = [&](int it) -> float * { return saved ? tape.cacheV + (size_t)it * N * 4 * H : nullptr; };
  auto cacheE = [&](int it) -> float * { return saved ? tape.cacheE + (size_t)it * E * 4 * H : nullptr; };
  // State buffers: with `saved` every iteration writes into its own slot (no ping-pong copy);
  // otherwise ping-pong in the workspace and write the last iteration straight to the outputs.
  auto vbuf = [&](int it) -> float * {
    if (saved) return saved + (size_t)it * (vN + eN);
    return it == T ? V_out : s.V[it & 1];
  };
  auto ebuf = [&](int it) -> float * {
    if (saved) return saved + (size_t)it * (vN + eN) + vN;
    return it == T ? E_out : s.Eh[it & 1];
  };
  int rc;
  cudaStream_t sb = side_stream(st, 0), sc = side_stream(st, 1);
  const bool par = sb != nullptr && sc != nullptr && N > 0 && E > 0;
  cudaStream_t sn = par ? sb : st;                 // stream of the object branch (ctx, node GRU)
  cudaStream_t sp = par ? sc : st;                 // stream of P = V W_ih^T (needs only V: starts at the iteration boundary)
  if (par && !obj_on_side && (rc = stream_order(st, sb))) return rc;
  if (!par && obj_on_side && sb != nullptr && (rc = stream_order(sb, st))) return rc;   // caller forked, we run serial
  {  // hx = 0 initial step (:68-72)
    GruArgs a{}; a.x = obj_rep; a.h = nullptr; a.w_ih = w->node_w_ih; a.w_hh = w->node_w_hh; a.b_ih = w->node_b_ih;
    a.b_hh = w->node_b_hh; a.out = vbuf(0); a.cache = cacheV(0); a.M = N; a.H = H;
    if (w->node_w_ih_split) {
      if ((rc = tc_gru(0, obj_rep, nullptr, w->node_w_ih_split, nullptr, w->node_b_ih, w->node_b_hh, nullptr, nullptr,
                       nullptr, nullptr, vbuf(0), cacheV(0), N, H, sn))) return rc;
    } else if ((rc = launch_gru<GRU_INIT>(a, sn))) return rc;
    GruArgs b{}; b.x = rel_rep; b.h = nullptr; b.w_ih = w->edge_w_ih; b.w_hh = w->edge_w_hh; b.b_ih = w->edge_b_ih;
    b.b_hh = w->edge_b_hh; b.out = ebuf(0); b.cache = cacheE(0); b.M = E; b.H = H;
    if (w->edge_w_ih_split) {
      if ((rc = tc_gru(0, rel_rep, nullptr, w->edge_w_ih_split, nullptr, w->edge_b_ih, w->edge_b_hh, nullptr, nullptr,
                       nullptr, nullptr, ebuf(0), cacheE(0), E, H, st))) return rc;
    } else if ((rc = launch_gru<GRU_INIT>(b, st))) return rc;
  }
  for (int it = 0; it < T; ++it) {
    const float *V = vbuf(it), *Eh = ebuf(it);
    if (saved) {   // training: per-iteration intermediates go to the tape instead of the scratch
      s.g = tape.gates + (size_t)it * E * 4;
      s.ctx = tape.ctx + (size_t)it * N * H;
      s.P = tape.P + (size_t)it * N * 3 * H;
    }
    if (par) {     // iteration boundary: every branch sees the others' previous-iteration results / reads
      if ((rc = stream_order(sb, st))) return rc;     // V_it ready for the gates; old gates / ctx no longer read
      if ((rc = stream_order(st, sb))) return rc;     // Eh_it ready for ctx
      if ((rc = stream_order(sb, sp))) return rc;     // V_it ready for P
      if ((rc = stream_order(st, sp))) return rc;     // previous edge GRU no longer reads P
    }
    if (N > 0) {
      k_gate_node<<<(N * 32 + 255) / 256, 256, 0, st>>>(V, N, H, w->gate_w[0], w->gate_w[1], w->gate_w[2],
                                                        w->gate_w[3], s.a);
      SGG_RETURN_IF_LAUNCH_FAILED("k_gate_node");
    }
    if (E > 0) {
      k_gate_edge<<<(int)(((size_t)E * 32 + 255) / 256), 256, 0, st>>>(
          Eh, E, H, w->gate_w[0], w->gate_w[1], w->gate_w[2], w->gate_w[3], w->gate_b[0], w->gate_b[1], w->gate_b[2],
          w->gate_b[3], s.a, g.subj, g.obj, s.g);
      SGG_RETURN_IF_LAUNCH_FAILED("k_gate_edge");
      if (w->edge_w_ih_split) {
        if ((rc = tc_linear(V, w->edge_w_ih_split, nullptr, s.P, N, 3 * H, H, 0, s.lin_ws, sp))) return rc;
      } else if ((rc = launch_linear(V, w->edge_w_ih, nullptr, s.P, N, 3 * H, H, 0, sp))) return rc;
    }
    if (par) {
      if ((rc = stream_order(st, sb))) return rc;     // gates -> ctx
      if ((rc = stream_order(sp, st))) return rc;     // P -> edge GRU
    }
    if (N > 0) {
      k_ctx<<<N, (H / 4 < 128 ? H / 4 : 128), 0, sn>>>(Eh, s.g, g.out_ptr, g.out_idx, g.in_ptr, g.in_idx, H, s.ctx);
      SGG_RETURN_IF_LAUNCH_FAILED("k_ctx");
    }
    {
      GruArgs a{}; a.h = Eh; a.w_ih = w->edge_w_ih; a.w_hh = w->edge_w_hh; a.b_ih = w->edge_b_ih;
      a.b_hh = w->edge_b_hh; a.P = s.P; a.gates = s.g; a.subj = g.subj; a.obj = g.obj; a.out = ebuf(it + 1);
      a.cache = cacheE(it + 1);
      a.M = E; a.H = H;
      if (w->edge_w_hh_split) {
        if ((rc = tc_gru(2, nullptr, Eh, nullptr, w->edge_w_hh_split, w->edge_b_ih, w->edge_b_hh, s.P, s.g, g.subj,
                         g.obj, ebuf(it + 1), cacheE(it + 1), E, H, st))) return rc;
      } else if ((rc = launch_gru<GRU_EDGE>(a, st))) return rc;
    }
    {
      GruArgs a{}; a.x = s.ctx; a.h = V; a.w_ih = w->node_w_ih; a.w_hh = w->node_w_hh; a.b_ih = w->node_b_ih;
      a.b_hh = w->node_b_hh; a.out = vbuf(it + 1); a.cache = cacheV(it + 1); a.M = N; a.H = H;
      if (w->node_w_ih_split && w->node_w_hh_split) {
        if ((rc = tc_gru(1, s.ctx, V, w->node_w_ih_split, w->node_w_hh_split, w->node_b_ih, w->node_b_hh, nullptr,
                         nullptr, nullptr, nullptr, vbuf(it + 1), cacheV(it + 1), N, H, sn))) return rc;
      } else if ((rc = launch_gru<GRU_NODE>(a, sn))) return rc;
    }
  }
  if (par && (rc = stream_order(sb, st))) return rc;    // join: V_T (and everything on the side stream) -> caller's stream
  if (saved) {   // outputs are the last saved slot
    if (N > 0) SGG_CUDA_TRY(cudaMemcpyAsync(V_out, vbuf(T), vN * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (E > 0) SGG_CUDA_TRY(cudaMemcpyAsync(E_out, ebuf(T), eN * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

size_t mp_workspace_bytes(int N, int E, int H) {
  MpScratch s;
  return mp_layout(&s, nullptr, N, E, H);
}

}  // namespace sgg

// One edge-GRU update in isolation (the dominant kernel of the path): Eh' = GRU(g_s P[s] + g_o P[o] + b_ih,
// Eh W_hh^T + b_hh, Eh).  Exposed for unit tests and for bench.py's per-kernel roofline timing.
extern "C" int sgg_edge_gru_forward(const float *Eh, const float *P, const float *gates, const void *graph_ws,
                                    const float *w_ih, const float *w_hh, const float *w_hh_split, const float *b_ih,
                                    const float *b_hh, int N, int E, int H, float *out, void *stream) {
  using namespace sgg;
  if (E <= 0) return 0;
  if (!Eh || !P || !gates || !graph_ws || !w_hh || !b_ih || !b_hh || !out || (H % BN) != 0)
    return sgg_set_err(SGG_E_BADARG, "edge_gru_forward: bad argument");
  SggGraphView g = sgg_graph_view(graph_ws, N, E);
  cudaStream_t st = (cudaStream_t)stream;
  if (w_hh_split) return tc_gru(2, nullptr, Eh, nullptr, w_hh_split, b_ih, b_hh, P, gates, g.subj, g.obj, out, nullptr, E, H, st);
  GruArgs a{}; a.h = Eh; a.w_ih = w_ih; a.w_hh = w_hh; a.b_ih = b_ih; a.b_hh = b_hh; a.P = P; a.gates = gates;
  a.subj = g.subj; a.obj = g.obj; a.out = out; a.M = E; a.H = H;
  return launch_gru<GRU_EDGE>(a, st);
}

// Measurement probe (bench.py roofline): ONE launch of the fused message-passing schedule, iteration 0, on a workspace
// that a preceding sgg_mp_forward call with the same arguments has initialised.  which: 0 = INIT launch (k_mp_gru),
// 1 = launch A (k_mp_pre: P/Q tiles + ctx gather), 2 = launch B (k_mp_gru: edge + node GRU tiles).
extern "C" int sgg_mp_probe_launch(int which, const float *obj_rep, const float *rel_rep, const void *graph_ws,
                                   const sgg_mp_weights *w, int N, int E, int H, int T, float *V_out, float *E_out,
                                   void *ws, size_t ws_bytes, void *stream) {
  using namespace sgg;
  if (which < 0 || which > 2 || !w || !graph_ws || !ws) return sgg_set_err(SGG_E_BADARG, "mp_probe_launch: bad argument");
  if (!mpf::supported(w, N, E, H)) return sgg_set_err(SGG_E_BADARG, "mp_probe_launch: fused path not active for this shape / engine");
  MpScratch s;
  const size_t need = mp_layout(&s, ws, N, E, H);
  if (need > ws_bytes) return sgg_set_err(SGG_E_WORKSPACE, "mp_probe_launch: workspace %zu < %zu", ws_bytes, need);
  return mpf::forward(obj_rep, rel_rep, nullptr, nullptr, graph_ws, w, N, E, H, T, V_out, E_out, nullptr, s.fused,
                      s.fused_bytes, (cudaStream_t)stream, nullptr, which);
}

extern "C" size_t sgg_mp_tape_bytes(int N, int E, int H, int T) {
  return sgg::mp_tape_view(nullptr, N < 0 ? 0 : N, E < 0 ? 0 : E, H, T < 0 ? 0 : T).floats * sizeof(float);
}

extern "C" size_t sgg_mp_workspace_bytes(int N, int E, int H, int T) {
  (void)T;
  return sgg::mp_workspace_bytes(N < 0 ? 0 : N, E < 0 ? 0 : E, H);
}

extern "C" int sgg_mp_forward(const float *obj_rep, const float *rel_rep, const void *graph_ws,
                              const sgg_mp_weights *w, int N, int E, int H, int T, float *V_out, float *E_out,
                              float *saved, void *ws, size_t ws_bytes, void *stream) {
  return sgg::mp_forward(obj_rep, rel_rep, graph_ws, w, N, E, H, T, V_out, E_out, saved, ws, ws_bytes,
                         (cudaStream_t)stream, false);
}

// ---- L1: 4096-d features -> dists (rel_model_stanford.py:103-107 without roi_fmap*) ----
namespace sgg {
struct L1Scratch {
  float *obj_rep, *rel_rep, *V, *Eh, *ws_obj, *ws_edge; void *mp; size_t mp_bytes;
  mpf::Planes obj_pl, rel_pl;             // fp16 [hi | lo] planes of the unary outputs (fused message passing)
};
static size_t l1_layout(L1Scratch *s, void *ws, int N, int E, int H, int D = 4096, int n_cls = 151, int n_rel = 51) {
  SggArena ar(ws, (size_t)-1);
  const size_t n1 = N > 0 ? N : 1, e1 = E > 0 ? E : 1;
  s->obj_rep = ar.take<float>(n1 * H); s->rel_rep = ar.take<float>(e1 * H);
  s->V = ar.take<float>(n1 * H); s->Eh = ar.take<float>(e1 * H);
  s->obj_pl.hi = ar.take<__half>(n1 * H); s->obj_pl.lo = ar.take<__half>(n1 * H);
  s->rel_pl.hi = ar.take<__half>(e1 * H); s->rel_pl.lo = ar.take<__half>(e1 * H);
  const size_t o1 = tc_linear_workspace_floats(N, H, D), o2 = tc_linear_workspace_floats(N, n_cls, H);
  const size_t e1w = tc_linear_workspace_floats(E, H, D), e2w = tc_linear_workspace_floats(E, n_rel, H);
  s->ws_obj = ar.take<float>((o1 > o2 ? o1 : o2) + 4);      // object-branch and edge-branch linears may overlap
  s->ws_edge = ar.take<float>((e1w > e2w ? e1w : e2w) + 4);
  s->mp_bytes = mp_workspace_bytes(N, E, H);
  s->mp = ar.take<char>(s->mp_bytes);
  return ar.off;
}
}  // namespace sgg

extern "C" size_t sgg_l1_workspace_bytes(int N, int E, int H, int T) {
  (void)T;
  sgg::L1Scratch s;
  return sgg::l1_layout(&s, nullptr, N < 0 ? 0 : N, E < 0 ? 0 : E, H);
}

static int l1_forward_impl(const float *obj_feat, const float *edge_feat, const int64_t *rel_inds, int64_t row_stride,
                           int col_subj, int col_obj, void *graph_ws, size_t graph_ws_bytes,
                           const sgg_head_weights *hw, const sgg_mp_weights *w, int N, int E, int D, int H, int T,
                           int n_cls, int n_rel, float *obj_dists, float *rel_dists, void *ws, size_t ws_bytes,
                           void *stream) {
  if (!hw || !w || !ws || !graph_ws) return sgg_set_err(SGG_E_BADARG, "l1_forward: null pointer");
  if (N < 0 || E < 0 || H <= 0 || (H % sgg::BN) != 0) return sgg_set_err(SGG_E_BADARG, "l1_forward: bad shape");
  sgg::L1Scratch s;
  const size_t need = sgg::l1_layout(&s, ws, N, E, H);
  if (need > ws_bytes) return sgg_set_err(SGG_E_WORKSPACE, "l1_forward: workspace %zu < %zu", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  auto lin = [&](const float *x, const float *wt, const float *wsplit, const float *b, float *y, int M, int No, int K,
                 int relu, float *lws, cudaStream_t s2) -> int {
    return wsplit ? sgg::tc_linear(x, wsplit, b, y, M, No, K, relu, lws, s2) : sgg::launch_linear(x, wt, b, y, M, No, K, relu, s2);
  };
  cudaStream_t sb = sgg::side_stream(st);
  const bool par = sb != nullptr && N > 0 && E > 0;
  cudaStream_t sn = par ? sb : st;
  if (par && (rc = sgg::stream_order(st, sb))) return rc;         // fork: inputs (and a prebuilt graph) are ready on `st`
  // graph index on the object-branch stream: it overlaps the edge-unary GEMM; every consumer is either on that
  // stream (k_ctx) or behind the iteration-0 join of mp_forward (gates, edge GRU)
  if (rel_inds && (rc = sgg_graph_build(rel_inds, row_stride, col_subj, col_obj, N, E, graph_ws, graph_ws_bytes, sn))) return rc;
  // fused message passing (3xFP16): the unary epilogues also emit the fp16 operand planes the INIT launch reads via TMA
  const bool planes = sgg::mpf::supported(w, N, E, H) && hw->obj_unary_w_split && hw->edge_unary_w_split && (D % 8) == 0;
  if (planes) {
    if ((rc = sgg::tc16::linear_planes(obj_feat, hw->obj_unary_w_split, hw->obj_unary_b, s.obj_rep, s.obj_pl.hi, s.obj_pl.lo,
                                       N, H, D, 0, s.ws_obj, sn))) return rc;
    if ((rc = sgg::tc16::linear_planes(edge_feat, hw->edge_unary_w_split, hw->edge_unary_b, s.rel_rep, s.rel_pl.hi,
                                       s.rel_pl.lo, E, H, D, 1, s.ws_edge, st))) return rc;
  } else {
    if ((rc = lin(obj_feat, hw->obj_unary_w, hw->obj_unary_w_split, hw->obj_unary_b, s.obj_rep, N, H, D, 0, s.ws_obj, sn))) return rc;
    if ((rc = lin(edge_feat, hw->edge_unary_w, hw->edge_unary_w_split, hw->edge_unary_b, s.rel_rep, E, H, D, 1, s.ws_edge, st))) return rc;
  }
  // ... and the classifier heads read the final states through planes emitted by the last GRU launch: one launch
  const bool fheads = planes && hw->obj_fc_w_split && hw->rel_fc_w_split;
  sgg::mpf::Planes lastp[2];
  if ((rc = sgg::mp_forward(s.obj_rep, s.rel_rep, graph_ws, w, N, E, H, T, s.V, s.Eh, nullptr, s.mp, s.mp_bytes, st, par,
                            planes ? &s.obj_pl : nullptr, planes ? &s.rel_pl : nullptr, fheads ? lastp : nullptr)))
    return rc;
  if (fheads) return sgg::mpf::heads(lastp[0], lastp[1], hw, N, E, H, n_cls, n_rel, obj_dists, rel_dists, st);
  // heads: mp_forward joined the side stream into `st`; fork again for the two classifiers
  if (par && (rc = sgg::stream_order(st, sb))) return rc;
  if ((rc = lin(s.V, hw->obj_fc_w, hw->obj_fc_w_split, hw->obj_fc_b, obj_dists, N, n_cls, H, 0, s.ws_obj, sn))) return rc;
  if ((rc = lin(s.Eh, hw->rel_fc_w, hw->rel_fc_w_split, hw->rel_fc_b, rel_dists, E, n_rel, H, 0, s.ws_edge, st))) return rc;
  if (par && (rc = sgg::stream_order(sb, st))) return rc;
  return 0;
}

extern "C" int sgg_l1_forward(const float *obj_feat, const float *edge_feat, const void *graph_ws,
                              const sgg_head_weights *hw, const sgg_mp_weights *w, int N, int E, int D, int H, int T,
                              int n_cls, int n_rel, float *obj_dists, float *rel_dists, void *ws, size_t ws_bytes,
                              void *stream) {
  return l1_forward_impl(obj_feat, edge_feat, nullptr, 0, 0, 0, const_cast<void *>(graph_ws), 0, hw, w, N, E, D, H, T,
                         n_cls, n_rel, obj_dists, rel_dists, ws, ws_bytes, stream);
}

extern "C" int sgg_l1_forward_rel(const float *obj_feat, const float *edge_feat, const int64_t *rel_inds,
                                  int64_t row_stride, int col_subj, int col_obj, void *graph_ws, size_t graph_ws_bytes,
                                  const sgg_head_weights *hw, const sgg_mp_weights *w, int N, int E, int D, int H, int T,
                                  int n_cls, int n_rel, float *obj_dists, float *rel_dists, void *ws, size_t ws_bytes,
                                  void *stream) {
  if (E > 0 && !rel_inds) return sgg_set_err(SGG_E_BADARG, "l1_forward_rel: null rel_inds");
  return l1_forward_impl(obj_feat, edge_feat, rel_inds, row_stride, col_subj, col_obj, graph_ws, graph_ws_bytes, hw, w, N,
                         E, D, H, T, n_cls, n_rel, obj_dists, rel_dists, ws, ws_bytes, stream);
}
