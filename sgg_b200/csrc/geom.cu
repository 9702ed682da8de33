// Union-box geometry: draw_union_boxes (the reference's only native code,
// lib/draw_rectangles/draw_rectangles.pyx:12-67, CPU Cython with a D2H/H2D round
// trip at lib/get_union_boxes.py:66-67) and the conv branch of
// UnionBoxesAndFeats (lib/get_union_boxes.py:51-59,101) computed on the device.
//
// Because of the reference's stride-capture quirk (both convs run with stride 16,
// get_union_boxes.py:40-43) the branch reduces, per edge, to
//   conv7x7/s16/p3 at 4 output positions -> ReLU -> BN1 -> max over the 4 positions
//   -> centre tap of the 3x3 conv (a [C, C/2] matvec) -> ReLU -> BN2  => geom[E, C]
// which is broadcast-added to the pooled union features.  The 27x27 masks are
// rebuilt in shared memory from the boxes and never touch HBM.
#include "common.cuh"
#include "kernels.h"

namespace sgg {

constexpr int GP = 27;   // mask size = pooling_size*4-1 (get_union_boxes.py:67)

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

struct PairBoxes { float b[8]; };

__device__ __forceinline__ PairBoxes load_pair(const float *rois, const int64_t *ui, int64_t stride, int cs, int co,
                                               int e) {
  PairBoxes p;
  const float *rs = rois + (size_t)ui[e * stride + cs] * 5 + 1;
  const float *ro = rois + (size_t)ui[e * stride + co] * 5 + 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) { p.b[i] = rs[i]; p.b[4 + i] = ro[i]; }
  return p;
}

// Same float32 operation order as the pyx (:45-66); *_rn intrinsics forbid FMA contraction
// so the masks are bit-identical to the Cython output.
__device__ __forceinline__ float mask_value(const PairBoxes &p, int ch, int j, int k, int P) {
  const float x1u = fminf(p.b[0], p.b[4]), y1u = fminf(p.b[1], p.b[5]);
  const float x2u = fmaxf(p.b[2], p.b[6]), y2u = fmaxf(p.b[3], p.b[7]);
  const float w = __fsub_rn(x2u, x1u), h = __fsub_rn(y2u, y1u);
  const float fp = (float)P;
  const float x1 = __fdiv_rn(__fmul_rn(__fsub_rn(p.b[0 + 4 * ch], x1u), fp), w);
  const float y1 = __fdiv_rn(__fmul_rn(__fsub_rn(p.b[1 + 4 * ch], y1u), fp), h);
  const float x2 = __fdiv_rn(__fmul_rn(__fsub_rn(p.b[2 + 4 * ch], x1u), fp), w);
  const float y2 = __fdiv_rn(__fmul_rn(__fsub_rn(p.b[3 + 4 * ch], y1u), fp), h);
  const float yc = __fmul_rn(clamp01(__fsub_rn((float)(j + 1), y1)), clamp01(__fsub_rn(y2, (float)j)));
  const float xc = __fmul_rn(clamp01(__fsub_rn((float)(k + 1), x1)), clamp01(__fsub_rn(x2, (float)k)));
  return __fmul_rn(xc, yc);
}

__global__ void k_draw_union_boxes(const float *__restrict__ rois, const int64_t *__restrict__ ui, int64_t stride,
                                   int cs, int co, int E, int P, float sub, float *__restrict__ out) {
  const size_t total = (size_t)E * 2 * P * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int k = (int)(i % P), j = (int)((i / P) % P), ch = (int)((i / ((size_t)P * P)) % 2), e = (int)(i / ((size_t)2 * P * P));
    PairBoxes p = load_pair(rois, ui, stride, cs, co, e);
    out[i] = mask_value(p, ch, j, k, P) - sub;
  }
}

// conv1 weights [C1,2,7,7] -> w1t[98][C1]; conv2 centre taps [C,C1,3,3] -> w2c[C][C1].
__global__ void k_geom_prep(const float *__restrict__ w1, const float *__restrict__ w2, int C1, int C,
                            float *__restrict__ w1t, float *__restrict__ w2c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 98 * C1) { int tap = i / C1, c = i % C1; w1t[i] = w1[(size_t)c * 98 + tap]; }
  if (i < C * C1) w2c[i] = w2[(size_t)i * 9 + 4];
}

// Post-ReLU conv1 outputs at the 4 positions: c1out[e][pos][c1].  One CTA = GE edges, thread = channel.
constexpr int GE = 4;
__global__ void __launch_bounds__(256) k_geom_conv1(const float *__restrict__ rois, const int64_t *__restrict__ ui,
                                                    int64_t stride, int cs, int co, int E, int C1,
                                                    const float *__restrict__ w1t, const float *__restrict__ b1,
                                                    float *__restrict__ c1out) {
  __shared__ float rect[GE][2][GP][GP + 1];
  const int e0 = blockIdx.x * GE;
  for (int i = threadIdx.x; i < GE * 2 * GP * GP; i += blockDim.x) {
    int k = i % GP, j = (i / GP) % GP, ch = (i / (GP * GP)) % 2, ge = i / (2 * GP * GP);
    float v = 0.f;
    if (e0 + ge < E) {
      PairBoxes p = load_pair(rois, ui, stride, cs, co, e0 + ge);
      v = mask_value(p, ch, j, k, GP) - 0.5f;
    }
    rect[ge][ch][j][k] = v;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C1; c += blockDim.x) {
    float acc[GE][4];
#pragma unroll
    for (int ge = 0; ge < GE; ++ge)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[ge][q] = 0.f;
    for (int ch = 0; ch < 2; ++ch)
      for (int ky = 0; ky < 7; ++ky)
        for (int kx = 0; kx < 7; ++kx) {
          const float wv = w1t[(size_t)((ch * 7 + ky) * 7 + kx) * C1 + c];
#pragma unroll
          for (int oy = 0; oy < 2; ++oy) {
            const int y = oy * 16 - 3 + ky;
            if (y < 0 || y >= GP) continue;
#pragma unroll
            for (int ox = 0; ox < 2; ++ox) {
              const int x = ox * 16 - 3 + kx;
              if (x < 0 || x >= GP) continue;
#pragma unroll
              for (int ge = 0; ge < GE; ++ge) acc[ge][oy * 2 + ox] = fmaf(wv, rect[ge][ch][y][x], acc[ge][oy * 2 + ox]);
            }
          }
        }
    const float bb = b1[c];
#pragma unroll
    for (int ge = 0; ge < GE; ++ge)
      if (e0 + ge < E)
#pragma unroll
        for (int q = 0; q < 4; ++q) c1out[((size_t)(e0 + ge) * 4 + q) * C1 + c] = fmaxf(acc[ge][q] + bb, 0.f);
  }
}

// im2col of the 4 live conv1 windows: patches[e][pos][ch*49 + ky*7 + kx] (zero where the window hangs over the
// padding) — training path: conv1 then runs as a linear op with autograd.
__global__ void k_geom_patches(const float *__restrict__ rois, const int64_t *__restrict__ ui, int64_t stride, int cs,
                               int co, int E, float *__restrict__ out) {
  const size_t total = (size_t)E * 4 * 98;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 98), pos = (int)((i / 98) % 4), e = (int)(i / (4 * 98));
    const int ch = tap / 49, ky = (tap % 49) / 7, kx = tap % 7;
    const int y = (pos >> 1) * 16 - 3 + ky, x = (pos & 1) * 16 - 3 + kx;
    float v = 0.f;
    if (y >= 0 && y < GP && x >= 0 && x < GP) {
      PairBoxes p = load_pair(rois, ui, stride, cs, co, e);
      v = mask_value(p, ch, y, x, GP) - 0.5f;
    }
    out[i] = v;
  }
}

// hid[e][c] = max_pos BN1(c1out[e][pos][c])      (BatchNorm then MaxPool2d(3,2,1) on a 2x2 map)
__global__ void k_geom_pool(const float *__restrict__ c1out, int E, int C1, const float *__restrict__ g,
                            const float *__restrict__ b, const float *__restrict__ mean,
                            const float *__restrict__ var, float eps, float *__restrict__ hid) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)E * C1) return;
  const int c = (int)(i % C1); const size_t e = i / C1;
  const float sc = g[c] / sqrtf(var[c] + eps), sh = b[c] - mean[c] * sc;
  float m = -INFINITY;
#pragma unroll
  for (int q = 0; q < 4; ++q) m = fmaxf(m, fmaf(c1out[(e * 4 + q) * C1 + c], sc, sh));
  hid[i] = m;
}

// out = (pools +) BN2(t)   with t[e][c] = relu(conv2 centre tap)
__global__ void k_geom_finish(const float *__restrict__ t, int E, int C, int spatial, const float *__restrict__ g,
                              const float *__restrict__ b, const float *__restrict__ mean,
                              const float *__restrict__ var, float eps, const float *__restrict__ pools,
                              float *__restrict__ out) {
  const size_t total = (size_t)E * C * spatial;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t ec = i / spatial;
    const int c = (int)(ec % C);
    const float sc = g[c] / sqrtf(var[c] + eps);
    const float v = fmaf(t[ec] - mean[c], sc, b[c]);
    out[i] = pools ? pools[i] + v : v;
  }
}

struct GeomScratch { float *w1t, *w2c, *c1out, *hid, *t; };
static size_t geom_layout(GeomScratch *s, void *ws, int E, int C) {
  SggArena ar(ws, (size_t)-1);
  const size_t e1 = E > 0 ? E : 1; const int C1 = C / 2;
  s->w1t = ar.take<float>((size_t)98 * C1);
  s->w2c = ar.take<float>((size_t)C * C1);
  s->c1out = ar.take<float>(e1 * 4 * C1);
  s->hid = ar.take<float>(e1 * C1);
  s->t = ar.take<float>(e1 * C);
  return ar.off;
}

}  // namespace sgg

extern "C" int sgg_draw_union_boxes(const float *rois, const int64_t *union_inds, int64_t row_stride, int col_subj,
                                    int col_obj, int E, int P, int sub_half, float *out, void *stream) {
  if (E < 0 || P <= 0) return sgg_set_err(SGG_E_BADARG, "draw_union_boxes: bad shape");
  if (E == 0) return 0;
  if (!rois || !union_inds || !out) return sgg_set_err(SGG_E_BADARG, "draw_union_boxes: null pointer");
  const size_t total = (size_t)E * 2 * P * P;
  int blocks = (int)((total + 255) / 256);
  const int cap = sgg_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  sgg::k_draw_union_boxes<<<blocks, 256, 0, (cudaStream_t)stream>>>(rois, union_inds, row_stride, col_subj, col_obj, E,
                                                                     P, sub_half ? 0.5f : 0.f, out);
  SGG_RETURN_IF_LAUNCH_FAILED("k_draw_union_boxes");
  return 0;
}

extern "C" int sgg_geom_patches(const float *rois, const int64_t *union_inds, int64_t row_stride, int col_subj,
                                int col_obj, int E, float *out, void *stream) {
  if (E <= 0) return 0;
  if (!rois || !union_inds || !out) return sgg_set_err(SGG_E_BADARG, "geom_patches: null pointer");
  const size_t total = (size_t)E * 4 * 98;
  int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  sgg::k_geom_patches<<<blocks, 256, 0, (cudaStream_t)stream>>>(rois, union_inds, row_stride, col_subj, col_obj, E, out);
  SGG_RETURN_IF_LAUNCH_FAILED("k_geom_patches");
  return 0;
}

extern "C" size_t sgg_union_geom_workspace_bytes(int E, int C) {
  sgg::GeomScratch s;
  return sgg::geom_layout(&s, nullptr, E < 0 ? 0 : E, C);
}

extern "C" int sgg_union_geom_forward(const float *rois, const int64_t *union_inds, int64_t row_stride, int col_subj,
                                      int col_obj, int E, int C, const sgg_geom_weights *gw, const float *union_pools,
                                      float *out, void *ws, size_t ws_bytes, void *stream) {
  if (E < 0 || C <= 0 || (C % 32) != 0) return sgg_set_err(SGG_E_BADARG, "union_geom: bad shape (C %% 32 == 0)");
  if (E == 0) return 0;
  if (!rois || !union_inds || !gw || !out || !ws) return sgg_set_err(SGG_E_BADARG, "union_geom: null pointer");
  sgg::GeomScratch s;
  const size_t need = sgg::geom_layout(&s, ws, E, C);
  if (need > ws_bytes) return sgg_set_err(SGG_E_WORKSPACE, "union_geom: workspace %zu < %zu", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int C1 = C / 2;
  const float eps = 1e-5f;   // nn.BatchNorm2d default (get_union_boxes.py:54,58)
  {
    int n = C * C1 > 98 * C1 ? C * C1 : 98 * C1;
    sgg::k_geom_prep<<<(n + 255) / 256, 256, 0, st>>>(gw->conv1_w, gw->conv2_w, C1, C, s.w1t, s.w2c);
    SGG_RETURN_IF_LAUNCH_FAILED("k_geom_prep");
  }
  sgg::k_geom_conv1<<<(E + sgg::GE - 1) / sgg::GE, 256, 0, st>>>(rois, union_inds, row_stride, col_subj, col_obj, E, C1,
                                                                  s.w1t, gw->conv1_b, s.c1out);
  SGG_RETURN_IF_LAUNCH_FAILED("k_geom_conv1");
  sgg::k_geom_pool<<<(int)(((size_t)E * C1 + 255) / 256), 256, 0, st>>>(s.c1out, E, C1, gw->bn1_w, gw->bn1_b,
                                                                         gw->bn1_rm, gw->bn1_rv, eps, s.hid);
  SGG_RETURN_IF_LAUNCH_FAILED("k_geom_pool");
  int rc = sgg::launch_linear(s.hid, s.w2c, gw->conv2_b, s.t, E, C, C1, 1, st);
  if (rc) return rc;
  const int spatial = union_pools ? 49 : 1;
  const size_t total = (size_t)E * C * spatial;
  int blocks = (int)((total + 255) / 256);
  const int cap = sgg_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  sgg::k_geom_finish<<<blocks, 256, 0, st>>>(s.t, E, C, spatial, gw->bn2_w, gw->bn2_b, gw->bn2_rm, gw->bn2_rv, eps,
                                             union_pools, out);
  SGG_RETURN_IF_LAUNCH_FAILED("k_geom_finish");
  return 0;
}
