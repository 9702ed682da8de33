// node_edge_features (sgg_models/rel_model_base.py:245-260): RoIAlign of the objects
// and of the UNION boxes of every candidate pair.  Arithmetic follows torchvision's
// roi_align (aligned=False) as called through MultiScaleRoIAlign(['0'], 7,
// sampling_ratio=2) at rel_model_base.py:97-99,258-259; the union box
// (min x1y1, max x2y2 — :248-250) is computed on the fly, so the [E,5] union_rois
// tensor and the per-image roi lists of convert_roi_to_list (:303-307, a host sync)
// never exist.
#include "common.cuh"

namespace sgg {

__device__ __forceinline__ float bilinear(const float *__restrict__ f, int Hf, int Wf, float y, float x) {
  if (y < -1.0f || y > (float)Hf || x < -1.0f || x > (float)Wf) return 0.f;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= Hf - 1) { yh = yl = Hf - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= Wf - 1) { xh = xl = Wf - 1; x = (float)xl; } else xh = xl + 1;
  const float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
  const float v1 = f[yl * Wf + xl], v2 = f[yl * Wf + xh], v3 = f[yh * Wf + xl], v4 = f[yh * Wf + xh];
  return hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
}

// One thread per output element (r, c, ph, pw); r < N are objects, r >= N union boxes.
__global__ void __launch_bounds__(256)
k_roi_align(const float *__restrict__ fmap, int C, int Hf, int Wf, const float *__restrict__ rois, int N,
            const int64_t *__restrict__ ui, int64_t stride, int cs, int co, int E, float scale, int pool, int sr,
            float *__restrict__ node_out, float *__restrict__ edge_out, int do_node, int do_edge) {
  const int pp = pool * pool;
  const size_t per = (size_t)C * pp;
  const size_t r_begin = do_node ? 0 : (size_t)N, r_end = do_edge ? (size_t)N + E : (size_t)N;
  const size_t total = (r_end - r_begin) * per;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = r_begin + i / per;
    const int rem = (int)(i % per);
    const int c = rem / pp, ph = (rem % pp) / pool, pw = rem % pool;
    float x1, y1, x2, y2; int b;
    float *dst;
    if (r < (size_t)N) {
      const float *q = rois + r * 5;
      b = (int)q[0]; x1 = q[1]; y1 = q[2]; x2 = q[3]; y2 = q[4];
      dst = node_out + r * per + rem;
    } else {
      const size_t e = r - N;
      const float *qs = rois + (size_t)ui[e * stride + cs] * 5, *qo = rois + (size_t)ui[e * stride + co] * 5;
      b = (int)qs[0];
      x1 = fminf(qs[1], qo[1]); y1 = fminf(qs[2], qo[2]); x2 = fmaxf(qs[3], qo[3]); y2 = fmaxf(qs[4], qo[4]);
      dst = edge_out + e * per + rem;
    }
    const float sw = x1 * scale, sh = y1 * scale;
    const float rw = fmaxf(x2 * scale - sw, 1.f), rh = fmaxf(y2 * scale - sh, 1.f);
    const float bw = rw / (float)pool, bh = rh / (float)pool;
    const float *f = fmap + ((size_t)b * C + c) * Hf * Wf;
    float acc = 0.f;
    for (int iy = 0; iy < sr; ++iy) {
      const float y = sh + ph * bh + (iy + 0.5f) * bh / (float)sr;
      for (int ix = 0; ix < sr; ++ix) {
        const float x = sw + pw * bw + (ix + 0.5f) * bw / (float)sr;
        acc += bilinear(f, Hf, Wf, y, x);
      }
    }
    *dst = acc / (float)(sr * sr);
  }
}

}  // namespace sgg

extern "C" int sgg_node_edge_features(const float *fmap, int B, int C, int Hf, int Wf, const float *rois, int N,
                                      const int64_t *union_inds, int64_t row_stride, int col_subj, int col_obj, int E,
                                      float spatial_scale, int pool, int sampling_ratio, float *node_feat,
                                      float *edge_feat, void *stream) {
  if (B <= 0 || C <= 0 || Hf <= 0 || Wf <= 0 || N < 0 || E < 0 || pool <= 0 || sampling_ratio <= 0)
    return sgg_set_err(SGG_E_BADARG, "node_edge_features: bad shape");
  const int do_node = node_feat != nullptr && N > 0, do_edge = edge_feat != nullptr && E > 0;
  if (!do_node && !do_edge) return 0;
  if (!fmap || !rois || (do_edge && !union_inds)) return sgg_set_err(SGG_E_BADARG, "node_edge_features: null pointer");
  const size_t total = ((size_t)(do_node ? N : 0) + (do_edge ? E : 0)) * C * pool * pool;
  int blocks = (int)((total + 255) / 256 < (size_t)sgg_num_sms() * 32 ? (total + 255) / 256 : (size_t)sgg_num_sms() * 32);
  sgg::k_roi_align<<<blocks, 256, 0, (cudaStream_t)stream>>>(fmap, C, Hf, Wf, rois, N, union_inds, row_stride,
                                                             col_subj, col_obj, E, spatial_scale, pool,
                                                             sampling_ratio, node_feat, edge_feat, do_node, do_edge);
  SGG_RETURN_IF_LAUNCH_FAILED("k_roi_align");
  return 0;
}
