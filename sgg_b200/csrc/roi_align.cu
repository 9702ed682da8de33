// node_edge_features (sgg_models/rel_model_base.py:245-260): RoIAlign of the objects
// and of the UNION boxes of every candidate pair.  Arithmetic follows torchvision's
// roi_align (aligned=False) as called through MultiScaleRoIAlign(['0'], 7,
// sampling_ratio=2) at rel_model_base.py:97-99,258-259; the union box
// (min x1y1, max x2y2 — :248-250) is computed on the fly, so the [E,5] union_rois
// tensor and the per-image roi lists of convert_roi_to_list (:303-307, a host sync)
// never exist.
#include "common.cuh"
#include "tc16_common.cuh"

namespace sgg {
__device__ unsigned int g_roi_overflow = 0;   // sticky: an emitted operand plane left the fp16 range (sgg_tc16_overflow, bit 0)

// Indices come from the caller (rel_inds / im_inds): clamp them so a bad value can never become an out-of-bounds read
// (graph.cu flags the same situation for the message-passing graph; here the result for such a row is simply that of
// the clamped index).
__device__ __forceinline__ int clamp_idx(long long v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : (int)v); }


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
k_roi_align(const float *__restrict__ fmap, int B, int C, int Hf, int Wf, const float *__restrict__ rois, int N,
            const int64_t *__restrict__ ui, int64_t stride, int cs, int co, int E, float scale, int pool, int sr,
            float *__restrict__ node_out, float *__restrict__ edge_out, int do_node, int do_edge,
            const float *__restrict__ edge_add) {
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
      b = clamp_idx((long long)q[0], B); x1 = q[1]; y1 = q[2]; x2 = q[3]; y2 = q[4];
      dst = node_out + r * per + rem;
    } else {
      const size_t e = r - N;
      const float *qs = rois + (size_t)clamp_idx(ui[e * stride + cs], N) * 5, *qo = rois + (size_t)clamp_idx(ui[e * stride + co], N) * 5;
      b = clamp_idx((long long)qs[0], B);
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
    float v = acc / (float)(sr * sr);
    if (edge_add != nullptr && r >= (size_t)N) v += edge_add[(r - N) * C + c];
    *dst = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// Fast path: channel-last feature map.  A CTA owns one RoI; threads own channels, so every bilinear corner
// fetch is a coalesced read of consecutive channels, the separable sample geometry (14 y-samples x 14
// x-samples for 7x7 bins, sampling_ratio 2) is computed once per RoI in shared memory, and the [C,7,7] result
// is staged in shared memory and written out as one contiguous, coalesced block.
__global__ void k_nchw_to_nhwc(const float *__restrict__ in, int C, int HW, float *__restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float *src = in + (size_t)b * C * HW;
  float *dst = out + (size_t)b * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, pp = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && pp < HW) ? src[(size_t)c * HW + pp] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, c = c0 + threadIdx.x;
    if (pp < HW && c < C) dst[(size_t)pp * C + c] = tile[threadIdx.x][i];
  }
}

constexpr int RA_MAXS = 32;   // max samples per axis (pool * sampling_ratio)

__global__ void __launch_bounds__(256)
k_roi_align_nhwc(const float *__restrict__ fmap /*[B,Hf,Wf,C]*/, int B, int C, int Hf, int Wf, const float *__restrict__ rois,
                 int N, const int64_t *__restrict__ ui, int64_t stride, int cs, int co, int E, float scale, int pool,
                 int sr, float *__restrict__ node_out, float *__restrict__ edge_out, int r_begin,
                 const float *__restrict__ edge_add) {
  extern __shared__ __align__(16) float s_out[];          // [C][pool*pool]
  __shared__ int s_lo[2][RA_MAXS], s_hi[2][RA_MAXS];
  __shared__ float s_l[2][RA_MAXS], s_h[2][RA_MAXS];       // lerp weights (l = frac, h = 1 - frac); 0,0 when invalid
  const int r = r_begin + blockIdx.x;
  const int pp = pool * pool, ns = pool * sr;
  float x1, y1, x2, y2; int b;
  float *dst;
  if (r < N) {
    const float *q = rois + (size_t)r * 5;
    b = clamp_idx((long long)q[0], B); x1 = q[1]; y1 = q[2]; x2 = q[3]; y2 = q[4];
    dst = node_out + (size_t)r * C * pp;
  } else {
    const size_t e = (size_t)(r - N);
    const float *qs = rois + (size_t)clamp_idx(ui[e * stride + cs], N) * 5, *qo = rois + (size_t)clamp_idx(ui[e * stride + co], N) * 5;
    b = clamp_idx((long long)qs[0], B);
    x1 = fminf(qs[1], qo[1]); y1 = fminf(qs[2], qo[2]); x2 = fmaxf(qs[3], qo[3]); y2 = fmaxf(qs[4], qo[4]);
    dst = edge_out + e * C * pp;
  }
  if (threadIdx.x < 2 * ns) {
    const int axis = threadIdx.x / ns, i = threadIdx.x % ns;     // axis 0 = y, 1 = x
    const float start = (axis == 0 ? y1 : x1) * scale;
    const float len = fmaxf((axis == 0 ? y2 : x2) * scale - start, 1.f);
    const float bin = len / (float)pool;
    const int size = axis == 0 ? Hf : Wf;
    float v = start + (i / sr) * bin + ((i % sr) + 0.5f) * bin / (float)sr;
    int lo = 0, hi = 0; float l = 0.f, h = 0.f;
    if (!(v < -1.0f || v > (float)size)) {
      if (v <= 0.f) v = 0.f;
      lo = (int)v;
      if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else hi = lo + 1;
      l = v - lo; h = 1.f - l;
    } else {
      lo = -1;                                                   // marks "sample contributes 0"
    }
    s_lo[axis][i] = lo; s_hi[axis][i] = hi; s_l[axis][i] = l; s_h[axis][i] = h;
  }
  __syncthreads();
  const float *f = fmap + (size_t)b * Hf * Wf * C;
  const float inv = 1.f / (float)(sr * sr);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float addv = (edge_add != nullptr && r >= N) ? edge_add[(size_t)(r - N) * C + c] : 0.f;
    for (int ph = 0; ph < pool; ++ph)
      for (int pw = 0; pw < pool; ++pw) {
        float acc = 0.f;
        for (int iy = 0; iy < sr; ++iy) {
          const int sy = ph * sr + iy;
          const int yl = s_lo[0][sy], yh = s_hi[0][sy];
          const float ly = s_l[0][sy], hy = s_h[0][sy];
          for (int ix = 0; ix < sr; ++ix) {
            const int sx = pw * sr + ix;
            const int xl = s_lo[1][sx], xh = s_hi[1][sx];
            if (yl < 0 || xl < 0) continue;
            const float lx = s_l[1][sx], hx = s_h[1][sx];
            const float v1 = f[((size_t)yl * Wf + xl) * C + c], v2 = f[((size_t)yl * Wf + xh) * C + c];
            const float v3 = f[((size_t)yh * Wf + xl) * C + c], v4 = f[((size_t)yh * Wf + xh) * C + c];
            acc += hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
          }
        }
        s_out[c * pp + ph * pool + pw] = acc * inv + addv;
      }
  }
  __syncthreads();
  const int total = C * pp;
  if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const float4 *s4 = reinterpret_cast<const float4 *>(s_out);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int i = threadIdx.x; i < total / 4; i += blockDim.x) d4[i] = s4[i];
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = s_out[i];
  }
}

// Same CTA-per-RoI scheme with 16-byte channel-quad fetches: thread = (channel quad q, bin group g); a thread walks
// the bins g, g + G, ... and for each of the SR x SR samples issues four float4 corner loads (all independent, so
// up to 16 are in flight per thread).  4x fewer load instructions and 4x the bytes in flight of the scalar kernel;
// the per-channel arithmetic (and its order) is unchanged.  Needs C % 4 == 0 and C / 4 <= blockDim.x.
template <int SR>
__global__ void __launch_bounds__(256)
k_roi_align_nhwc4(const float *__restrict__ fmap /*[B,Hf,Wf,C]*/, int B, int C, int Hf, int Wf, const float *__restrict__ rois,
                  int N, const int64_t *__restrict__ ui, int64_t stride, int cs, int co, int E, float scale, int pool,
                  float *__restrict__ node_out, float *__restrict__ edge_out, int r_begin,
                  const float *__restrict__ edge_add, __half *__restrict__ node_planes, __half *__restrict__ edge_planes) {
  extern __shared__ __align__(16) float s_out[];          // [C][pool*pool]
  __shared__ int s_lo[2][RA_MAXS], s_hi[2][RA_MAXS];
  __shared__ float s_l[2][RA_MAXS], s_h[2][RA_MAXS];
  const int r = r_begin + blockIdx.x;
  const int pp = pool * pool, ns = pool * SR;
  float x1, y1, x2, y2; int b;
  float *dst;
  __half *pl_hi = nullptr, *pl_lo = nullptr;             // fp16 [hi | lo * 2^11] planes of the same row (nullable)
  if (r < N) {
    const float *q = rois + (size_t)r * 5;
    b = clamp_idx((long long)q[0], B); x1 = q[1]; y1 = q[2]; x2 = q[3]; y2 = q[4];
    dst = node_out != nullptr ? node_out + (size_t)r * C * pp : nullptr;       // fp32 rows are optional when planes are emitted
    if (node_planes != nullptr) { pl_hi = node_planes + (size_t)r * C * pp; pl_lo = pl_hi + (size_t)N * C * pp; }
  } else {
    const size_t e = (size_t)(r - N);
    const float *qs = rois + (size_t)clamp_idx(ui[e * stride + cs], N) * 5, *qo = rois + (size_t)clamp_idx(ui[e * stride + co], N) * 5;
    b = clamp_idx((long long)qs[0], B);
    x1 = fminf(qs[1], qo[1]); y1 = fminf(qs[2], qo[2]); x2 = fmaxf(qs[3], qo[3]); y2 = fmaxf(qs[4], qo[4]);
    dst = edge_out != nullptr ? edge_out + e * C * pp : nullptr;
    if (edge_planes != nullptr) { pl_hi = edge_planes + e * C * pp; pl_lo = pl_hi + (size_t)E * C * pp; }
  }
  if (threadIdx.x < 2 * ns) {
    const int axis = threadIdx.x / ns, i = threadIdx.x % ns;     // axis 0 = y, 1 = x
    const float start = (axis == 0 ? y1 : x1) * scale;
    const float len = fmaxf((axis == 0 ? y2 : x2) * scale - start, 1.f);
    const float bin = len / (float)pool;
    const int size = axis == 0 ? Hf : Wf;
    float v = start + (i / SR) * bin + ((i % SR) + 0.5f) * bin / (float)SR;
    int lo = 0, hi = 0; float l = 0.f, h = 0.f;
    if (!(v < -1.0f || v > (float)size)) {
      if (v <= 0.f) v = 0.f;
      lo = (int)v;
      if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else hi = lo + 1;
      l = v - lo; h = 1.f - l;
    }   // else: sample outside the map contributes 0 — index 0 with both weights 0, so the main loop has no branch
        // (all 16 corner loads of a bin are issued back to back)
    s_lo[axis][i] = lo; s_hi[axis][i] = hi; s_l[axis][i] = l; s_h[axis][i] = h;
  }
  __syncthreads();
  const int C4 = C >> 2;
  const int G = blockDim.x / C4;                           // bin groups working in parallel
  const int q = threadIdx.x % C4, g = threadIdx.x / C4;
  const float4 *f4 = reinterpret_cast<const float4 *>(fmap + (size_t)b * Hf * Wf * C) + q;
  const float inv = 1.f / (float)(SR * SR);
  if (g < G) {
    // union-box geometry embedding (lib/get_union_boxes.py:101: union_pools + conv(...), broadcast over the 7x7 bins)
    const float4 add = (edge_add != nullptr && r >= N)
                           ? __ldg(reinterpret_cast<const float4 *>(edge_add + (size_t)(r - N) * C) + q)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int bin = g; bin < pp; bin += G) {
      const int ph = bin / pool, pw = bin - ph * pool;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int iy = 0; iy < SR; ++iy) {
        const int sy = ph * SR + iy;
        const int yl = s_lo[0][sy], yh = s_hi[0][sy];
        const float ly = s_l[0][sy], hy = s_h[0][sy];
#pragma unroll
        for (int ix = 0; ix < SR; ++ix) {
          const int sx = pw * SR + ix;
          const int xl = s_lo[1][sx], xh = s_hi[1][sx];
          const float lx = s_l[1][sx], hx = s_h[1][sx];
          const float4 v1 = __ldg(f4 + ((size_t)yl * Wf + xl) * C4), v2 = __ldg(f4 + ((size_t)yl * Wf + xh) * C4);
          const float4 v3 = __ldg(f4 + ((size_t)yh * Wf + xl) * C4), v4 = __ldg(f4 + ((size_t)yh * Wf + xh) * C4);
          acc.x += hy * hx * v1.x + hy * lx * v2.x + ly * hx * v3.x + ly * lx * v4.x;
          acc.y += hy * hx * v1.y + hy * lx * v2.y + ly * hx * v3.y + ly * lx * v4.y;
          acc.z += hy * hx * v1.z + hy * lx * v2.z + ly * hx * v3.z + ly * lx * v4.z;
          acc.w += hy * hx * v1.w + hy * lx * v2.w + ly * hx * v3.w + ly * lx * v4.w;
        }
      }
      float *so = s_out + (size_t)(4 * q) * pp + bin;
      so[0] = acc.x * inv + add.x; so[pp] = acc.y * inv + add.y;
      so[2 * pp] = acc.z * inv + add.z; so[3 * pp] = acc.w * inv + add.w;
    }
  }
  __syncthreads();
  const int total = C * pp;
  if (dst == nullptr) {
    // planes only
  } else if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const float4 *s4 = reinterpret_cast<const float4 *>(s_out);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int i = threadIdx.x; i < total / 4; i += blockDim.x) __stcs(d4 + i, s4[i]);
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = s_out[i];
  }
  if (pl_hi != nullptr) {
    // the operand planes of the fc layer that consumes this row (lin16p.cu): x = hi + 2^-11 lo
    unsigned int ovf = 0;
    if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(pl_hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(pl_lo) & 7) == 0) {
      const float4 *s4 = reinterpret_cast<const float4 *>(s_out);
      for (int i = threadIdx.x; i < total / 4; i += blockDim.x) {
        const float4 v = s4[i];
        uint2 hi, lo;
        tc16::split2(v.x, v.y, hi.x, lo.x); tc16::split2(v.z, v.w, hi.y, lo.y);
        ovf |= tc16::f16x2_nonfinite(hi.x) | tc16::f16x2_nonfinite(hi.y);
        *reinterpret_cast<uint2 *>(pl_hi + 4 * (size_t)i) = hi;
        *reinterpret_cast<uint2 *>(pl_lo + 4 * (size_t)i) = lo;
      }
    } else {
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const float v = s_out[i];
        const __half h = __float2half_rn(v);
        pl_hi[i] = h;
        pl_lo[i] = __float2half_rn((v - __half2float(h)) * tc16::LO_SCALE);
        ovf |= (unsigned int)(__hisinf(h) || __hisnan(h));
      }
    }
    if (ovf) atomicOr(&g_roi_overflow, 1u);
  }
}

}  // namespace sgg

extern "C" size_t sgg_node_edge_features_workspace_bytes(int B, int C, int Hf, int Wf) {
  return (size_t)B * C * Hf * Wf * sizeof(float) + 256;
}

extern "C" int sgg_node_edge_features(const float *fmap, int B, int C, int Hf, int Wf, const float *rois, int N,
                                      const int64_t *union_inds, int64_t row_stride, int col_subj, int col_obj, int E,
                                      float spatial_scale, int pool, int sampling_ratio, float *node_feat,
                                      float *edge_feat, void *ws, size_t ws_bytes, void *stream) {
  return sgg_node_edge_features_add(fmap, B, C, Hf, Wf, rois, N, union_inds, row_stride, col_subj, col_obj, E,
                                    spatial_scale, pool, sampling_ratio, nullptr, node_feat, edge_feat, ws, ws_bytes,
                                    stream);
}

extern "C" int sgg_node_edge_features_add(const float *fmap, int B, int C, int Hf, int Wf, const float *rois, int N,
                                          const int64_t *union_inds, int64_t row_stride, int col_subj, int col_obj,
                                          int E, float spatial_scale, int pool, int sampling_ratio,
                                          const float *edge_add, float *node_feat, float *edge_feat, void *ws,
                                          size_t ws_bytes, void *stream) {
  return sgg_node_edge_features_planes(fmap, B, C, Hf, Wf, rois, N, union_inds, row_stride, col_subj, col_obj, E,
                                       spatial_scale, pool, sampling_ratio, edge_add, node_feat, edge_feat, nullptr,
                                       nullptr, ws, ws_bytes, stream);
}

/* Same, and the fp16 [hi | lo * 2^11] operand planes of the rows (node_planes: 2 * N * C * pool^2 halves, edge_planes:
 * 2 * E * C * pool^2 halves; each nullable) for sgg_tc16_linear_pre.  Planes need the channel-last fast path
 * (workspace given, sampling_ratio = 2, C % 4 == 0). */
extern "C" int sgg_node_edge_features_planes(const float *fmap, int B, int C, int Hf, int Wf, const float *rois, int N,
                                             const int64_t *union_inds, int64_t row_stride, int col_subj, int col_obj,
                                             int E, float spatial_scale, int pool, int sampling_ratio,
                                             const float *edge_add, float *node_feat, float *edge_feat,
                                             void *node_planes, void *edge_planes, void *ws, size_t ws_bytes,
                                             void *stream) {
  if (B <= 0 || C <= 0 || Hf <= 0 || Wf <= 0 || N < 0 || E < 0 || pool <= 0 || sampling_ratio <= 0)
    return sgg_set_err(SGG_E_BADARG, "node_edge_features: bad shape");
  // a side is produced when its fp32 rows or its planes are requested (planes alone: the fp32 rows are not written)
  const int do_node = (node_feat != nullptr || node_planes != nullptr) && N > 0;
  const int do_edge = (edge_feat != nullptr || edge_planes != nullptr) && E > 0;
  if (!do_node && !do_edge) return 0;
  if (!fmap || !rois || (do_edge && !union_inds)) return sgg_set_err(SGG_E_BADARG, "node_edge_features: null pointer");
  const size_t smem_need = (size_t)C * pool * pool * sizeof(float);
  if (ws && ws_bytes >= sgg_node_edge_features_workspace_bytes(B, C, Hf, Wf) && smem_need <= 200 * 1024 &&
      pool * sampling_ratio <= sgg::RA_MAXS) {
    // channel-last fast path
    cudaStream_t st = (cudaStream_t)stream;
    float *nhwc = (float *)ws;
    const int HW = Hf * Wf;
    dim3 tg((HW + 31) / 32, (C + 31) / 32, B);
    sgg::k_nchw_to_nhwc<<<tg, dim3(32, 8), 0, st>>>(fmap, C, HW, nhwc);
    SGG_RETURN_IF_LAUNCH_FAILED("k_nchw_to_nhwc");
    static bool attr = false;
    if (!attr) {
      SGG_CUDA_TRY(cudaFuncSetAttribute(sgg::k_roi_align_nhwc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = true;
    }
    const int r0 = do_node ? 0 : N, r1 = do_edge ? N + E : N;
    if (sampling_ratio == 2 && (C & 3) == 0 && C / 4 <= 256) {
      static bool attr4 = false;
      if (!attr4) {
        SGG_CUDA_TRY(cudaFuncSetAttribute(sgg::k_roi_align_nhwc4<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr4 = true;
      }
      sgg::k_roi_align_nhwc4<2><<<r1 - r0, 256, smem_need, st>>>(nhwc, B, C, Hf, Wf, rois, N, union_inds, row_stride, col_subj,
                                                              col_obj, E, spatial_scale, pool, node_feat, edge_feat, r0, edge_add,
                                                              (__half *)node_planes, (__half *)edge_planes);
      SGG_RETURN_IF_LAUNCH_FAILED("k_roi_align_nhwc4");
      return 0;
    }
    if (node_planes || edge_planes) return sgg_set_err(SGG_E_BADARG, "node_edge_features: planes need sampling_ratio 2 and C %% 4 == 0");
    sgg::k_roi_align_nhwc<<<r1 - r0, 256, smem_need, st>>>(nhwc, B, C, Hf, Wf, rois, N, union_inds, row_stride, col_subj,
                                                          col_obj, E, spatial_scale, pool, sampling_ratio, node_feat,
                                                          edge_feat, r0, edge_add);
    SGG_RETURN_IF_LAUNCH_FAILED("k_roi_align_nhwc");
    return 0;
  }
  if (node_planes || edge_planes) return sgg_set_err(SGG_E_BADARG, "node_edge_features: planes need the channel-last path (workspace)");
  const size_t total = ((size_t)(do_node ? N : 0) + (do_edge ? E : 0)) * C * pool * pool;
  int blocks = (int)((total + 255) / 256 < (size_t)sgg_num_sms() * 32 ? (total + 255) / 256 : (size_t)sgg_num_sms() * 32);
  sgg::k_roi_align<<<blocks, 256, 0, (cudaStream_t)stream>>>(fmap, B, C, Hf, Wf, rois, N, union_inds, row_stride,
                                                             col_subj, col_obj, E, spatial_scale, pool,
                                                             sampling_ratio, node_feat, edge_feat, do_node, do_edge, edge_add);
  SGG_RETURN_IF_LAUNCH_FAILED("k_roi_align");
  return 0;
}

namespace sgg {
int roi_overflow_flag(int reset, unsigned int *out) {
  unsigned int v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&v, g_roi_overflow, sizeof(v)) != cudaSuccess) return -1;
  if (reset) {
    const unsigned int z = 0;
    if (cudaMemcpyToSymbol(g_roi_overflow, &z, sizeof(z)) != cudaSuccess) return -1;
  }
  *out = v;
  return 0;
}
}  // namespace sgg
