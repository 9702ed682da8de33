// Training-mode pieces of the union-box geometry branch (UnionBoxesAndFeats.conv, lib/get_union_boxes.py:51-59):
//   Conv(2 -> C/2, 7x7, s16) -> ReLU -> BatchNorm2d(C/2, momentum 0.01) -> MaxPool2d(3, 2, 1) -> Conv(C/2 -> C, 3x3, s16)
//   -> ReLU -> BatchNorm2d(C)
// With the reference's stride-16 quirk the two convolutions are plain linear maps on [4E, 98] patches / [E, C/2] rows
// (geom.cu, linear kernels); what remains for training is BatchNorm with BATCH statistics (+ running-stat update,
// forward and backward, the preceding ReLU folded in) and the 2x2 -> 1 max-pool, both as deterministic CUDA kernels
// (fixed-order column reductions, no atomics).  Round 1 ran these through ATen.
#include "common.cuh"
#include "tc16_common.cuh"
#include "kernels.h"

namespace sgg {

constexpr int BN_MAX_PARTS = 128;

static int bn_parts(int rows) {
  int n = (rows + 63) / 64;
  if (n > BN_MAX_PARTS) n = BN_MAX_PARTS;
  if (n < 1) n = 1;
  return n;
}

// stage 1 of a column reduction over a row slice: out[part][c] = sum_r f(x[r][c])   (and optionally a second sum)
// MODE 0: v = act(x)                       -> p0 = sum v
// MODE 1: v = act(x) - mean[c]             -> p0 = sum v*v
// MODE 2: backward: p0 = sum dy, p1 = sum dy * xhat,  xhat = (act(x) - mean) * invstd
template <int MODE>
__global__ void __launch_bounds__(128) k_bn_partial(const float *__restrict__ x, const float *__restrict__ dy, int M, int C,
                                                    int relu_in, const float *__restrict__ mean,
                                                    const float *__restrict__ invstd, int rows_per, float *__restrict__ p0,
                                                    float *__restrict__ p1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  const float mu = MODE >= 1 ? mean[c] : 0.f, is = MODE == 2 ? invstd[c] : 0.f;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;              // two independent chains per sum
  int r = r0;
  for (; r + 1 < r1; r += 2) {
    float v0 = x[(size_t)r * C + c], v1 = x[(size_t)(r + 1) * C + c];
    if (relu_in) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
    if (MODE == 0) { a0 += v0; a1 += v1; }
    if (MODE == 1) { v0 -= mu; v1 -= mu; a0 += v0 * v0; a1 += v1 * v1; }
    if (MODE == 2) {
      const float d0 = dy[(size_t)r * C + c], d1 = dy[(size_t)(r + 1) * C + c];
      a0 += d0; a1 += d1;
      b0 += d0 * ((v0 - mu) * is); b1 += d1 * ((v1 - mu) * is);
    }
  }
  if (r < r1) {
    float v0 = x[(size_t)r * C + c];
    if (relu_in) v0 = fmaxf(v0, 0.f);
    if (MODE == 0) a0 += v0;
    if (MODE == 1) { v0 -= mu; a0 += v0 * v0; }
    if (MODE == 2) { const float d0 = dy[(size_t)r * C + c]; a0 += d0; b0 += d0 * ((v0 - mu) * is); }
  }
  p0[(size_t)blockIdx.y * C + c] = a0 + a1;
  if (MODE == 2) p1[(size_t)blockIdx.y * C + c] = b0 + b1;
}

// stage 2 (forward statistics): WHAT 0: mean; WHAT 1: biased variance -> invstd, running-stat update
// (torch BatchNorm: running_var takes the UNBIASED batch variance, running = (1 - momentum) * running + momentum * new)
template <int WHAT>
__global__ void k_bn_stat_finish(const float *__restrict__ part, int nparts, int M, int C, float eps, float momentum,
                                 float *__restrict__ mean, float *__restrict__ invstd, float *__restrict__ running_mean,
                                 float *__restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * C + c];
  if (WHAT == 0) {
    mean[c] = s / (float)M;
  } else {
    const float var = s / (float)M;
    invstd[c] = 1.0f / sqrtf(var + eps);
    if (running_mean != nullptr) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean[c];
    if (running_var != nullptr) {
      const float unbiased = M > 1 ? s / (float)(M - 1) : var;
      running_var[c] = (1.0f - momentum) * running_var[c] + momentum * unbiased;
    }
  }
}

__global__ void k_bn_sum_finish(const float *__restrict__ p0, const float *__restrict__ p1, int nparts, int C,
                                float *__restrict__ o0, float *__restrict__ o1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s0 = 0.f, s1 = 0.f;
  for (int p = 0; p < nparts; ++p) { s0 += p0[(size_t)p * C + c]; s1 += p1[(size_t)p * C + c]; }
  o0[c] = s0; o1[c] = s1;
}

// y = (act(x) - mean) * invstd * gamma + beta
__global__ void k_bn_apply(const float *__restrict__ x, size_t n, int C, int relu_in, const float *__restrict__ mean,
                           const float *__restrict__ invstd, const float *__restrict__ gamma,
                           const float *__restrict__ beta, float *__restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float v = x[i];
    if (relu_in) v = fmaxf(v, 0.f);
    y[i] = (v - mean[c]) * invstd[c] * gamma[c] + beta[c];
  }
}

// dx = gamma * invstd * (dy - dbeta / M - xhat * dgamma / M), times the ReLU mask of the input when relu_in
__global__ void k_bn_bwd_apply(const float *__restrict__ x, const float *__restrict__ dy, size_t n, int M, int C, int relu_in,
                               const float *__restrict__ mean, const float *__restrict__ invstd,
                               const float *__restrict__ gamma, const float *__restrict__ dbeta,
                               const float *__restrict__ dgamma, float *__restrict__ dx) {
  const float inv_m = 1.0f / (float)M;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float raw = x[i];
    const float v = relu_in ? fmaxf(raw, 0.f) : raw;
    const float xhat = (v - mean[c]) * invstd[c];
    float d = gamma[c] * invstd[c] * (dy[i] - dbeta[c] * inv_m - xhat * dgamma[c] * inv_m);
    if (relu_in && !(raw > 0.f)) d = 0.f;
    dx[i] = d;
  }
}

// MaxPool2d(3, 2, 1) over the 2x2 map the stride-16 7x7 conv leaves: max over the 4 positions of an edge, first maximum
// wins (the reference's max-pool backward routes the gradient to the arg-max)
__global__ void k_max4_fwd(const float *__restrict__ x, size_t n_out, int C, float *__restrict__ y, unsigned char *__restrict__ idx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i / C;
    const int c = (int)(i % C);
    const float *p = x + (e * 4) * C + c;
    float best = p[0];
    int bi = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = p[(size_t)k * C];
      if (v > best) { best = v; bi = k; }
    }
    y[i] = best;
    idx[i] = (unsigned char)bi;
  }
}
__global__ void k_max4_bwd(const float *__restrict__ dy, const unsigned char *__restrict__ idx, size_t n_out, int C,
                           float *__restrict__ dx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i / C;
    const int c = (int)(i % C);
    const float d = dy[i];
    const int bi = idx[i];
    float *p = dx + (e * 4) * C + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) p[(size_t)k * C] = (k == bi) ? d : 0.f;
  }
}

__device__ unsigned int g_bcast_overflow = 0;   // sticky: an emitted operand plane left the fp16 range (sgg_tc16_overflow, bit 0)
// out[e, c, s] = pools[e, c, s] + geom[e, c]   (lib/get_union_boxes.py:101, the broadcast add of the training path)
// pl_hi / pl_lo (nullable): also the fp16 [hi | lo * 2^11] operand planes of `out` for the pre-split fc6 GEMM (lin16p.cu)
__global__ void k_bcast_add(const float4 *__restrict__ pools, const float *__restrict__ geom, size_t n4, int S4,
                            float4 *__restrict__ out, uint2 *__restrict__ pl_hi, uint2 *__restrict__ pl_lo,
                            unsigned int *__restrict__ ovf_flag) {
  unsigned int ovf = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = pools[i];
    // element 4i..4i+3 of [E*C, S]: row = (4i + k) / S
    const size_t base = 4 * i;
    const int S = S4;
    const float g0 = geom[base / S], g1 = geom[(base + 1) / S], g2 = geom[(base + 2) / S], g3 = geom[(base + 3) / S];
    v.x += g0; v.y += g1; v.z += g2; v.w += g3;
    out[i] = v;
    if (pl_hi != nullptr) {
      uint2 hi, lo;
      tc16::split2(v.x, v.y, hi.x, lo.x); tc16::split2(v.z, v.w, hi.y, lo.y);
      ovf |= tc16::f16x2_nonfinite(hi.x) | tc16::f16x2_nonfinite(hi.y);
      pl_hi[i] = hi; pl_lo[i] = lo;
    }
  }
  if (ovf) atomicOr(ovf_flag, 1u);
}
// dx = dy where y > 0 else 0   (ReLU backward from the saved output)
__global__ void k_relu_bwd(const float4 *__restrict__ dy, const float4 *__restrict__ y, size_t n4, float4 *__restrict__ dx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 d = dy[i], v = y[i];
    dx[i] = make_float4(v.x > 0.f ? d.x : 0.f, v.y > 0.f ? d.y : 0.f, v.z > 0.f ? d.z : 0.f, v.w > 0.f ? d.w : 0.f);
  }
}
// out[g] = sum_{s < S} x[g * S + s]: one warp per 32 groups would be uncoalesced; here a warp walks one group at a time
__global__ void k_group_sum(const float *__restrict__ x, size_t groups, int S, float *__restrict__ out) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (size_t g = warp; g < groups; g += nwarps) {
    float s = 0.f;
    for (int k = lane; k < S; k += 32) s += x[g * S + k];
    s = sgg_warp_sum(s);
    if (lane == 0) out[g] = s;
  }
}

static int ew_grid(size_t n) {
  size_t b = (n + 255) / 256;
  const size_t cap = (size_t)sgg_num_sms() * 8;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace sgg

extern "C" size_t sgg_bn_workspace_bytes(int M, int C) {
  (void)M;
  return sgg_align_up((size_t)2 * sgg::BN_MAX_PARTS * (C > 0 ? C : 1) * sizeof(float)) + 256;
}

// y [M,C] = BatchNorm(act(x)) with batch statistics over the M rows; act = ReLU when relu_in.  save_mean / save_invstd
// [C] are kept for the backward pass; running_mean / running_var (nullable) receive the momentum update.
extern "C" int sgg_bn_train_forward(const float *x, int M, int C, int relu_in, const float *gamma, const float *beta,
                                    float *running_mean, float *running_var, float momentum, float eps, float *y,
                                    float *save_mean, float *save_invstd, void *ws, size_t ws_bytes, void *stream) {
  using namespace sgg;
  if (M <= 0 || C <= 0) return M == 0 ? 0 : sgg_set_err(SGG_E_BADARG, "bn_train_forward: bad shape");
  if (!x || !gamma || !beta || !y || !save_mean || !save_invstd || !ws) return sgg_set_err(SGG_E_BADARG, "bn_train_forward: null pointer");
  if (ws_bytes < sgg_bn_workspace_bytes(M, C)) return sgg_set_err(SGG_E_WORKSPACE, "bn_train_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int nparts = bn_parts(M);
  const int rows_per = (M + nparts - 1) / nparts;
  nparts = (M + rows_per - 1) / rows_per;
  float *part = (float *)ws;
  dim3 grid((C + 127) / 128, nparts);
  k_bn_partial<0><<<grid, 128, 0, st>>>(x, nullptr, M, C, relu_in, nullptr, nullptr, rows_per, part, nullptr);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bn_partial<0>");
  k_bn_stat_finish<0><<<(C + 127) / 128, 128, 0, st>>>(part, nparts, M, C, eps, momentum, save_mean, save_invstd, nullptr, nullptr);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bn_stat_finish<0>");
  k_bn_partial<1><<<grid, 128, 0, st>>>(x, nullptr, M, C, relu_in, save_mean, nullptr, rows_per, part, nullptr);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bn_partial<1>");
  k_bn_stat_finish<1><<<(C + 127) / 128, 128, 0, st>>>(part, nparts, M, C, eps, momentum, save_mean, save_invstd, running_mean,
                                                       running_var);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bn_stat_finish<1>");
  const size_t n = (size_t)M * C;
  k_bn_apply<<<ew_grid(n), 256, 0, st>>>(x, n, C, relu_in, save_mean, save_invstd, gamma, beta, y);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bn_apply");
  return 0;
}

// Backward of sgg_bn_train_forward: dx [M,C] (gradient w.r.t. the PRE-activation x when relu_in), dgamma / dbeta [C]
// (overwritten).
extern "C" int sgg_bn_train_backward(const float *x, const float *dy, int M, int C, int relu_in, const float *gamma,
                                     const float *save_mean, const float *save_invstd, float *dx, float *dgamma,
                                     float *dbeta, void *ws, size_t ws_bytes, void *stream) {
  using namespace sgg;
  if (M <= 0 || C <= 0) return M == 0 ? 0 : sgg_set_err(SGG_E_BADARG, "bn_train_backward: bad shape");
  if (!x || !dy || !gamma || !save_mean || !save_invstd || !dx || !dgamma || !dbeta || !ws)
    return sgg_set_err(SGG_E_BADARG, "bn_train_backward: null pointer");
  if (ws_bytes < sgg_bn_workspace_bytes(M, C)) return sgg_set_err(SGG_E_WORKSPACE, "bn_train_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int nparts = bn_parts(M);
  const int rows_per = (M + nparts - 1) / nparts;
  nparts = (M + rows_per - 1) / rows_per;
  float *p0 = (float *)ws, *p1 = p0 + (size_t)BN_MAX_PARTS * C;
  dim3 grid((C + 127) / 128, nparts);
  k_bn_partial<2><<<grid, 128, 0, st>>>(x, dy, M, C, relu_in, save_mean, save_invstd, rows_per, p0, p1);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bn_partial<2>");
  k_bn_sum_finish<<<(C + 127) / 128, 128, 0, st>>>(p0, p1, nparts, C, dbeta, dgamma);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bn_sum_finish");
  const size_t n = (size_t)M * C;
  k_bn_bwd_apply<<<ew_grid(n), 256, 0, st>>>(x, dy, n, M, C, relu_in, save_mean, save_invstd, gamma, dbeta, dgamma, dx);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bn_bwd_apply");
  return 0;
}

// x [E,4,C] -> y [E,C] = max over the 4 positions, idx [E,C] = arg-max (first maximum)
extern "C" int sgg_max4_forward(const float *x, int E, int C, float *y, unsigned char *idx, void *stream) {
  if (E <= 0 || C <= 0) return 0;
  if (!x || !y || !idx) return sgg_set_err(SGG_E_BADARG, "max4_forward: null pointer");
  const size_t n = (size_t)E * C;
  sgg::k_max4_fwd<<<sgg::ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, n, C, y, idx);
  SGG_RETURN_IF_LAUNCH_FAILED("k_max4_fwd");
  return 0;
}
extern "C" int sgg_max4_backward(const float *dy, const unsigned char *idx, int E, int C, float *dx, void *stream) {
  if (E <= 0 || C <= 0) return 0;
  if (!dy || !idx || !dx) return sgg_set_err(SGG_E_BADARG, "max4_backward: null pointer");
  const size_t n = (size_t)E * C;
  sgg::k_max4_bwd<<<sgg::ew_grid(n), 256, 0, (cudaStream_t)stream>>>(dy, idx, n, C, dx);
  SGG_RETURN_IF_LAUNCH_FAILED("k_max4_bwd");
  return 0;
}

// out [E,C,S] = pools [E,C,S] + geom [E,C] broadcast over the S = 7 x 7 positions (E*C*S % 4 == 0)
extern "C" int sgg_bcast_add(const float *pools, const float *geom, long long rows, int S, float *out, void *stream) {
  if (rows <= 0 || S <= 0) return 0;
  const size_t n = (size_t)rows * S;
  if (!pools || !geom || !out || (n & 3)) return sgg_set_err(SGG_E_BADARG, "bcast_add: bad argument");
  sgg::k_bcast_add<<<sgg::ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((const float4 *)pools, geom, n / 4, S, (float4 *)out,
                                                                          nullptr, nullptr, nullptr);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bcast_add");
  return 0;
}
/* Same, and the fp16 [hi | lo * 2^11] operand planes of out (2 * rows * S halves) for sgg_tc16_linear_pre: the train-mode
 * fc6 forward of rel_model_stanford.py:100-101 then runs on pre-split operands while out (fp32) stays for the backward. */
extern "C" int sgg_bcast_add_planes(const float *pools, const float *geom, long long rows, int S, float *out, void *out_planes,
                                    void *stream) {
  if (rows <= 0 || S <= 0) return 0;
  const size_t n = (size_t)rows * S;
  if (!pools || !geom || !out || !out_planes || (n & 3)) return sgg_set_err(SGG_E_BADARG, "bcast_add_planes: bad argument");
  unsigned int *flag = nullptr;
  SGG_CUDA_TRY(cudaGetSymbolAddress((void **)&flag, sgg::g_bcast_overflow));
  uint2 *hi = (uint2 *)out_planes, *lo = (uint2 *)((__half *)out_planes + n);
  sgg::k_bcast_add<<<sgg::ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>((const float4 *)pools, geom, n / 4, S, (float4 *)out, hi, lo,
                                                                          flag);
  SGG_RETURN_IF_LAUNCH_FAILED("k_bcast_add");
  return 0;
}
namespace sgg {
int bcast_overflow_flag(int reset, unsigned int *out) {
  unsigned int v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&v, g_bcast_overflow, sizeof(v)) != cudaSuccess) return -1;
  if (reset) {
    const unsigned int z = 0;
    if (cudaMemcpyToSymbol(g_bcast_overflow, &z, sizeof(z)) != cudaSuccess) return -1;
  }
  *out = v;
  return 0;
}
}  // namespace sgg
extern "C" int sgg_relu_backward(const float *dy, const float *y, long long n, float *dx, void *stream) {
  if (n <= 0) return 0;
  if (!dy || !y || !dx || (n & 3)) return sgg_set_err(SGG_E_BADARG, "relu_backward: bad argument");
  sgg::k_relu_bwd<<<sgg::ew_grid((size_t)n / 4), 256, 0, (cudaStream_t)stream>>>((const float4 *)dy, (const float4 *)y, (size_t)n / 4,
                                                                               (float4 *)dx);
  SGG_RETURN_IF_LAUNCH_FAILED("k_relu_bwd");
  return 0;
}
// out [groups] = sum over S consecutive elements of x [groups, S]
extern "C" int sgg_group_sum(const float *x, long long groups, int S, float *out, void *stream) {
  if (groups <= 0 || S <= 0) return 0;
  if (!x || !out) return sgg_set_err(SGG_E_BADARG, "group_sum: null pointer");
  const size_t threads = (size_t)groups * 32;
  sgg::k_group_sum<<<sgg::ew_grid(threads), 256, 0, (cudaStream_t)stream>>>(x, (size_t)groups, S, out);
  SGG_RETURN_IF_LAUNCH_FAILED("k_group_sum");
  return 0;
}
