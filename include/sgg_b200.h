/* sgg_b200 — C-ABI of the B200-native IMP relation-model hot path.
 *
 * Drop-in boundary for the message-passing relation model of bknyaz/sgg
 * (sgg_models/rel_model_stanford.py, lib/get_union_boxes.py,
 * lib/draw_rectangles/draw_rectangles.pyx, sgg_models/rel_model_base.py).
 * The reference is pure Python/PyTorch (+ one Cython op); the binding a
 * maintainer adds is a ctypes stub or the mirror classes in
 * sgg_b200/rel_model_stanford.py.
 *
 * Conventions (all entry points):
 *  - plain C: pointers + sizes only, no torch / C++ types;
 *  - every pointer is a DEVICE pointer to row-major contiguous memory owned by
 *    the caller (torch), unless documented as "host";
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *    synchronises, nothing allocates — the caller passes workspaces sized by
 *    the matching *_workspace_bytes function;
 *  - return value: 0 = ok, otherwise a cudaError_t (or SGG_E_* below); the
 *    text is available from sgg_last_error(); nothing throws;
 *  - fp32 data, int64 indices at the boundary (what the reference holds),
 *    int32 internally.
 */
#ifndef SGG_B200_H
#define SGG_B200_H
#include <stddef.h>
#include <stdint.h>

/* Process model: ONE process per GPU (as torchrun launches them).  The library keeps a few process-global caches that are
 * not keyed by device (side streams / events of the fork-join helper, the SM count, per-kernel shared-memory opt-ins):
 * driving two devices from one process, or calling in from several host threads at once, is not supported. */
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SGG_API __attribute__((visibility("default")))
#else
#define SGG_API
#endif

#define SGG_ABI_VERSION 2
#define SGG_E_BADARG 10001   /* shape / alignment / null pointer          */
#define SGG_E_WORKSPACE 10002 /* workspace too small                      */
#define SGG_E_INDEX 10003    /* rel_inds out of [0, N) (reported by sgg_graph_check) */

SGG_API int sgg_abi_version(void);
SGG_API const char *sgg_last_error(void);
/* sm count etc. of the current device: {sm_count, cc_major, cc_minor, l2_bytes} */
SGG_API int sgg_device_info(int out[4]);
/* number of kernels this library has launched in this process (bench.py gpu_launches) */
SGG_API unsigned long long sgg_launch_count(void);

/* ---- message-passing weights: rel_model_stanford.py:36-45 ------------------
 * edge_gru / node_gru: nn.GRUCell(H, H): weight_ih, weight_hh [3H,H] (r,z,n),
 * bias_ih, bias_hh [3H].  gate_w[k]: Linear(2H,1).weight [2H] = [vertex half |
 * edge half]; gate_b[k]: [1].  k = 0 sub_vert_w_fc, 1 obj_vert_w_fc,
 * 2 out_edge_w_fc, 3 in_edge_w_fc. */
typedef struct {
  const float *edge_w_ih, *edge_w_hh, *edge_b_ih, *edge_b_hh;
  const float *node_w_ih, *node_w_hh, *node_b_ih, *node_b_hh;
  const float *gate_w[4];
  const float *gate_b[4];
  /* optional tensor-core operands: [hi | lo] splits of the four GRU matrices made by
   * sgg_tc_split_weights (2 * 3H*H floats each).  Non-NULL => that GEMM runs on tcgen05 (3xTF32),
   * NULL => fp32 SIMT tiles. */
  const float *edge_w_ih_split, *edge_w_hh_split, *node_w_ih_split, *node_w_hh_split;
} sgg_mp_weights;

/* ---- ragged graph index (replaces the dense [N,E] incidence matrices of
 * rel_model_stanford.py:58-66) -----------------------------------------------
 * rel_inds: int64 rows with stride `row_stride` elements; columns col_subj /
 * col_obj hold GLOBAL object ids (rel_model_stanford.py:76-77).  Writes into
 * graph_ws: subj32[E], obj32[E], CSR by subject (out_ptr[N+1], out_idx[E]) and
 * by object (in_ptr[N+1], in_idx[E]); per-node lists are sorted by edge id so
 * every reduction order is deterministic.  Edges need not be sorted; duplicate
 * edges are kept (they add, as in the reference). */
SGG_API size_t sgg_graph_workspace_bytes(int N, int E);
SGG_API int sgg_graph_build(const int64_t *rel_inds, int64_t row_stride, int col_subj, int col_obj,
                    int N, int E, void *graph_ws, size_t graph_ws_bytes, void *stream);
/* host-synchronising validation helper (debug / tests): returns SGG_E_INDEX if any id
 * was out of range during the last sgg_graph_build on this workspace. */
SGG_API int sgg_graph_check(const void *graph_ws, int N, int E, void *stream);

/* ---- a1-a3: RelModelStanford.message_pass (rel_model_stanford.py:48-94) -----
 * obj_rep [N,H], rel_rep [E,H] -> V_out [N,H], E_out [E,H] after T iterations.
 * saved (nullable): training tape of sgg_mp_tape_bytes(N,E,H,T) bytes; it starts with the states
 * V_0,E_0 ... V_T,E_T ((T+1) * (N+E) * H floats) followed by the GRU gate caches, edge gates, vertex
 * contexts and P matrices sgg_mp_backward needs.  H must be a multiple of 64. */
SGG_API size_t sgg_mp_tape_bytes(int N, int E, int H, int T);
SGG_API size_t sgg_mp_workspace_bytes(int N, int E, int H, int T);
SGG_API int sgg_mp_forward(const float *obj_rep, const float *rel_rep, const void *graph_ws,
                   const sgg_mp_weights *w, int N, int E, int H, int T,
                   float *V_out, float *E_out, float *saved,
                   void *ws, size_t ws_bytes, void *stream);

/* Measurement probe: ONE launch of the fused 3xFP16 message-passing schedule (iteration 0) on a workspace that a
 * preceding sgg_mp_forward call with the same arguments initialised.  which: 0 = INIT launch, 1 = launch A (P/Q GEMM
 * tiles + vertex-context gather), 2 = launch B (edge + node GRU tiles).  Used by bench.py for the per-kernel roofline. */
SGG_API int sgg_mp_probe_launch(int which, const float *obj_rep, const float *rel_rep, const void *graph_ws,
                        const sgg_mp_weights *w, int N, int E, int H, int T, float *V_out, float *E_out,
                        void *ws, size_t ws_bytes, void *stream);

/* One edge-GRU update in isolation (rel_model_stanford.py:83 with the gathered, gate-scaled input of :76-81
 * expressed through P = V W_ih^T [N,3H] and gates [E,4] = (g_sub, g_obj, g_out, g_in)):
 * out[E,H] = GRUCell(g_sub P[s] + g_obj P[o] + b_ih, Eh W_hh^T + b_hh, Eh).  w_hh_split (nullable) selects tcgen05. */
SGG_API int sgg_edge_gru_forward(const float *Eh, const float *P, const float *gates, const void *graph_ws,
                         const float *w_ih, const float *w_hh, const float *w_hh_split,
                         const float *b_ih, const float *b_hh, int N, int E, int H, float *out, void *stream);

/* Backward of sgg_mp_forward (BPTT over the tape).  `grads` has the layout of the first 16 pointers of
 * sgg_mp_weights; every non-NULL buffer is ACCUMULATED into (+=).  d_obj_rep [N,H] / d_rel_rep [E,H]
 * (grads of the forward inputs) are overwritten, nullable.  Deterministic (no float atomics). */
typedef struct {
  float *edge_w_ih, *edge_w_hh, *edge_b_ih, *edge_b_hh;
  float *node_w_ih, *node_w_hh, *node_b_ih, *node_b_hh;
  float *gate_w[4];
  float *gate_b[4];
} sgg_mp_grads;
SGG_API size_t sgg_mp_backward_workspace_bytes(int N, int E, int H, int T);
SGG_API int sgg_mp_backward(const float *obj_rep, const float *rel_rep, const void *graph_ws,
                    const sgg_mp_weights *w, const float *tape, int N, int E, int H, int T,
                    const float *dV_T, const float *dE_T, const sgg_mp_grads *grads,
                    float *d_obj_rep, float *d_rel_rep, void *ws, size_t ws_bytes, void *stream);

/* ---- a5/a6: nn.Linear (+ReLU): y[M,Nout] = act(x[M,K] @ w[Nout,K]^T + b) ----
 * (obj_unary / edge_unary / obj_fc / rel_fc rel_model_stanford.py:29-33,
 *  roi_fmap / roi_fmap_obj rel_model_base.py:110-111).  b nullable.  Any K (unaligned
 *  operands fall back to scalar loads). */
SGG_API int sgg_linear_forward(const float *x, const float *w, const float *b, float *y,
                       int M, int Nout, int K, int relu, void *stream);

/* nn.Linear backward.  dy [M,Nout] must already be masked by the ReLU derivative when the forward had
 * relu=1.  dx [M,K] is overwritten (nullable); dw [Nout,K] and db [Nout] are ACCUMULATED into (nullable). */
SGG_API size_t sgg_linear_backward_workspace_bytes(int M, int Nout, int K);
SGG_API int sgg_linear_backward(const float *x, const float *w, const float *dy, int M, int Nout, int K,
                        float *dx, float *dw, float *db, void *ws, size_t ws_bytes, void *stream);

/* same; accumulate = 0 overwrites dw / db instead of adding to them (no zero-fill of fresh gradient buffers needed) */
SGG_API int sgg_linear_backward_ex(const float *x, const float *w, const float *dy, int M, int Nout, int K,
                           float *dx, float *dw, float *db, int accumulate, void *ws, size_t ws_bytes, void *stream);
/* Tensor-core building blocks of the nn.Linear backward (3xTF32 engine, fp32 exponent range):
 * sgg_bwd_transpose: in [R,C] (row stride ldin) -> out [C,Rpad], rows R..Rpad-1 zero; split != 0 writes the engine's
 * [hi | lo] operand planes (2 * C * Rpad floats).  sgg_tc32_linear_forward = sgg_tc_linear_forward pinned to the 3xTF32
 * engine (w_split from sgg_bwd_transpose(split = 1) or from sgg_tc_split_weights in mode 0); K % 4 == 0.
 *   dX = dY W    : y = tc32_linear(x = dY [M,Nout],     w_split = split(W^T) [K,Nout])
 *   dW = dY^T X  : y = tc32_linear(x = dY^T [Nout,Mp],  w_split = split(X^T) [K,Mp]),  Mp = M rounded up to 32 */
SGG_API int sgg_bwd_transpose(const float *in, long long ldin, int R, int C, float *out, int Rpad, int split, void *stream);
SGG_API size_t sgg_tc32_linear_workspace_bytes(int M, int Nout, int K);
SGG_API int sgg_tc32_linear_forward(const float *x, const float *w_split, const float *b, float *y,
                            int M, int Nout, int K, int relu, void *ws, size_t ws_bytes, void *stream);

/* Scaled 3xFP16 variants (twice the MMA rate of 3xTF32): gradients are far below fp16's normal range, so the gradient
 * operand is multiplied by an exact power of two taken from its absolute maximum (sgg_pow2_scale: sc[0] = 2^k with
 * max|x| 2^k in [1024, 2048), sc[1] = 2^-k, device floats, no synchronisation) and the GEMM epilogue multiplies by sc[1].
 * sgg_bwd_transpose16: mode 0 = fp32 transposed and scaled by sc[0]; mode 1 = fp16 [hi | lo * 2^11] operand planes.
 * sgg_scale_by: y = sc[0] * x.  sgg_tc16_linear_scaled: y = out_scale[0] * (x w^T), w as fp16 planes; K % 8 == 0. */
SGG_API size_t sgg_pow2_scale_workspace_bytes(void);
SGG_API int sgg_pow2_scale(const float *x, long long n, float *sc, void *ws, size_t ws_bytes, void *stream);
SGG_API int sgg_bwd_transpose16(const float *in, long long ldin, int R, int C, void *out, int Rpad, int mode,
                        const float *sc, void *stream);
SGG_API int sgg_scale_by(const float *x, long long n, const float *sc, float *y, void *stream);
SGG_API size_t sgg_tc16_linear_workspace_bytes(int M, int Nout, int K);
SGG_API int sgg_tc16_linear_scaled(const float *x, const void *w_split16, float *y, int M, int Nout, int K,
                           const float *out_scale, void *ws, size_t ws_bytes, void *stream);
/* nn.Linear forward on tcgen05 with BOTH operands pre-split (lin16p.cu): x_planes = fp16 [hi | lo * 2^11] planes of
 * x [M,K] (2 * M * K halves, written by sgg_node_edge_features_planes or as y_planes of a previous call), w_split16 =
 * planes of w [Nout,K] (sgg_tc_split_weights, 3xFP16 layout).  y [M,Nout] fp32; y_planes nullable (2 * M * Nout halves).
 * K % 8 == 0, Nout % 4 == 0.  Replaces F.linear of rel_model_stanford.py:100-101 on RoIAlign rows.
 * out_scale (nullable device scalar) multiplies x w^T before bias / ReLU: the 1/s of a scaled backward GEMM. */
SGG_API int sgg_tc16_linear_pre(const void *x_planes, const void *w_split16, const float *bias, float *y, void *y_planes,
                                int M, int Nout, int K, int relu, const float *out_scale, void *stream);

/* ---- tensor-core (tcgen05 / TMEM / TMA) variants -----------------------------------------
 * fp32 in, fp32 out, fp32-grade accuracy through a 3-pass operand split.  Two engines:
 *   mode 0 "3xTF32": x = hi + lo (fp32 words, hi = top 19 bits), kind::tf32;     split buffer = 2n floats
 *   mode 1 "3xFP16": x = hi + 2^-11 lo (fp16 halves), kind::f16, |x| < 65504;    split buffer = 2n halves
 * sgg_tc_set_mode selects the engine for every later call of this process (default SGG_TC_DEFAULT_MODE, or the
 * SGG_TC_MODE environment variable); a split made in one mode must not be used in the other.
 * sgg_tc_split_weights: split = [hi(w) | lo(w)]; always pass a buffer of 2n floats; call once per weight version.
 * sgg_tc_linear_forward: same contract as sgg_linear_forward but takes the split weight.  K % 4 == 0 (mode 0),
 * K % 8 == 0 (mode 1). */
#define SGG_TC_DEFAULT_MODE 1
SGG_API int sgg_tc_set_mode(int mode);
SGG_API int sgg_tc_get_mode(void);
/* debug aid: 8 clock64 phase timestamps per CTA of the last 3xFP16 kernel (needs SGG_TC_TIMING=1); host_out holds
 * 8 * n_ctas values; synchronises the device. */
SGG_API int sgg_tc_debug_timing(long long *host_out, int n_ctas);
/* same for the last fused message-passing launch (csrc/mp_fused.cu): which = 0 k_mp_gru (launch B / INIT), 1 k_mp_pre (launch A) */
SGG_API int sgg_mpf_debug_timing(long long *host_out, int n_ctas, int which);
/* fp16 range guard of mode 1: the split / conversion / plane-emitting kernels raise a sticky device flag when an
 * operand is outside the fp16 range (|x| >= 65520) or non-finite.  Returns the flag (0 = clean; bit 0 activation operand,
 * bit 1 emitted plane, bit 2 weight), < 0 on a CUDA error; reset != 0 clears it.  Synchronises the device — call it
 * where the host waits anyway (the eval tail does) and fall back to mode 0 (3xTF32, fp32 range) when it fires. */
SGG_API int sgg_tc16_overflow(int reset);
SGG_API int sgg_tc_split_weights(const float *w, size_t n, float *split, void *stream);
SGG_API size_t sgg_tc_linear_workspace_bytes(int M, int Nout, int K);   /* split-K partials (0 = none needed) */
SGG_API int sgg_tc_linear_forward(const float *x, const float *w_split, const float *b, float *y,
                          int M, int Nout, int K, int relu, void *ws, size_t ws_bytes, void *stream);

/* ---- L1: 4096-d features -> obj_dists / rel_dists (rel_model_stanford.py:103-107
 * without roi_fmap*): obj_unary, relu(edge_unary), message_pass, obj_fc, rel_fc. */
typedef struct {
  const float *obj_unary_w, *obj_unary_b;   /* [H,D],[H] */
  const float *edge_unary_w, *edge_unary_b; /* [H,D],[H] */
  const float *obj_fc_w, *obj_fc_b;         /* [n_cls,H],[n_cls] */
  const float *rel_fc_w, *rel_fc_b;         /* [n_rel,H],[n_rel] */
  /* optional [hi | lo] splits (sgg_tc_split_weights) => tcgen05 path for that projection */
  const float *obj_unary_w_split, *edge_unary_w_split, *obj_fc_w_split, *rel_fc_w_split;
} sgg_head_weights;
SGG_API size_t sgg_l1_workspace_bytes(int N, int E, int H, int T);
SGG_API int sgg_l1_forward(const float *obj_feat, const float *edge_feat, const void *graph_ws,
                   const sgg_head_weights *hw, const sgg_mp_weights *w,
                   int N, int E, int D, int H, int T, int n_cls, int n_rel,
                   float *obj_dists, float *rel_dists,
                   void *ws, size_t ws_bytes, void *stream);
/* Same, starting from the int64 rel_inds the reference passes (rel_model_stanford.py:105): the graph index is
 * built INSIDE the call, on the object-branch side stream, concurrently with the edge-unary GEMM, into graph_ws
 * (sgg_graph_workspace_bytes(N, E)); one call = one step of the path. */
SGG_API int sgg_l1_forward_rel(const float *obj_feat, const float *edge_feat,
                   const int64_t *rel_inds, int64_t row_stride, int col_subj, int col_obj,
                   void *graph_ws, size_t graph_ws_bytes,
                   const sgg_head_weights *hw, const sgg_mp_weights *w,
                   int N, int E, int D, int H, int T, int n_cls, int n_rel,
                   float *obj_dists, float *rel_dists,
                   void *ws, size_t ws_bytes, void *stream);

/* ---- a8: draw_union_boxes (lib/draw_rectangles/draw_rectangles.pyx:12-67) ----
 * rois [N,5] (img,x1,y1,x2,y2) as in rel_model_base.py:147; union_inds int64
 * rows (stride row_stride) cols (col_subj,col_obj); out [E,2,P,P] f32.
 * `sub_half` != 0 subtracts 0.5 (lib/get_union_boxes.py:67). */
SGG_API int sgg_draw_union_boxes(const float *rois, const int64_t *union_inds, int64_t row_stride,
                         int col_subj, int col_obj, int E, int P, int sub_half,
                         float *out, void *stream);

/* ---- a7: UnionBoxesAndFeats geometry branch (lib/get_union_boxes.py:51-59,101)
 * geom[E,C] = BN2(ReLU(conv3s16(maxpool(BN1(ReLU(conv7s16(rects))))))) computed
 * straight from the boxes (the 27x27 masks never reach HBM); eval-mode BN
 * (running statistics).  C = 512 output channels, C/2 hidden.
 * If union_pools != NULL: out[E,C,7,7] = union_pools + geom (broadcast add, :101);
 * else out = geom [E,C]. */
typedef struct {
  const float *conv1_w, *conv1_b;              /* [C/2,2,7,7],[C/2] */
  const float *bn1_w, *bn1_b, *bn1_rm, *bn1_rv; /* [C/2] */
  const float *conv2_w, *conv2_b;              /* [C,C/2,3,3],[C] */
  const float *bn2_w, *bn2_b, *bn2_rm, *bn2_rv; /* [C] */
} sgg_geom_weights;
SGG_API size_t sgg_union_geom_workspace_bytes(int E, int C);
SGG_API int sgg_union_geom_forward(const float *rois, const int64_t *union_inds, int64_t row_stride,
                           int col_subj, int col_obj, int E, int C,
                           const sgg_geom_weights *gw, const float *union_pools,
                           float *out, void *ws, size_t ws_bytes, void *stream);

/* Training path of a7: the four live 7x7x2 conv1 windows of (draw_union_boxes - 0.5) as rows
 * patches[E,4,98] (tap = ch*49 + ky*7 + kx; zero over the padding), so conv1 becomes a linear op. */
SGG_API int sgg_geom_patches(const float *rois, const int64_t *union_inds, int64_t row_stride,
                     int col_subj, int col_obj, int E, float *out, void *stream);

/* ---- a7, training mode: BatchNorm with BATCH statistics (+ running-stat update) and the 2x2 max-pool of the
 * geometry branch (lib/get_union_boxes.py:51-59: Conv-ReLU-BN-MaxPool-Conv-ReLU-BN; the convolutions are linear maps
 * here, see sgg_geom_patches).  relu_in != 0 folds the preceding ReLU: y = BN(relu(x)), and the backward returns the
 * gradient w.r.t. the pre-activation x.  running_mean / running_var (nullable): running = (1 - momentum) * running +
 * momentum * batch (unbiased variance), as torch.nn.BatchNorm2d.  Deterministic (fixed-order column reductions). */
SGG_API size_t sgg_bn_workspace_bytes(int M, int C);
SGG_API int sgg_bn_train_forward(const float *x, int M, int C, int relu_in, const float *gamma, const float *beta,
                         float *running_mean, float *running_var, float momentum, float eps, float *y,
                         float *save_mean, float *save_invstd, void *ws, size_t ws_bytes, void *stream);
SGG_API int sgg_bn_train_backward(const float *x, const float *dy, int M, int C, int relu_in, const float *gamma,
                          const float *save_mean, const float *save_invstd, float *dx, float *dgamma, float *dbeta,
                          void *ws, size_t ws_bytes, void *stream);
/* x [E,4,C] -> y [E,C] = max over the 4 conv positions of an edge (MaxPool2d(3,2,1) on the 2x2 map), idx = arg-max
 * (first maximum, where the reference's max-pool backward sends the gradient); backward scatters dy to dx [E,4,C]. */
SGG_API int sgg_max4_forward(const float *x, int E, int C, float *y, unsigned char *idx, void *stream);
SGG_API int sgg_max4_backward(const float *dy, const unsigned char *idx, int E, int C, float *dx, void *stream);
/* Elementwise / small-reduction helpers of the training path (no ATen on it):
 * sgg_bcast_add: out[r,s] = pools[r,s] + geom[r], r < rows = E*C, s < S = 49 (lib/get_union_boxes.py:101);
 * sgg_relu_backward: dx = y > 0 ? dy : 0;  sgg_group_sum: out[g] = sum_s x[g,s] (e.g. the 7x7 taps of fc6's weight). */
SGG_API int sgg_bcast_add(const float *pools, const float *geom, long long rows, int S, float *out, void *stream);
/* sgg_bcast_add + the fp16 [hi | lo * 2^11] operand planes of out (2 * rows * S halves) for sgg_tc16_linear_pre */
SGG_API int sgg_bcast_add_planes(const float *pools, const float *geom, long long rows, int S, float *out, void *out_planes,
                                 void *stream);
SGG_API int sgg_relu_backward(const float *dy, const float *y, long long n, float *dx, void *stream);
SGG_API int sgg_group_sum(const float *x, long long groups, int S, float *out, void *stream);

/* ---- a10 / f2: the frozen VGG16 conv stack (rel_model_base.py:184, :310-321) as tcgen05 implicit GEMMs (3xFP16) ----
 * Activations between layers are NHWC fp16 planes [hi | lo * 2^11] (2 * B*H*W*C halves: hi plane then lo plane).
 * sgg_conv_weight_planes: w [Cout,Cin,3,3] fp32 -> planes in implicit-GEMM order [Cout][tap][Cin] (2 * Cout*9*Cin halves);
 * sgg_conv3x3_first: img [B,3,H,W] fp32 NCHW -> relu(conv(img)) planes, Cout = 64 (SIMT, K = 27);
 * sgg_conv3x3_tc: 3x3 / pad 1 / stride 1 conv + bias (+ReLU) (+fused 2x2 max-pool, H and W even) -> planes of
 *   [B,Ho,Wo,Cout], or fp32 NCHW [B,Cout,Ho,Wo] when out_f32_nchw != NULL (the fmap layout of the reference).
 *   Cin % 64 == 0, Cout % 64 == 0.  sgg_conv_overflow: sticky fp16 range flag of the emitted activations. */
SGG_API int sgg_conv_weight_planes(const float *w, int Cout, int Cin, void *planes, void *stream);
SGG_API int sgg_conv3x3_first(const float *img, const float *w, const float *bias, int B, int H, int W, int Cout,
                      void *out_planes, void *stream);
SGG_API int sgg_conv3x3_tc(const void *in_planes, const void *w_planes, const float *bias, int B, int H, int W, int Cin,
                   int Cout, int relu, int pool, void *out_planes, float *out_f32_nchw, void *stream);
SGG_API int sgg_conv_overflow(int reset);

/* ---- a9: node_edge_features (rel_model_base.py:245-260): torchvision
 * roi_align(aligned=False, sampling_ratio=2, 7x7, scale 1/16) for objects and
 * for union boxes computed on the fly from (rois, union_inds).
 * fmap [B,C,Hf,Wf]; node_feat [N,C,7,7]; edge_feat [E,C,7,7] (either may be NULL).
 * ws (nullable): sgg_node_edge_features_workspace_bytes => channel-last fast path (one CTA per RoI, coalesced
 * corner fetches, contiguous output block); without it a per-element kernel on the NCHW map is used. */
SGG_API size_t sgg_node_edge_features_workspace_bytes(int B, int C, int Hf, int Wf);
SGG_API int sgg_node_edge_features(const float *fmap, int B, int C, int Hf, int Wf,
                           const float *rois, int N,
                           const int64_t *union_inds, int64_t row_stride, int col_subj, int col_obj, int E,
                           float spatial_scale, int pool, int sampling_ratio,
                           float *node_feat, float *edge_feat, void *ws, size_t ws_bytes, void *stream);
/* Same, with the union-box geometry embedding folded into the edge rows: edge_feat[e,c,:,:] = RoIAlign(...) +
 * edge_add[e,c] (lib/get_union_boxes.py:101, union_pools + conv(rects) broadcast over the bins), so the [E,C,7,7]
 * tensor is written once instead of written, re-read and re-written.  edge_add [E,C] nullable. */
SGG_API int sgg_node_edge_features_add(const float *fmap, int B, int C, int Hf, int Wf,
                               const float *rois, int N,
                               const int64_t *union_inds, int64_t row_stride, int col_subj, int col_obj, int E,
                               float spatial_scale, int pool, int sampling_ratio, const float *edge_add,
                               float *node_feat, float *edge_feat, void *ws, size_t ws_bytes, void *stream);
/* Same, and the fp16 [hi | lo * 2^11] operand planes of the rows for sgg_tc16_linear_pre (the fc6 layers of
 * rel_model_stanford.py:100-101 / rel_model_base.py:166-170 then skip their in-kernel fp32 -> fp16 conversion):
 * node_planes 2 * N * C * pool^2 halves, edge_planes 2 * E * C * pool^2 halves, each nullable; with planes given the
 * fp32 rows of that side (node_feat / edge_feat) may be NULL and are then not written.  Planes need the
 * channel-last fast path: workspace given, sampling_ratio = 2, C % 4 == 0 (SGG_E_BADARG otherwise). */
SGG_API int sgg_node_edge_features_planes(const float *fmap, int B, int C, int Hf, int Wf,
                               const float *rois, int N,
                               const int64_t *union_inds, int64_t row_stride, int col_subj, int col_obj, int E,
                               float spatial_scale, int pool, int sampling_ratio, const float *edge_add,
                               float *node_feat, float *edge_feat, void *node_planes, void *edge_planes,
                               void *ws, size_t ws_bytes, void *stream);

/* ==== evaluation tail: lib/surgery.py:17-55 filter_dets ========================================
 * score[e] = max_{p>=1} prob[e,p] * obj_scores[subj] * obj_scores[obj]; edges are ordered by descending score (ties by
 * ascending edge id: deterministic, torch.sort in the reference is not) and rels_out [E,2] = (subj, obj) ids,
 * pred_out [E,P] = probabilities are written in that order.  rel_dists [E,P]: logits when apply_softmax != 0 (the
 * softmax of rel_model_stanford.py:206 is fused), probabilities otherwise.  col_img >= 0: rel_inds also holds an image
 * id in that column and the ranking is per image (rows grouped by ascending image id) — batched evaluation; -1: one
 * image.  score_out [E] / order_out [E] (original row of each output row) are optional. */
SGG_API size_t sgg_rank_relations_workspace_bytes(int E, int P);
SGG_API int sgg_rank_relations(const float *rel_dists, int apply_softmax, const float *obj_scores,
                       const int64_t *rel_inds, int64_t row_stride, int col_img, int col_subj, int col_obj,
                       int N, int E, int P, int64_t *rels_out, float *pred_out, float *score_out, int *order_out,
                       void *ws, size_t ws_bytes, void *stream);
/* host-synchronising: SGG_E_INDEX if an endpoint was outside [0,N) in the last call on this workspace */
SGG_API int sgg_rank_relations_check(const void *ws, void *stream);

/* ==== training tail: lib/losses.py, lib/pytorch_misc.py grad_clip / get_optim =============== */

/* ---- edge_losses / node_losses: lib/losses.py:5-74 ---------------------------
 * One call = per-row cross entropy, the reference's row weighting, the summed loss and (optionally) d loss/d logits.
 *   SGG_LOSS_MEAN       node_losses :73-74       CE(reduction='mean'), rows with label -100 ignored
 *   SGG_LOSS_BASELINE   edge_losses :40-44       gamma * CE / M          (requires alpha == beta == 1, as :42 asserts)
 *   SGG_LOSS_DNORM      edge_losses :46-62       FG rows alpha/M_FG, BG rows beta/M_FG (weights stay 1 where M_FG == 0)
 *   SGG_LOSS_DNORM_FGBG edge_losses :46-64       FG rows alpha/M_FG, BG rows beta/M_BG
 * logits [M,C] fp32; labels [M] int64; category (nullable) int8 [M]: 1 = FG row, 2 = BG row, 0 = neither — the
 * explicit idx_fg / idx_bg of :27-31; NULL => FG = label > 0, BG = label == 0.
 * loss: device float[1]; dlogits (nullable) [M,C]; counts_out (nullable) device int[4] = {M_FG, M_BG, rows counted by
 * the mean, rows whose label was outside [0,C) and not -100 (an error the caller may check)}. */
#define SGG_LOSS_MEAN 0
#define SGG_LOSS_BASELINE 1
#define SGG_LOSS_DNORM 2
#define SGG_LOSS_DNORM_FGBG 3
SGG_API size_t sgg_ce_loss_workspace_bytes(int M);
SGG_API int sgg_ce_loss(const float *logits, const int64_t *labels, const int8_t *category, int M, int C, int mode,
                float alpha, float beta, float gamma, float *loss, float *dlogits, int *counts_out,
                void *ws, size_t ws_bytes, void *stream);

/* ---- multi-tensor gradient norm / clipping / SGD step --------------------------
 * Replaces clip_grad_norm (lib/pytorch_misc.py:625-664, called by grad_clip :70-73) and optim.SGD(momentum,
 * weight_decay) built by get_optim (:130-157), main.py:118-120.  One table row per parameter tensor; the caller
 * fills a HOST array of sgg_mt_tensor and uploads it once (and again whenever a pointer / lr / flag changes). */
typedef struct {
  float *p;            /* parameter                                                                    */
  const float *g;      /* gradient; NULL = no gradient this step (tensor is skipped, as torch does)    */
  float *m;            /* momentum buffer (same size as p)                                             */
  void *split;         /* nullable: 2n fp16 [hi | lo*2^11] tensor-core operand of the NEW weight (n % 8 == 0) */
  long long n;         /* elements                                                                     */
  float lr, wd;        /* per-group learning rate / weight decay                                       */
  int flags;           /* bit 0: momentum buffer not initialised yet (first step: m = d)               */
  int reserved;
} sgg_mt_tensor;
SGG_API int sgg_mt_chunk_elems(void);
SGG_API size_t sgg_mt_table_bytes(int n_tensors);
SGG_API long long sgg_mt_total_chunks(const sgg_mt_tensor *host_table, int n_tensors);
SGG_API size_t sgg_mt_workspace_bytes(long long total_chunks);
SGG_API int sgg_mt_table_upload(const sgg_mt_tensor *host_table, int n_tensors, void *table, size_t table_bytes,
                        void *stream);
/* norm_out: device float[4] = {total_norm, clip_coef = max_norm / (total_norm + 1e-6), scale applied
 * (clip_coef if < 1 else 1; 1 when max_norm <= 0), sum of squares}.  Deterministic (fixed-order partials). */
SGG_API int sgg_mt_grad_norm(const void *table, int n_tensors, long long total_chunks, float max_norm, float *norm_out,
                     void *ws, size_t ws_bytes, void *stream);
/* same with a gradient pre-scale: the gradients in memory are the data-parallel SUM and the step means
 * grad_scale * g (grad_scale = 1 / world): norm_out[0] is the norm of the scaled gradients, norm_out[2] the clip factor
 * times grad_scale — the factor sgg_mt_sgd_step / sgg_mt_scale_grads apply.  max_norm <= 0: norm_out[2] = grad_scale. */
SGG_API int sgg_mt_grad_norm_scaled(const void *table, int n_tensors, long long total_chunks, float max_norm,
                            float grad_scale, float *norm_out, void *ws, size_t ws_bytes, void *stream);
/* in-place g *= norm[2] (the reference's clip_grad_norm(..., clip=True) side effect) */
SGG_API int sgg_mt_scale_grads(const void *table, int n_tensors, long long total_chunks, const float *norm, void *stream);
/* d = g*norm[2] + wd*p; m = first ? d : momentum*m + d; p -= lr*m (+ optional operand split, + optional write-back of
 * the clipped gradient).  norm NULL => no clipping. */
SGG_API int sgg_mt_sgd_step(const void *table, int n_tensors, long long total_chunks, const float *norm, float momentum,
                    int write_clipped_grads, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SGG_B200_H */
