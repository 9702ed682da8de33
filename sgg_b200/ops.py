"""Tensor-level entry points over the C-ABI (include/sgg_b200.h).

torch is plumbing here: it owns device memory and the current stream; every op
below validates its tensors, hands raw device pointers to ``libsgg_b200.so`` and
returns fresh torch tensors.  There is no CPU path — CPU tensors raise.
"""
import ctypes as C
import weakref
import torch

from . import _lib
from ._lib import MpWeights, HeadWeights, GeomWeights, check

MP_KEYS = ('edge_gru.weight_ih', 'edge_gru.weight_hh', 'edge_gru.bias_ih', 'edge_gru.bias_hh',
           'node_gru.weight_ih', 'node_gru.weight_hh', 'node_gru.bias_ih', 'node_gru.bias_hh')
GATE_KEYS = ('sub_vert_w_fc', 'obj_vert_w_fc', 'out_edge_w_fc', 'in_edge_w_fc')


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, name, shape=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.SggError('%s must be a CUDA tensor (sgg_b200 has no CPU path)' % name)
    if t.dtype != torch.float32:
        raise _lib.SggError('%s must be float32, got %s' % (name, t.dtype))
    t = t.detach()
    if not t.is_contiguous():
        t = t.contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise _lib.SggError('%s has shape %s, expected %s' % (name, tuple(t.shape), tuple(shape)))
    return t


def _i64_rows(t, name):
    """int64 index matrix, possibly a column-slice view of a wider row-major matrix."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.SggError('%s must be a CUDA tensor' % name)
    if t.dtype != torch.int64 or t.dim() != 2:
        raise _lib.SggError('%s must be a 2-d int64 tensor' % name)
    if t.shape[0] > 1 and (t.stride(1) != 1):
        t = t.contiguous()
    if t.shape[0] <= 1:
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else t.shape[1])


# ---- tensor-core mode ---------------------------------------------------------------------
# 'tc16' : GEMM-shaped stages on tcgen05 tensor cores, 3xFP16 split (kind::f16, csrc/tc16_gemm.cu)
# 'tc32' : same stages, 3xTF32 split (kind::tf32, csrc/tc_gemm.cu)
# 'tc'   : whichever of the two the library defaults to (SGG_TC_DEFAULT_MODE / env SGG_TC_MODE)
# 'simt' : fp32 FMA tile kernels
# All four are CUDA paths with fp32-grade accuracy; there is no CPU path.
_MODE = {'gemm': 'tc'}
_SPLIT_CACHE = {}


def set_gemm_mode(mode):
    if mode not in ('tc', 'tc16', 'tc32', 'simt'):
        raise ValueError(mode)
    if mode in ('tc16', 'tc32'):
        check(_lib.load().sgg_tc_set_mode(1 if mode == 'tc16' else 0), 'sgg_tc_set_mode')
        _SPLIT_CACHE.clear()        # a split is only valid for the engine it was made for
    _MODE['gemm'] = mode


def get_gemm_mode():
    return _MODE['gemm']


def _use_tc():
    return _MODE['gemm'] != 'simt'


def tc_engine():
    """'tc16' or 'tc32': the tensor-core engine the library currently dispatches to."""
    return 'tc16' if _lib.load().sgg_tc_get_mode() == 1 else 'tc32'


def tc16_overflow(reset=True):
    """Sticky fp16 range flag of the 3xFP16 engine (0 = clean).  Synchronises the device."""
    v = _lib.load().sgg_tc16_overflow(1 if reset else 0)
    if v < 0:
        raise _lib.SggError('sgg_tc16_overflow failed')
    return v


def run_range_guarded(fn):
    """Run ``fn()`` on the current engine; if the 3xFP16 engine saw an operand outside the fp16 range (|x| >= 65504,
    which it would silently turn into inf), re-run it on the 3xTF32 engine (fp32 exponent range) and keep that engine.
    Costs one device synchronisation: use it where the host waits for the result anyway (the evaluation tail)."""
    if not _use_tc() or tc_engine() != 'tc16':
        return fn()
    tc16_overflow(reset=True)
    out = fn()
    if tc16_overflow(reset=True):
        import warnings
        warnings.warn('sgg_b200: operand outside the fp16 range; switching the tensor-core engine to 3xTF32')
        set_gemm_mode('tc32')
        out = fn()
    return out


def _tc_k_ok(K):
    return K % (8 if _lib.load().sgg_tc_get_mode() == 1 else 4) == 0


def split_weight(w):
    """[hi | lo] 3xTF32 split of a weight tensor, cached per (storage, version) so it is recomputed only
    after the parameter changes (optimizer step / load_state_dict)."""
    key = id(w)
    ver = (w._version, w.data_ptr(), tuple(w.shape), _lib.load().sgg_tc_get_mode())
    hit = _SPLIT_CACHE.get(key)
    if hit is not None and hit[0] == ver and hit[2]() is w:      # same live tensor object, unchanged
        return hit[1]
    lib = _lib.load()
    wc = _f32(w, 'weight')
    out = torch.empty((2,) + tuple(wc.shape), dtype=torch.float32, device=wc.device)
    check(lib.sgg_tc_split_weights(_ptr(wc), wc.numel(), _ptr(out), _stream()), 'sgg_tc_split_weights')
    if len(_SPLIT_CACHE) > 256:
        _SPLIT_CACHE.clear()
    _SPLIT_CACHE[key] = (ver, out, weakref.ref(w))
    return out


def invalidate_split_cache(params=None):
    """Drop cached tensor-core operand splits (and conv weight planes).  The caches are keyed by the tensor object and
    ``_version``; writes that do NOT bump the version — ``p.data.copy_()/mul_()``, ``dist.broadcast(p.data)``, EMA /
    weight surgery through ``.data`` — leave a stale split behind, so call this after any such write (or pass the
    tensors that changed).  ``load_state_dict`` / ``optimizer.step`` bump versions and need nothing."""
    if params is None:
        _SPLIT_CACHE.clear(); _CONV_W_CACHE.clear()
        return
    for p in (params.values() if isinstance(params, dict) else params):
        _SPLIT_CACHE.pop(id(p), None); _CONV_W_CACHE.pop(id(p), None)


def split_cache_peek(w):
    """The cached operand-split buffer of the live tensor object ``w`` (whatever version it was made for), or None.
    Used by the fused optimizer (sgg_b200.optim), which rewrites the split in the same sweep that updates ``w``."""
    hit = _SPLIT_CACHE.get(id(w))
    if hit is None or hit[2]() is not w or hit[0][1:] != (w.data_ptr(), tuple(w.shape), _lib.load().sgg_tc_get_mode()):
        return None
    return hit[1]


def split_cache_commit(w):
    """Re-stamp the cached split of ``w`` with its current version (the caller has just rewritten the buffer)."""
    hit = _SPLIT_CACHE.get(id(w))
    if hit is not None and hit[2]() is w:
        _SPLIT_CACHE[id(w)] = ((w._version, w.data_ptr(), tuple(w.shape), _lib.load().sgg_tc_get_mode()), hit[1], hit[2])


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Graph(object):
    """Device-side ragged index of the candidate-edge graph (int32 endpoints + two CSRs)."""

    def __init__(self, ws, N, E):
        self.ws, self.N, self.E = ws, N, E

    def check(self):
        check(_lib.load().sgg_graph_check(_ptr(self.ws), self.N, self.E, _stream()), 'sgg_graph_check')
        return self


def build_graph(rel_inds, N, col_subj=0, col_obj=1, validate=False):
    """rel_inds: int64 [E, >=2] with GLOBAL object ids in columns (col_subj, col_obj)
    (the ``rel_inds[:, 1:3]`` the reference passes to message_pass, rel_model_stanford.py:105)."""
    lib = _lib.load()
    rel, stride = _i64_rows(rel_inds, 'rel_inds')
    E = rel.shape[0]
    nbytes = lib.sgg_graph_workspace_bytes(N, E)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=rel.device)
    check(lib.sgg_graph_build(_ptr(rel), stride, col_subj, col_obj, N, E, _ptr(ws), nbytes, _stream()),
          'sgg_graph_build')
    g = Graph(ws, N, E)
    if validate:
        g.check()
    return g


def mp_weights(params, device=None):
    """params: mapping state-dict key -> tensor (reference names).
    Returns (struct, keepalive list)."""
    w = MpWeights()
    keep = []
    for field, key in zip(('edge_w_ih', 'edge_w_hh', 'edge_b_ih', 'edge_b_hh',
                           'node_w_ih', 'node_w_hh', 'node_b_ih', 'node_b_hh'), MP_KEYS):
        t = _f32(params[key], key); keep.append(t)
        setattr(w, field, t.data_ptr())
    H = keep[0].shape[1]
    for i, k in enumerate(GATE_KEYS):
        tw = _f32(params[k + '.0.weight'], k + '.0.weight', (1, 2 * H)); tb = _f32(params[k + '.0.bias'], k + '.0.bias', (1,))
        keep += [tw, tb]
        w.gate_w[i] = tw.data_ptr(); w.gate_b[i] = tb.data_ptr()
    if _use_tc():
        for field, key in (('edge_w_ih_split', 'edge_gru.weight_ih'), ('edge_w_hh_split', 'edge_gru.weight_hh'),
                           ('node_w_ih_split', 'node_gru.weight_ih'), ('node_w_hh_split', 'node_gru.weight_hh')):
            sp = split_weight(params[key]); keep.append(sp)
            setattr(w, field, sp.data_ptr())
    return w, keep, H


def head_weights(params):
    hw = HeadWeights()
    keep = []
    for field, key in (('obj_unary_w', 'obj_unary.weight'), ('obj_unary_b', 'obj_unary.bias'),
                       ('edge_unary_w', 'edge_unary.weight'), ('edge_unary_b', 'edge_unary.bias'),
                       ('obj_fc_w', 'obj_fc.weight'), ('obj_fc_b', 'obj_fc.bias'),
                       ('rel_fc_w', 'rel_fc.weight'), ('rel_fc_b', 'rel_fc.bias')):
        t = _f32(params[key], key); keep.append(t)
        setattr(hw, field, t.data_ptr())
    if _use_tc():
        for field, key in (('obj_unary_w_split', 'obj_unary.weight'), ('edge_unary_w_split', 'edge_unary.weight'),
                           ('obj_fc_w_split', 'obj_fc.weight'), ('rel_fc_w_split', 'rel_fc.weight')):
            sp = split_weight(params[key]); keep.append(sp)
            setattr(hw, field, sp.data_ptr())
    return hw, keep


def message_pass(rel_rep, obj_rep, graph, params, mp_iter=3, save_states=False):
    """RelModelStanford.message_pass (rel_model_stanford.py:48-94).  Argument order follows the
    reference (rel_rep first).  Returns (V_T [N,H], E_T [E,H]) (+ saved states if requested)."""
    lib = _lib.load()
    w, keep, H = mp_weights(params)
    N, E = graph.N, graph.E
    obj_rep = _f32(obj_rep, 'obj_rep', (N, H)); rel_rep = _f32(rel_rep, 'rel_rep', (E, H))
    dev = obj_rep.device
    V = torch.empty((N, H), dtype=torch.float32, device=dev)
    Eo = torch.empty((E, H), dtype=torch.float32, device=dev)
    saved = (torch.empty((lib.sgg_mp_tape_bytes(N, E, H, mp_iter) // 4,), dtype=torch.float32, device=dev)
             if save_states else None)   # tape: states first, then caches (include/sgg_b200.h)
    nbytes = lib.sgg_mp_workspace_bytes(N, E, H, mp_iter)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib.sgg_mp_forward(_ptr(obj_rep), _ptr(rel_rep), _ptr(graph.ws), C.byref(w), N, E, H, mp_iter,
                             _ptr(V), _ptr(Eo), _ptr(saved), _ptr(ws), nbytes, _stream()), 'sgg_mp_forward')
    if save_states:
        return V, Eo, saved
    return V, Eo


# ---- VGG16 conv stack on tensor cores (csrc/conv_tc.cu) -------------------------------------------------------
_CONV_W_CACHE = {}


def conv_weight_planes(w):
    """[Cout,Cin,3,3] fp32 -> fp16 [hi | lo] planes in implicit-GEMM order, cached per (tensor object, version)."""
    key = id(w)
    ver = (w._version, w.data_ptr(), tuple(w.shape))
    hit = _CONV_W_CACHE.get(key)
    if hit is not None and hit[0] == ver and hit[2]() is w:
        return hit[1]
    wc = _f32(w, 'conv weight')
    Cout, Cin = wc.shape[0], wc.shape[1]
    out = torch.empty((2, Cout, 9, Cin), dtype=torch.float16, device=wc.device)
    check(_lib.load().sgg_conv_weight_planes(_ptr(wc), Cout, Cin, _ptr(out), _stream()), 'sgg_conv_weight_planes')
    if len(_CONV_W_CACHE) > 64:
        _CONV_W_CACHE.clear()
    _CONV_W_CACHE[key] = (ver, out, weakref.ref(w))
    return out


def vgg_layers(features):
    """torchvision VGG ``features`` Sequential (Conv2d 3x3 pad 1, ReLU, MaxPool2d(2,2)) -> [(weight, bias, pool_after)],
    or None if the module is not of that form."""
    import torch.nn as nn
    mods = list(features.children())
    out, i = [], 0
    while i < len(mods):
        m = mods[i]
        if not (isinstance(m, nn.Conv2d) and m.kernel_size == (3, 3) and m.padding == (1, 1) and m.stride == (1, 1)
                and m.dilation == (1, 1) and m.groups == 1 and m.bias is not None):
            return None
        if i + 1 >= len(mods) or not isinstance(mods[i + 1], nn.ReLU):
            return None
        i += 2
        pool = False
        if i < len(mods) and isinstance(mods[i], nn.MaxPool2d):
            mp = mods[i]
            if not (mp.kernel_size in (2, (2, 2)) and mp.stride in (2, (2, 2)) and mp.padding in (0, (0, 0)) and not mp.ceil_mode):
                return None
            pool = True
            i += 1
        out.append((m.weight, m.bias, pool))
    if not out or out[0][0].shape[1] != 3 or out[0][0].shape[0] != 64 or out[0][2]:
        return None
    if any(w.shape[1] % 64 or w.shape[0] % 64 for w, _, _ in out[1:]):
        return None
    return out


def vgg_features(images, layers):
    """images [B,3,H,W] fp32 (normalised, padded: GeneralizedRCNNTransform output) -> fmap [B,Cout,H',W'] fp32 NCHW.
    ``layers`` from ``vgg_layers``.  Every convolution after the first runs on tcgen05 (3xFP16, fp32-grade)."""
    lib = _lib.load()
    x = _f32(images, 'images')
    B, C3, H, W = x.shape
    if C3 != 3:
        raise _lib.SggError('vgg_features: images must be [B,3,H,W]')
    npool = sum(1 for _, _, p in layers if p)
    if H % (1 << npool) or W % (1 << npool):
        raise _lib.SggError('vgg_features: H and W must be multiples of %d' % (1 << npool))
    dev = x.device
    w0, b0, _ = layers[0]
    cur = torch.empty((2, B, H, W, 64), dtype=torch.float16, device=dev)
    check(lib.sgg_conv3x3_first(_ptr(x), _ptr(_f32(w0, 'w0')), _ptr(_f32(b0, 'b0')), B, H, W, 64, _ptr(cur), _stream()),
          'sgg_conv3x3_first')
    h, w_, cin = H, W, 64
    fmap = None
    for li, (wt, bs, pool) in enumerate(layers[1:], start=1):
        cout = wt.shape[0]
        wp = conv_weight_planes(wt)
        ho, wo = (h // 2, w_ // 2) if pool else (h, w_)
        last = li == len(layers) - 1
        if last:
            fmap = torch.empty((B, cout, ho, wo), dtype=torch.float32, device=dev)
            nxt = None
        else:
            nxt = torch.empty((2, B, ho, wo, cout), dtype=torch.float16, device=dev)
        check(lib.sgg_conv3x3_tc(_ptr(cur), _ptr(wp), _ptr(_f32(bs, 'bias')), B, h, w_, cin, cout, 1, 1 if pool else 0,
                                 _ptr(nxt), _ptr(fmap) if last else None, _stream()), 'sgg_conv3x3_tc')
        cur, h, w_, cin = nxt, ho, wo, cout
    return fmap


class MpProbe(object):
    """Per-launch timing probe of the fused message-passing schedule (bench.py roofline): runs the whole loop once on a
    private workspace, then ``launch(which)`` re-issues ONE launch of iteration 0 (0 = INIT, 1 = launch A, 2 = launch B)."""

    def __init__(self, rel_rep, obj_rep, graph, params, mp_iter=3):
        lib = _lib.load()
        self.w, self._keep, H = mp_weights(params)
        self.N, self.E, self.H, self.T, self.graph = graph.N, graph.E, H, mp_iter, graph
        self.obj = _f32(obj_rep, 'obj_rep', (self.N, H)); self.rel = _f32(rel_rep, 'rel_rep', (self.E, H))
        dev = self.obj.device
        self.V = torch.empty((self.N, H), dtype=torch.float32, device=dev)
        self.Eo = torch.empty((self.E, H), dtype=torch.float32, device=dev)
        self.nbytes = lib.sgg_mp_workspace_bytes(self.N, self.E, H, mp_iter)
        self.ws = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        check(lib.sgg_mp_forward(_ptr(self.obj), _ptr(self.rel), _ptr(graph.ws), C.byref(self.w), self.N, self.E, H, mp_iter,
                                 _ptr(self.V), _ptr(self.Eo), None, _ptr(self.ws), self.nbytes, _stream()), 'sgg_mp_forward')

    def launch(self, which):
        check(_lib.load().sgg_mp_probe_launch(which, _ptr(self.obj), _ptr(self.rel), _ptr(self.graph.ws), C.byref(self.w),
                                              self.N, self.E, self.H, self.T, _ptr(self.V), _ptr(self.Eo), _ptr(self.ws),
                                              self.nbytes, _stream()), 'sgg_mp_probe_launch')


def linear(x, weight, bias=None, relu=False, x_planes=None, out_planes=False):
    """nn.Linear forward (+ReLU) — y = act(x @ weight.T + bias).

    ``x_planes`` [2, M, K] float16 (optional): the fp16 [hi | lo * 2^11] operand planes of ``x`` written by its producer
    (``node_edge_features(planes=True)`` or a previous ``linear(out_planes=True)``); on the 3xFP16 engine the GEMM then
    runs on the pre-split kernel (csrc/lin16p.cu, no in-kernel conversion of the activations).  ``out_planes=True``
    returns ``(y, y_planes)`` — ``y_planes`` is None when the pre-split kernel was not used."""
    lib = _lib.load()
    w_obj = weight                 # the split cache is keyed by the caller's (long-lived) tensor object, not the detached view
    weight = _f32(weight, 'weight')
    if x is None:                  # planes-only input (node_edge_features(planes='only'))
        if x_planes is None or x_planes.dim() != 3:
            raise _lib.SggError('linear: x is None and no x_planes [2, M, K] given')
        if not (_use_tc() and lib.sgg_tc_get_mode() == 1 and x_planes.shape[2] % 8 == 0 and weight.shape[0] % 4 == 0):
            raise _lib.SggError('linear: planes-only input needs the 3xFP16 engine, K % 8 == 0 and Nout % 4 == 0')
        M, K = int(x_planes.shape[1]), int(x_planes.shape[2])
        dev = x_planes.device
        if weight.dim() != 2 or weight.shape[1] != K:
            raise _lib.SggError('linear: x_planes %s vs weight %s' % (tuple(x_planes.shape), tuple(weight.shape)))
    else:
        x = _f32(x, 'x')
        if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1]:
            raise _lib.SggError('linear: x %s vs weight %s' % (tuple(x.shape), tuple(weight.shape)))
        M, K = x.shape
        dev = x.device
    if bias is not None:
        bias = _f32(bias, 'bias', (weight.shape[0],))
    Nout = weight.shape[0]
    y = torch.empty((M, Nout), dtype=torch.float32, device=dev)
    ypl = None
    if M == 0:                     # pair-less batch: nothing to compute (and no fp32 x to fall back on when planes-only)
        if out_planes:
            return y, (torch.empty((2, 0, Nout), dtype=torch.float16, device=dev) if x_planes is not None else None)
        return y
    if (x_planes is not None and _use_tc() and lib.sgg_tc_get_mode() == 1 and K % 8 == 0 and Nout % 4 == 0 and M > 0):
        if (x_planes.dtype != torch.float16 or tuple(x_planes.shape) != (2, M, K) or not x_planes.is_contiguous()
                or x_planes.device != dev):
            raise _lib.SggError('linear: x_planes must be a contiguous float16 [2, %d, %d] tensor on %s' % (M, K, dev))
        sp = split_weight(w_obj)
        if out_planes:
            ypl = torch.empty((2, M, Nout), dtype=torch.float16, device=dev)
        check(lib.sgg_tc16_linear_pre(_ptr(x_planes), _ptr(sp), _ptr(bias), _ptr(y), _ptr(ypl), M, Nout, K,
                                      1 if relu else 0, None, _stream()), 'sgg_tc16_linear_pre')
    elif _use_tc() and _tc_k_ok(K):
        sp = split_weight(w_obj)
        nb = lib.sgg_tc_linear_workspace_bytes(M, Nout, K)
        ws = torch.empty(nb, dtype=torch.uint8, device=x.device) if nb else None
        check(lib.sgg_tc_linear_forward(_ptr(x), _ptr(sp), _ptr(bias), _ptr(y), M, Nout, K, 1 if relu else 0,
                                        _ptr(ws), nb, _stream()), 'sgg_tc_linear_forward')
    else:
        check(lib.sgg_linear_forward(_ptr(x), _ptr(weight), _ptr(bias), _ptr(y), M, Nout, K, 1 if relu else 0,
                                     _stream()), 'sgg_linear_forward')
    return (y, ypl) if out_planes else y


class L1Plan(object):
    """Preallocated buffers + weight structs for repeated L1 forwards of one (N, E) shape;
    ``run`` enqueues exactly one C-ABI call and is CUDA-graph capturable."""

    def __init__(self, params, N, E, D=4096, mp_iter=3, device='cuda'):
        lib = _lib.load()
        self.hw, self._k1 = head_weights(params)
        self.w, self._k2, self.H = mp_weights(params)
        self.N, self.E, self.D, self.T = N, E, D, mp_iter
        self.n_cls = params['obj_fc.weight'].shape[0]; self.n_rel = params['rel_fc.weight'].shape[0]
        self.obj_dists = torch.empty((N, self.n_cls), dtype=torch.float32, device=device)
        self.rel_dists = torch.empty((E, self.n_rel), dtype=torch.float32, device=device)
        self.nbytes = lib.sgg_l1_workspace_bytes(N, E, self.H, mp_iter)
        self.ws = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self._fn = lib.sgg_l1_forward
        self.graph_ws, self.graph_bytes = None, 0

    def refresh(self, params):
        """Re-read the weights (and re-split them if they changed): call after the parameters were updated by anything
        other than ``FusedSGD`` in 3xFP16 mode (which rewrites the cached splits in place), or after
        ``invalidate_split_cache()``."""
        self.hw, self._k1 = head_weights(params)
        self.w, self._k2, self.H = mp_weights(params)

    def run(self, obj_feat, edge_feat, graph):
        """graph: a prebuilt ``Graph``, or the int64 rel_inds [E, 2] themselves (global subject / object ids) — then the
        graph index is built inside the same C-ABI call, overlapped with the edge-unary GEMM."""
        if isinstance(graph, Graph):
            check(self._fn(_ptr(obj_feat), _ptr(edge_feat), _ptr(graph.ws), C.byref(self.hw), C.byref(self.w),
                           self.N, self.E, self.D, self.H, self.T, self.n_cls, self.n_rel,
                           _ptr(self.obj_dists), _ptr(self.rel_dists), _ptr(self.ws), self.nbytes, _stream()),
                  'sgg_l1_forward')
            return self.obj_dists, self.rel_dists
        rel, stride = _i64_rows(graph, 'rel_inds')
        if rel.shape[0] != self.E:
            raise _lib.SggError('rel_inds has %d rows, plan was made for E=%d' % (rel.shape[0], self.E))
        if self.graph_ws is None:
            self.graph_bytes = _lib.load().sgg_graph_workspace_bytes(self.N, self.E)
            self.graph_ws = torch.empty(self.graph_bytes, dtype=torch.uint8, device=self.obj_dists.device)
        check(_lib.load().sgg_l1_forward_rel(_ptr(obj_feat), _ptr(edge_feat), _ptr(rel), stride, 0, 1, _ptr(self.graph_ws),
                                             self.graph_bytes, C.byref(self.hw), C.byref(self.w), self.N, self.E, self.D,
                                             self.H, self.T, self.n_cls, self.n_rel, _ptr(self.obj_dists),
                                             _ptr(self.rel_dists), _ptr(self.ws), self.nbytes, _stream()),
              'sgg_l1_forward_rel')
        self._rel_keep = rel
        return self.obj_dists, self.rel_dists


def l1_forward(obj_feat, edge_feat, graph, params, mp_iter=3):
    """4096-d features -> (obj_dists, rel_dists): rel_model_stanford.py:103-107 without roi_fmap*."""
    obj_feat = _f32(obj_feat, 'obj_feat'); edge_feat = _f32(edge_feat, 'edge_feat')
    D = params['obj_unary.weight'].shape[1]
    if obj_feat.shape != (graph.N, D) or edge_feat.shape != (graph.E, D):
        raise _lib.SggError('l1_forward: feature shapes %s %s do not match graph (%d, %d) x %d'
                            % (tuple(obj_feat.shape), tuple(edge_feat.shape), graph.N, graph.E, D))
    plan = L1Plan(params, graph.N, graph.E, D, mp_iter, obj_feat.device)
    return plan.run(obj_feat, edge_feat, graph)


def draw_union_boxes(rois, union_inds, pooling_size=27, sub_half=False):
    """lib/draw_rectangles/draw_rectangles.pyx:12-67 on the device.  rois [N,5], union_inds int64 [E,2]."""
    lib = _lib.load()
    rois = _f32(rois, 'rois'); ui, stride = _i64_rows(union_inds, 'union_inds')
    E = ui.shape[0]
    out = torch.empty((E, 2, pooling_size, pooling_size), dtype=torch.float32, device=rois.device)
    check(lib.sgg_draw_union_boxes(_ptr(rois), _ptr(ui), stride, 0, 1, E, pooling_size, 1 if sub_half else 0,
                                   _ptr(out), _stream()), 'sgg_draw_union_boxes')
    return out


def geom_patches(rois, union_inds):
    """[E,4,98] conv1 windows of draw_union_boxes(...) - 0.5 (training path of the geometry branch)."""
    lib = _lib.load()
    rois = _f32(rois, 'rois'); ui, stride = _i64_rows(union_inds, 'union_inds')
    E = ui.shape[0]
    out = torch.empty((E, 4, 98), dtype=torch.float32, device=rois.device)
    check(lib.sgg_geom_patches(_ptr(rois), _ptr(ui), stride, 0, 1, E, _ptr(out), _stream()), 'sgg_geom_patches')
    return out


def bn_train_forward(x, gamma, beta, running_mean, running_var, momentum, eps, relu_in=True):
    """y = BatchNorm(relu(x)) with batch statistics over the rows of x [M, C] (+ in-place running-stat update).
    Returns (y, save_mean, save_invstd)."""
    lib = _lib.load()
    x = _f32(x, 'x')
    M, Cc = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(Cc, dtype=torch.float32, device=x.device); invstd = torch.empty_like(mean)
    nb = lib.sgg_bn_workspace_bytes(M, Cc)
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
    check(lib.sgg_bn_train_forward(_ptr(x), M, Cc, 1 if relu_in else 0, _ptr(_f32(gamma, 'gamma', (Cc,))),
                                   _ptr(_f32(beta, 'beta', (Cc,))), _ptr(running_mean), _ptr(running_var), float(momentum),
                                   float(eps), _ptr(y), _ptr(mean), _ptr(invstd), _ptr(ws), nb, _stream()),
          'sgg_bn_train_forward')
    return y, mean, invstd


def bn_train_backward(x, dy, gamma, mean, invstd, relu_in=True):
    """-> (dx w.r.t. the pre-activation x, dgamma, dbeta)"""
    lib = _lib.load()
    x = _f32(x, 'x'); dy = _f32(dy, 'dy', tuple(x.shape))
    M, Cc = x.shape
    dx = torch.empty_like(x)
    dgamma = torch.empty(Cc, dtype=torch.float32, device=x.device); dbeta = torch.empty_like(dgamma)
    nb = lib.sgg_bn_workspace_bytes(M, Cc)
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
    check(lib.sgg_bn_train_backward(_ptr(x), _ptr(dy), M, Cc, 1 if relu_in else 0, _ptr(_f32(gamma, 'gamma', (Cc,))),
                                    _ptr(mean), _ptr(invstd), _ptr(dx), _ptr(dgamma), _ptr(dbeta), _ptr(ws), nb, _stream()),
          'sgg_bn_train_backward')
    return dx, dgamma, dbeta


def max4_forward(x):
    """x [E, 4, C] -> (max over the 4 positions [E, C], arg-max uint8 [E, C])"""
    lib = _lib.load()
    x = _f32(x, 'x')
    E, four, Cc = x.shape
    assert four == 4
    y = torch.empty((E, Cc), dtype=torch.float32, device=x.device)
    idx = torch.empty((E, Cc), dtype=torch.uint8, device=x.device)
    check(lib.sgg_max4_forward(_ptr(x), E, Cc, _ptr(y), _ptr(idx), _stream()), 'sgg_max4_forward')
    return y, idx


def max4_backward(dy, idx):
    lib = _lib.load()
    dy = _f32(dy, 'dy')
    E, Cc = dy.shape
    dx = torch.empty((E, 4, Cc), dtype=torch.float32, device=dy.device)
    check(lib.sgg_max4_backward(_ptr(dy), _ptr(idx), E, Cc, _ptr(dx), _stream()), 'sgg_max4_backward')
    return dx


def bcast_add(pools, geom, planes=False):
    """pools [E,C,P,P] + geom [E,C] broadcast over the P x P positions (lib/get_union_boxes.py:101).
    ``planes=True``: returns (out, out_planes [2, E, C*P*P] float16) — the operand planes for ``linear(x_planes=...)``."""
    lib = _lib.load()
    pools = _f32(pools, 'pools'); geom = _f32(geom, 'geom', (pools.shape[0], pools.shape[1]))
    out = torch.empty_like(pools)
    S = pools.shape[2] * pools.shape[3]
    if planes:
        pl = torch.empty((2, pools.shape[0], pools.shape[1] * S), dtype=torch.float16, device=pools.device)
        check(lib.sgg_bcast_add_planes(_ptr(pools), _ptr(geom), pools.shape[0] * pools.shape[1], S, _ptr(out), _ptr(pl),
                                       _stream()), 'sgg_bcast_add_planes')
        return out, pl
    check(lib.sgg_bcast_add(_ptr(pools), _ptr(geom), pools.shape[0] * pools.shape[1], S, _ptr(out), _stream()), 'sgg_bcast_add')
    return out


def relu_backward(dy, y):
    lib = _lib.load()
    dy = _f32(dy, 'dy'); y = _f32(y, 'y', tuple(dy.shape))
    if dy.numel() % 4:
        return dy * (y > 0)
    dx = torch.empty_like(dy)
    check(lib.sgg_relu_backward(_ptr(dy), _ptr(y), dy.numel(), _ptr(dx), _stream()), 'sgg_relu_backward')
    return dx


def group_sum(x, S):
    """x [..., S] contiguous -> sum over the last S elements, shape x.shape[:-1]"""
    lib = _lib.load()
    x = _f32(x, 'x')
    assert x.shape[-1] == S
    out = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
    check(lib.sgg_group_sum(_ptr(x), out.numel(), S, _ptr(out), _stream()), 'sgg_group_sum')
    return out


def geom_weights(params, prefix='union_boxes.conv.'):
    gw = GeomWeights()
    keep = []
    for field, key in (('conv1_w', '0.weight'), ('conv1_b', '0.bias'), ('bn1_w', '2.weight'), ('bn1_b', '2.bias'),
                       ('bn1_rm', '2.running_mean'), ('bn1_rv', '2.running_var'), ('conv2_w', '4.weight'),
                       ('conv2_b', '4.bias'), ('bn2_w', '6.weight'), ('bn2_b', '6.bias'),
                       ('bn2_rm', '6.running_mean'), ('bn2_rv', '6.running_var')):
        t = _f32(params[prefix + key], prefix + key); keep.append(t)
        setattr(gw, field, t.data_ptr())
    C_out = params[prefix + '4.weight'].shape[0]
    return gw, keep, C_out


def union_geom(rois, union_inds, params, union_pools=None, prefix='union_boxes.conv.'):
    """UnionBoxesAndFeats.forward, edge_model='motifs', eval-mode BN (lib/get_union_boxes.py:63-101).
    Returns union_pools + geom (broadcast) if union_pools is given, else geom [E, C]."""
    lib = _lib.load()
    rois = _f32(rois, 'rois'); ui, stride = _i64_rows(union_inds, 'union_inds')
    gw, keep, Cc = geom_weights(params, prefix)
    E = ui.shape[0]
    if union_pools is not None:
        union_pools = _f32(union_pools, 'union_pools', (E, Cc, 7, 7))
        out = torch.empty_like(union_pools)
    else:
        out = torch.empty((E, Cc), dtype=torch.float32, device=rois.device)
    nbytes = lib.sgg_union_geom_workspace_bytes(E, Cc)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=rois.device)
    check(lib.sgg_union_geom_forward(_ptr(rois), _ptr(ui), stride, 0, 1, E, Cc, C.byref(gw), _ptr(union_pools),
                                     _ptr(out), _ptr(ws), nbytes, _stream()), 'sgg_union_geom_forward')
    return out


def node_edge_features(fmap, rois, union_inds, spatial_scale=1.0 / 16, pool=7, sampling_ratio=2,
                       want_node=True, want_edge=True, fast=True, edge_add=None, planes=False):
    """RelModelBase.node_edge_features (rel_model_base.py:245-260).  ``edge_add`` [E,C] (optional): the union-box
    geometry embedding, added to every bin of the edge rows in the same kernel (lib/get_union_boxes.py:101).
    ``planes=True`` also returns the fp16 operand planes of both outputs ([2, rows, C*pool*pool] float16 each) for
    ``linear(..., x_planes=...)``: (node, edge, node_planes, edge_planes).  ``planes='only'``: the fp32 rows are not
    written at all (node and edge are None) — the eval forward, where nothing but the fc6 layers reads them."""
    lib = _lib.load()
    fmap = _f32(fmap, 'fmap'); rois = _f32(rois, 'rois'); ui, stride = _i64_rows(union_inds, 'union_inds')
    B, Cc, Hf, Wf = fmap.shape
    N, E = rois.shape[0], ui.shape[0]
    if edge_add is not None:
        edge_add = _f32(edge_add, 'edge_add', (E, Cc))
    rows32 = planes != 'only'
    node = torch.empty((N, Cc, pool, pool), dtype=torch.float32, device=fmap.device) if (want_node and rows32) else None
    edge = torch.empty((E, Cc, pool, pool), dtype=torch.float32, device=fmap.device) if (want_edge and rows32) else None
    nb = lib.sgg_node_edge_features_workspace_bytes(B, Cc, Hf, Wf) if fast else 0
    ws = torch.empty(nb, dtype=torch.uint8, device=fmap.device) if nb else None
    if planes:
        if not (fast and sampling_ratio == 2 and Cc % 4 == 0):
            raise _lib.SggError('node_edge_features: planes need the channel-last path (fast, sampling_ratio 2, C % 4 == 0)')
        per = Cc * pool * pool
        npl = torch.empty((2, N, per), dtype=torch.float16, device=fmap.device) if want_node else None
        epl = torch.empty((2, E, per), dtype=torch.float16, device=fmap.device) if want_edge else None
        check(lib.sgg_node_edge_features_planes(_ptr(fmap), B, Cc, Hf, Wf, _ptr(rois), N, _ptr(ui), stride, 0, 1, E,
                                                float(spatial_scale), pool, sampling_ratio, _ptr(edge_add), _ptr(node),
                                                _ptr(edge), _ptr(npl), _ptr(epl), _ptr(ws), nb, _stream()),
              'sgg_node_edge_features_planes')
        return node, edge, npl, epl
    check(lib.sgg_node_edge_features_add(_ptr(fmap), B, Cc, Hf, Wf, _ptr(rois), N, _ptr(ui), stride, 0, 1, E,
                                         float(spatial_scale), pool, sampling_ratio, _ptr(edge_add), _ptr(node),
                                         _ptr(edge), _ptr(ws), nb, _stream()), 'sgg_node_edge_features_add')
    return node, edge


def rank_relations(rel_dists, obj_scores, rel_inds, logits=True, per_image=False, validate=False):
    """Device-side filter_dets core (lib/surgery.py:42-52).  rel_inds: int64 [E,2] (subj, obj) or, with
    ``per_image``, [E,3] (img, subj, obj) — then every image's edges are ranked separately in one launch.
    Returns (rels [E,2] int64, pred_scores [E,P] probabilities, scores [E], order [E] int32), all in ranked order."""
    lib = _lib.load()
    x = _f32(rel_dists, 'rel_dists'); sc = _f32(obj_scores, 'obj_scores')
    ri, stride = _i64_rows(rel_inds, 'rel_inds')
    E, P = x.shape
    want_cols = 3 if per_image else 2
    if ri.shape[0] != E or ri.shape[1] != want_cols:
        raise _lib.SggError('rank_relations: rel_inds %s vs rel_dists %s' % (tuple(ri.shape), tuple(x.shape)))
    off = 1 if per_image else 0
    dev = x.device
    rels = torch.empty((E, 2), dtype=torch.int64, device=dev)
    pred = torch.empty((E, P), dtype=torch.float32, device=dev)
    score = torch.empty((E,), dtype=torch.float32, device=dev)
    order = torch.empty((E,), dtype=torch.int32, device=dev)
    nb = lib.sgg_rank_relations_workspace_bytes(E, P)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    check(lib.sgg_rank_relations(_ptr(x), 1 if logits else 0, _ptr(sc), _ptr(ri), stride, 0 if per_image else -1, off,
                                 off + 1, sc.shape[0], E, P, _ptr(rels), _ptr(pred), _ptr(score), _ptr(order),
                                 _ptr(ws), nb, _stream()), 'sgg_rank_relations')
    if validate and E > 0:
        check(lib.sgg_rank_relations_check(_ptr(ws), _stream()), 'sgg_rank_relations_check')
    return rels, pred, score, order


# ---- backward entry points -----------------------------------------------------------------------
# Gradient sinks: a data-parallel reducer (sgg_b200.parallel.FlatGradReducer) registers, per weight matrix, the view of
# its flat gradient buffer the weight gradient should be written INTO (no separate dW tensor, no copy) together with the
# row chunks it all-reduces separately; linear_backward then produces dW chunk by chunk and notifies after each one.
_GRAD_SINKS = {}


def register_grad_sink(param, view, chunks, notify, owner=None):
    _GRAD_SINKS[param.data_ptr()] = (weakref.ref(param), view, chunks, notify, owner)


def clear_grad_sinks(owner=None):
    for k in [k for k, v in _GRAD_SINKS.items() if owner is None or v[4] is owner]:
        del _GRAD_SINKS[k]


def grad_sink(weight):
    hit = _GRAD_SINKS.get(weight.data_ptr())
    if hit is None:
        return None
    p = hit[0]()
    if p is None or tuple(p.shape) != tuple(weight.shape):
        return None
    return hit


# below min_rows / min_red the SIMT tiles (split-K) are faster.  engine: 'tc16' = 3xFP16 with an exact power-of-two scale
# of the gradient operand (twice the MMA rate), 'tc32' = 3xTF32 (fp32 exponent range, no scale needed)
import os as _os
_TC_BWD = {'min_rows': 256, 'min_red': 256, 'enabled': True, 'engine': _os.environ.get('SGG_BWD_ENGINE', 'tc16')}


def _pad32(n):
    return (n + 31) // 32 * 32


def _transpose(inp, R, C, ldin, Rpad, split):
    """inp [R, C] (row stride ldin) -> [C, Rpad] (zero rows beyond R); split: 3xTF32 [hi | lo] planes [2, C, Rpad]."""
    lib = _lib.load()
    out = torch.empty(((2, C, Rpad) if split else (C, Rpad)), dtype=torch.float32, device=inp.device)
    check(lib.sgg_bwd_transpose(_ptr(inp), ldin, R, C, _ptr(out), Rpad, 1 if split else 0, _stream()), 'sgg_bwd_transpose')
    return out


def _tc32_linear(x, w_split, M, Nout, K, out=None):
    """y [M, Nout] = x [M, K] @ w^T on the 3xTF32 tcgen05 engine (fp32 exponent range: safe for gradients)."""
    lib = _lib.load()
    y = out if out is not None else torch.empty((M, Nout), dtype=torch.float32, device=x.device)
    nb = lib.sgg_tc32_linear_workspace_bytes(M, Nout, K)
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device) if nb else None
    check(lib.sgg_tc32_linear_forward(_ptr(x), _ptr(w_split), None, _ptr(y), M, Nout, K, 0, _ptr(ws), nb, _stream()),
          'sgg_tc32_linear_forward')
    return y


def _pow2_scale(x):
    """device floats [s, 1/s], s = 2^k with max|x| * s in [1024, 2048) — no synchronisation"""
    lib = _lib.load()
    sc = torch.empty(2, dtype=torch.float32, device=x.device)
    nb = lib.sgg_pow2_scale_workspace_bytes()
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
    check(lib.sgg_pow2_scale(_ptr(x), x.numel(), _ptr(sc), _ptr(ws), nb, _stream()), 'sgg_pow2_scale')
    return sc


def _transpose16(inp, R, C, ldin, Rpad, planes, sc=None):
    """inp [R, C] (row stride ldin) -> [C, Rpad], scaled by sc[0] when given: planes=False: fp32; planes=True: fp16 [hi | lo]
    planes [2, C, Rpad]"""
    lib = _lib.load()
    if planes:
        out = torch.empty((2, C, Rpad), dtype=torch.float16, device=inp.device)
    else:
        out = torch.empty((C, Rpad), dtype=torch.float32, device=inp.device)
    check(lib.sgg_bwd_transpose16(_ptr(inp), ldin, R, C, _ptr(out), Rpad, 1 if planes else 0, _ptr(sc), _stream()),
          'sgg_bwd_transpose16')
    return out


def _tc16_linear_scaled(x, w_planes, M, Nout, K, inv_scale, out=None):
    """y [M, Nout] = inv_scale[0] * (x [M, K] @ w^T) on the 3xFP16 engine; w_planes fp16 [2, Nout, K]"""
    lib = _lib.load()
    y = out if out is not None else torch.empty((M, Nout), dtype=torch.float32, device=x.device)
    nb = lib.sgg_tc16_linear_workspace_bytes(M, Nout, K)
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device) if nb else None
    check(lib.sgg_tc16_linear_scaled(_ptr(x), _ptr(w_planes), _ptr(y), M, Nout, K, _ptr(inv_scale), _ptr(ws), nb, _stream()),
          'sgg_tc16_linear_scaled')
    return y


def matmul_nn(a, b):
    """a [M,K] @ b [K,N] (fp32-grade): 3xTF32 tensor cores through a transposed, split copy of b when large, else the
    fp32 SIMT tiles."""
    lib = _lib.load()
    a = _f32(a, 'a'); b = _f32(b, 'b')
    M, K = a.shape
    N = b.shape[1]
    if _TC_BWD['enabled'] and _use_tc() and M >= _TC_BWD['min_rows'] and K % 4 == 0 and K >= 64 and N >= 64:
        bT = _transpose(b, K, N, N, K, split=True)                            # [2, N, K]
        return _tc32_linear(a, bT, M, N, K)
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    # dX-type SIMT GEMM: out = a @ b with b [K, N] read "column-wise" (the dx route of sgg_linear_backward: dy W)
    nbytes = lib.sgg_linear_backward_workspace_bytes(M, K, N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
    check(lib.sgg_linear_backward_ex(None, _ptr(b), _ptr(a), M, K, N, _ptr(out), None, None, 0, _ptr(ws), nbytes, _stream()),
          'sgg_linear_backward_ex')
    return out


def linear_backward(x, weight, dy, need_dx=True, need_dw=True, need_db=True):
    """nn.Linear backward through the C-ABI: returns (dx, dw, db); dy must already carry the ReLU mask.

    Large GEMMs run on the 3xTF32 tcgen05 engine (dX = dY W with x = dY, w = split(W^T); dW = dY^T X with x = dY^T,
    w = split(X^T)); small ones on the fp32 SIMT tiles.  dW is written (not accumulated) into a fresh tensor or, when a
    reducer registered a sink for ``weight``, straight into its flat gradient buffer, row chunk by row chunk."""
    lib = _lib.load()
    x = _f32(x, 'x'); w_obj = weight; weight = _f32(weight, 'weight'); dy = _f32(dy, 'dy')
    M, K = x.shape
    Nout = weight.shape[0]
    dev = x.device
    use_tc = _TC_BWD['enabled'] and _use_tc()
    dx = dw = db = None
    sink = grad_sink(w_obj) if need_dw else None
    # ---- dX [M,K] = dY [M,Nout] W [Nout,K]
    tc_dx = need_dx and use_tc and M >= _TC_BWD['min_rows'] and Nout % 8 == 0 and Nout >= 64 and K >= 64
    tc_dw = need_dw and use_tc and M >= _TC_BWD['min_red'] and Nout >= 64 and K >= 64
    tc16 = _TC_BWD['engine'] == 'tc16' and M % 4 == 0
    sc = _pow2_scale(dy) if tc16 and (tc_dx or tc_dw) else None
    if tc_dx and tc16:
        wT = _transpose16(weight, Nout, K, K, Nout, planes=True)          # fp16 planes [2, K, Nout]
        dys = torch.empty_like(dy)
        check(lib.sgg_scale_by(_ptr(dy), dy.numel(), _ptr(sc), _ptr(dys), _stream()), 'sgg_scale_by')
        dx = _tc16_linear_scaled(dys, wT, M, K, Nout, sc[1:])
        del wT, dys
    elif tc_dx:
        wT = _transpose(weight, Nout, K, K, Nout, split=True)             # [2, K, Nout]
        dx = _tc32_linear(dy, wT, M, K, Nout)
        del wT
    # ---- dW [Nout,K] = dY^T X (reduction over the M rows)
    if need_dw:
        dw = sink[1] if sink is not None else torch.empty((Nout, K), dtype=torch.float32, device=dev)
    if tc_dw:
        Mp = _pad32(M)
        chunks = sink[2] if (sink is not None and sink[2]) else [(0, Nout)]
        if tc16 and sc is not None:
            xT = _transpose16(x, M, K, K, Mp, planes=True)                # fp16 planes [2, K, Mp], shared by every row chunk
            # both operands pre-split (csrc/lin16p.cu) when the output has enough 128 x 128 tiles to fill the GPU; small
            # outputs (classifier heads: 4 tiles, 150 k-blocks each) stay on the split-K LINEAR engine
            def pre_ok(rows):
                return K % 4 == 0 and ((rows + 127) // 128) * ((K + 127) // 128) >= 96
            for r0, r1 in chunks:
                if pre_ok(r1 - r0):
                    dyT = _transpose16(dy[:, r0:], M, r1 - r0, Nout, Mp, planes=True, sc=sc)   # planes of (s dY)^T [2, r1-r0, Mp]
                    check(lib.sgg_tc16_linear_pre(_ptr(dyT), _ptr(xT), None, _ptr(dw[r0:r1]), None, r1 - r0, K, Mp, 0,
                                                  _ptr(sc[1:]), _stream()), 'sgg_tc16_linear_pre')
                else:
                    dyT = _transpose16(dy[:, r0:], M, r1 - r0, Nout, Mp, planes=False, sc=sc)  # (s dY)^T [r1-r0, Mp]
                    _tc16_linear_scaled(dyT, xT, r1 - r0, K, Mp, sc[1:], out=dw[r0:r1])
                if sink is not None:
                    sink[3](sink[0](), r0, r1)
        else:
            xT = _transpose(x, M, K, K, Mp, split=True)                   # [2, K, Mp], shared by every row chunk
            for r0, r1 in chunks:
                dyT = _transpose(dy[:, r0:], M, r1 - r0, Nout, Mp, split=False)    # [r1-r0, Mp]
                _tc32_linear(dyT, xT, r1 - r0, K, Mp, out=dw[r0:r1])
                if sink is not None:
                    sink[3](sink[0](), r0, r1)
        del xT
    # ---- remaining pieces on the SIMT tiles (one call; accumulate = 0: plain stores, no zero-fill)
    rest_dx, rest_dw = need_dx and not tc_dx, need_dw and not tc_dw
    if rest_dx or rest_dw or need_db:
        if rest_dx:
            dx = torch.empty((M, K), dtype=torch.float32, device=dev)
        if need_db:
            db = torch.empty((Nout,), dtype=torch.float32, device=dev)
        nbytes = lib.sgg_linear_backward_workspace_bytes(M, Nout, K)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib.sgg_linear_backward_ex(_ptr(x), _ptr(weight), _ptr(dy), M, Nout, K, _ptr(dx) if rest_dx else None,
                                         _ptr(dw) if rest_dw else None, _ptr(db), 0, _ptr(ws), nbytes, _stream()),
              'sgg_linear_backward_ex')
        if rest_dw and sink is not None:
            sink[3](sink[0](), 0, Nout)
    if sink is not None:
        dw = dw.view(Nout, K)      # fresh alias: autograd adopts it as p.grad without a copy (it is the only reference)
    return dx, dw, db


def message_pass_train(rel_rep, obj_rep, graph, params, mp_iter=3):
    """Forward that records the training tape.  Returns (V_T, E_T, tape)."""
    lib = _lib.load()
    w, keep, H = mp_weights(params)
    N, E = graph.N, graph.E
    obj_rep = _f32(obj_rep, 'obj_rep', (N, H)); rel_rep = _f32(rel_rep, 'rel_rep', (E, H))
    dev = obj_rep.device
    V = torch.empty((N, H), dtype=torch.float32, device=dev)
    Eo = torch.empty((E, H), dtype=torch.float32, device=dev)
    tape = torch.empty((lib.sgg_mp_tape_bytes(N, E, H, mp_iter) // 4,), dtype=torch.float32, device=dev)
    nbytes = lib.sgg_mp_workspace_bytes(N, E, H, mp_iter)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib.sgg_mp_forward(_ptr(obj_rep), _ptr(rel_rep), _ptr(graph.ws), C.byref(w), N, E, H, mp_iter,
                             _ptr(V), _ptr(Eo), _ptr(tape), _ptr(ws), nbytes, _stream()), 'sgg_mp_forward')
    return V, Eo, tape


def message_pass_backward(rel_rep, obj_rep, graph, params, tape, dV, dE, mp_iter=3):
    """BPTT through message_pass.  Returns (d_rel_rep, d_obj_rep, grads dict keyed like ``params``)."""
    lib = _lib.load()
    w, keep, H = mp_weights(params)
    N, E = graph.N, graph.E
    obj_rep = _f32(obj_rep, 'obj_rep', (N, H)); rel_rep = _f32(rel_rep, 'rel_rep', (E, H))
    dV = _f32(dV, 'dV', (N, H)); dE = _f32(dE, 'dE', (E, H))
    dev = obj_rep.device
    # the kernels ACCUMULATE over the T iterations: ONE zero-filled slab holds all 16 gradient tensors (one memset
    # instead of 16 fill launches), each view 256-byte aligned
    shapes = [(key, tuple(params[key].shape)) for key in MP_KEYS]
    for k in GATE_KEYS:
        shapes += [(k + '.0.weight', (1, 2 * H)), (k + '.0.bias', (1,))]
    offs, total = [], 0
    for _, shp in shapes:
        n = 1
        for d in shp:
            n *= d
        offs.append((total, n)); total += (n + 63) // 64 * 64
    slab = torch.zeros(total, dtype=torch.float32, device=dev)
    grads = {key: slab[o:o + n].view(shp) for (key, shp), (o, n) in zip(shapes, offs)}
    gs = _lib.MpGrads()
    for field, key in zip(('edge_w_ih', 'edge_w_hh', 'edge_b_ih', 'edge_b_hh',
                           'node_w_ih', 'node_w_hh', 'node_b_ih', 'node_b_hh'), MP_KEYS):
        setattr(gs, field, grads[key].data_ptr())
    for i, k in enumerate(GATE_KEYS):
        gs.gate_w[i] = grads[k + '.0.weight'].data_ptr(); gs.gate_b[i] = grads[k + '.0.bias'].data_ptr()
    d_obj = torch.empty((N, H), dtype=torch.float32, device=dev)
    d_rel = torch.empty((E, H), dtype=torch.float32, device=dev)
    nbytes = lib.sgg_mp_backward_workspace_bytes(N, E, H, mp_iter)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib.sgg_mp_backward(_ptr(obj_rep), _ptr(rel_rep), _ptr(graph.ws), C.byref(w), _ptr(tape), N, E, H, mp_iter,
                              _ptr(dV), _ptr(dE), C.byref(gs), _ptr(d_obj), _ptr(d_rel), _ptr(ws), nbytes, _stream()),
          'sgg_mp_backward')
    return d_rel, d_obj, grads


def edge_gru(Eh, P, gates, graph, params):
    """One edge-GRU update (dominant kernel) in isolation: see sgg_edge_gru_forward in include/sgg_b200.h."""
    lib = _lib.load()
    Eh = _f32(Eh, 'Eh'); P = _f32(P, 'P'); gates = _f32(gates, 'gates')
    E, H = Eh.shape
    w_hh = _f32(params['edge_gru.weight_hh'], 'w_hh'); w_ih = _f32(params['edge_gru.weight_ih'], 'w_ih')
    b_ih = _f32(params['edge_gru.bias_ih'], 'b_ih'); b_hh = _f32(params['edge_gru.bias_hh'], 'b_hh')
    sp = split_weight(params['edge_gru.weight_hh']) if _use_tc() else None
    out = torch.empty_like(Eh)
    check(lib.sgg_edge_gru_forward(_ptr(Eh), _ptr(P), _ptr(gates), _ptr(graph.ws), _ptr(w_ih), _ptr(w_hh), _ptr(sp),
                                   _ptr(b_ih), _ptr(b_hh), graph.N, E, H, _ptr(out), _stream()), 'sgg_edge_gru_forward')
    return out
