"""ctypes binding of the C-ABI library (include/sgg_b200.h).

The product has NO CPU fallback: if ``libsgg_b200.so`` cannot be loaded every op
raises.  ``load()`` builds the library first when nvcc is present (build box);
on the GPU box the prebuilt in-tree ``.so`` is used.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'libsgg_b200.so')
_lib = None

c_f = C.c_void_p       # device pointers are passed as integers (tensor.data_ptr())
c_i64p = C.c_void_p


class MpWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('edge_w_ih', 'edge_w_hh', 'edge_b_ih', 'edge_b_hh',
                                          'node_w_ih', 'node_w_hh', 'node_b_ih', 'node_b_hh')] + \
               [('gate_w', C.c_void_p * 4), ('gate_b', C.c_void_p * 4)] + \
               [(n, C.c_void_p) for n in ('edge_w_ih_split', 'edge_w_hh_split', 'node_w_ih_split', 'node_w_hh_split')]


class MpGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('edge_w_ih', 'edge_w_hh', 'edge_b_ih', 'edge_b_hh',
                                          'node_w_ih', 'node_w_hh', 'node_b_ih', 'node_b_hh')] + \
               [('gate_w', C.c_void_p * 4), ('gate_b', C.c_void_p * 4)]


class HeadWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('obj_unary_w', 'obj_unary_b', 'edge_unary_w', 'edge_unary_b',
                                          'obj_fc_w', 'obj_fc_b', 'rel_fc_w', 'rel_fc_b',
                                          'obj_unary_w_split', 'edge_unary_w_split', 'obj_fc_w_split',
                                          'rel_fc_w_split')]


class GeomWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('conv1_w', 'conv1_b', 'bn1_w', 'bn1_b', 'bn1_rm', 'bn1_rv',
                                          'conv2_w', 'conv2_b', 'bn2_w', 'bn2_b', 'bn2_rm', 'bn2_rv')]


class MtTensor(C.Structure):
    """sgg_mt_tensor: one row of the multi-tensor optimizer table (include/sgg_b200.h)."""
    _fields_ = [('p', C.c_void_p), ('g', C.c_void_p), ('m', C.c_void_p), ('split', C.c_void_p), ('n', C.c_longlong),
                ('lr', C.c_float), ('wd', C.c_float), ('flags', C.c_int), ('reserved', C.c_int)]


# name -> (restype, argtypes); must list every symbol include/sgg_b200.h declares (tests check this).
SIGNATURES = {
    'sgg_abi_version': (C.c_int, []),
    'sgg_last_error': (C.c_char_p, []),
    'sgg_device_info': (C.c_int, [C.POINTER(C.c_int)]),
    'sgg_launch_count': (C.c_ulonglong, []),
    'sgg_graph_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'sgg_graph_build': (C.c_int, [c_i64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                  C.c_void_p]),
    'sgg_graph_check': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'sgg_edge_gru_forward': (C.c_int, [c_f, c_f, c_f, C.c_void_p, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f,
                                       C.c_void_p]),
    'sgg_mp_tape_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    'sgg_mp_backward_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    'sgg_mp_backward': (C.c_int, [c_f, c_f, C.c_void_p, C.POINTER(MpWeights), c_f, C.c_int, C.c_int, C.c_int, C.c_int,
                                  c_f, c_f, C.POINTER(MpGrads), c_f, c_f, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_linear_backward_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'sgg_linear_backward': (C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    'sgg_linear_backward_ex': (C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, C.c_int, C.c_void_p,
                                         C.c_size_t, C.c_void_p]),
    'sgg_bwd_transpose': (C.c_int, [c_f, C.c_longlong, C.c_int, C.c_int, c_f, C.c_int, C.c_int, C.c_void_p]),
    'sgg_tc32_linear_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'sgg_tc32_linear_forward': (C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                          C.c_void_p]),
    'sgg_mp_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    'sgg_mp_forward': (C.c_int, [c_f, c_f, C.c_void_p, C.POINTER(MpWeights), C.c_int, C.c_int, C.c_int, C.c_int,
                                 c_f, c_f, c_f, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_linear_forward': (C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'sgg_tc_set_mode': (C.c_int, [C.c_int]),
    'sgg_tc_get_mode': (C.c_int, []),
    'sgg_tc_debug_timing': (C.c_int, [C.POINTER(C.c_longlong), C.c_int]),
    'sgg_mp_probe_launch': (C.c_int, [C.c_int, c_f, c_f, C.c_void_p, C.POINTER(MpWeights), C.c_int, C.c_int, C.c_int, C.c_int,
                                      c_f, c_f, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_conv_weight_planes': (C.c_int, [c_f, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'sgg_conv3x3_first': (C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'sgg_conv3x3_tc': (C.c_int, [C.c_void_p, C.c_void_p, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, c_f, C.c_void_p]),
    'sgg_conv_overflow': (C.c_int, [C.c_int]),
    'sgg_pow2_scale_workspace_bytes': (C.c_size_t, []),
    'sgg_pow2_scale': (C.c_int, [c_f, C.c_longlong, c_f, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_bwd_transpose16': (C.c_int, [c_f, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, c_f, C.c_void_p]),
    'sgg_scale_by': (C.c_int, [c_f, C.c_longlong, c_f, c_f, C.c_void_p]),
    'sgg_tc16_linear_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'sgg_tc16_linear_scaled': (C.c_int, [c_f, C.c_void_p, c_f, C.c_int, C.c_int, C.c_int, c_f, C.c_void_p, C.c_size_t,
                                         C.c_void_p]),
    'sgg_tc16_linear_pre': (C.c_int, [C.c_void_p, C.c_void_p, c_f, c_f, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      c_f, C.c_void_p]),
    'sgg_tc16_overflow': (C.c_int, [C.c_int]),
    'sgg_bn_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'sgg_bn_train_forward': (C.c_int, [c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, c_f, C.c_float, C.c_float, c_f, c_f,
                                       c_f, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_bn_train_backward': (C.c_int, [c_f, c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, c_f, c_f, c_f, C.c_void_p,
                                        C.c_size_t, C.c_void_p]),
    'sgg_max4_forward': (C.c_int, [c_f, C.c_int, C.c_int, c_f, C.c_void_p, C.c_void_p]),
    'sgg_max4_backward': (C.c_int, [c_f, C.c_void_p, C.c_int, C.c_int, c_f, C.c_void_p]),
    'sgg_bcast_add': (C.c_int, [c_f, c_f, C.c_longlong, C.c_int, c_f, C.c_void_p]),
    'sgg_bcast_add_planes': (C.c_int, [c_f, c_f, C.c_longlong, C.c_int, c_f, C.c_void_p, C.c_void_p]),
    'sgg_relu_backward': (C.c_int, [c_f, c_f, C.c_longlong, c_f, C.c_void_p]),
    'sgg_group_sum': (C.c_int, [c_f, C.c_longlong, C.c_int, c_f, C.c_void_p]),
    'sgg_mpf_debug_timing': (C.c_int, [C.POINTER(C.c_longlong), C.c_int, C.c_int]),
    'sgg_tc_split_weights': (C.c_int, [c_f, C.c_size_t, c_f, C.c_void_p]),
    'sgg_tc_linear_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'sgg_tc_linear_forward': (C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    'sgg_l1_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    'sgg_l1_forward': (C.c_int, [c_f, c_f, C.c_void_p, C.POINTER(HeadWeights), C.POINTER(MpWeights),
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 c_f, c_f, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_l1_forward_rel': (C.c_int, [c_f, c_f, c_i64p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                     C.POINTER(HeadWeights), C.POINTER(MpWeights),
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     c_f, c_f, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_draw_union_boxes': (C.c_int, [c_f, c_i64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_f,
                                       C.c_void_p]),
    'sgg_geom_patches': (C.c_int, [c_f, c_i64p, C.c_int64, C.c_int, C.c_int, C.c_int, c_f, C.c_void_p]),
    'sgg_union_geom_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'sgg_union_geom_forward': (C.c_int, [c_f, c_i64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(GeomWeights), c_f, c_f, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_node_edge_features_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    'sgg_node_edge_features': (C.c_int, [c_f, C.c_int, C.c_int, C.c_int, C.c_int, c_f, C.c_int, c_i64p, C.c_int64,
                                         C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, c_f, c_f,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_node_edge_features_add': (C.c_int, [c_f, C.c_int, C.c_int, C.c_int, C.c_int, c_f, C.c_int, c_i64p, C.c_int64,
                                             C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, c_f, c_f, c_f,
                                             C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_node_edge_features_planes': (C.c_int, [c_f, C.c_int, C.c_int, C.c_int, C.c_int, c_f, C.c_int, c_i64p, C.c_int64,
                                                C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, c_f, c_f, c_f,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_rank_relations_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'sgg_rank_relations': (C.c_int, [c_f, C.c_int, c_f, c_i64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, c_i64p, c_f, c_f, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_rank_relations_check': (C.c_int, [C.c_void_p, C.c_void_p]),
    'sgg_ce_loss_workspace_bytes': (C.c_size_t, [C.c_int]),
    'sgg_ce_loss': (C.c_int, [c_f, c_i64p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                              c_f, c_f, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_mt_chunk_elems': (C.c_int, []),
    'sgg_mt_table_bytes': (C.c_size_t, [C.c_int]),
    'sgg_mt_total_chunks': (C.c_longlong, [C.POINTER(MtTensor), C.c_int]),
    'sgg_mt_workspace_bytes': (C.c_size_t, [C.c_longlong]),
    'sgg_mt_table_upload': (C.c_int, [C.POINTER(MtTensor), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    'sgg_mt_grad_norm': (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_float, c_f, C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    'sgg_mt_grad_norm_scaled': (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_float, c_f, C.c_void_p,
                                          C.c_size_t, C.c_void_p]),
    'sgg_mt_scale_grads': (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, c_f, C.c_void_p]),
    'sgg_mt_sgd_step': (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, c_f, C.c_float, C.c_int, C.c_void_p]),
}


class SggError(RuntimeError):
    pass


def load():
    """Load (building first if possible) the C-ABI library.  Raises loudly if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    from . import _build
    path = _build.ensure_built()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.sgg_abi_version() != 2:
        raise SggError('libsgg_b200.so ABI version mismatch')
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().sgg_last_error().decode(errors='replace')
        raise SggError('%s failed (code %d): %s' % (what, rc, msg))
