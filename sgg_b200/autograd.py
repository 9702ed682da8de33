"""torch.autograd wrappers: forward AND backward of every stage run in libsgg_b200.so.

The model (``sgg_b200.model``) composes these exactly where the reference composes
nn.Linear / nn.GRUCell / index ops, so ``loss.backward()`` in the reference's
``main.py:116-120`` works unchanged.  Contract (SURVEY.md §8b "Autograd"): gradients for all
trainable tensors and for the ``node_feat`` / ``edge_feat`` inputs of ``predict``; ``fmap`` and
the detector never need gradients (main.py:62-63, rel_model_stanford.py:125-131).
"""
import torch
import torch.nn.functional as F

from . import ops
from .ops import MP_KEYS, GATE_KEYS

_MP_PARAM_KEYS = tuple(MP_KEYS) + tuple(k + s for k in GATE_KEYS for s in ('.0.weight', '.0.bias'))


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        y = ops.linear(x, weight, bias, relu=relu)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.relu:
            dy = ops.relu_backward(dy, y)           # elementwise ReLU mask (CUDA)
        dx, dw, db = ops.linear_backward(x, weight, dy, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                         ctx.has_bias and ctx.needs_input_grad[2])
        return dx, dw, db, None


def linear(x, weight, bias=None, relu=False, x_planes=None, out_planes=False):
    """nn.Linear (+ReLU) with CUDA forward/backward.  Inference (no grad) skips the autograd graph; there ``x_planes`` /
    ``out_planes`` (ops.linear) route the GEMM to the pre-split tcgen05 kernel.  With gradients the planes are ignored
    (``out_planes`` then returns ``(y, None)``)."""
    if x is None:                  # planes-only activations (eval): there is nothing to differentiate through
        if torch.is_grad_enabled() and (weight.requires_grad or (bias is not None and bias.requires_grad)):
            raise NotImplementedError('linear: planes-only input inside an autograd region')
        return ops.linear(None, weight, bias, relu=relu, x_planes=x_planes, out_planes=out_planes)
    x = x.contiguous()
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad)):
        y = _LinearFn.apply(x, weight, bias, relu)
        return (y, None) if out_planes else y
    return ops.linear(x, weight, bias, relu=relu, x_planes=x_planes, out_planes=out_planes)


class _MessagePassFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rel_rep, obj_rep, rel_inds, mp_iter, *plist):
        params = dict(zip(_MP_PARAM_KEYS, plist))
        graph = ops.build_graph(rel_inds, obj_rep.shape[0])
        v, e, tape = ops.message_pass_train(rel_rep, obj_rep, graph, params, mp_iter)
        ctx.graph, ctx.mp_iter, ctx.tape = graph, mp_iter, tape
        ctx.save_for_backward(rel_rep, obj_rep, *plist)
        return v, e

    @staticmethod
    def backward(ctx, dv, de):
        rel_rep, obj_rep = ctx.saved_tensors[:2]
        params = dict(zip(_MP_PARAM_KEYS, ctx.saved_tensors[2:]))
        dv = torch.zeros_like(obj_rep) if dv is None else dv.contiguous()
        de = torch.zeros_like(rel_rep) if de is None else de.contiguous()
        d_rel, d_obj, grads = ops.message_pass_backward(rel_rep, obj_rep, ctx.graph, params, ctx.tape, dv, de, ctx.mp_iter)
        ctx.tape = None
        return (d_rel, d_obj, None, None) + tuple(grads[k] for k in _MP_PARAM_KEYS)


def message_pass(rel_rep, obj_rep, rel_inds, params, mp_iter=3):
    """RelModelStanford.message_pass (rel_model_stanford.py:48-94): rel_inds int64 [E,2] global ids."""
    rel_rep, obj_rep = rel_rep.contiguous(), obj_rep.contiguous()
    plist = [params[k] for k in _MP_PARAM_KEYS]
    need = torch.is_grad_enabled() and (rel_rep.requires_grad or obj_rep.requires_grad or any(p.requires_grad for p in plist))
    if need:
        return _MessagePassFn.apply(rel_rep, obj_rep, rel_inds, mp_iter, *plist)
    graph = ops.build_graph(rel_inds, obj_rep.shape[0])
    return ops.message_pass(rel_rep, obj_rep, graph, params, mp_iter)


def node_edge_features(fmap, rois, union_inds, spatial_scale, pool=7, sampling_ratio=2, edge_add=None, planes=False):
    """RoIAlign of objects and union boxes.  ``fmap`` is produced under no_grad and detached by the caller
    (rel_model_stanford.py:125-131), so no backward is defined; a differentiable fmap (GAN ``-attachG``
    path, out of scope) is rejected loudly rather than silently dropping its gradient."""
    if torch.is_grad_enabled() and fmap.requires_grad:
        raise NotImplementedError('node_edge_features: gradient w.r.t. fmap (GAN -attachG path) is not implemented')
    return ops.node_edge_features(fmap.detach(), rois.detach(), union_inds, spatial_scale, pool, sampling_ratio,
                                  edge_add=edge_add, planes=planes)


def _conv_params(conv):
    p = {}
    for idx in ('0', '4'):
        p['union_boxes.conv.%s.weight' % idx] = getattr(conv, idx).weight
        p['union_boxes.conv.%s.bias' % idx] = getattr(conv, idx).bias
    for idx in ('2', '6'):
        bn = getattr(conv, idx)
        p['union_boxes.conv.%s.weight' % idx] = bn.weight; p['union_boxes.conv.%s.bias' % idx] = bn.bias
        p['union_boxes.conv.%s.running_mean' % idx] = bn.running_mean
        p['union_boxes.conv.%s.running_var' % idx] = bn.running_var
    return p


class _ReluBnTrainFn(torch.autograd.Function):
    """BatchNorm(relu(x)) in training mode (batch statistics over the rows, running-stat update), CUDA fwd + bwd."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps):
        y, mean, invstd = ops.bn_train_forward(x, gamma, beta, running_mean, running_var, momentum, eps, relu_in=True)
        ctx.save_for_backward(x, gamma, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, invstd = ctx.saved_tensors
        dx, dgamma, dbeta = ops.bn_train_backward(x, dy.contiguous(), gamma, mean, invstd, relu_in=True)
        return dx, dgamma, dbeta, None, None, None, None


class _Max4Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y, idx = ops.max4_forward(x)
        ctx.save_for_backward(idx)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        return ops.max4_backward(dy.contiguous(), idx)


def relu_bn_train(x, bn):
    y = _ReluBnTrainFn.apply(x.contiguous(), bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps)
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return y


def union_geom(rois, union_inds, conv, training):
    """Geometry embedding [E, dim] of UnionBoxesAndFeats (lib/get_union_boxes.py:51-59,66-67).
    eval: one fused CUDA pipeline straight from the boxes (running BN statistics).
    train: the conv windows are materialised by a CUDA kernel ([4E, 98] patches of the 27x27 masks), the two
    convolutions run through the CUDA linear op (fwd+bwd), ReLU + BatchNorm with batch statistics (running-stat
    update, fwd+bwd) and the 2x2 max-pool are CUDA kernels too (csrc/bn.cu)."""
    rois = rois.detach()
    if not training:
        return ops.union_geom(rois, union_inds, _conv_params(conv))
    c0, bn1, c1, bn2 = conv[0], conv[2], conv[4], conv[6]
    patches = ops.geom_patches(rois, union_inds)                                  # [E,4,98]
    E = patches.shape[0]
    x = linear(patches.view(E * 4, 98), c0.weight.view(c0.weight.shape[0], -1), c0.bias)      # ReLU folded into the BN op
    x = relu_bn_train(x, bn1)
    x = _Max4Fn.apply(x.view(E, 4, -1))                                           # MaxPool2d(3,2,1) over the 2x2 map
    x = linear(x, c1.weight[:, :, 1, 1].contiguous(), c1.bias)                    # only the centre tap sees data
    return relu_bn_train(x, bn2)


class _BcastAddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pools, geom):
        ctx.S = pools.shape[2] * pools.shape[3]
        return ops.bcast_add(pools, geom)

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()
        dgeom = ops.group_sum(d.view(d.shape[0], d.shape[1], ctx.S), ctx.S) if ctx.needs_input_grad[1] else None
        return (d if ctx.needs_input_grad[0] else None), dgeom


def broadcast_add(union_pools, geom):
    """union_pools [E,C,7,7] + geom [E,C] (lib/get_union_boxes.py:101), CUDA forward and backward."""
    return _BcastAddFn.apply(union_pools.contiguous(), geom.contiguous())


class _FcBroadcastFn(torch.autograd.Function):
    """relu(fc6(pools + geom broadcast over the 7x7 positions)) — rel_model_stanford.py:100-101 in training mode, where
    only ``geom`` (the union-box geometry embedding) needs a gradient on the input side.  The generic backward would form
    dX [E, C*49] = dY W (a 2 TFLOP GEMM at 9600 edges) just to sum it over the 49 positions; since the sum commutes,
    dgeom = dY (sum_p W[:, c*49 + p]) is a [E,4096] x [4096,C] GEMM, 49x cheaper, exact up to fp32 reassociation."""

    @staticmethod
    def forward(ctx, pools, geom, weight, bias):
        E = pools.shape[0]
        if ops._use_tc() and ops.tc_engine() == 'tc16' and pools.numel() % 4 == 0 and weight.shape[0] % 4 == 0:
            # the add also emits the fp16 operand planes: the forward GEMM runs on the pre-split kernel (no in-kernel
            # conversion of the 9600 x 25088 activations); the fp32 x stays for the weight gradient
            x, xpl = ops.bcast_add(pools, geom, planes=True)
            x = x.view(E, -1)
            y = ops.linear(x, weight, bias, relu=True, x_planes=xpl)
            del xpl
        else:
            x = ops.bcast_add(pools, geom).view(E, -1)
            y = ops.linear(x, weight, bias, relu=True)
        ctx.S = pools.shape[2] * pools.shape[3]
        ctx.save_for_backward(x, weight, y)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        dym = ops.relu_backward(dy.contiguous(), y)
        _, dw, db = ops.linear_backward(x, weight, dym, False, ctx.needs_input_grad[2], ctx.has_bias and ctx.needs_input_grad[3])
        dgeom = None
        if ctx.needs_input_grad[1]:
            Nout, K = weight.shape
            Cc = K // ctx.S
            wsum = ops.group_sum(weight.detach().view(Nout, Cc, ctx.S), ctx.S)       # [Nout, C]
            dgeom = ops.matmul_nn(dym, wsum)                                         # [E, C] = dym [E,Nout] @ wsum [Nout,C]
        return None, dgeom, dw, db


def fc_broadcast(pools, geom, weight, bias):
    """relu(linear(pools + geom[:, :, None, None], weight, bias)); ``pools`` must not require a gradient."""
    if pools.requires_grad and torch.is_grad_enabled():
        return linear(broadcast_add(pools, geom).reshape(pools.shape[0], -1), weight, bias, relu=True)
    return _FcBroadcastFn.apply(pools.contiguous(), geom.contiguous(), weight, bias)
