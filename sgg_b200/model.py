"""Drop-in ``RelModelStanford`` (IMP relation model) for bknyaz/sgg, B200-native.

Mirrors the reference's plugin surface — constructor kwargs, ``forward(batch)``,
``predict``, ``message_pass``, ``node_edge_features``, ``get_rel_inds``,
``get_scaled_boxes``, ``set_box_score_thresh``, attributes ``detector / edge_dim /
pool_sz / fmap_sz / mode / hidden_dim / mp_iter`` and STATE-DICT KEYS
— so ``main.py`` / ``lib/eval.py`` / checkpoints of the reference work against it:

    reference                                   here
    sgg_models/rel_model_base.py:22-122         RelModelBase.__init__
    sgg_models/rel_model_base.py:143-165        RelModelBase.get_rel_inds   (host.get_rel_inds)
    sgg_models/rel_model_base.py:175-242        RelModelBase.faster_rcnn
    sgg_models/rel_model_base.py:245-260        RelModelBase.node_edge_features  -> CUDA RoIAlign w/ on-the-fly union boxes
    lib/get_union_boxes.py:17-101               UnionBoxesAndFeats               -> CUDA geometry branch (no host trip)
    sgg_models/rel_model_stanford.py:48-94      RelModelStanford.message_pass    -> CUDA CSR message passing
    sgg_models/rel_model_stanford.py:97-107     RelModelStanford.predict
    sgg_models/rel_model_stanford.py:110-207    RelModelStanford.forward

The nn.Module children exist to own the parameters under the reference's names
(``wandb.watch`` / ``get_optim`` / ``load_checkpoint`` iterate them); the arithmetic of
every stage after the frozen detector backbone runs in ``libsgg_b200.so`` through
``sgg_b200.autograd`` — there is no PyTorch/CPU fallback for those stages.
The detector (image transform + VGG16 conv stack, frozen) is
torchvision library code exactly as in the reference.
"""
import copy
import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import heads, host, ops
from . import autograd as K
from .host import Result, IM_SCALE, BATCHNORM_MOMENTUM


def _vgg16_parts():
    """torchvision VGG16 split the way rel_model_base.py:310-321 ``load_vgg`` does: conv stack without the
    last max-pool, classifier without the class layer."""
    import torchvision
    vgg = torchvision.models.vgg16(weights=None)
    feats = nn.Sequential(*list(vgg.features.children())[:-1])
    feats.out_channels = 512
    cls = list(vgg.classifier.children())       # Linear ReLU Dropout Linear ReLU Dropout Linear
    return feats, cls


class UnionBoxesAndFeats(nn.Module):
    """lib/get_union_boxes.py:17-101, edge_model='motifs'.  ``conv`` holds the parameters (keys
    union_boxes.conv.{0,2,4,6}.*); both convs are built with stride 16 — the reference's lambda captures the
    constructor's ``stride`` (:40-43) — which is what makes the branch collapse to one [E, dim] vector."""

    def __init__(self, edge_model='motifs', pooling_size=7, stride=16, dim=256, concat=False, use_feats=True):
        super().__init__()
        if edge_model not in ('motifs',):
            raise NotImplementedError(edge_model)   # 'raw_boxes' (grid_sample variant) is not on the IMP default path
        self.edge_model, self.pooling_size, self.stride, self.dim = edge_model, pooling_size, stride, dim
        self.concat, self.use_feats = concat, use_feats
        self.conv = nn.Sequential(
            nn.Conv2d(2, dim // 2, kernel_size=7, stride=stride, padding=3, bias=True),
            nn.ReLU(inplace=True),
            nn.BatchNorm2d(dim // 2, momentum=BATCHNORM_MOMENTUM),
            nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            nn.Conv2d(dim // 2, dim, kernel_size=3, stride=stride, padding=1, bias=True),
            nn.ReLU(inplace=True),
            nn.BatchNorm2d(dim, momentum=BATCHNORM_MOMENTUM))

    def geometry(self, rois, union_inds):
        """[E, dim] geometry embedding (the conv branch applied to draw_union_boxes(...) - 0.5)."""
        return K.union_geom(rois, union_inds, self.conv, self.training)

    def forward(self, union_pools, rois, union_inds, im_sizes=None):
        if not self.training and not self.concat:      # fused: geometry + broadcast add in one CUDA pipeline
            from . import ops
            return ops.union_geom(rois.detach(), union_inds, K._conv_params(self.conv), union_pools)
        geom = self.geometry(rois, union_inds)
        if self.concat:
            return torch.cat((union_pools, geom[:, :, None, None].expand(-1, -1, *union_pools.shape[2:])), 1)
        return K.broadcast_add(union_pools, geom)


class RelModelBase(nn.Module):
    def __init__(self, train_data, mode='sgcls', require_overlap_det=True, use_bias=False, test_bias=False,
                 backbone='vgg16', RELS_PER_IMG=1024, min_size=None, max_size=None, edge_model='motifs'):
        super().__init__()
        self.classes = train_data.ind_to_classes
        self.rel_classes = train_data.ind_to_predicates
        self.mode, self.backbone, self.RELS_PER_IMG = mode, backbone, RELS_PER_IMG
        self.pool_sz, self.stride = 7, 16
        self.dropout_p = 0.5            # torchvision's VGG classifier dropout (roi_fmap / roi_fmap_obj); tests set it to 0
        self.use_bias, self.test_bias = use_bias, test_bias
        self.require_overlap = require_overlap_det and self.mode == 'sgdet'
        if backbone != 'vgg16':
            # resnet50 needs COCO weights from the network; vgg16_old is removed in the reference too (:126-127)
            raise NotImplementedError(backbone)
        from torchvision.models.detection import FasterRCNN
        from torchvision.models.detection.faster_rcnn import TwoMLPHead, FastRCNNPredictor
        from torchvision.models.detection.rpn import AnchorGenerator
        from torchvision.ops import MultiScaleRoIAlign
        self.obj_dim, self.fmap_sz = 4096, 38
        min_size = IM_SCALE if min_size is None else min_size
        max_size = IM_SCALE if max_size is None else max_size
        feats, cls_edge = _vgg16_parts()
        _, cls_node = _vgg16_parts()
        self.detector = FasterRCNN(
            feats, min_size=min_size, max_size=max_size,
            rpn_anchor_generator=AnchorGenerator(sizes=((32, 64, 128, 256, 512),), aspect_ratios=((0.5, 1.0, 2.0),)),
            box_head=TwoMLPHead(512 * self.pool_sz ** 2, self.obj_dim),
            box_predictor=FastRCNNPredictor(self.obj_dim, len(self.classes)),
            box_roi_pool=MultiScaleRoIAlign(featmap_names=['0'], output_size=self.pool_sz, sampling_ratio=2),
            box_detections_per_img=50, box_score_thresh=0.2)
        # edge head: fc6, ReLU, Dropout, fc7 (no final ReLU); node head: fc6, ReLU, Dropout, fc7, ReLU, Dropout
        self.roi_fmap = nn.Sequential(nn.Flatten(), nn.Sequential(*cls_edge[:4]))
        self.roi_fmap_obj = nn.Sequential(*cls_node[:6])
        self.roi_pool = copy.deepcopy(self.detector.roi_heads.box_roi_pool)   # parameter-free; kept for API parity
        self.edge_dim = self.detector.backbone.out_channels
        self.union_boxes = UnionBoxesAndFeats(pooling_size=self.pool_sz, stride=self.stride, dim=self.edge_dim,
                                              edge_model=edge_model)
        if self.use_bias:
            fg, bg = host.dataset_counts(train_data, must_overlap=True)       # lib/sparse_targets.py:16
            self.freq_bias = host.FrequencyBias(fg, bg)

    num_classes = property(lambda self: len(self.classes))
    num_rels = property(lambda self: len(self.rel_classes))

    def predict(self, node_feat, edge_feat, rel_inds, rois, im_sizes):
        raise NotImplementedError('predict')

    def forward(self, batch):
        raise NotImplementedError('forward')

    def get_rel_inds(self, rel_labels, im_inds, box_priors):
        return host.get_rel_inds(im_inds.detach(), rel_labels, self.training, box_priors, self.require_overlap)

    def set_box_score_thresh(self, box_score_thresh):
        self.detector.roi_heads.score_thresh = box_score_thresh

    def _device(self):
        return self.rel_fc.weight.device

    def faster_rcnn(self, x, gt_boxes, gt_classes, gt_rels):
        """Image transform + backbone (+ RPN / RoI heads in SGDet) — rel_model_base.py:175-242."""
        dev = self._device()
        segs = host.image_segments(gt_classes[:, 0])
        targets, imgs, org_sizes = [], [], []
        for i, s, e in segs:
            targets.append({'boxes': gt_boxes[s:e].detach().clone(), 'labels': gt_classes[s:e, 1].long()})
            imgs.append(x[i].to(dev).squeeze())
            org_sizes.append(tuple(x[i].shape[-2:]))
        images, targets = self.detector.transform(imgs, targets)
        fmaps = self._backbone(images.tensors)
        if isinstance(fmaps, torch.Tensor):
            fmaps = OrderedDict([('0', fmaps)])
        if self.mode != 'sgdet':
            rois, obj_labels, rel_labels = self.gt_labels(gt_boxes, gt_classes, gt_rels)
            result = Result(od_obj_labels=obj_labels, rm_obj_labels=obj_labels,
                            rm_box_priors=torch.cat([t['boxes'] for t in targets]),
                            rel_labels=rel_labels, im_inds=rois[:, 0].long())
            result.rm_box_priors_org = gt_boxes
        else:
            proposals, _ = self.detector.rpn(images, fmaps, targets)
            dets, _ = self.detector.roi_heads(fmaps, proposals, images.image_sizes, targets)
            kept = copy.deepcopy(dets)
            dets_org = self.detector.transform.postprocess(dets, images.image_sizes, org_sizes)
            for d in kept:
                if len(d['boxes']) <= 1:
                    raise ValueError('at least two objects must be detected to build relationships, make sure '
                                     'the detector is properly pretrained', kept)
            im_inds = torch.cat([torch.full((len(d['boxes']),), i, dtype=torch.long) for i, d in enumerate(kept)])
            result = Result(rm_obj_labels=torch.cat([d['labels'] for d in dets_org]).view(-1),
                            rm_box_priors=torch.cat([d['boxes'] for d in kept]), im_inds=im_inds.to(dev))
            result.rel_labels = None
            result.rm_box_priors_org = torch.cat([d['boxes'] for d in dets_org])
            if len(result.rm_box_priors) <= 1:
                raise ValueError('at least two objects must be detected to build relationships')
        result.im_sizes_org = org_sizes
        result.im_sizes = images.image_sizes
        result.fmap = fmaps[list(fmaps.keys())[-1]]
        result.rois = torch.cat((result.im_inds.float()[:, None], result.rm_box_priors), 1)
        return result

    def _backbone(self, x):
        """The frozen conv stack (rel_model_base.py:184).  With the 3xFP16 tensor-core engine active, the VGG16 stack
        runs as tcgen05 implicit GEMMs (csrc/conv_tc.cu, fp32-grade results); otherwise torchvision / cuDNN in fp32
        (TF32 off: the 1e-4 parity bar).  SGG_BACKBONE=cudnn forces the library path."""
        import os
        bb = self.detector.backbone
        if (x.is_cuda and os.environ.get('SGG_BACKBONE', 'tc') != 'cudnn' and isinstance(bb, nn.Sequential)
                and ops._use_tc() and ops.tc_engine() == 'tc16'):
            layers = getattr(self, '_vgg_layers', None)
            if layers is None:
                layers = self._vgg_layers = ops.vgg_layers(bb) or False
            if layers and x.shape[2] % 16 == 0 and x.shape[3] % 16 == 0:
                return ops.vgg_features(x, layers)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):     # fp32 parity with the reference (1e-4)
            return bb(x)

    def _spatial_scale(self, fmap, im_sizes):
        """MultiScaleRoIAlign's scale inference: 2 ** round(log2(feature size / padded-input size))."""
        if im_sizes is None:
            return 1.0 / self.stride
        mh = max(int(s[0]) for s in im_sizes); mw = max(int(s[1]) for s in im_sizes)
        scales = [2.0 ** round(math.log2(float(fs) / float(os_))) for fs, os_ in ((fmap.shape[-2], mh), (fmap.shape[-1], mw))]
        return scales[0]

    def node_edge_features(self, fmap, rois, union_inds, im_sizes, edge_add=None, planes=False):
        """rel_model_base.py:245-260: 7x7 RoIAlign features of the objects and of the union box of every pair.
        ``edge_add`` [E,C] (internal, eval forward only): geometry embedding folded into the edge rows.
        ``planes`` (internal, eval forward only): also return the fp16 operand planes of both outputs for the fc6 layers."""
        assert union_inds.shape[1] == 2, union_inds.shape
        if isinstance(fmap, dict):
            fmap = fmap['0']
        return K.node_edge_features(fmap, rois, union_inds, self._spatial_scale(fmap, im_sizes), self.pool_sz, 2,
                                    edge_add=edge_add, planes=planes)

    def get_scaled_boxes(self, boxes, im_inds, im_sizes):
        """rel_model_base.py:263-274: boxes / (w, h, w, h) of their image."""
        wh = torch.tensor([[s[1], s[0], s[1], s[0]] for s in im_sizes], dtype=boxes.dtype, device=boxes.device)
        scaled = boxes / wh[im_inds.long()]
        assert scaled.max() <= 1 + 1e-3, (scaled.max(), boxes.max(), im_sizes)
        return scaled

    def gt_labels(self, gt_boxes, gt_classes, gt_rels=None, sample_factor=-1):
        """rel_model_base.py:277-300."""
        assert gt_boxes is not None
        rois = torch.cat((gt_classes[:, 0].float()[:, None], gt_boxes), 1)
        if gt_rels is not None and self.training:
            return host.proposal_assignments_gtbox(rois.detach(), gt_boxes.detach(), gt_classes.detach(),
                                                   gt_rels.detach(), 0, self.RELS_PER_IMG, sample_factor=sample_factor)
        return rois, gt_classes[:, 1], None


class RelModelStanford(RelModelBase):
    """Iterative Message Passing (Xu et al. 2017) relation model — sgg_models/rel_model_stanford.py."""

    def __init__(self, train_data, hidden_dim=512, mp_iter=3, **kwargs):
        super().__init__(train_data, **kwargs)
        self.hidden_dim, self.mp_iter = hidden_dim, mp_iter
        heads.add_imp_heads(self, hidden_dim, self.obj_dim, self.num_classes, self.num_rels)     # :27-45, reference order

    def _mp_params(self):
        return heads.mp_params(self)

    def message_pass(self, rel_rep, obj_rep, rel_inds):
        """rel_rep [E,H], obj_rep [N,H], rel_inds [E,2] global (subject, object) ids -> (V_T, E_T)."""
        return K.message_pass(rel_rep, obj_rep, rel_inds, self._mp_params(), self.mp_iter)

    def predict(self, node_feat, edge_feat, rel_inds, rois, im_sizes):
        """rel_model_stanford.py:97-107."""
        E = edge_feat.shape[0]
        ub = self.union_boxes
        pools = edge_feat.view(E, -1, self.pool_sz, self.pool_sz)
        if (self.training and isinstance(ub, UnionBoxesAndFeats) and not ub.concat and ub.use_feats
                and not (pools.requires_grad and torch.is_grad_enabled())):
            # training: pools + broadcast(geom) feeds fc6 directly; the fused op's backward sends only the 7x7-summed
            # gradient to the geometry branch (49x cheaper than materialising dX of fc6, autograd.fc_broadcast)
            geom = ub.geometry(rois, rel_inds[:, 1:])
            return self._predict_pooled(node_feat, pools, rel_inds, geom=geom)
        edge_feat = ub(pools, rois, rel_inds[:, 1:], im_sizes)
        return self._predict_pooled(node_feat, edge_feat, rel_inds)

    def _predict_pooled(self, node_feat, edge_feat, rel_inds, geom=None, planes=None):
        """rel_model_stanford.py:100-107: everything after ``self.union_boxes`` (edge_feat already carries the
        union-box geometry, or ``geom`` [E,C] is still to be broadcast-added to it).  ``planes`` = (node_planes,
        edge_planes) from ``node_edge_features(planes=True)`` (eval): fc6 and fc7 of both heads then run on pre-split
        operands, each layer's epilogue emitting the planes of the next."""
        fo, fe = self.roi_fmap_obj, self.roi_fmap[1]
        drop = self.training
        npl, epl = planes if (planes is not None and not drop) else (None, None)
        # eval with planes: the fp32 RoIAlign rows may not exist at all (node_edge_features(planes='only'))
        n, npl = K.linear(None if node_feat is None else node_feat.reshape(node_feat.shape[0], -1), fo[0].weight,
                          fo[0].bias, relu=True, x_planes=npl, out_planes=True)
        n = F.dropout(n, self.dropout_p, drop)
        n = K.linear(n, fo[3].weight, fo[3].bias, relu=True, x_planes=npl)
        n = F.dropout(n, self.dropout_p, drop)
        if geom is not None:
            e = K.fc_broadcast(edge_feat, geom, fe[0].weight, fe[0].bias)
            epl = None
        else:
            e, epl = K.linear(None if edge_feat is None else edge_feat.reshape(edge_feat.shape[0], -1), fe[0].weight,
                              fe[0].bias, relu=True, x_planes=epl, out_planes=True)
        e = F.dropout(e, self.dropout_p, drop)
        e = K.linear(e, fe[3].weight, fe[3].bias, relu=False, x_planes=None if drop else epl)
        n = K.linear(n, self.obj_unary.weight, self.obj_unary.bias)
        e = K.linear(e, self.edge_unary.weight, self.edge_unary.bias, relu=True)
        v, eh = self.message_pass(e, n, rel_inds[:, 1:3])
        return (K.linear(v, self.obj_fc.weight, self.obj_fc.bias), K.linear(eh, self.rel_fc.weight, self.rel_fc.bias))

    def forward(self, batch):
        """batch: indexable, len 1; batch[0] = (imgs, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels,
        proposals, [None,] fns) (dataloaders/blob.py:244-249).  train -> Result; eval -> filter_dets 5-tuple."""
        assert len(batch) == 1, ('single GPU is only supported in this code', len(batch))
        x, gt_boxes, gt_classes, gt_rels = batch[0][0], batch[0][3], batch[0][4], batch[0][5]
        dev = self._device()
        gt_boxes, gt_classes = gt_boxes.to(dev), gt_classes.to(dev)
        gt_rels = gt_rels.to(dev) if gt_rels is not None else None
        with torch.no_grad():
            result = self.faster_rcnn(x, gt_boxes, gt_classes, gt_rels)
        result.fmap = result.fmap.detach()
        im_inds, boxes = result.im_inds, result.rm_box_priors
        if self.training and getattr(result, 'rel_labels', None) is None:
            assert self.mode == 'sgdet'
            result.rel_labels = host.rel_assignments(im_inds, boxes, result.rm_obj_labels, gt_boxes, gt_classes,
                                                     gt_rels, 0, filter_non_overlap=True, num_sample_per_gt=1)
        elif not hasattr(result, 'rel_labels'):
            result.rel_labels = None
        rel_inds = self.get_rel_inds(result.rel_labels if self.training else None, im_inds, boxes)
        result.rel_inds = rel_inds
        rois = torch.cat((im_inds[:, None].float(), boxes), 1)
        ub = self.union_boxes
        if not self.training and isinstance(ub, UnionBoxesAndFeats) and not ub.concat and ub.use_feats:
            # eval: nobody reads the raw union-box features (only the train-mode Result exposes them), so the geometry
            # embedding [E,C] is computed first and added inside the RoIAlign kernel: the [E,C,7,7] tensor is written
            # once instead of written, re-read and re-written (lib/get_union_boxes.py:101)
            geom = ops.union_geom(rois, rel_inds[:, 1:], K._conv_params(ub.conv))
            # (planes-only activations cannot enter an autograd graph: without torch.no_grad() the fp32 path is kept)
            use_planes = (ops._use_tc() and ops.tc_engine() == 'tc16' and result.fmap.shape[1] % 4 == 0
                          and not torch.is_grad_enabled())
            # with planes, the 963 MB of fp32 union-box rows are not written either: only the fc6 layers read them
            feats = self.node_edge_features(result.fmap, rois, rel_inds[:, 1:], im_sizes=result.im_sizes, edge_add=geom,
                                            planes='only' if use_planes else False)
            result.node_feat, result.edge_feat = feats[0], feats[1]
            result.rm_obj_dists, result.rel_dists = self._predict_pooled(result.node_feat, result.edge_feat, rel_inds,
                                                                         planes=feats[2:] if use_planes else None)
        else:
            result.node_feat, result.edge_feat = self.node_edge_features(result.fmap, rois, rel_inds[:, 1:],
                                                                         im_sizes=result.im_sizes)
            result.rm_obj_dists, result.rel_dists = self.predict(result.node_feat, result.edge_feat, rel_inds,
                                                                 rois=rois, im_sizes=result.im_sizes)
        if self.use_bias:
            if self.mode == 'predcls':
                result.obj_preds = gt_classes[:, 1]
            else:
                result.obj_preds = F.softmax(result.rm_obj_dists, dim=1)[:, 1:].argmax(1) + 1
            freq = self.freq_bias.index_with_labels(torch.stack((result.obj_preds[rel_inds[:, 1]],
                                                                 result.obj_preds[rel_inds[:, 2]]), 1))
            result.rel_dists = freq if self.test_bias else result.rel_dists + freq
        if self.training:
            result.rois = rois
            return result
        if self.mode == 'predcls':
            result.obj_scores = result.rm_obj_dists.new_ones(gt_classes.shape[0])
            result.obj_preds = gt_classes[:, 1]
        elif self.mode in ('sgcls', 'sgdet'):
            sc, idx = F.softmax(result.rm_obj_dists.detach(), dim=1)[:, 1:].max(1)
            result.obj_scores, result.obj_preds = sc, idx + 1
        else:
            raise NotImplementedError(self.mode)
        # rel_model_stanford.py:206-207: softmax + filter_dets; the softmax is fused into the ranking kernel
        return host.filter_dets(result.rm_box_priors_org, result.obj_scores, result.obj_preds, rel_inds[:, 1:],
                                result.rel_dists.detach(), logits=True)
