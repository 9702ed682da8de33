"""``ImpHeads``: the relation model from the "identical precomputed 4096-d features" boundary on
(sgg_models/rel_model_stanford.py:27-45 heads, :103-107 forward) as a standalone nn.Module with the
reference's parameter names — used for feature-cached training / serving and by the multi-GPU train bench."""
from collections import OrderedDict

import torch.nn as nn

from . import autograd as K


class ImpHeads(nn.Module):
    def __init__(self, hidden_dim=512, obj_dim=4096, num_classes=151, num_rels=51, mp_iter=3):
        super().__init__()
        self.hidden_dim, self.mp_iter = hidden_dim, mp_iter
        self.rel_fc = nn.Linear(hidden_dim, num_rels)
        self.obj_fc = nn.Linear(hidden_dim, num_classes)
        self.obj_unary = nn.Linear(obj_dim, hidden_dim)
        self.edge_unary = nn.Linear(obj_dim, hidden_dim)
        self.edge_gru = nn.GRUCell(hidden_dim, hidden_dim)
        self.node_gru = nn.GRUCell(hidden_dim, hidden_dim)
        gate = lambda: nn.Sequential(nn.Linear(hidden_dim * 2, 1), nn.Sigmoid())
        self.sub_vert_w_fc, self.obj_vert_w_fc = gate(), gate()
        self.out_edge_w_fc, self.in_edge_w_fc = gate(), gate()

    def mp_params(self):
        p = OrderedDict()
        for g in ('edge_gru', 'node_gru'):
            for k in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh'):
                p[g + '.' + k] = getattr(getattr(self, g), k)
        for g in ('sub_vert_w_fc', 'obj_vert_w_fc', 'out_edge_w_fc', 'in_edge_w_fc'):
            p[g + '.0.weight'] = getattr(self, g)[0].weight
            p[g + '.0.bias'] = getattr(self, g)[0].bias
        return p

    def forward(self, obj_feat, edge_feat, rel_inds):
        """obj_feat [N,4096], edge_feat [E,4096], rel_inds int64 [E,2] global (subject, object) -> (obj_dists, rel_dists)."""
        n = K.linear(obj_feat, self.obj_unary.weight, self.obj_unary.bias)
        e = K.linear(edge_feat, self.edge_unary.weight, self.edge_unary.bias, relu=True)
        v, eh = K.message_pass(e, n, rel_inds, self.mp_params(), self.mp_iter)
        return K.linear(v, self.obj_fc.weight, self.obj_fc.bias), K.linear(eh, self.rel_fc.weight, self.rel_fc.bias)
