"""The IMP heads (sgg_models/rel_model_stanford.py:27-45): parameter-owning submodules under the reference's names, shared by
the drop-in ``RelModelStanford`` (sgg_b200/model.py) and by ``ImpHeads`` — the relation model from the "identical precomputed
4096-d features" boundary on (:103-107) as a standalone nn.Module, used for feature-cached training / serving."""
from collections import OrderedDict

import torch.nn as nn

from . import autograd as K

GATES = ('sub_vert_w_fc', 'obj_vert_w_fc', 'out_edge_w_fc', 'in_edge_w_fc')


def add_imp_heads(m, hidden_dim, obj_dim, num_classes, num_rels):
    """Register the heads on module ``m`` in the reference's order (state-dict keys and optimizer groups depend on it)."""
    m.rel_fc = nn.Linear(hidden_dim, num_rels)
    m.obj_fc = nn.Linear(hidden_dim, num_classes)
    m.obj_unary = nn.Linear(obj_dim, hidden_dim)
    m.edge_unary = nn.Linear(obj_dim, hidden_dim)
    m.edge_gru = nn.GRUCell(input_size=hidden_dim, hidden_size=hidden_dim)
    m.node_gru = nn.GRUCell(input_size=hidden_dim, hidden_size=hidden_dim)
    for g in GATES:
        setattr(m, g, nn.Sequential(nn.Linear(hidden_dim * 2, 1), nn.Sigmoid()))


def mp_params(m):
    """state-dict-keyed view of the message-passing parameters (what sgg_b200.autograd.message_pass takes)"""
    p = OrderedDict()
    for g in ('edge_gru', 'node_gru'):
        for k in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh'):
            p[g + '.' + k] = getattr(getattr(m, g), k)
    for g in GATES:
        p[g + '.0.weight'] = getattr(m, g)[0].weight
        p[g + '.0.bias'] = getattr(m, g)[0].bias
    return p


class ImpHeads(nn.Module):
    def __init__(self, hidden_dim=512, obj_dim=4096, num_classes=151, num_rels=51, mp_iter=3):
        super().__init__()
        self.hidden_dim, self.mp_iter = hidden_dim, mp_iter
        add_imp_heads(self, hidden_dim, obj_dim, num_classes, num_rels)

    def mp_params(self):
        return mp_params(self)

    def forward(self, obj_feat, edge_feat, rel_inds):
        """obj_feat [N,4096], edge_feat [E,4096], rel_inds int64 [E,2] global (subject, object) -> (obj_dists, rel_dists)."""
        n = K.linear(obj_feat, self.obj_unary.weight, self.obj_unary.bias)
        e = K.linear(edge_feat, self.edge_unary.weight, self.edge_unary.bias, relu=True)
        v, eh = K.message_pass(e, n, rel_inds, mp_params(self), self.mp_iter)
        return K.linear(v, self.obj_fc.weight, self.obj_fc.bias), K.linear(eh, self.rel_fc.weight, self.rel_fc.bias)
