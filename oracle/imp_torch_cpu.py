"""CPU baseline port (torch, all host threads) of the reference's IMP path.

TEST / BENCH INFRASTRUCTURE ONLY — used by ``bench.py`` as the ``cpu_baseline``
and ``--impl reference`` arm (the real reference is Python and cannot travel to
the GPU box; /root/reference does not exist there) and pinned against the golden
vectors in tests/test_oracle_golden.py.  Never imported by the product.

Unlike ``imp_numpy`` (a from-the-maths restatement used as the checker), this
port keeps the reference's *operation sequence* so that its timing is a fair
stand-in for the reference's PyTorch-CPU path: nn.GRUCell / nn.Linear modules,
the two dense [N,E] incidence matrices and their matmuls
(sgg_models/rel_model_stanford.py:58-66,91), the cat + Linear(2H,1) + Sigmoid
gates (:78-81,86-89).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class ImpCpu(nn.Module):
    """Heads of RelModelStanford.__init__ (rel_model_stanford.py:27-45)."""

    def __init__(self, H=512, D=4096, n_cls=151, n_rel=51, mp_iter=3):
        super().__init__()
        self.H, self.mp_iter = H, mp_iter
        self.rel_fc = nn.Linear(H, n_rel)
        self.obj_fc = nn.Linear(H, n_cls)
        self.obj_unary = nn.Linear(D, H)
        self.edge_unary = nn.Linear(D, H)
        self.edge_gru = nn.GRUCell(H, H)
        self.node_gru = nn.GRUCell(H, H)
        mk = lambda: nn.Sequential(nn.Linear(2 * H, 1), nn.Sigmoid())
        self.sub_vert_w_fc, self.obj_vert_w_fc = mk(), mk()
        self.out_edge_w_fc, self.in_edge_w_fc = mk(), mk()

    def load_numpy(self, p):
        sd = self.state_dict()
        for k, v in p.items():
            if k in sd:
                sd[k].copy_(torch.from_numpy(v))
        return self

    def message_pass(self, rel_rep, obj_rep, rel_inds):
        """rel_model_stanford.py:48-94 (dense incidence formulation, as the reference runs it)."""
        E, N = rel_rep.shape[0], obj_rep.shape[0]
        ar = torch.arange(E)
        a_out = rel_rep.new_zeros(N, E); a_out.view(-1)[rel_inds[:, 0] * E + ar] = 1
        a_in = rel_rep.new_zeros(N, E); a_in.view(-1)[rel_inds[:, 1] * E + ar] = 1
        v = self.node_gru(obj_rep, obj_rep.new_zeros(N, self.H))
        e = self.edge_gru(rel_rep, rel_rep.new_zeros(E, self.H))
        for _ in range(self.mp_iter):
            sv, ov = v[rel_inds[:, 0]], v[rel_inds[:, 1]]
            ws = self.sub_vert_w_fc(torch.cat((sv, e), 1)) * sv
            wo = self.obj_vert_w_fc(torch.cat((ov, e), 1)) * ov
            e_new = self.edge_gru(ws + wo, e)
            po = self.out_edge_w_fc(torch.cat((sv, e), 1)) * e
            pi = self.in_edge_w_fc(torch.cat((ov, e), 1)) * e
            v = self.node_gru(a_out @ po + a_in @ pi, v)
            e = e_new
        return v, e

    def l1_forward(self, obj_feat, edge_feat, rel_inds):
        """rel_model_stanford.py:103-107 without roi_fmap* (the 4096-d feature boundary)."""
        nf = self.obj_unary(obj_feat)
        ef = F.relu(self.edge_unary(edge_feat))
        v, e = self.message_pass(ef, nf, rel_inds)
        return self.obj_fc(v), self.rel_fc(e)
