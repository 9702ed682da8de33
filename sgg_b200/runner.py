"""Host-buffer front end of the L1 hot path (4096-d features -> obj_dists / rel_dists).

This is the call a user of the reference makes at the "identical precomputed
features" boundary (sgg_models/rel_model_stanford.py:103-107): inputs live in
HOST memory (numpy / pinned torch), outputs come back to the host.  Each
submission does H2D copies of its inputs, one fused L1 forward (graph index built inside it)
and a D2H copy of the logits, on a small ring of in-flight slots so that the
copies of step k+1 overlap the kernels of step k (copy engine + SMs).
"""
import ctypes as C
import numpy as np
import torch

from . import _lib, ops


class _Slot(object):
    def __init__(self, N, E, D, n_cls, n_rel, device):
        pin = dict(dtype=torch.float32, pin_memory=True)
        self.h_obj = torch.empty((N, D), **pin); self.h_edge = torch.empty((E, D), **pin)
        self.h_rel = torch.empty((E, 2), dtype=torch.int64, pin_memory=True)
        self.h_od = torch.empty((N, n_cls), **pin); self.h_rd = torch.empty((E, n_rel), **pin)
        self.d_obj = torch.empty((N, D), dtype=torch.float32, device=device)
        self.d_edge = torch.empty((E, D), dtype=torch.float32, device=device)
        self.d_rel = torch.empty((E, 2), dtype=torch.int64, device=device)
        self.stream = torch.cuda.Stream(device=device)
        self.done = torch.cuda.Event()
        self.busy = False


class ImpL1Runner(object):
    """params: state-dict-keyed tensors/ndarrays (reference names).  Fixed (N, E) per runner
    (ragged batches: one runner per shape bucket, or use sgg_b200.ops directly)."""

    def __init__(self, params, N, E, mp_iter=3, slots=2, device='cuda'):
        self.device = torch.device(device)
        self.params = {k: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v).to(self.device)
                       for k, v in params.items()}
        self.N, self.E, self.T = N, E, mp_iter
        self.D = self.params['obj_unary.weight'].shape[1]
        n_cls = self.params['obj_fc.weight'].shape[0]; n_rel = self.params['rel_fc.weight'].shape[0]
        self.slots = [_Slot(N, E, self.D, n_cls, n_rel, self.device) for _ in range(slots)]
        self.plans = [ops.L1Plan(self.params, N, E, self.D, mp_iter, self.device) for _ in range(slots)]
        # the parameter upload and weight splits above ran on the current stream; the slot streams are non-blocking and
        # must not start before them
        cur = torch.cuda.current_stream(self.device)
        for sl in self.slots:
            sl.stream.wait_stream(cur)
        self._next = 0
        self.h2d_bytes = (N + E) * self.D * 4 + E * 2 * 8
        self.d2h_bytes = (N * n_cls + E * n_rel) * 4

    def submit(self, obj_feat, edge_feat, rel_inds):
        """Host arrays in, returns a slot handle; results are valid after ``wait(handle)``."""
        i = self._next; self._next = (self._next + 1) % len(self.slots)
        s = self.slots[i]
        if s.busy:
            s.done.synchronize()
        # stage into pinned memory only if the caller's buffers are not already pinned torch tensors
        ho = obj_feat if (isinstance(obj_feat, torch.Tensor) and obj_feat.is_pinned()) else s.h_obj.copy_(torch.as_tensor(obj_feat))
        he = edge_feat if (isinstance(edge_feat, torch.Tensor) and edge_feat.is_pinned()) else s.h_edge.copy_(torch.as_tensor(edge_feat))
        hr = rel_inds if (isinstance(rel_inds, torch.Tensor) and rel_inds.is_pinned()) else s.h_rel.copy_(torch.as_tensor(rel_inds))
        with torch.cuda.stream(s.stream):
            s.d_obj.copy_(ho, non_blocking=True)
            s.d_edge.copy_(he, non_blocking=True)
            s.d_rel.copy_(hr, non_blocking=True)
            od, rd = self.plans[i].run(s.d_obj, s.d_edge, s.d_rel)      # graph index built inside the call
            s.h_od.copy_(od, non_blocking=True)
            s.h_rd.copy_(rd, non_blocking=True)
            s.done.record(s.stream)
        s.busy = True
        return i

    def wait(self, handle):
        """Returns the slot's pinned result buffers (views, valid until the slot is reused by a later ``submit``: copy them
        if they must outlive ``len(slots)`` further submissions; ``__call__`` returns copies)."""
        s = self.slots[handle]
        s.done.synchronize()
        s.busy = False
        return s.h_od, s.h_rd

    def __call__(self, obj_feat, edge_feat, rel_inds):
        od, rd = self.wait(self.submit(obj_feat, edge_feat, rel_inds))
        return od.numpy().copy(), rd.numpy().copy()

    def drain(self):
        for i, s in enumerate(self.slots):
            if s.busy:
                self.wait(i)
