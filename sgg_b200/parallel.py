"""Data-parallel training across the GPUs of one node — new relative to the reference,
which asserts num_gpus == 1 (config.py:71, rel_model_stanford.py:121).

Every image is an independent graph (rel_inds only connect objects of one image), so the batch shards by image with
NO forward/backward communication; the only collective is the gradient all-reduce over the ~248 M trainable
parameters (991 MB fp32, of which roi_fmap / roi_fmap_obj are 96 %).

``FlatGradReducer`` keeps ONE flat gradient buffer; every ``p.grad`` is a view into it, laid out in the order the
gradients become ready in backward.  The buffer is cut into buckets (large tensors into row chunks); a bucket is
all-reduced IN PLACE, asynchronously, as soon as every element of it has been written and all earlier buckets have been
launched (fixed order => the same NCCL sequence on every rank whatever the arrival order).  The backward GEMMs of the
big layers write their weight gradient straight into the flat buffer chunk by chunk (``ops.register_grad_sink``), so the
all-reduce of chunk k overlaps the GEMM of chunk k+1; nothing is flattened, copied back or divided afterwards — the
1/world average is folded into the fused clip + SGD sweep (``FusedSGD.step(grad_scale=...)``), which reads the same
buffer.  The reference's gradient clipping (lib/pytorch_misc.py:625-664) sees the global-batch norm because it runs
after ``finish()``.
"""
import torch
import torch.distributed as dist


def shard_images(num_images, rank, world):
    """Contiguous, balanced image range [lo, hi) of this rank (gt_classes[:, 0] is sorted by image).  Every rank gets
    at least one image; a batch smaller than the world is rejected (an empty shard would skip the collectives)."""
    if num_images < world:
        raise ValueError('batch of %d images cannot be sharded over %d ranks' % (num_images, world))
    return (num_images * rank) // world, (num_images * (rank + 1)) // world


def shard_batch(batch0, rank, world):
    """Split one ``batch[0]`` tuple (dataloaders/blob.py:244-249) by image.  Image indices in gt_classes /
    gt_rels are re-based to start at 0 on every rank; box ids in gt_rels are image-local already."""
    imgs, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels = batch0[:6]
    rest = batch0[6:]
    B = len(imgs)
    lo, hi = shard_images(B, rank, world)
    keep_o = (gt_classes[:, 0] >= lo) & (gt_classes[:, 0] < hi)
    cls = gt_classes[keep_o].clone(); cls[:, 0] -= lo
    rels = None
    if gt_rels is not None:
        keep_r = (gt_rels[:, 0] >= lo) & (gt_rels[:, 0] < hi)
        rels = gt_rels[keep_r].clone(); rels[:, 0] -= lo
    fns = rest[-1][lo:hi] if (len(rest) and rest[-1] is not None) else None
    sizes = im_sizes[lo:hi] if im_sizes is not None else None
    return (imgs[lo:hi], sizes, image_offset, gt_boxes[keep_o], cls, rels) + tuple(rest[:-1]) + (fns,)


class _Bucket(object):
    __slots__ = ('lo', 'hi', 'left', 'work')

    def __init__(self, lo, hi):
        self.lo, self.hi, self.left, self.work = lo, hi, hi - lo, None


class FlatGradReducer(object):
    """Usage per step::

        red.begin()            # before backward: p.grad = None, counters reset
        loss.backward()        # hooks + gradient sinks fill the flat buffer and launch buckets
        red.finish()           # zero-fill parameters without a gradient, launch the rest, wait (stream-level)
        opt.step(max_norm=c, grad_scale=red.grad_scale)     # or average=True and a plain optimizer

    ``order``: parameter names (or tensors) in expected readiness order; default = the order observed during the first
    backward (until then: reverse registration order).  ``average=True`` divides the buffer by world in ``finish``
    (one extra sweep; for optimizers without a gradient pre-scale)."""

    def __init__(self, module, bucket_bytes=64 << 20, group=None, average=False, register_sinks=True, auto_begin=False):
        self.group = group
        self.auto_begin = auto_begin
        self.comm_enabled = True           # False: buckets are tracked but no collective is launched (compute-only timing)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.average = average
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        self.params = [p for _, p in self.named]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev, dt = self.params[0].device, self.params[0].dtype
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError('FlatGradReducer: parameters must share one device and dtype')
        # every tensor starts on a 256-byte boundary of the flat buffer (vectorised optimizer / GEMM stores need 16)
        self.total = sum(self._pad(p.numel()) for p in self.params)
        self.flat = torch.zeros(self.total, dtype=dt, device=dev)
        self.register_sinks = register_sinks and dev.type == 'cuda'
        self._learned = False
        self._arrival = []
        self._layout(list(reversed(self.params)))
        self.hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._in_step = False
        self.launched_bytes = 0

    @staticmethod
    def _pad(n):
        return (n + 63) // 64 * 64

    # ---- layout ----------------------------------------------------------------------------
    def _layout(self, ordered):
        """Place the parameters in ``ordered`` into the flat buffer, cut buckets, (re)create the gradient views."""
        self.offset, self.view, self.chunks = {}, {}, {}
        self.buckets = []
        off = 0
        cur_lo = 0
        for p in ordered:
            n = p.numel()
            self.offset[p] = off
            self.view[p] = self.flat[off:off + n].view_as(p)
            if n > self.bucket_elems and p.dim() == 2:
                # big matrix: close the running bucket, then one bucket per row chunk (chunks follow the GEMM's row order)
                if off > cur_lo:
                    self.buckets.append(_Bucket(cur_lo, off))
                rows, cols = p.shape
                per = max(128, (self.bucket_elems // cols) // 128 * 128)
                ch = [(r, min(rows, r + per)) for r in range(0, rows, per)]
                self.chunks[p] = ch
                for r0, r1 in ch:
                    self.buckets.append(_Bucket(off + r0 * cols, off + (r1 * cols if r1 < rows else self._pad(n))))
                cur_lo = off + self._pad(n)
            elif off + self._pad(n) - cur_lo >= self.bucket_elems:
                self.buckets.append(_Bucket(cur_lo, off + self._pad(n)))
                cur_lo = off + self._pad(n)
            off += self._pad(n)
        if off > cur_lo:
            self.buckets.append(_Bucket(cur_lo, off))
        self.starts = [b.lo for b in self.buckets]
        for p in ordered:
            if p.grad is not None:                    # keep already computed gradients across a re-layout
                self.view[p].copy_(p.grad)
                p.grad = self.view[p]
        if self.register_sinks:
            from . import ops
            ops.clear_grad_sinks(owner=self)
            for p in ordered:
                if p.dim() == 2:
                    ops.register_grad_sink(p, self.view[p], self.chunks.get(p), self._notify, owner=self)

    @property
    def grad_scale(self):
        return 1.0 if self.average else 1.0 / self.world

    # ---- per step --------------------------------------------------------------------------
    def begin(self, clear=True):
        """clear=True: p.grad = None, so autograd adopts the produced gradient tensors (views of the flat buffer when the
        producer wrote into the sink) without a zero-fill or an accumulate pass.  clear=False keeps existing gradients
        (``zero_grad(set_to_none=False)`` protocol): autograd accumulates in place into the flat buffer."""
        if clear:
            for p in self.params:
                p.grad = None
        self._marked = {p: 0 for p in self.params}
        self._arrived = set()
        for b in self.buckets:
            b.left, b.work = b.hi - b.lo, None
        self._next = 0
        self._in_step = True
        self.launched_bytes = 0

    def _mark(self, lo, hi):
        """elements [lo, hi) of the flat buffer are final for this step"""
        import bisect
        i = bisect.bisect_right(self.starts, lo) - 1
        while i < len(self.buckets) and self.buckets[i].lo < hi:
            b = self.buckets[i]
            b.left -= min(hi, b.hi) - max(lo, b.lo)
            i += 1
        self._launch_ready()

    def _launch_ready(self):
        while self._next < len(self.buckets) and self.buckets[self._next].left <= 0:
            b = self.buckets[self._next]
            if self.world > 1 and self.comm_enabled:
                b.work = dist.all_reduce(self.flat[b.lo:b.hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.launched_bytes += (b.hi - b.lo) * 4
            self._next += 1

    def _notify(self, p, r0, r1):
        """called by the backward GEMM after rows [r0, r1) of p's gradient were written into the flat buffer"""
        if not self._in_step:
            return
        rows, cols = p.shape
        self._marked[p] += (r1 - r0) * cols
        hi = r1 * cols if r1 < rows else self._pad(p.numel())      # the last chunk carries the alignment padding
        self._mark(self.offset[p] + r0 * cols, self.offset[p] + hi)

    def _on_grad(self, p):
        if not self._in_step:
            if not self.auto_begin:
                raise RuntimeError('FlatGradReducer: call begin() before backward()')
            self.begin(clear=False)
        v = self.view[p]
        g = p.grad
        if g.data_ptr() != v.data_ptr():
            v.copy_(g)                        # generic path: the producer did not write into the sink
            p.grad = v
        if not self._learned:
            self._arrival.append(p)
        if id(p) in self._arrived:
            raise RuntimeError('FlatGradReducer: gradient of a parameter arrived twice in one step (call begin() per backward)')
        self._arrived.add(id(p))
        n = p.numel()
        done = self._marked[p]                 # elements already announced by the producing GEMM (gradient sink)
        if done < n:
            self._marked[p] = n
            self._mark(self.offset[p] + done, self.offset[p] + self._pad(n))

    def finish(self):
        """After backward: parameters that received no gradient contribute zeros (every rank launches the same
        collectives), remaining buckets are launched in order, and the current stream waits for all of them."""
        if not self._in_step:
            raise RuntimeError('FlatGradReducer: finish() without begin()')
        for p in self.params:
            if self._marked[p] == 0:
                self.view[p].zero_()
                p.grad = self.view[p]
                self._marked[p] = p.numel()
                self._mark(self.offset[p], self.offset[p] + self._pad(p.numel()))
        assert self._next == len(self.buckets), 'unlaunched buckets'
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()
                b.work = None
        if self.average and self.world > 1:
            self.flat.div_(self.world)
        self._in_step = False
        if not self._learned:
            # first step done: re-lay the buffer in the observed readiness order (same on every rank: same graph)
            self._learned = True
            seen = set()
            order = [p for p in self._arrival if not (id(p) in seen or seen.add(id(p)))]
            order += [p for p in reversed(self.params) if id(p) not in seen]
            self._arrival = []
            if self.world > 1:                 # ranks may see different arrival orders (unused parameters): rank 0 decides
                pos = {id(p): i for i, p in enumerate(self.params)}
                idx = torch.tensor([pos[id(p)] for p in order], dtype=torch.int64, device=self.flat.device)
                dist.broadcast(idx, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
                order = [self.params[i] for i in idx.tolist()]
            if [id(p) for p in order] != [id(p) for p in sorted(self.params, key=lambda q: self.offset[q])]:
                old = self.flat
                self.flat = torch.empty_like(old)
                self._layout(order)
                del old

    def remove(self):
        for h in self.hooks:
            h.remove()
        if self.register_sinks:
            from . import ops
            ops.clear_grad_sinks(owner=self)


# round-1 name, kept for callers / tests
class GradAllReducer(FlatGradReducer):
    def __init__(self, module, bucket_bytes=32 << 20, group=None):
        super().__init__(module, bucket_bytes=bucket_bytes, group=group, average=True, auto_begin=True)

    def finish(self):
        if not self._in_step:                         # backward produced no gradient at all on this rank
            self.begin(clear=False)
        super().finish()


def clip_grad_norm(named_parameters, max_norm, clip=True):
    """Global-norm clipping with the reference's semantics (lib/pytorch_misc.py:625-664): one norm over all
    gradients, scale by max_norm / (norm + 1e-6) when that is < 1.  Single fused norm instead of a Python loop
    of .norm() calls; returns the total norm.  (torch ops: used by the gloo CPU tests; the CUDA path is
    ``sgg_b200.optim.clip_grad_norm`` / ``FusedSGD.step(max_norm=...)``.)"""
    grads = [p.grad for _, p in named_parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = float(max_norm) / (total + 1e-6)
    if clip and coef < 1:
        for g in grads:
            g.mul_(coef)
    return total
