"""Data-parallel training across the GPUs of one node — new relative to the reference,
which asserts num_gpus == 1 (config.py:71, rel_model_stanford.py:121).

Every image is an independent graph (rel_inds only connect objects of one image), so the batch shards
by image with NO forward/backward communication; the only collective is the gradient all-reduce
(sum, then / world) over the ~248 M trainable parameters, bucketed and launched from autograd hooks so
that it overlaps the rest of the backward pass (NCCL over NVLink/NVSwitch on the box; gloo in CPU tests).
The reference's gradient clipping (lib/pytorch_misc.py:625-664) must run AFTER ``finish()`` so that it
sees the global-batch gradient norm.
"""
import torch
import torch.distributed as dist


def shard_images(num_images, rank, world):
    """Contiguous image range [lo, hi) of this rank (gt_classes[:, 0] is sorted by image)."""
    per = (num_images + world - 1) // world
    lo = min(rank * per, num_images)
    return lo, min(lo + per, num_images)


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


class GradAllReducer(object):
    """Bucketed gradient all-reduce driven by post-accumulate-grad hooks.

    Buckets are filled in reverse parameter order (roughly the order gradients become ready); when the last
    gradient of a bucket lands, the bucket is flattened into a persistent buffer and all-reduced
    asynchronously.  ``finish()`` waits, averages and scatters the results back into ``p.grad``."""

    def __init__(self, module, bucket_bytes=32 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p); size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur); cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=b[0].dtype, device=b[0].device) for b in self.buckets]
        self.where = {}
        for bi, b in enumerate(self.buckets):
            for p in b:
                self.where[p] = bi
        self.pending = [len(b) for b in self.buckets]
        self.works = []
        self.hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _launch(self, bi):
        flat, off = self.flat[bi], 0
        for p in self.buckets[bi]:
            n = p.numel()
            flat[off:off + n].copy_(p.grad.reshape(-1) if p.grad is not None else torch.zeros_like(p).reshape(-1))
            off += n
        if self.world > 1:
            self.works.append((bi, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)))
        else:
            self.works.append((bi, None))

    def _on_grad(self, p):
        bi = self.where[p]
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def finish(self):
        """Call after loss.backward(): flushes buckets whose parameters got no gradient, waits for the
        collectives, writes the averaged gradients back."""
        for bi, left in enumerate(self.pending):
            if left > 0:
                self._launch(bi)
        for bi, w in self.works:
            if w is not None:
                w.wait()
            flat, off = self.flat[bi], 0
            if self.world > 1:
                flat.div_(self.world)
            for p in self.buckets[bi]:
                n = p.numel()
                if p.grad is not None:
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n
        self.works = []
        self.pending = [len(b) for b in self.buckets]

    def remove(self):
        for h in self.hooks:
            h.remove()


def clip_grad_norm(named_parameters, max_norm, clip=True):
    """Global-norm clipping with the reference's semantics (lib/pytorch_misc.py:625-664): one norm over all
    gradients, scale by max_norm / (norm + 1e-6) when that is < 1.  Single fused norm instead of a Python loop
    of .norm() calls; returns the total norm."""
    grads = [p.grad for _, p in named_parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = float(max_norm) / (total + 1e-6)
    if clip and coef < 1:
        for g in grads:
            g.mul_(coef)
    return total
