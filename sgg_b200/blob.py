"""The batch wire format on the caller's side of the path: ``Blob`` / ``vg_collate`` of
``dataloaders/blob.py:17-261`` and ``dataloaders/visual_genome.py:681-688``, plus a synthetic Visual-Genome-shaped
dataset so ``main.py``-style loops (``for batch in loader: sgg_model(batch.scatter())``) run without the 60 GB dataset.

What a caller relies on is kept: constructor arguments, ``append`` of the dataset's entry dicts
(``visual_genome.py:439-449``: img, img_size, gt_boxes, gt_classes, gt_relations, scale, fn[, proposals]), ``reduce``,
``scatter``, ``len(blob) == num_gpus`` and the tuple ``blob[i]`` returns —

    train: (imgs, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels, proposals, None, fns)
    eval : (imgs, im_sizes, image_offset, gt_boxes, gt_classes, gt_rels, proposals, fns)

with ``gt_boxes`` scaled by ``entry['scale']``, ``gt_classes`` rows ``(img, class)`` and ``gt_rels`` rows
``(img, subj_local, obj_local, predicate)``.  B200-first differences, none visible in the tuple: the flat host tensors
are pinned so ``scatter()``'s ``non_blocking`` copies really overlap the previous step; with ``num_gpus > 1`` the blob
is sliced per rank on the host (one process per GPU, ``blob[rank]``) instead of ``torch.nn.parallel`` scatter of a
single-process DataParallel, which the reference declares but cannot run (``config.py:71`` asserts one GPU).
"""
import os

import numpy as np
import torch

from . import synth
from .host import IM_SCALE

BOX_SCALE = 1024                      # config.py:30


class Blob(object):
    def __init__(self, mode='det', is_train=False, num_gpus=1, primary_gpu=0, batch_size_per_gpu=3,
                 torch_detector=False, is_cuda=True):
        assert mode in ('det', 'rel')
        assert num_gpus >= 1
        self.mode, self.is_train, self.num_gpus = mode, is_train, num_gpus
        self.batch_size_per_gpu, self.primary_gpu = batch_size_per_gpu, primary_gpu
        self.torch_detector, self.is_cuda = torch_detector, is_cuda
        self.fns, self.imgs, self.im_sizes = [], [], []
        self.gt_boxes = self.gt_classes = self.gt_rels = self.proposals = None
        self.gt_box_chunks = self.gt_rel_chunks = self.proposal_chunks = None
        self._entries = []                # per-image raw arrays; the (image, ...) index columns are built in reduce()

    @property
    def is_rel(self):
        return self.mode == 'rel'

    @property
    def volatile(self):
        return not self.is_train

    def append(self, d):
        """Add one dataset entry (visual_genome.py:439-449 dict; wire format of dataloaders/blob.py:77-124)."""
        self.fns.append(os.path.basename(d['fn']))
        self.imgs.append(d['img'])
        self.im_sizes.append(tuple(d['img_size']))            # (h, w, scale)
        sc = d['scale']
        self._entries.append({
            'boxes': np.asarray(d['gt_boxes'], np.float32) * sc,
            'classes': np.asarray(d['gt_classes'], np.int64).reshape(-1),
            'rels': np.asarray(d['gt_relations'], np.int64).reshape(-1, 3) if self.is_rel else None,
            'props': sc * np.asarray(d['proposals'], np.float32) if 'proposals' in d else None,
        })

    def _pin(self, t):
        if self.is_cuda and torch.cuda.is_available():
            return t.pin_memory()
        return t

    def _flatten(self, key, index_dtype, out_dtype):
        """Rows of every image stacked, optionally prefixed by the image-index column; plus rows per GPU.
        Returns (tensor or 0 when there are no rows at all — as the reference does —, per-GPU row counts)."""
        rows = [e[key] for e in self._entries]
        counts = np.array([r.shape[0] for r in rows], dtype=np.int64)
        per_gpu = counts.reshape(self.num_gpus, self.batch_size_per_gpu).sum(1).tolist()
        if counts.sum() == 0:
            return 0, per_gpu
        flat = np.concatenate([r.reshape(r.shape[0], -1) for r in rows], 0)
        if index_dtype is not None:
            img = np.repeat(np.arange(len(rows)), counts).astype(index_dtype)
            flat = np.concatenate((img[:, None], flat.astype(index_dtype)), 1)
        return self._pin(torch.from_numpy(np.ascontiguousarray(flat)).to(out_dtype)), per_gpu

    def reduce(self):
        """Per-image lists -> flat pinned tensors + per-GPU row counts (dataloaders/blob.py:145-169)."""
        want = self.batch_size_per_gpu * self.num_gpus
        if len(self.imgs) != want:
            raise ValueError('Wrong batch size? imgs len {} bsize/gpu {} numgpus {}'.format(
                len(self.imgs), self.batch_size_per_gpu, self.num_gpus))
        if not self.torch_detector:
            self.imgs = self._pin(torch.stack(self.imgs, 0))
        self.im_sizes = np.asarray(self.im_sizes).reshape(self.num_gpus, self.batch_size_per_gpu, 3)
        self.gt_boxes, self.gt_box_chunks = self._flatten('boxes', None, torch.float32)
        self.gt_classes, _ = self._flatten('classes', np.int64, torch.int64)
        if self.is_rel:
            self.gt_rels, self.gt_rel_chunks = self._flatten('rels', np.int64, torch.int64)
        else:
            self.gt_rels = []
        if any(e['props'] is not None for e in self._entries):
            self.proposals, self.proposal_chunks = self._flatten('props', np.float32, torch.float32)
        else:
            self.proposals = []

    def _to_device(self, x, device=None):
        if not self.is_cuda or not isinstance(x, torch.Tensor):
            return x
        return x.cuda(self.primary_gpu if device is None else device, non_blocking=True)

    def _split(self, x, chunk_sizes):
        """Rows of ``x`` per GPU, image / box indices untouched (the consumer re-bases them, parallel.shard_batch)."""
        if not isinstance(x, torch.Tensor):
            return [x] * self.num_gpus
        out, lo = [], 0
        for n in chunk_sizes:
            out.append(x[lo:lo + n]); lo += n
        return out

    def scatter(self):
        """Move everything to the device(s) (dataloaders/blob.py:180-212); returns self."""
        if self.num_gpus == 1:
            if not self.torch_detector:
                self.imgs = self._to_device(self.imgs)
            self.gt_classes_primary = self.gt_classes = self._to_device(self.gt_classes)
            self.gt_boxes_primary = self.gt_boxes = self._to_device(self.gt_boxes)
            if self.is_rel:
                self.gt_rels = self._to_device(self.gt_rels)
            if self.proposal_chunks is not None:
                self.proposals = self._to_device(self.proposals)
            return self
        # one process per GPU: keep the host tensors, slice per rank in __getitem__
        bs = self.batch_size_per_gpu
        self.gt_classes_primary, self.gt_boxes_primary = self.gt_classes, self.gt_boxes
        self.gt_classes = self._split(self.gt_classes, self.gt_box_chunks)
        self.gt_boxes = self._split(self.gt_boxes, self.gt_box_chunks)
        if self.is_rel:
            self.gt_rels = self._split(self.gt_rels, self.gt_rel_chunks)
        if self.torch_detector:
            self.imgs = [self.imgs[i * bs:(i + 1) * bs] for i in range(self.num_gpus)]
        else:
            self.imgs = [self.imgs[i * bs:(i + 1) * bs] for i in range(self.num_gpus)]
        self.fns = [self.fns[i * bs:(i + 1) * bs] for i in range(self.num_gpus)]
        return self

    def __len__(self):
        return len(self.im_sizes)

    def __getitem__(self, index):
        """dataloaders/blob.py:219-261."""
        if index not in list(range(self.num_gpus)):
            raise ValueError('Out of bounds with index {} and {} gpus'.format(index, self.num_gpus))
        proposals = None if self.proposal_chunks is None else self.proposals
        if index == 0 and self.num_gpus == 1:
            rels = self.gt_rels if self.is_rel else None
            if self.is_train:
                return (self.imgs, self.im_sizes[0], 0, self.gt_boxes, self.gt_classes, rels, proposals, None, self.fns)
            return self.imgs, self.im_sizes[0], 0, self.gt_boxes, self.gt_classes, rels, proposals, self.fns
        assert proposals is None
        image_offset = self.batch_size_per_gpu * index
        rels_i = self.gt_rels[index] if self.is_rel else None
        if self.is_train:
            return (self.imgs[index], self.im_sizes[index], image_offset, self.gt_boxes[index], self.gt_classes[index],
                    rels_i, None, None, self.fns[index])
        return (self.imgs[index], self.im_sizes[index], image_offset, self.gt_boxes[index], self.gt_classes[index],
                rels_i, None, self.fns[index])


def vg_collate(data, num_gpus=1, is_train=False, mode='det', torch_detector=False, is_cuda=True):
    """dataloaders/visual_genome.py:681-688."""
    assert mode in ('det', 'rel')
    blob = Blob(mode=mode, is_train=is_train, num_gpus=num_gpus, batch_size_per_gpu=len(data) // num_gpus,
                torch_detector=torch_detector, is_cuda=is_cuda)
    for d in data:
        blob.append(d)
    blob.reduce()
    return blob


class SyntheticVG(torch.utils.data.Dataset):
    """Visual-Genome-shaped synthetic dataset: entries have the keys / dtypes / coordinate conventions of
    ``VG.__getitem__`` (dataloaders/visual_genome.py:439-449): boxes in the BOX_SCALE = 1024 frame with
    ``scale = IM_SCALE / BOX_SCALE``, image tensors [3,h,w] in [0,1] whose longer side is IM_SCALE, image-local
    relation triples (subj, obj, predicate).  Box count per image follows VG's statistics (mean 11.6, sd 5.8, 2..62,
    Zero_Shot_VG.ipynb) unless ``n_box`` is given.  Exposes ``ind_to_classes`` / ``ind_to_predicates`` (all the model
    constructor needs, rel_model_base.py:45-46)."""

    def __init__(self, num_images=64, n_box=None, n_rel=5, seed=0, im_hw=(IM_SCALE, IM_SCALE), with_images=True):
        self.num_images, self.n_box, self.n_rel, self.seed = num_images, n_box, n_rel, seed
        self.im_hw, self.with_images = im_hw, with_images
        self.ind_to_classes = ['__background__'] + ['c%d' % i for i in range(1, synth.NUM_CLASSES)]
        self.ind_to_predicates = ['__background__'] + ['p%d' % i for i in range(1, synth.NUM_RELS)]
        self.filenames = ['synthetic_%06d.jpg' % i for i in range(num_images)]

    def __len__(self):
        return self.num_images

    def __getitem__(self, index):
        rng = np.random.default_rng(self.seed * 1000003 + index)
        nb = self.n_box if self.n_box is not None else int(np.clip(np.rint(rng.normal(11.6, 5.8)), 2, 62))
        h, w = self.im_hw
        factor = float(IM_SCALE) / max(h, w)
        bs = float(BOX_SCALE) / IM_SCALE                              # boxes live in the 1024 frame
        xy = rng.random((nb, 2)) * np.array([0.6 * w, 0.6 * h])
        wh = rng.random((nb, 2)) * np.array([0.4 * w - 16, 0.4 * h - 16]) + 16
        gt_boxes = (np.concatenate((xy, xy + wh), 1) * factor * bs).astype(np.float32)
        gt_classes = rng.integers(1, synth.NUM_CLASSES, nb).astype(np.int64)
        ii, jj = np.nonzero(~np.eye(nb, dtype=bool))
        nr = min(self.n_rel, ii.shape[0])
        sel = np.sort(rng.choice(ii.shape[0], nr, replace=False))
        gt_rels = np.stack((ii[sel], jj[sel], rng.integers(1, synth.NUM_RELS, nr)), 1).astype(np.int64)
        if h > w:
            im_size = (IM_SCALE, int(w * factor), factor)
        elif h < w:
            im_size = (int(h * factor), IM_SCALE, factor)
        else:
            im_size = (IM_SCALE, IM_SCALE, factor)
        img = (torch.from_numpy(rng.random((3, im_size[0], im_size[1]), dtype=np.float32)) if self.with_images
               else torch.zeros(3, 1, 1))
        return {'img': img, 'img_size': im_size, 'gt_boxes': gt_boxes, 'gt_classes': gt_classes,
                'gt_relations': gt_rels, 'scale': IM_SCALE / BOX_SCALE, 'index': index, 'flipped': False,
                'fn': self.filenames[index]}


def synthetic_loader(dataset, batch_size=1, num_gpus=1, is_train=False, mode='rel', shuffle=False, num_workers=0,
                     is_cuda=True):
    """``VGDataLoader.splits``-style loader (dataloaders/visual_genome.py:691-735): batches are ``Blob``s built by
    ``vg_collate`` with ``torch_detector=True`` (images stay a list of [3,h,w] tensors for the detector transform)."""
    return torch.utils.data.DataLoader(
        dataset, batch_size=batch_size * num_gpus, shuffle=shuffle, num_workers=num_workers, drop_last=True,
        collate_fn=lambda x: vg_collate(x, mode=mode, num_gpus=num_gpus, is_train=is_train, torch_detector=True,
                                        is_cuda=is_cuda))
