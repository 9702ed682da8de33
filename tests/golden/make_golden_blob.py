#!/usr/bin/env python
"""Golden vectors for the batch wire format, made by RUNNING THE REFERENCE's ``dataloaders/blob.py`` Blob and
``dataloaders/visual_genome.py`` vg_collate on synthetic dataset entries (``sgg_b200.blob.SyntheticVG``).

Build container only (needs /root/reference):   python tests/golden/make_golden_blob.py
Stores, for each case, every member of the tuple ``blob[0]`` the reference returns (CPU, is_cuda=False)."""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

CASES = {   # name: (mode, is_train, n_images, torch_detector, with_proposals, dataset kwargs)
    'rel_train': ('rel', True, 3, True, False, dict(n_box=None, n_rel=5, seed=1)),
    'rel_eval': ('rel', False, 1, True, False, dict(n_box=10, n_rel=4, seed=2)),
    'det_eval_stacked': ('det', False, 2, False, False, dict(n_box=6, n_rel=3, seed=3)),
    'rel_train_proposals': ('rel', True, 2, True, True, dict(n_box=7, n_rel=2, seed=4)),
    'rel_eval_no_rels': ('rel', False, 2, True, False, dict(n_box=3, n_rel=0, seed=5)),
}


def entries(case):
    from sgg_b200.blob import SyntheticVG
    mode, is_train, n, td, props, kw = CASES[case]
    ds = SyntheticVG(num_images=n, im_hw=(592, 592) if not td else (480, 592), with_images=False, **kw)
    out = []
    for i in range(n):
        d = ds[i]
        if props:
            rng = np.random.default_rng(100 + i)
            d['proposals'] = (rng.random((4 + i, 4)) * 500).astype(np.float32)
        out.append(d)
    return out


def flatten(tup):
    out = {}
    for k, v in enumerate(tup):
        if v is None:
            out['m%d_none' % k] = np.zeros(0)
        elif isinstance(v, (list, tuple)) and len(v) and isinstance(v[0], str):
            out['m%d_strs' % k] = np.array(v)
        elif isinstance(v, (list, tuple)):
            out['m%d_len' % k] = np.array(len(v))
            for j, t in enumerate(v):
                out['m%d_%d' % (k, j)] = np.asarray(t)
        elif isinstance(v, int):
            out['m%d_int' % k] = np.array(v)
        else:
            out['m%d' % k] = np.asarray(v.detach().numpy() if hasattr(v, 'detach') else v)
    return out


def main():
    import_reference()
    from dataloaders.blob import Blob as RefBlob   # noqa: F401  (h5py is stubbed by import_reference)
    sys.modules.setdefault('pycocotools', type(sys)('pycocotools'))
    try:
        from dataloaders.visual_genome import vg_collate as ref_collate
    except Exception as ex:                        # heavy optional imports missing: the function is 6 lines around Blob
        print('vg_collate import failed (%s); driving Blob directly as vg_collate does' % ex)

        def ref_collate(data, num_gpus=1, is_train=False, mode='det', torch_detector=False, is_cuda=True):
            blob = RefBlob(mode=mode, is_train=is_train, num_gpus=num_gpus, batch_size_per_gpu=len(data) // num_gpus,
                           torch_detector=torch_detector, is_cuda=is_cuda)
            for d in data:
                blob.append(d)
            blob.reduce()
            return blob
    out = {}
    for name, (mode, is_train, n, td, props, kw) in CASES.items():
        b = ref_collate(entries(name), num_gpus=1, is_train=is_train, mode=mode, torch_detector=td, is_cuda=False)
        b = b.scatter()
        assert len(b) == 1
        for k, v in flatten(b[0]).items():
            out['%s__%s' % (name, k)] = v
        out['%s__tuple_len' % name] = np.array(len(b[0]))
    np.savez_compressed(os.path.join(HERE, 'blob.npz'), **out)
    print('wrote blob.npz with %d arrays' % len(out))


if __name__ == '__main__':
    main()
