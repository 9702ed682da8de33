#!/usr/bin/env python
"""Generate golden vectors by RUNNING THE REFERENCE ITSELF (bknyaz/sgg) on CPU.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Recipe: copy the reference to a writable temp dir,
build its Cython ``lib/draw_rectangles`` there, stub ``h5py`` (not installed;
only imported at module top by lib/pytorch_misc.py:7), import
``sgg_models.rel_model_stanford.RelModelStanford`` and drive its own
``message_pass`` / ``predict`` / ``forward`` / ``union_boxes`` / ``roi_pool``.

Inputs and weights come from ``sgg_b200.synth`` (numpy ``default_rng`` seeds), so
the fixtures store only seeds + shapes + an input digest + the reference's
outputs (row-subsampled for big cases, plus float64 column sums of the full
output).  Tests regenerate the inputs from the seeds and compare.
"""
import os, sys, types, shutil, subprocess, tempfile, argparse
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from sgg_b200 import synth  # noqa: E402


def import_reference(ref_src='/root/reference'):
    tmp = os.environ.get('SGG_REF_COPY', os.path.join(tempfile.gettempdir(), 'sgg_ref_copy'))
    if not os.path.exists(os.path.join(tmp, 'main.py')):
        shutil.copytree(ref_src, tmp, dirs_exist_ok=True)
        subprocess.check_call(['chmod', '-R', 'u+w', tmp])
    dr = os.path.join(tmp, 'lib', 'draw_rectangles')
    if not any(f.endswith('.so') for f in os.listdir(dr)):
        if os.path.exists(os.path.join(dr, 'draw_rectangles.c')):
            os.remove(os.path.join(dr, 'draw_rectangles.c'))      # stale Cython-0.29 output
        subprocess.check_call([sys.executable, 'setup.py', 'build_ext', '--inplace'], cwd=dr)
    sys.path.insert(0, tmp)
    sys.modules.setdefault('h5py', types.ModuleType('h5py'))
    import torch
    torch.manual_seed(111)
    from sgg_models.rel_model_stanford import RelModelStanford
    return RelModelStanford


class FakeData:
    ind_to_classes = ['__background__'] + ['c%d' % i for i in range(150)]
    ind_to_predicates = ['__background__'] + ['p%d' % i for i in range(50)]


def load_params(model, p):
    import torch
    sd = model.state_dict()
    for k, v in p.items():
        assert k in sd and tuple(sd[k].shape) == v.shape, (k, v.shape)
        sd[k].copy_(torch.from_numpy(v))


def subsample(a, seed, k=48):
    rng = np.random.default_rng(seed)
    if a.shape[0] <= k:
        return np.arange(a.shape[0]), a
    rows = np.sort(rng.choice(a.shape[0], k, replace=False))
    return rows, a[rows]


def special_graph():
    """Edge cases the reference handles implicitly: unsorted rel_inds, duplicate
    edges (they simply add), isolated nodes (4, 6), a self-contained 2-node image."""
    rel = np.array([[0, 3, 1], [0, 0, 1], [0, 0, 1], [0, 2, 0], [0, 1, 3], [0, 3, 2],
                    [1, 5, 7], [1, 7, 5], [0, 1, 0], [0, 2, 3], [0, 0, 2]], np.int64)
    return 8, rel


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=HERE)
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    RelModelStanford = import_reference()
    model = RelModelStanford(train_data=FakeData(), mode='predcls')
    model.eval()
    tt = torch.from_numpy

    def want(name):
        return not args.only or name in args.only.split(',')

    def save(name, **kw):
        np.savez_compressed(os.path.join(args.out, name + '.npz'), **kw)
        print('wrote', name, {k: getattr(v, 'shape', v) for k, v in kw.items()})

    # ---------------- L0: message_pass ---------------------------------------
    l0_cases = [  # name, B, n_box, n_edge, all_pairs, ragged, T, scale, seed
        ('l0_cfg1', 1, 10, 90, True, False, 3, 1.0, 1235),
        ('l0_cfg2_s3', 8, 30, 300, False, False, 3, 3.0, 1236),
        ('l0_ragged_t6', 5, 12, 40, False, True, 6, 2.0, 1237),
    ]
    for name, B, nb, ne, ap_, rg, T, scale, seed in l0_cases:
        if not want(name):
            continue
        g = synth.synth_graph(B, nb, ne, seed, ragged=rg, all_pairs=ap_)
        N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
        obj, rel = synth.synth_l0_states(N, E, seed)
        p = synth.synth_params(seed, scale=scale, level='l0')
        load_params(model, p)
        model.mp_iter = T
        with torch.no_grad():
            v, e = model.message_pass(tt(rel), tt(obj), tt(g['rel_inds'][:, 1:3]))
        v, e = v.numpy(), e.numpy()
        vr, vs = subsample(v, seed); er, es = subsample(e, seed + 1)
        save(name, B=B, n_box=nb, n_edge=ne, all_pairs=ap_, ragged=rg, T=T, scale=scale, seed=seed,
             N=N, E=E, digest=synth.digest(obj, rel, g['rel_inds'], p['edge_gru.weight_hh']),
             v_rows=vr, v=vs, e_rows=er, e=es, v_colsum=v.astype(np.float64).sum(0),
             e_colsum=e.astype(np.float64).sum(0))
    if want('l0_special'):
        seed, T, scale = 1238, 3, 2.0
        N, relg = special_graph()
        E = relg.shape[0]
        obj, rel = synth.synth_l0_states(N, E, seed)
        p = synth.synth_params(seed, scale=scale, level='l0')
        load_params(model, p)
        model.mp_iter = T
        with torch.no_grad():
            v, e = model.message_pass(tt(rel), tt(obj), tt(relg[:, 1:3]))
        save('l0_special', T=T, scale=scale, seed=seed, N=N, E=E, rel_inds=relg,
             digest=synth.digest(obj, rel, relg, p['edge_gru.weight_hh']), v=v.numpy(), e=e.numpy())

    # ---------------- L1: 4096-d feats -> dists ------------------------------
    l1_cases = [('l1_cfg1', 1, 10, 90, True, 3, 1.0, 2235), ('l1_cfg2', 8, 30, 300, False, 3, 1.0, 2236),
                ('l1_cfg2_s3', 8, 30, 300, False, 3, 3.0, 2237)]
    for name, B, nb, ne, ap_, T, scale, seed in l1_cases:
        if not want(name):
            continue
        g = synth.synth_graph(B, nb, ne, seed, all_pairs=ap_)
        N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
        of, ef = synth.synth_l1_feats(N, E, seed)
        p = synth.synth_params(seed, scale=scale, level='l1')
        load_params(model, p)
        model.mp_iter = T
        with torch.no_grad():
            nf = model.obj_unary(tt(of)); efu = torch.relu(model.edge_unary(tt(ef)))
            v, e = model.message_pass(efu, nf, tt(g['rel_inds'][:, 1:3]))
            od, rd = model.obj_fc(v).numpy(), model.rel_fc(e).numpy()
        orow, osub = subsample(od, seed, 64); rrow, rsub = subsample(rd, seed + 1, 128)
        save(name, B=B, n_box=nb, n_edge=ne, all_pairs=ap_, T=T, scale=scale, seed=seed, N=N, E=E,
             digest=synth.digest(of, ef, g['rel_inds'], p['obj_unary.weight']),
             obj_rows=orow, obj_dists=osub, rel_rows=rrow, rel_dists=rsub,
             obj_colsum=od.astype(np.float64).sum(0), rel_colsum=rd.astype(np.float64).sum(0))

    # ---------------- L1 gradients (torch.autograd through the reference's modules) ------------
    # grad_l1_cfg2 / grad_l1_shard32: BASELINE cfg2 and the cfg4 per-GPU shard (N = 960, E = 9600) — sizes at which the
    # CUDA backward runs on the tensor-core GEMMs (round-1 verdict: gradient parity only existed at toy sizes)
    for name, B, nb, ne, ap_, T, scale, seed in [('grad_l1_cfg1', 1, 10, 90, True, 3, 1.0, 7235),
                                                  ('grad_l1_small_s2', 3, 9, 30, False, 3, 2.0, 7236),
                                                  ('grad_l1_cfg2', 8, 30, 300, False, 3, 1.0, 7237),
                                                  ('grad_l1_shard32', 32, 30, 300, False, 3, 1.0, 7238)]:
        if not want(name):
            continue
        g = synth.synth_graph(B, nb, ne, seed, all_pairs=ap_)
        N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
        of, ef = synth.synth_l1_feats(N, E, seed)
        p = synth.synth_params(seed, scale=scale, level='l1')
        load_params(model, p)
        model.mp_iter = T
        rng = np.random.default_rng(seed + 5)
        r1 = rng.standard_normal((N, 151), dtype=np.float32); r2 = rng.standard_normal((E, 51), dtype=np.float32)
        oft, eft = tt(of).requires_grad_(), tt(ef).requires_grad_()
        model.zero_grad()
        nf = model.obj_unary(oft); efu = torch.relu(model.edge_unary(eft))
        v, e = model.message_pass(efu, nf, tt(g['rel_inds'][:, 1:3]))
        loss = (model.obj_fc(v) * tt(r1)).sum() + (model.rel_fc(e) * tt(r2)).sum()
        loss.backward()
        out = dict(B=B, n_box=nb, n_edge=ne, all_pairs=ap_, T=T, scale=scale, seed=seed, N=N, E=E,
                   digest=synth.digest(of, ef, g['rel_inds'], p['obj_unary.weight']), loss=float(loss))
        sd = dict(model.named_parameters())
        gr = {k: sd[k].grad.numpy() for k in p.keys()}
        gr['obj_feat'] = oft.grad.numpy(); gr['edge_feat'] = eft.grad.numpy()
        for k, gv in gr.items():
            flat = gv.reshape(-1)
            idx = np.sort(np.random.default_rng(seed + 11).choice(flat.shape[0], min(512, flat.shape[0]), replace=False))
            kk = k.replace('.', '__')
            out['idx__' + kk] = idx; out['val__' + kk] = flat[idx]
            out['sum__' + kk] = np.float64(flat.astype(np.float64).sum())
            out['asum__' + kk] = np.float64(np.abs(flat).astype(np.float64).sum())
        save(name, **out)

    # ---------------- a8: draw_union_boxes (Cython) ---------------------------
    if want('draw_union_boxes'):
        from lib.draw_rectangles.draw_rectangles import draw_union_boxes
        g = synth.synth_graph(2, 9, 30, 3235)
        rois = g['rois']; ui = g['rel_inds'][:, 1:]
        pair = np.concatenate((rois[:, 1:][ui[:, 0]], rois[:, 1:][ui[:, 1]]), 1).astype(np.float32)
        # add touching / nested / identical boxes
        extra = np.array([[10, 10, 50, 50, 10, 10, 50, 50], [0, 0, 100, 80, 20, 20, 40, 30],
                          [5, 5, 25, 25, 25, 25, 60, 70], [1.5, 2.25, 300.75, 17.5, 290.1, 3.3, 591.9, 400.2]], np.float32)
        pair = np.concatenate((pair, extra))
        out = draw_union_boxes(pair, 27)
        save('draw_union_boxes', pairs=pair, out=out.astype(np.float32))

    # ---------------- a7: union-box geometry conv, eval + train BN ------------
    if want('union_geom'):
        seed = 4235
        g = synth.synth_graph(3, 8, 20, seed)
        p = synth.synth_params(seed, level='l2', scale=1.5)
        pg = {k: v for k, v in p.items() if k.startswith('union_boxes.')}
        load_params(model, pg)
        rois = tt(g['rois']); ui = tt(g['rel_inds'][:, 1:])
        E = ui.shape[0]
        pools = torch.zeros(E, 512, 7, 7)
        with torch.no_grad():
            model.union_boxes.eval()
            ev = model.union_boxes(pools, rois, ui, None)[:, :, 0, 0].numpy()
            model.union_boxes.train()
            rm0 = model.union_boxes.conv[2].running_mean.clone()
            tr = model.union_boxes(pools, rois, ui, None)[:, :, 0, 0].numpy()
            rm1 = model.union_boxes.conv[2].running_mean.clone().numpy()
            rv1 = model.union_boxes.conv[2].running_var.clone().numpy()
            rm2 = model.union_boxes.conv[6].running_mean.clone().numpy()
            rv2 = model.union_boxes.conv[6].running_var.clone().numpy()
            model.union_boxes.eval()
        save('union_geom', seed=seed, E=E, scale=1.5, digest=synth.digest(g['rois'], g['rel_inds'], p['union_boxes.conv.0.weight']),
             out_eval=ev, out_train=tr, rm1=rm1, rv1=rv1, rm2=rm2, rv2=rv2)

    # ---------------- a7 (train mode): forward with batch statistics + gradients of the conv / BN parameters --------
    if want('union_geom_train'):
        seed = 4236
        g = synth.synth_graph(3, 9, 40, seed)
        p = synth.synth_params(seed, level='l2', scale=1.5)
        pg = {k: v for k, v in p.items() if k.startswith('union_boxes.')}
        load_params(model, pg)
        rois = tt(g['rois']); ui = tt(g['rel_inds'][:, 1:])
        E = ui.shape[0]
        r = np.random.default_rng(seed + 5).standard_normal((E, 512), dtype=np.float32)
        ub = model.union_boxes
        ub.train()
        for q in ub.parameters():
            q.grad = None
        out = ub(torch.zeros(E, 512, 7, 7), rois, ui, None)[:, :, 0, 0]
        loss = (out * tt(r)).sum()
        loss.backward()
        rec = dict(seed=seed, E=E, scale=1.5, loss=float(loss), out_train=out.detach().numpy(),
                   digest=synth.digest(g['rois'], g['rel_inds'], p['union_boxes.conv.0.weight']),
                   rm1=ub.conv[2].running_mean.numpy().copy(), rv1=ub.conv[2].running_var.numpy().copy(),
                   rm2=ub.conv[6].running_mean.numpy().copy(), rv2=ub.conv[6].running_var.numpy().copy())
        for idx in ('0', '2', '4', '6'):
            for nm in ('weight', 'bias'):
                gq = getattr(ub.conv[int(idx)], nm).grad.numpy()
                key = 'g%s_%s' % (idx, nm)
                if idx == '4' and nm == 'weight':          # only the centre tap of the 3x3 kernel sees data (stride-16 quirk)
                    rec['g4_offcentre_absmax'] = float(np.abs(np.delete(gq.reshape(512, 256, 9), 4, axis=2)).max())
                    gq = gq[:, :, 1, 1]
                    sel = np.sort(np.random.default_rng(seed + 11).choice(gq.size, 4096, replace=False))
                    rec[key + '_idx'] = sel; rec[key + '_val'] = gq.reshape(-1)[sel]
                    rec[key + '_asum'] = np.float64(np.abs(gq).astype(np.float64).sum())
                else:
                    rec[key] = gq
        ub.eval()
        save('union_geom_train', **rec)

    # ---------------- a9: node_edge_features (RoIAlign) -----------------------
    if want('roi_align'):
        seed = 5235
        g = synth.synth_graph(2, 5, 8, seed)
        # push some boxes across the image border / make tiny boxes (edge rules of roi_align)
        g['rois'][0, 1:] = [-20.0, -10.0, 30.0, 25.0]
        g['rois'][3, 1:] = [580.0, 570.0, 640.0, 650.0]
        g['rois'][6, 1:] = [100.0, 100.0, 104.0, 103.0]
        fmap = synth.synth_fmap(2, seed, C=64)
        with torch.no_grad():
            nf, ef = model.node_edge_features(tt(fmap), tt(g['rois']), tt(g['rel_inds'][:, 1:]),
                                              [(592, 592), (592, 592)])
        save('roi_align', seed=seed, rois=g['rois'], union_inds=g['rel_inds'][:, 1:],
             digest=synth.digest(fmap), node_feat=nf.numpy(), edge_feat=ef.numpy())

    # ---------------- state-dict contract ------------------------------------------------
    if want('state_dict'):
        import json
        sd = model.state_dict()
        with open(os.path.join(args.out, 'state_dict_keys.json'), 'w') as f:
            json.dump({k: list(v.shape) for k, v in sd.items()}, f, indent=0, sort_keys=True)
        print('wrote state_dict_keys.json', len(sd))

    # ---------------- L3: forward(batch) eval, predcls + sgcls ------------------------------
    if want('l3_forward'):
        seed = 8235
        sizes, boxes, gt_classes, gt_rels = l3_case(seed)
        imgs = synth.synth_images(sizes, seed)
        p = synth.synth_params(seed, level='l2', scale=1.0)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith('detector.backbone.')}
        p.update(synth.synth_backbone_params(shapes, seed))
        load_params(model, p)
        model.mp_iter = 3
        out = dict(seed=seed, digest=synth.digest(imgs[0], imgs[1], boxes, p['detector.backbone.0.weight']))
        for mode in ('predcls', 'sgcls'):
            model.mode = mode
            model.eval()
            batch = [([tt(i)[None] for i in imgs], None, 0, tt(boxes), tt(gt_classes), tt(gt_rels), None, ['a', 'b'])]
            with torch.no_grad():
                b, oc, os_, rels, ps = model(batch)
            out.update({mode + '_boxes': b, mode + '_obj_classes': oc, mode + '_obj_scores': os_,
                        mode + '_rels': rels, mode + '_pred_scores': ps})
        model.mode = 'predcls'
        save('l3_forward', **out)

    # ---------------- L2: predict ---------------------------------------------
    if want('l2_predict'):
        seed = 6235
        g = synth.synth_graph(2, 4, 10, seed)
        N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
        nfe, efe = synth.synth_pooled(N, E, seed)
        p = synth.synth_params(seed, level='l2', scale=1.0)
        load_params(model, p)
        model.mp_iter = 3
        with torch.no_grad():
            od, rd = model.predict(tt(nfe), tt(efe), tt(g['rel_inds']), tt(g['rois']), None)
        save('l2_predict', seed=seed, N=N, E=E, digest=synth.digest(nfe, efe, p['roi_fmap.1.0.weight'][:64]),
             obj_dists=od.numpy(), rel_dists=rd.numpy())


    # ---------------- L2 in TRAINING mode: predict() forward + autograd through every trainable tensor --------------
    # (batch-statistics BN in the geometry branch, fc6 on pools + broadcast geometry, fc7, unary, T x message passing,
    # heads; dropout probability set to 0 on the live modules so the run is deterministic)
    if want('l2_train_grad'):
        seed = 6236
        g = synth.synth_graph(3, 6, 14, seed)
        N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
        nfe, efe = synth.synth_pooled(N, E, seed)
        p = synth.synth_params(seed, level='l2', scale=1.0)
        load_params(model, p)
        model.mp_iter = 3
        model.train()
        drops = [m for m in model.modules() if isinstance(m, torch.nn.Dropout)]
        old_p = [m.p for m in drops]
        for m in drops:
            m.p = 0.0
        for q in model.parameters():
            q.grad = None
        rng = np.random.default_rng(seed + 5)
        r1 = rng.standard_normal((N, 151), dtype=np.float32); r2 = rng.standard_normal((E, 51), dtype=np.float32)
        od, rd = model.predict(tt(nfe), tt(efe), tt(g['rel_inds']), tt(g['rois']), None)
        loss = (od * tt(r1)).sum() + (rd * tt(r2)).sum()
        loss.backward()
        out = dict(seed=seed, N=N, E=E, digest=synth.digest(nfe, efe, p['roi_fmap.1.0.weight'][:64]), loss=float(loss.detach()),
                   obj_dists=od.detach().numpy(), rel_dists=rd.detach().numpy(),
                   rm1=model.union_boxes.conv[2].running_mean.numpy().copy(), rv2=model.union_boxes.conv[6].running_var.numpy().copy())
        names = []
        for k, q in model.named_parameters():
            if k.startswith('detector.') or q.grad is None:
                continue
            flat = q.grad.numpy().reshape(-1)
            idx = np.sort(np.random.default_rng(seed + 11).choice(flat.shape[0], min(512, flat.shape[0]), replace=False))
            kk = k.replace('.', '__')
            out['idx__' + kk] = idx; out['val__' + kk] = flat[idx]
            out['asum__' + kk] = np.float64(np.abs(flat).astype(np.float64).sum())
            names.append(k)
        out['names'] = np.array(names)
        for m, pp in zip(drops, old_p):
            m.p = pp
        model.eval()
        save('l2_train_grad', **out)


def l3_case(seed=8235):
    """Shared by make_golden and the tests: 2 images (one non-square), GT boxes inside the image."""
    sizes = [(592, 592), (400, 592)]
    g = synth.synth_graph(2, 7, 10, seed)
    boxes = g['boxes'].copy()
    n0 = int((g['gt_classes'][:, 0] == 0).sum())
    boxes[n0:, [1, 3]] *= np.float32(400.0 / 592.0)            # second image is 400 px high
    boxes[n0:, 3] = np.maximum(boxes[n0:, 3], boxes[n0:, 1] + 8)
    return sizes, boxes.astype(np.float32), g['gt_classes'], g['gt_rels']


if __name__ == '__main__':
    main()
