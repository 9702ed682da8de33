"""Shared helpers: regenerate the inputs of a golden fixture from its seeds."""
import os
import numpy as np
from sgg_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: z[k] for k in z.files}


def l0_inputs(fx):
    seed = int(fx['seed'])
    if 'rel_inds' in fx:                       # explicit (special) graph
        rel = fx['rel_inds']; N = int(fx['N'])
    else:
        g = synth.synth_graph(int(fx['B']), int(fx['n_box']), int(fx['n_edge']), seed,
                              ragged=bool(fx['ragged']), all_pairs=bool(fx['all_pairs']))
        rel = g['rel_inds']; N = g['boxes'].shape[0]
    E = rel.shape[0]
    obj, relrep = synth.synth_l0_states(N, E, seed)
    p = synth.synth_params(seed, scale=float(fx['scale']), level='l0')
    assert synth.digest(obj, relrep, rel, p['edge_gru.weight_hh']) == str(fx['digest']), 'generator drift'
    return obj, relrep, rel, p, int(fx['T'])


def l1_inputs(fx):
    seed = int(fx['seed'])
    g = synth.synth_graph(int(fx['B']), int(fx['n_box']), int(fx['n_edge']), seed, all_pairs=bool(fx['all_pairs']))
    N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
    of, ef = synth.synth_l1_feats(N, E, seed)
    p = synth.synth_params(seed, scale=float(fx['scale']), level='l1')
    assert synth.digest(of, ef, g['rel_inds'], p['obj_unary.weight']) == str(fx['digest']), 'generator drift'
    return of, ef, g['rel_inds'], p, int(fx['T'])


def l2_inputs(fx):
    seed = int(fx['seed'])
    g = synth.synth_graph(2, 4, 10, seed)
    N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
    nfe, efe = synth.synth_pooled(N, E, seed)
    p = synth.synth_params(seed, level='l2', scale=1.0)
    assert synth.digest(nfe, efe, p['roi_fmap.1.0.weight'][:64]) == str(fx['digest']), 'generator drift'
    return nfe, efe, g['rel_inds'], g['rois'], p


def geom_inputs(fx):
    seed = int(fx['seed'])
    g = synth.synth_graph(3, 8, 20, seed)
    p = synth.synth_params(seed, level='l2', scale=float(fx['scale']))
    assert synth.digest(g['rois'], g['rel_inds'], p['union_boxes.conv.0.weight']) == str(fx['digest'])
    p = {k: v for k, v in p.items() if k.startswith('union_boxes.')}
    return g['rois'], g['rel_inds'][:, 1:], p


def roi_inputs(fx):
    fmap = synth.synth_fmap(2, int(fx['seed']), C=64)
    assert synth.digest(fmap) == str(fx['digest'])
    return fmap, fx['rois'], fx['union_inds']


def check_rows(full, rows, expect, colsum, atol, name=''):
    got = full[rows]
    err = np.abs(got - expect).max()
    assert err <= atol, '%s rows: max|d|=%.3e > %.1e' % (name, err, atol)
    cs = full.astype(np.float64).sum(0)
    scale = max(1.0, np.abs(full).astype(np.float64).sum(0).max())
    cerr = np.abs(cs - colsum).max()
    assert cerr <= atol * full.shape[0] * 0.25 + 1e-9 * scale, '%s colsum: %.3e' % (name, cerr)
