#!/usr/bin/env python
"""Benchmark of the IMP relation-model hot path (BASELINE.json metric: images/sec, PredCls,
3 MP iterations; MP/L1-kernel HBM GB/s vs measured peak).

  python bench.py --gpus 1 --steps 50 --warmup 5            # this repo's CUDA path
  python bench.py --impl reference --steps 3 --warmup 1     # reference's CPU path (oracle port)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload at N=1 = BASELINE.json configs[1]: PredCls batch=8, 30 boxes/img, 300 candidate
edges/img, 4096-d features (the L1 boundary: rel_model_stanford.py:103-107 without roi_fmap*),
3 MP iterations.  A "step" = one L1 forward over one batch of 8 synthetic images.  N>1: every
rank runs the same-sized independent batch (images shard naturally; no data-path collective),
weak scaling, value = images of all ranks / max-over-ranks time.

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from sgg_b200 import synth  # noqa: E402

L2_BYTES_FALLBACK = 126 * 1024 * 1024


_T0 = time.time()


def log(msg):
    """progress line on stderr (stdout carries exactly one JSON line); rank 0 only"""
    if os.environ.get('RANK', '0') != '0':
        return
    sys.stderr.write('[bench %7.1fs] %s\n' % (time.time() - _T0, msg)); sys.stderr.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--boxes', type=int, default=30)
    ap.add_argument('--edges', type=int, default=300)
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='l1', choices=['l1', 'train'],
                    help="headline workload: 'l1' = BASELINE.json configs[1] forward (default); 'train' = configs[3] train step")
    ap.add_argument('--no-other-configs', action='store_true', help='skip the cfg3 / cfg5 single-GPU measurements')
    ap.add_argument('--no-train', action='store_true', help='skip the train-step measurement nested under "train_step"')
    ap.add_argument('--train-batch', type=int, default=32, help='images per GPU of the train step (configs[3]: 256 / 8)')
    ap.add_argument('--train-steps', type=int, default=8)
    ap.add_argument('--no-graph', action='store_true', help='do not replay the step from a CUDA graph')
    ap.add_argument('--cpu-baseline-only', action='store_true', help='internal: print the cpu_baseline object and exit')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), float(d.get('bf16_tflops', 0)), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1590.0, 'fallback (B200_PROFILING.md)'


def alg_bytes_l1(N, E, H, D, T, n_cls=151, n_rel=51, idx_bytes=8):
    """SURVEY.md §8d 'L1 per forward' minimum: feature read + logits out + weights once (states
    assumed L2-resident across iterations) + the int64 rel_inds the boundary receives."""
    w_mp = 4 * (2 * (2 * 3 * H * H + 2 * 3 * H) + 4 * (2 * H + 1))
    w_unary = 2 * 4 * (H * D + H)
    w_heads = 4 * (n_cls * H + n_cls + n_rel * H + n_rel)
    return 4 * D * (N + E) + 4 * (n_cls * N + n_rel * E) + idx_bytes * 2 * E + w_mp + w_unary + w_heads


def alg_flops_l1(N, E, H, D, T, n_cls=151, n_rel=51):
    """SURVEY.md §8d canonical flop count (no linearity shortcut)."""
    gru = 2 * (H * 3 * H) * 2
    return (2 * D * H * (N + E) + (gru // 2) * (N + E) + T * (gru * (N + E) + 8 * H * 2 * E)
            + 2 * H * (n_cls * N + n_rel * E))


class ClockSampler(object):
    """nvidia-smi sampler (B200_PROFILING.md 'clocks line') running during the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def make_case(args, rank, slot):
    seed = 1234 + 2 + 1000 * rank + slot
    g = synth.synth_graph(args.batch, args.boxes, args.edges, seed)
    N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
    of, ef = synth.synth_l1_feats(N, E, seed)
    return dict(N=N, E=E, obj=of, edge=ef, rel=np.ascontiguousarray(g['rel_inds'][:, 1:3]))


def cpu_baseline_subprocess(args, timeout_s=100):
    """Run the CPU-baseline leg in a fresh process (no CUDA context, clean OpenMP state) under a hard timeout, so a
    slow or wedged host-thread pool can never stall the benchmark itself."""
    cmd = [sys.executable, os.path.abspath(__file__), '--cpu-baseline-only', '--batch', str(args.batch), '--boxes',
           str(args.boxes), '--edges', str(args.edges), '--iters', str(args.iters)]
    env = dict(os.environ); env['CUDA_VISIBLE_DEVICES'] = ''
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)
        return {'value': None, 'unit': 'images/s', 'cores': None, 'kind': 'port', 'sample': 'failed: ' + r.stderr[-200:]}
    except subprocess.TimeoutExpired:
        return {'value': None, 'unit': 'images/s', 'cores': None, 'kind': 'port',
                'sample': 'timed out after %d s (see --impl reference for the reference arm)' % timeout_s}


def cpu_pick_threads(args, cases, params):
    """The reference's CPU path gets the thread count it runs FASTEST with on this host (intra-op
    oversubscription on a many-core box makes "all threads" much slower than a moderate count)."""
    from oracle.imp_torch_cpu import ImpCpu
    ncpu = os.cpu_count() or 1
    cands = sorted(set(t for t in (ncpu, ncpu // 2, 64, 32, 16, 8) if 1 <= t <= ncpu))
    m = ImpCpu(mp_iter=args.iters).load_numpy(params).eval()
    o, e, r = (torch.from_numpy(cases[0][k]) for k in ('obj', 'edge', 'rel'))
    best, best_t = None, None
    with torch.no_grad():
        for t in cands:               # ascending; stop as soon as more threads make it clearly slower
            torch.set_num_threads(t)
            t0 = time.perf_counter(); m.l1_forward(o, e, r); dt = time.perf_counter() - t0
            if dt < 3.0:                  # second (warm) pass only when the first was not hopeless
                t0 = time.perf_counter(); m.l1_forward(o, e, r); dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, best_t = dt, t
            elif dt > 1.5 * best:
                break
    return best_t, cands


def cpu_reference_run(args, cases, params, steps, warmup, threads):
    """The reference's PyTorch-CPU path (oracle/imp_torch_cpu.py port: same op sequence incl. the dense
    [N,E] incidence matmuls), eval mode, no_grad."""
    from oracle.imp_torch_cpu import ImpCpu
    torch.set_num_threads(threads)
    m = ImpCpu(mp_iter=args.iters).load_numpy(params).eval()
    ins = [(torch.from_numpy(c['obj']), torch.from_numpy(c['edge']), torch.from_numpy(c['rel'])) for c in cases]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            o, e, r = ins[i % len(ins)]
            t0 = time.perf_counter()
            m.l1_forward(o, e, r)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times


def measure_other_configs(args, dev, dparams):
    """BASELINE.json configs[2] and configs[4] on one GPU (rank 0 only, a few hundred ms each), so that the driver's
    record carries them too:
      cfg3: SGCls-shaped batch of 32 images with the VGG16 RoIAlign feature head — RoIAlign of objects + union boxes on a
            [32,512,38,38] feature map, union-box geometry, fc6/fc7 of both heads, then the L1 path (the frozen conv
            stack in front of it is torchvision/cuDNN library code and is not timed here);
      cfg5: dense-graph stress, 8 images x 64 boxes x 2000 candidate edges, 6 message-passing iterations (L1 boundary)."""
    from sgg_b200 import ops
    out = {}

    def timeit(fn, reps):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # ---- cfg5
    try:
        g = synth.synth_graph(8, 64, 2000, 1239)
        N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
        of, ef = synth.synth_l1_feats(N, E, 1239)
        o, e = torch.from_numpy(of).to(dev), torch.from_numpy(ef).to(dev)
        rel = torch.from_numpy(np.ascontiguousarray(g['rel_inds'][:, 1:3])).to(dev)
        plan = ops.L1Plan(dparams, N, E, 4096, 6, dev)
        ms = timeit(lambda: plan.run(o, e, rel), 10)
        out['cfg5_dense_stress'] = {'workload': 'L1, 8 img x 64 boxes x 2000 edges, 6 MP iters (N=%d, E=%d)' % (N, E),
                                    'ms_per_step': ms, 'images_per_s': 8 / (ms * 1e-3),
                                    'fp32_equiv_tflops': alg_flops_l1(N, E, 512, 4096, 6) / (ms * 1e-3) / 1e12}
        del plan, o, e
    except Exception as ex:
        out['cfg5_dense_stress'] = {'error': str(ex)[:200]}
    # ---- cfg3
    try:
        B = 32
        g = synth.synth_graph(B, 30, 300, 77)
        N, E = g['boxes'].shape[0], g['rel_inds'].shape[0]
        gen = torch.Generator(device=dev).manual_seed(0)
        fmap = torch.relu(torch.randn(B, 512, 38, 38, device=dev, generator=gen))
        rois = torch.from_numpy(np.ascontiguousarray(g['rois'])).to(dev)
        rel = torch.from_numpy(np.ascontiguousarray(g['rel_inds'])).to(dev)
        P = dict(dparams)

        def u(*shape, fan):
            return (torch.rand(*shape, device=dev, generator=gen) * 2 - 1) / fan ** 0.5
        for pre in ('roi_fmap.1.', 'roi_fmap_obj.'):
            P[pre + '0.weight'] = u(4096, 25088, fan=25088); P[pre + '0.bias'] = u(4096, fan=25088)
            P[pre + '3.weight'] = u(4096, 4096, fan=4096); P[pre + '3.bias'] = u(4096, fan=4096)
        gp = synth.synth_params(5, level='l2')
        for k, v in gp.items():
            if k.startswith('union_boxes.'):
                P[k] = torch.from_numpy(v).to(dev)
        gr = ops.build_graph(rel[:, 1:3], N)
        st = {}

        def step():
            geom = ops.union_geom(rois, rel[:, 1:3], P)
            # RoIAlign emits the fp16 operand planes of its rows, fc6 those of its output: both fc layers of both heads
            # run on pre-split operands (csrc/lin16p.cu), exactly what RelModelStanford.forward does in eval mode
            _, _, npl, epl = ops.node_edge_features(fmap, rois, rel[:, 1:3], edge_add=geom, planes='only')
            h, hpl = ops.linear(None, P['roi_fmap.1.0.weight'], P['roi_fmap.1.0.bias'], relu=True, x_planes=epl, out_planes=True)
            e4096 = ops.linear(h, P['roi_fmap.1.3.weight'], P['roi_fmap.1.3.bias'], x_planes=hpl)
            hn, hnpl = ops.linear(None, P['roi_fmap_obj.0.weight'], P['roi_fmap_obj.0.bias'], relu=True, x_planes=npl,
                                  out_planes=True)
            n4096 = ops.linear(hn, P['roi_fmap_obj.3.weight'], P['roi_fmap_obj.3.bias'], relu=True, x_planes=hnpl)
            return ops.l1_forward(n4096, e4096, gr, P, 3)
        step()
        ms = timeit(step, 5)
        nf, ef2, npl, epl = ops.node_edge_features(fmap, rois, rel[:, 1:3], planes=True)
        st['fc6_edge_ms'] = timeit(lambda: ops.linear(ef2.view(E, -1), P['roi_fmap.1.0.weight'], P['roi_fmap.1.0.bias'], relu=True,
                                                      x_planes=epl), 3)
        st['fc6_edge_fp32_input_ms'] = timeit(lambda: ops.linear(ef2.view(E, -1), P['roi_fmap.1.0.weight'],
                                                                 P['roi_fmap.1.0.bias'], relu=True), 3)
        st['roi_align_ms'] = timeit(lambda: ops.node_edge_features(fmap, rois, rel[:, 1:3], planes='only'), 3)
        out['cfg3_feature_head'] = {'workload': 'B=32 x 30 boxes x 300 edges: RoIAlign (objects + union boxes) + geometry + '
                                                'fc6/fc7 (both heads) + L1, fmap [32,512,38,38] resident (N=%d, E=%d)' % (N, E),
                                    'ms_per_step': ms, 'images_per_s': B / (ms * 1e-3), 'stage_ms': st,
                                    'fc6_edge_fp32_equiv_tflops': 2.0 * E * 25088 * 4096 / (st['fc6_edge_ms'] * 1e-3) / 1e12}
        del P, fmap, nf, ef2
    except Exception as ex:
        out['cfg3_feature_head'] = {'error': str(ex)[:200]}
    # ---- cfg3 end to end through the drop-in class: forward(batch) with HOST images in, the 5 numpy arrays out
    try:
        from sgg_b200 import trainstep
        from sgg_b200.model import RelModelStanford
        B = 8
        torch.manual_seed(0)
        with torch.device(dev):
            m = RelModelStanford(train_data=trainstep.FakeData(), mode='sgcls').eval()
        g = synth.synth_graph(B, 30, 300, 78)
        imgs = [torch.rand(3, 592, 592).pin_memory() for _ in range(B)]
        batch = [(imgs, None, 0, torch.from_numpy(np.ascontiguousarray(g['boxes'])), torch.from_numpy(np.ascontiguousarray(g['gt_classes'])),
                  None, None, None)]
        with torch.no_grad():
            res = m(batch)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 4
            for _ in range(reps):
                res = m(batch)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x = torch.rand(B, 3, 608, 608, device=dev)
            m._backbone(x); a.record(); m._backbone(x); b.record(); torch.cuda.synchronize()
        out['cfg3_l3_forward_e2e'] = {
            'workload': 'RelModelStanford(mode=sgcls).forward(batch), eval: %d host images 3x592x592 -> transform -> VGG16 conv '
                        'stack (tcgen05) -> RoIAlign -> geometry -> fc6/fc7 -> IMP (all %d ordered pairs) -> ranking -> 5 numpy arrays'
                        % (B, res[3].shape[0]),
            'ms_per_batch': dt * 1e3, 'images_per_s': B / dt, 'h2d_bytes_per_batch': B * 3 * 592 * 592 * 4,
            'd2h_bytes_per_batch': int(sum(np.asarray(r).nbytes for r in res)),
            'backbone_ms': a.elapsed_time(b), 'backbone_fp32_equiv_tflops': 2 * 15.35e9 * (608 / 224) ** 2 * B / (a.elapsed_time(b) * 1e-3) / 1e12}
        del m, x
    except Exception as ex:
        out['cfg3_l3_forward_e2e'] = {'error': '%s: %s' % (type(ex).__name__, str(ex)[:200])}
    torch.cuda.empty_cache()
    return out


def measure_train(args, dev, rank, world, dist, barrier):
    """BASELINE.json configs[3]: the PredCls TRAIN step (sgg_b200/trainstep.py) with ``--train-batch`` images per GPU,
    sharded over the ranks, gradients all-reduced with NCCL.  Runs on every rank; returns the record (rank 0 prints it)."""
    from sgg_b200 import trainstep

    def rmax(x):
        if dist is None:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K = max(2, args.train_steps)
    ts = trainstep.TrainStep(dev, B=args.train_batch, boxes=args.boxes, edges=args.edges, mp_iter=args.iters, rank=rank)
    log('train: model built (%.1f M trainable parameters), N=%d E=%d per rank' % (ts.trainable / 1e6, ts.N, ts.E))
    first = float(ts.step(0))                                   # also learns the gradient arrival order (re-layout)
    ms = rmax(trainstep.timed_steps(ts, K, 3, barrier))
    last = float(ts.last_loss)
    log('train: %.2f ms/step' % ms)
    rec = {'workload': 'PredCls train step from the predict boundary (pooled 512x7x7 object / union-box features -> geometry '
                       'branch + fc6/fc7 -> unary -> %d x message passing -> heads -> node + edge CE losses -> backward -> '
                       'gradient all-reduce -> global-norm clip -> SGD momentum), %d images/GPU x %d boxes x %d edges; frozen '
                       'detector (no gradients, no communication) outside the step'
                       % (args.iters, args.train_batch, args.boxes, args.edges),
           'n_gpus': world, 'images_per_gpu': args.train_batch, 'steps': K, 'warmup': 3,
           'ms_per_step': ms, 'images_per_s': args.train_batch * world / (ms * 1e-3), 'scaling': 'weak',
           'trainable_params': ts.trainable, 'allreduce_bytes_per_step': ts.trainable * 4 if world > 1 else 0,
           'buckets': len(ts.red.buckets), 'loss_first': first, 'loss_last': last,
           'finite': bool(np.isfinite([first, last]).all())}
    if world > 1:
        ts.red.comm_enabled = False                             # same step without the collectives: what overlap must hide
        ms_nc = rmax(trainstep.timed_steps(ts, K, 1, barrier))
        ts.red.comm_enabled = True
        ar = trainstep.allreduce_probe(ts.trainable * 4, dev)
        exposed = max(0.0, ms - ms_nc)
        rec.update({'ms_per_step_no_collectives': ms_nc, 'exposed_comm_ms': exposed,
                    'allreduce_standalone': ar,
                    'comm_hidden_frac': max(0.0, 1.0 - exposed / ar['ms']) if ar['ms'] > 0 else None,
                    'note': 'buckets are all-reduced in place on the flat gradient buffer while backward runs; fc6 weight '
                            'gradients are produced and reduced in row chunks'})
    # ---- end to end: every step uploads its pooled features from pinned host memory (copy stream, double-buffered
    # device slots) and reads the loss back
    d0 = ts.batches[0][0]
    keys = ('node_feat', 'edge_feat', 'rois', 'rel_inds', 'obj_labels', 'rel_labels')
    host = {k: d0[k].cpu().pin_memory() for k in keys}
    slots = [{k: torch.empty_like(d0[k]) for k in keys} for _ in range(2)]
    copy_s = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    h_loss = torch.zeros(1, dtype=torch.float32).pin_memory()
    h2d = sum(host[k].numel() * host[k].element_size() for k in keys)

    def upload(i):
        sl = i % 2
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(freed[sl])
            for k in keys:
                slots[sl][k].copy_(host[k], non_blocking=True)
            ready[sl].record(copy_s)

    def run(n):
        cur = torch.cuda.current_stream(dev)
        for sl in range(2):
            freed[sl].record(cur)
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            cur.wait_event(ready[i % 2])
            loss = ts.step(i, inputs=slots[i % 2])
            freed[i % 2].record(cur)
            h_loss.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.synchronize()

    run(2)
    barrier()
    t0 = time.perf_counter()
    run(K)
    barrier()
    e2e_s = rmax(time.perf_counter() - t0)
    rec['e2e'] = {'images_per_s': args.train_batch * world * K / e2e_s, 'ms_per_step': e2e_s * 1e3 / K,
                  'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                  'api': 'sgg_b200.trainstep.TrainStep.step on host-resident pooled features (pinned, uploaded on a copy '
                         'stream one step ahead), loss read back every step'}
    log('train e2e: %.2f ms/step' % (e2e_s * 1e3 / K))
    del ts, slots, host
    torch.cuda.empty_cache()
    return rec


def main():
    args = parse()
    if os.environ.get('SGG_BENCH_WATCHDOG'):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['SGG_BENCH_WATCHDOG']), exit=True)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    H, D = 512, 4096
    workload = ('PredCls L1 batch=%d, %d boxes/img, %d edges/img, 4096-d feats, %d MP iters'
                % (args.batch, args.boxes, args.edges, args.iters))
    config = {'workload': workload, 'batch_per_gpu': args.batch, 'boxes_per_img': args.boxes,
              'edges_per_img': args.edges, 'mp_iter': args.iters, 'boundary': 'L1 (4096-d features -> dists)',
              'parallelism': 'dp%d (independent image shards, no data-path collective)' % max(world, 1),
              'l2_policy': 'CUDA arm: inputs cycle through a ring larger than L2 (features are read from HBM every step)'}
    run_info = {}
    if args.impl == 'native' and not args.cpu_baseline_only:
        torch.set_num_threads(min(8, os.cpu_count() or 1))     # host side of the CUDA arm is plumbing only
    params = synth.synth_params(111, level='l1')

    if args.cpu_baseline_only:
        cases = [make_case(args, 0, s) for s in range(2)]
        threads, cands = cpu_pick_threads(args, cases, params)
        tms = cpu_reference_run(args, cases, params, 5, 1, threads)
        print(json.dumps({'value': args.batch / float(np.median(tms)), 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                          'sample': '5 steps (+1 warm-up) of the same workload, median; oracle/imp_torch_cpu.py in a subprocess; '
                                    'host has %d cpus, thread count auto-tuned over %s (fastest used)' % (os.cpu_count() or 1, cands)}))
        return 0

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == 'reference':
        if rank != 0:
            return 0
        import faulthandler
        faulthandler.dump_traceback_later(540, exit=True)      # last-resort guard: never hang the driver
        cases = [make_case(args, 0, s) for s in range(2)]
        threads, cands = cpu_pick_threads(args, cases, params)
        times = cpu_reference_run(args, cases, params, args.steps, args.warmup, threads)
        ms = float(np.mean(times) * 1e3)
        val = args.batch / (ms / 1e3)
        line = {'impl': 'reference', 'metric': 'images/sec (PredCls, 3 MP iters)', 'value': val, 'unit': 'images/s',
                'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                                 'sample': '%d steps of the full workload (one batch of %d images each); host has %d '
                                           'cpus, thread count auto-tuned over %s (fastest used)'
                                           % (args.steps, args.batch, os.cpu_count() or 1, cands)},
                'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ native arm (CUDA)
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl native needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    from sgg_b200 import ops, _lib
    from sgg_b200.runner import ImpL1Runner
    lib = _lib.load()
    log('library loaded, tensor-core engine %s' % ops.tc_engine())

    info = (torch.cuda.get_device_properties(dev))
    l2 = getattr(info, 'L2_cache_size', L2_BYTES_FALLBACK) or L2_BYTES_FALLBACK
    c0 = make_case(args, rank, 0)
    N, E = c0['N'], c0['E']
    in_bytes = (N + E) * D * 4
    ring = max(2, int(np.ceil(1.5 * l2 / in_bytes)))          # input ring > L2 so features come from HBM
    cases = [c0] + [make_case(args, rank, s) for s in range(1, ring)]
    run_info['l2_policy'] = 'input ring of %d x %.1f MB > %.0f MB L2 (features read from HBM every step)' % (
        ring, in_bytes / 1e6, l2 / 1e6)

    log('synthetic cases ready (ring=%d)' % ring)
    dparams = {k: torch.from_numpy(v).to(dev) for k, v in params.items()}
    d_in = [(torch.from_numpy(c['obj']).to(dev), torch.from_numpy(c['edge']).to(dev),
             torch.from_numpy(c['rel']).to(dev)) for c in cases]
    plan = ops.L1Plan(dparams, N, E, D, args.iters, dev)

    def step_eager(i):
        o, e, r = d_in[i % ring]
        plan.run(o, e, r)            # one C-ABI call per step: graph index + unary + message passing + heads

    # CUDA graphs: one captured step per ring slot (static shapes), replayed in the loop.
    launches_per_step = None
    cuda_graphs = None
    if not args.no_graph:
        try:
            for i in range(ring):
                step_eager(i)
            torch.cuda.synchronize()
            cuda_graphs = []
            side = torch.cuda.Stream()
            for i in range(ring):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    n0 = lib.sgg_launch_count()
                    step_eager(i)
                    launches_per_step = lib.sgg_launch_count() - n0
                cuda_graphs.append(gr)
            run_info['launch'] = 'cuda-graph replay (1 graph/step, %d kernels)' % launches_per_step
        except Exception as ex:   # capture unsupported: fall back to eager launches (still the CUDA path)
            cuda_graphs = None
            run_info['launch'] = 'eager (graph capture failed: %s)' % str(ex)[:80]
    if cuda_graphs is None:
        n0 = lib.sgg_launch_count(); step_eager(0); launches_per_step = lib.sgg_launch_count() - n0
        run_info.setdefault('launch', 'eager')

    def step(i):
        if cuda_graphs is not None:
            cuda_graphs[i % ring].replay()
        else:
            step_eager(i)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    log('launch mode: %s' % run_info.get('launch'))
    # ---- device-resident throughput ("value")
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    log('device-resident: %.3f ms/step' % ms_step)

    # ---- end-to-end through the host-buffer API ("e2e"): pinned host inputs, H2D + D2H inside the timed region
    runner = ImpL1Runner(dparams, N, E, args.iters, slots=3, device=dev)
    h2d_bytes, d2h_bytes = runner.h2d_bytes, runner.d2h_bytes
    h_in = [(torch.from_numpy(c['obj']).pin_memory(), torch.from_numpy(c['edge']).pin_memory(),
             torch.from_numpy(c['rel']).pin_memory()) for c in cases[:max(2, min(ring, 4))]]
    for i in range(max(args.warmup, 3)):
        runner.wait(runner.submit(*h_in[i % len(h_in)]))
    barrier()
    t0 = time.perf_counter()
    pending = []
    for i in range(args.steps):
        pending.append(runner.submit(*h_in[i % len(h_in)]))
        if len(pending) >= 3:
            runner.wait(pending.pop(0))
    while pending:
        runner.wait(pending.pop(0))
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if sampler is not None else None
    log('e2e: %.3f ms/step' % (e2e_s * 1e3 / args.steps))
    # ---- the box's host->device ceiling: every rank streams a pinned 256 MB buffer with plain cudaMemcpyAsync at the same
    # time (what the e2e numbers above are bounded by: one PCIe link per GPU, shared host memory / root complexes)
    hbuf = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
    dbuf = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    dbuf.copy_(hbuf, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        dbuf.copy_(hbuf, non_blocking=True)
    torch.cuda.synchronize()
    h2d_plain = 4 * hbuf.numel() * 4 / (time.perf_counter() - t0) / 1e9
    if dist is not None:
        t = torch.tensor([h2d_plain], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        h2d_plain = float(t.item())
    del hbuf, dbuf

    # ---- BASELINE.json configs[3]: the train step with the gradient all-reduce (every rank takes part).  It runs LAST
    # (after rank 0 has taken its single-GPU stage timings) so that a failure in it can never cost the headline line.
    def run_train():
        if args.workload != 'train' and args.no_train:
            return None
        torch.cuda.empty_cache()
        try:
            return measure_train(args, dev, rank, world, dist, barrier)
        except Exception as ex:
            if args.workload == 'train':
                raise
            log('train step failed: %s: %s' % (type(ex).__name__, str(ex)[:300]))
            return {'error': '%s: %s' % (type(ex).__name__, str(ex)[:300])}

    del runner
    if rank != 0:
        run_train()
        try:
            dist.destroy_process_group()
        except Exception:
            pass
        return 0

    # ---- roofline of the dominant stage, timed live with CUDA events on the launching stream
    hbm_peak, tf_peak, peak_src = peaks()
    g0 = ops.build_graph(d_in[0][2], N)
    stages = {}

    def time_stage(name, fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        stages[name] = a.elapsed_time(b) / reps

    def time_stage_graph(name, fn, reps=20):
        """one kernel launch per fn() call: ``reps`` launches captured into a CUDA graph, so the figure is device time per
        launch (an eager loop of ~15 us kernels would measure the host's launch path instead)"""
        fn(); torch.cuda.synchronize()
        side = torch.cuda.Stream()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            for _ in range(reps):
                fn()
        for _ in range(2):
            gr.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            gr.replay()
        b.record(); torch.cuda.synchronize()
        stages[name] = a.elapsed_time(b) / (5 * reps)

    obj_rep = ops.linear(d_in[0][0], dparams['obj_unary.weight'], dparams['obj_unary.bias'])
    rel_rep = ops.linear(d_in[0][1], dparams['edge_unary.weight'], dparams['edge_unary.bias'], relu=True)
    k = [0]

    def unary():
        k[0] += 1
        o, e, _ = d_in[k[0] % ring]
        ops.linear(e, dparams['edge_unary.weight'], dparams['edge_unary.bias'], relu=True)

    time_stage('edge_unary_linear', unary)
    time_stage('message_pass_T%d' % args.iters, lambda: ops.message_pass(rel_rep, obj_rep, g0, dparams, args.iters))
    time_stage('graph_build', lambda: ops.build_graph(d_in[0][2], N))
    # dominant kernels in isolation, ONE launch each, CUDA events on the launching stream (sgg_mp_probe_launch):
    #   launch B = k_mp_gru: the edge-GRU and node-GRU tiles of one message-passing iteration (T launches per step, + INIT)
    #   launch A = k_mp_pre: P / Q GEMM tiles + vertex-context gather (T launches per step)
    w_mp = 4 * (2 * (2 * 3 * H * H + 2 * 3 * H) + 4 * (2 * H + 1))
    mp_bytes_iter = 4 * H * 2 * (N + E) + 8 * 2 * E      # SURVEY section 8d bytes_iter without weights
    mp_bytes = args.iters * mp_bytes_iter + w_mp + 4 * H * 2 * (N + E)   # + initial-step read/write
    unary_bytes = 4 * D * E + 4 * (H * D + H) + 4 * H * E
    gru_flops = 2 * (H * 3 * H) * 2
    mp_flops = (gru_flops // 2) * (N + E) + args.iters * (gru_flops * (N + E) + 8 * H * 2 * E)
    unary_flops = 2 * D * H * E
    engine = ops.tc_engine()
    fused = True
    try:
        probe = ops.MpProbe(rel_rep, obj_rep, g0, dparams, args.iters)
        time_stage_graph('mp_launch_B_gru', lambda: probe.launch(2))
        time_stage_graph('mp_launch_A_pre', lambda: probe.launch(1))
        time_stage_graph('mp_launch_init', lambda: probe.launch(0))
    except Exception as ex:                     # tc32 / simt engines: the 7-launch schedule, time its edge-GRU kernel
        fused = False
        log('fused probe unavailable (%s); timing the stand-alone edge-GRU kernel' % str(ex)[:80])
        P_t = ops.linear(obj_rep, dparams['edge_gru.weight_ih'])
        gates_t = torch.rand(E, 4, device=dev)
        time_stage('mp_launch_B_gru', lambda: ops.edge_gru(rel_rep, P_t, gates_t, g0, dparams), reps=30)
    # launch B, algorithmic (fp32, unsplit operands): states read + written, int32 endpoints, the P rows
    # (edge side) and Q rows + ctx (node side), both GRU weight matrices + biases, gate partials
    gruB_bytes = (4 * H * 2 * (N + E) + 8 * E + 4 * 3 * H * N * 2 + 4 * H * N + 2 * 4 * (3 * H * H + 6 * H) + 2 * 7 * 16 * (N + E))
    gruB_flops = 2 * 3 * H * H * (N + E) + 2 * 2 * 3 * H * E           # both GEMMs + the gated gather combine
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get('k_mp_gru_dram_bytes_per_launch')
    t_b = stages['mp_launch_B_gru'] * 1e-3
    mma_tf = 3 * gruB_flops / t_b / 1e12                              # 3 MMA passes per fp32-equivalent flop
    roofline = {'bound': 'tensor',
                'kernel': ('sgg::mpf::k_mp_gru (launch B: edge + node GRU tiles of one iteration; %d launches/step + the INIT '
                           'launch)' % args.iters) if fused else 'edge-GRU kernel of the 7-launch schedule',
                'engine': engine,
                'achieved': mma_tf, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': mma_tf / tf_peak if tf_peak else None,
                'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_flops': gruB_flops, 'algorithmic_bytes': gruB_bytes, 'ms_per_launch': stages['mp_launch_B_gru'],
                'note': 'achieved = fp16 MMA-pass TFLOP/s (3 tcgen05 passes per fp32-equivalent flop: the 3xFP16 operand split '
                        'that buys fp32-grade results) against the measured dense bf16/fp16 peak; the compulsory DRAM traffic '
                        'of the launch is ~7 MB (weights + first touch, states are L2-resident), so HBM is not the binding term',
                'hbm': {'achieved_gbs': gruB_bytes / t_b / 1e9, 'peak_gbs': hbm_peak, 'frac': gruB_bytes / t_b / 1e9 / hbm_peak},
                'fp32_equiv_tflops': gruB_flops / t_b / 1e12,
                'stage_ms': stages,
                'stage_tflops_fp32_equiv': {'edge_unary_linear': unary_flops / (stages['edge_unary_linear'] * 1e-3) / 1e12,
                                            'message_pass': mp_flops / (stages['message_pass_T%d' % args.iters] * 1e-3) / 1e12},
                'stage_algorithmic_gbs': {'edge_unary_linear': unary_bytes / (stages['edge_unary_linear'] * 1e-3) / 1e9,
                                          'message_pass': mp_bytes / (stages['message_pass_T%d' % args.iters] * 1e-3) / 1e9},
                'whole_step': {'algorithmic_bytes': alg_bytes_l1(N, E, H, D, args.iters),
                               'achieved_gbs': alg_bytes_l1(N, E, H, D, args.iters) / (ms_step * 1e-3) / 1e9,
                               'frac_of_hbm_peak': alg_bytes_l1(N, E, H, D, args.iters) / (ms_step * 1e-3) / 1e9 / hbm_peak,
                               'algorithmic_flops': alg_flops_l1(N, E, H, D, args.iters),
                               'fp32_equiv_tflops': alg_flops_l1(N, E, H, D, args.iters) / (ms_step * 1e-3) / 1e12,
                               'mma_pass_tflops_frac_of_peak': 3 * alg_flops_l1(N, E, H, D, args.iters) / (ms_step * 1e-3) / 1e12 / tf_peak if tf_peak else None}}

    # ---- CPU baseline: the reference's CPU path (oracle port), bounded sample, rank 0 only at N=1
    log('stage timings done: %s' % {k: round(v, 4) for k, v in stages.items()})
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_subprocess(args)
        log('cpu baseline done')

    other = None
    if world == 1 and not args.no_other_configs:
        other = measure_other_configs(args, dev, dparams)
        log('other configs: %s' % {k: round(v.get('ms_per_step', -1), 3) for k, v in other.items()})
    train = run_train()
    images = args.batch * world
    line = {'metric': 'images/sec (PredCls, 3 MP iters)', 'value': images / (ms_step * 1e-3), 'unit': 'images/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config, 'clocks': clocks,
            'e2e': {'value': images * args.steps / e2e_s, 'unit': 'images/s', 'h2d_bytes_per_step': h2d_bytes,
                    'd2h_bytes_per_step': d2h_bytes, 'ms_per_step': e2e_s * 1e3 / args.steps,
                    'h2d_gbs_per_rank': h2d_bytes * args.steps / e2e_s / 1e9,
                    'h2d_plain_memcpy_gbs_per_rank': h2d_plain,
                    'h2d_note': 'h2d_plain_memcpy = all ranks copying a pinned 256 MB buffer concurrently with cudaMemcpyAsync '
                                '(slowest rank): the host->device ceiling of this box at this rank count; the e2e step is '
                                'bound by it (%.0f %% of it reached)' % (100.0 * h2d_bytes * args.steps / e2e_s / 1e9 / h2d_plain),
                    'api': 'sgg_b200.runner.ImpL1Runner.submit/wait (3 in-flight slots, pinned host buffers)'},
            'gpu_launches': int(launches_per_step) * args.steps,
            'gpu_launches_per_step': int(launches_per_step),
            'run': run_info, 'roofline': roofline, 'cpu_baseline': cpu, 'train_step': train, 'other_configs': other}
    if args.workload == 'train':
        # headline = BASELINE.json configs[3]; the L1 forward numbers stay in the line under 'l1_forward'
        line['l1_forward'] = {k: line[k] for k in ('value', 'ms_per_step', 'e2e', 'gpu_launches_per_step')}
        line['config'] = dict(config, workload=train['workload'], batch_per_gpu=args.train_batch,
                              boundary='predict (pooled features -> dists) + losses + backward + all-reduce + clip + SGD',
                              parallelism='dp%d (images sharded by rank, NCCL all-reduce on gradients only)' % world)
        line.update({'value': train['images_per_s'], 'ms_per_step': train['ms_per_step'], 'steps': train['steps'],
                     'warmup': train['warmup'],
                     'e2e': {'value': train['e2e']['images_per_s'], 'unit': 'images/s',
                             'h2d_bytes_per_step': train['e2e']['h2d_bytes_per_step'],
                             'd2h_bytes_per_step': train['e2e']['d2h_bytes_per_step'],
                             'ms_per_step': train['e2e']['ms_per_step'], 'api': train['e2e']['api']}})
    print(json.dumps(line))
    sys.stdout.flush()
    if dist is not None:
        try:
            dist.destroy_process_group()
        except Exception:
            pass
    return 0


if __name__ == '__main__':
    sys.exit(main())
