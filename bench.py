#!/usr/bin/env python
"""Benchmark of the dense-head hot path (BASELINE.json metric: images/sec for
assign + loss fwd/bwd + decode/NMS at COCO shape, per B200, and HBM % of peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload coco] [--impl reference]

One "step" = one pass of the whole hot path over one batch of synthetic images per GPU:
K1 assign+encode -> K2/K3 fused focal + Huber loss forward+backward -> K4/K5 decode, select,
class-wise NMS.  Weak scaling: every rank owns `batch` images; the only exchange is the
all-reduce of the float64 statistics vector.  Rank 0 prints ONE JSON line.

`--impl reference` times the CPU restatement of the reference (oracle/rn_oracle.py; TensorFlow
is not installable here) on the host cores for the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'images/sec (assign+loss+decode/NMS, COCO 800x1333) at 1/8 B200; HBM % of peak'
UNIT = 'images/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='coco', choices=['shapes', 'voc', 'coco', 'crowded'])
    ap.add_argument('--batch', type=int, default=None, help='images per GPU (default: the workload batch)')
    ap.add_argument('--cls-mode', default='focal')
    ap.add_argument('--e2e-steps', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--logits-dtype', default='f32', choices=['f32', 'bf16'],
                    help='bf16: extension (SURVEY 8f-3), NOT the headline metric, which stays float32')
    ap.add_argument('--exchange', default='auto', choices=['auto', 'peer', 'nccl'],
                    help='multi-GPU statistics exchange: own kernel over NVLink peer memory, or NCCL')
    ap.add_argument('--mode', default='train_step', choices=['train_step', 'infer'],
                    help='infer: BASELINE config 4 (decode + top-k per level + per-class NMS only; not the headline metric)')
    ap.add_argument('--topk', type=int, default=1000)
    ap.add_argument('--no-overlap', action='store_true', help='keep the detection tail on the main stream')
    ap.add_argument('--cpu-images', type=int, default=None, help='images in the CPU sample (default: one per core)')
    ap.add_argument('--flip', action='store_true',
                    help="the reference's [orig, flipped] batches (dataset.py:182-204): --batch counts source images, the "
                         'step runs on twice as many')
    ap.add_argument('--graph', action='store_true',
                    help='replay the step as one CUDA graph (DenseHeadPath.capture); single GPU, for launch-bound '
                         'small workloads')
    ap.add_argument('--logit-sigma', type=float, default=1.0,
                    help='std of the synthetic background logits around the prior (SURVEY 8d text says 1.5, which makes '
                         '8.5%% of anchors fire instead of the ~1%% it asks for; default 1.0 gives ~1%%)')
    ap.add_argument('--opt', action='append', default=[], help='rn_set_option name=value (tuning knob)')
    ap.add_argument('--ref-budget-s', type=float, default=120.0, help='wall-clock budget of the reference arm')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle (a port of the reference's TF graph code) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_one_image(task):
    """Worker: the reference path for ONE image, dense formulation, float32 NumPy.
    build_labels -> process_labels_and_logits -> loss fwd+bwd -> boxes_decode + nms_classwise."""
    workload, image_index, cls_mode = task
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import numpy as np
    import rn_oracle as oracle
    from retinanet_tensorflow_b200 import synthetic
    wl = synthetic.WORKLOADS[workload]
    boxes, classes = synthetic.gt_for_image(wl, image_index)
    levels = oracle.build_levels()
    image_size = np.array([wl.image_h, wl.image_w], np.int32)
    rng = np.random.default_rng(image_index)
    t0 = time.perf_counter()
    cls, reg, mask = oracle.build_labels(image_size, classes, boxes, levels, wl.num_classes)
    t1 = time.perf_counter()
    logits_c, logits_r = {}, {}
    for pn in levels:  # synthetic head outputs (not timed)
        x = rng.normal(synthetic.PRIOR_LOGIT, 1.0, cls[pn].shape).astype(np.float32)
        x += cls[pn] * rng.normal(6.0, 1.0, cls[pn].shape[:-1] + (1,)).astype(np.float32)
        logits_c[pn] = x[None]
        logits_r[pn] = (reg[pn] + rng.normal(0, 0.1, reg[pn].shape)).astype(np.float32)[None]
    t2 = time.perf_counter()
    labels = {'image': np.zeros((1, wl.image_h, wl.image_w, 1), np.float32),
              'trainable_masks': {k: v[None] for k, v in mask.items()},
              'detection': {'classifications': {k: v[None] for k, v in cls.items()},
                            'regressions': {k: v[None] for k, v in reg.items()}}}
    logits = {'detection': {'classifications': logits_c, 'regressions': logits_r}}
    lab, pred = oracle.process_labels_and_logits(labels, logits, levels)
    lt, pt = lab['detection_trainable'], pred['detection_trainable']
    oracle.loss(lt, pt, mode=cls_mode)
    fg = oracle.classmap_decode(lt.classification.prob).fg_mask
    oracle.classification_loss_grad(lt.classification.prob, pt.classification.unscaled, fg, mode=cls_mode)
    oracle.regression_loss_grad(lt.regression, pt.regression, fg)
    t3 = time.perf_counter()
    det = pred['detection']
    decoded = [oracle.boxes_decode(det.classification.prob[pn][0], det.regression_postprocessed[pn][0]) for pn in levels]
    out = oracle.nms_classwise(oracle.merge_boxes_decoded(decoded), wl.num_classes)
    t4 = time.perf_counter()
    return {'assign': t1 - t0, 'loss': t3 - t2, 'decode_nms': t4 - t3, 'detections': int(out.scores.shape[0])}


def cpu_pass(workload, n_images, cls_mode, pool, first_image=0):
    """One bounded sample: n_images through the oracle on the pool's workers. -> (images/s, stages)."""
    t0 = time.perf_counter()
    res = pool.map(_cpu_one_image, [(workload, first_image + i, cls_mode) for i in range(n_images)])
    dt = time.perf_counter() - t0
    stages = {k: sum(r[k] for r in res) / len(res) for k in ('assign', 'loss', 'decode_nms')}
    return n_images / dt, stages


def make_pool(cores):
    import multiprocessing as mp
    return mp.get_context('spawn').Pool(cores)


def oracle_mode(cls_mode):
    return cls_mode if cls_mode in ('focal', 'bce_dice', 'bce', 'balanced_bce', 'balanced_bce_dice', 'focal_dice') \
        else 'focal'


def run_reference(args):
    """`--impl reference`: the CPU implementation of the path on all host cores.

    One step = a bounded sample of n images of the workload (n <= cores).  A COCO image costs the
    dense NumPy port ~7-10 s on one core, so the sample size is derived from a wall-clock budget
    (--ref-budget-s, default 120 s for the K timed steps): one untimed calibration wave of `cores`
    images (the warm-up), then the K x n timed images are fed to the worker pool back to back --
    the pool pipelines images the way the reference's tf.data map threads do (dataset.py:231)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from retinanet_tensorflow_b200 import synthetic
    wl = synthetic.WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    mode = oracle_mode(args.cls_mode)
    pool = make_pool(cores)
    try:
        t0 = time.perf_counter()
        cpu_pass(args.workload, cores, mode, pool, first_image=0)  # calibration + warm-up wave
        t_wave = time.perf_counter() - t0
        steps = max(args.steps, 1)
        n_images = args.cpu_images or int(max(1, min(cores, args.ref_budget_s * cores / (t_wave * steps))))
        total = n_images * steps
        t0 = time.perf_counter()
        _, stages = cpu_pass(args.workload, total, mode, pool, first_image=cores)
        dt = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    value = total / dt
    sample = ('%d %s images per step (%dx%d, C=%d, <=%d GT) x %d steps, pipelined over %d worker processes; '
              'warm-up = one untimed wave of %d images (%.1f s)' % (
                  n_images, wl.name, wl.image_h, wl.image_w, wl.num_classes, wl.max_gt, steps, cores, cores, t_wave))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(wl, args, wl.batch if args.batch is None else args.batch),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                         'stage_seconds_per_image': stages},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'TensorFlow 1.x is not installable in this image; this arm times oracle/rn_oracle.py, the float32 '
                'NumPy port of the reference graph code, with one process per host core',
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
def workload_config(wl, args, batch):
    return {
        'workload': '%s: %dx%d, %d classes, <=%d GT, P3-P7 x 9 anchors, batch %d images per GPU, %s + smooth-L1 '
                    'fwd+bwd, decode + score>0.5 + per-class NMS IoU 0.5' % (
                        wl.name, wl.image_h, wl.image_w, wl.num_classes, wl.max_gt, batch, args.cls_mode),
        'images_per_gpu': batch, 'cls_mode': args.cls_mode,
        'orig_plus_flipped': bool(getattr(args, 'flip', False)),
        'logits': 'N(-4.595, %.2f^2) + N(6,1) on (positive anchor, class)' % args.logit_sigma,
        'l2': 'inputs (%.1f GB logits per step) exceed the 126 MB L2; no flush needed' % (
            batch * 4e-9 * wl.num_classes * anchors_of(wl)),
    }


def anchors_of(wl):
    import math
    return 9 * sum(math.ceil(wl.image_h / 2 ** l) * math.ceil(wl.image_w / 2 ** l) for l in range(3, 8))


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None

    def ready(self):
        """True once nvidia-smi has printed its first sample (it can take seconds on an 8-GPU box)."""
        try:
            return self.p is None or os.path.getsize(self.f.name) > 0
        except OSError:
            return True

    def _kill(self):
        if self.p is not None and self.p.poll() is None:
            self.p.kill()

    def start(self):
        import atexit
        atexit.register(self._kill)   # never leave an orphan nvidia-smi behind, whatever happens to the bench
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '50'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self, t0=None, t1=None):
        """Median SM clock / throttle reasons of the samples taken in the wall-clock window [t0, t1] (the
        warm-up + timed steps).  nvidia-smi is started long before that window, because on an 8-GPU box it
        needs more than the timed region to print its first line."""
        import datetime
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(', ') for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        if t0 is not None:
            inside = []
            for r in rows:
                try:
                    ts = datetime.datetime.strptime(r[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                except ValueError:
                    continue
                if t0 - 0.05 <= ts <= t1 + 0.05:
                    inside.append(r)
            out['window'] = 'warm-up + timed steps' if inside else 'whole process (no sample fell inside the timed window)'
            rows = inside or rows
        sm, reasons, mx, pw = [], set(), None, []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.strip().lower() == 'active':
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw) if pw else None)
        return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(key):
    """Per-launch DRAM traffic of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def run_b200(args):
    import torch
    import torch.distributed as dist
    from retinanet_tensorflow_b200 import _lib, synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device('cuda', torch.cuda.current_device())

    for kv in args.opt:
        k, v = kv.split('=')
        _lib.call('rn_set_option', k.encode(), int(v))
    wl = synthetic.WORKLOADS[args.workload]
    batch = wl.batch if args.batch is None else args.batch

    # CPU baseline first (rank 0, N=1 only): bounded sample on all host cores
    cpu_baseline = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_img = args.cpu_images or cores
        pool = make_pool(cores)
        try:
            v, stages = cpu_pass(args.workload, n_img, oracle_mode(args.cls_mode), pool)
        finally:
            pool.close()
            pool.join()
        cpu_baseline = {
            'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '%d %s images (one per worker process), dense float32 NumPy port of the reference: '
                      'build_labels + process_labels_and_logits + loss fwd/bwd + boxes_decode + nms_classwise' % (
                          n_img, wl.name),
            'stage_seconds_per_image': stages}

    # nvidia-smi starts now (it needs seconds to come up on an 8-GPU box); only the samples inside the warm-up +
    # timed window are used
    sampler = ClockSampler('GPU-' + str(torch.cuda.get_device_properties(dev).uuid))
    if rank == 0:
        sampler.start()

    # ---- inputs resident in HBM
    src_batch = batch
    if args.flip:
        batch = 2 * src_batch     # images per step: [orig, flipped]
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, batch, wl.max_gt, cls_mode=args.cls_mode,
                         capacity=8192, process_group=None if world > 1 else False, device=dev,
                         exchange=args.exchange)
    gt_boxes_h, gt_classes_h, gt_counts_h = synthetic.gt_batch(wl, src_batch, first_image=rank * src_batch)
    gt_boxes = torch.from_numpy(gt_boxes_h).to(dev)
    gt_classes = torch.from_numpy(gt_classes_h).to(dev)
    gt_counts = torch.from_numpy(gt_counts_h).to(dev)
    path.assign(gt_boxes, gt_classes, gt_counts, with_flip=args.flip)
    logits, regs = synthetic.head_outputs(path.label, path.deltas, wl.num_classes, seed=1000 + rank, device=dev,
                                           logit_sigma=args.logit_sigma)
    if args.logits_dtype == 'bf16':
        logits = logits.bfloat16()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.mode == 'infer':
        # BASELINE configs[3]: inference post-processing only (standalone select -> top-k -> NMS)
        path.topk_per_level = args.topk
        t_load0 = time.time()
        for _ in range(args.warmup):
            path.detect(logits, regs)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            path.detect(logits, regs)
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        clocks = sampler.stop(t_load0, time.time()) if rank == 0 else None
        if rank == 0:
            peak, _ = measured_peaks()
            sel_bytes = batch * path.anchors * 4 * wl.num_classes
            print(json.dumps({
                'metric': 'images/sec (decode + score>0.5 + top-%d per level + per-class NMS IoU 0.5)' % args.topk,
                'value': world * batch * args.steps / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(wl, args, batch), 'clocks': clocks,
                'alg_bytes_per_step': sel_bytes, 'frac_of_hbm_whole_step': sel_bytes / (ms / args.steps * 1e-3) / 1e9 / peak,
                'detections': int(path.det_count.sum().item()), 'candidates': int(path.cand_count.sum().item())}))
        return 0

    # per-kernel CUDA events (on the launching stream) for the roofline object
    stage_events = {k: [] for k in ('assign', 'cls_sums', 'cls_loss', 'huber', 'select', 'nms')}
    record = {'on': False}

    def timed(name, fn):
        if not record['on']:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        stage_events[name].append((a, b))
        return out

    path.stage_hook = timed
    overlap = not args.no_overlap
    graph = None
    if args.graph:
        if world > 1:
            raise SystemExit('--graph is single-GPU')
        graph = path.capture(gt_boxes, gt_classes, gt_counts, logits, regs, with_flip=args.flip)
        overlap = False

    def one_step():
        if graph is not None:
            graph.replay()
        else:
            path.step(gt_boxes, gt_classes, gt_counts, logits, regs, overlap_detect=overlap, with_flip=args.flip)

    # extra untimed warm-up until the clock sampler delivers (rank 0 decides, every rank runs the same steps)
    warmup_extra = 0
    for _ in range(200):
        ready = torch.tensor([1 if (rank != 0 or sampler.ready()) else 0], device=dev)
        if world > 1:
            dist.broadcast(ready, src=0)
        if int(ready.item()):
            break
        for _ in range(25):
            one_step()
        warmup_extra += 25
        torch.cuda.synchronize()
    t_load0 = time.time()
    for _ in range(args.warmup):
        one_step()
    barrier()
    launches0 = _lib.lib.rn_launch_count()
    record['on'] = True
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        one_step()
    path.join()          # the timed region ends when the last step's detections are done too
    end.record()
    barrier()
    launches = _lib.lib.rn_launch_count() - launches0
    if graph is not None:
        # kernels inside a replayed graph are not visible to the host-side counter or the stage events: count
        # and time them on a few eager steps after the timed region
        l0 = _lib.lib.rn_launch_count()
        for _ in range(10):
            path.step(gt_boxes, gt_classes, gt_counts, logits, regs, with_flip=args.flip)
        torch.cuda.synchronize()
        launches = (_lib.lib.rn_launch_count() - l0) // 10 * args.steps
    record['on'] = False
    clocks = sampler.stop(t_load0, time.time()) if rank == 0 else None
    ms = start.elapsed_time(end)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    path.check_status(ignore=_lib.RN_FLAG_CAND_OVERFLOW | _lib.RN_FLAG_DET_OVERFLOW)
    loss_vals = path.loss_out.cpu().tolist()
    det_total = int(path.det_count.sum().item())
    cand_total = int(path.cand_count.sum().item())

    stage_ms = {k: (sum(a.elapsed_time(b) for a, b in v) / len(v) if v else None) for k, v in stage_events.items()}

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 10))
    h_logits = torch.empty(logits.shape, dtype=logits.dtype, pin_memory=True)
    h_regs = torch.empty(regs.shape, dtype=regs.dtype, pin_memory=True)
    h_logits.copy_(logits)
    h_regs.copy_(regs)
    h_gt = [torch.from_numpy(x).pin_memory() for x in (gt_boxes_h, gt_classes_h, gt_counts_h)]
    h_out = {
        'loss': torch.empty((4,), dtype=torch.float32, pin_memory=True),
        'det_boxes': torch.empty(path.det_boxes.shape, dtype=torch.float32, pin_memory=True),
        'det_scores': torch.empty(path.det_scores.shape, dtype=torch.float32, pin_memory=True),
        'det_classes': torch.empty(path.det_classes.shape, dtype=torch.int32, pin_memory=True),
        'det_count': torch.empty(path.det_count.shape, dtype=torch.int32, pin_memory=True),
    }
    d_logits, d_regs = torch.empty_like(logits), torch.empty_like(regs)
    del logits, regs
    h2d = sum(x.numel() * x.element_size() for x in [h_logits, h_regs] + h_gt)
    d2h = sum(x.numel() * x.element_size() for x in h_out.values())

    def e2e_step():
        b_, c_, n_ = (x.to(dev, non_blocking=True) for x in h_gt)
        d_logits.copy_(h_logits, non_blocking=True)
        d_regs.copy_(h_regs, non_blocking=True)
        path.step(b_, c_, n_, d_logits, d_regs, with_flip=args.flip)
        h_out['loss'].copy_(path.loss_out, non_blocking=True)
        h_out['det_boxes'].copy_(path.det_boxes, non_blocking=True)
        h_out['det_scores'].copy_(path.det_scores, non_blocking=True)
        h_out['det_classes'].copy_(path.det_classes, non_blocking=True)
        h_out['det_count'].copy_(path.det_count, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    if rank == 0:
        A = path.anchors
        C = wl.num_classes
        peak, peak_src = measured_peaks()
        esz = 2 if args.logits_dtype == 'bf16' else 4
        cls_bytes = batch * A * (2 * esz * C + 4)    # read logits + write dlogits + read label
        bytes_per_kernel = {
            'assign': batch * (24 * A + 20 * wl.max_gt),
            'cls_sums': batch * A * (4 * C + 4),
            'cls_loss': cls_bytes,
            'huber': batch * A * 20,                  # label + dreg (+32 B per positive, ~1%)
            'select': None if path.fused_select_supported() else batch * A * 4 * C,  # fused: finalize only
            'nms': None,
        }
        kernels = {}
        for k, v in stage_ms.items():
            if v:
                gbs = bytes_per_kernel[k] / (v * 1e-3) / 1e9 if bytes_per_kernel.get(k) else None
                kernels[k] = {'ms': v, 'alg_GB': (bytes_per_kernel[k] or 0) / 1e9, 'GBps': gbs,
                              'frac_of_hbm': (gbs / peak if gbs else None)}
        achieved = cls_bytes / (stage_ms['cls_loss'] * 1e-3) / 1e9 if stage_ms['cls_loss'] else None
        value = world * batch * args.steps / (ms_total * 1e-3)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'warmup_extra': warmup_extra, 'ms_per_step': ms_total / args.steps,
            'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32' if args.logits_dtype == 'f32' else 'f32 math on bf16 logits',
            'data': 'synthetic',
            'config': workload_config(wl, args, batch),
            'clocks': clocks,
            'e2e': {'value': world * batch * e2e_steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
                    'note': 'pinned host buffers -> H2D of GT + logits + regression preds, whole path, D2H of losses '
                            '+ detections; PCIe-bound (%.2f GB in per step)' % (h2d / 1e9)},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'kernel': 'cls_fast_kernel (fused sigmoid-focal fwd+bwd + candidate pre-selection)',
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': (achieved / peak) if achieved else None,
                         'traffic': ncu_traffic('%s_b%d_%s' % (wl.name, batch, args.cls_mode)),
                         'alg_bytes_per_launch': cls_bytes, 'ms_per_launch': stage_ms['cls_loss'],
                         'peak_source': peak_src},
            'kernels': kernels,
            'fused_loss_select': bool(path.fused_select_supported()),
            'stats_exchange': ('peer-memory kernel' if path._peer is not None else 'nccl') if world > 1 else None,
            'detect_tail_on_side_stream': bool(overlap and path.fused_select_supported()),
            'cuda_graph': graph is not None,
            'cpu_baseline': cpu_baseline,
            'check': {'class_loss': loss_vals[0], 'regr_loss': loss_vals[1], 'num_fg': loss_vals[2],
                      'candidates': cand_total, 'detections': det_total},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_b200(args)


if __name__ == '__main__':
    sys.exit(main())
