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
    ap.add_argument('--pipeline', action='store_true',
                    help='launch the assignment one step ahead on a second stream (DenseHeadPath.prefetch_assign); '
                         'measured: no gain on one GPU, the persistent loss kernel leaves no room on the SMs '
                         '(profiles/r02_schedule.md)')
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
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --batch images per GPU (default); strong: the workload batch is the GLOBAL batch')
    ap.add_argument('--api', default='auto', choices=['auto', 'modules', 'none'],
                    help='modules: also time the reference call chain through the drop-in modules (default at N=1)')
    ap.add_argument('--quick', action='store_true', help='headline numbers only (no sigma-1.5 / strong / modules legs)')
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
        'images_per_step_timed': n_images,   # the bounded sample: `config` names the workload, not this sample size
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


def ncu_kernels(key):
    """Pipe utilisation per stage from the committed ncu capture (profiles/ncu_kernels.json), or {}."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_kernels.json')) as f:
            return json.load(f).get(key) or {}
    except Exception:
        return {}


def bind_to_gpu_numa_node(dev_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE pinned host memory is allocated
    (first touch decides where the pages live; a remote node halves H2D bandwidth on a 2-socket box).
    Best effort: returns the node number or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(dev_index)
        bus = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
        return node
    except Exception:
        return None


class Harness(object):
    """One DenseHeadPath + its resident inputs + the timing loop (used for the headline run and the variants)."""

    def __init__(self, args, wl, batch, rank, world, dev, *, sigma, group, capacity=8192, seed_base=1000, first_image=None):
        import torch
        from retinanet_tensorflow_b200 import synthetic
        from retinanet_tensorflow_b200.pipeline import DenseHeadPath
        self.args, self.wl, self.rank, self.world, self.dev = args, wl, rank, world, dev
        self.src_batch = batch
        self.batch = 2 * batch if args.flip else batch
        self.path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, self.batch, wl.max_gt,
                                  cls_mode=args.cls_mode, capacity=capacity, process_group=group, device=dev,
                                  exchange=args.exchange)
        first = rank * batch if first_image is None else first_image
        self.gt_host = synthetic.gt_batch(wl, batch, first_image=first)
        self.gt = [torch.from_numpy(x).to(dev) for x in self.gt_host]
        self.path.assign(*self.gt, with_flip=args.flip)
        self.logits, self.regs = synthetic.head_outputs(self.path.label, self.path.deltas, wl.num_classes,
                                                         seed=seed_base + rank, device=dev, logit_sigma=sigma)
        if args.logits_dtype == 'bf16':
            self.logits = self.logits.bfloat16()
        torch.cuda.synchronize()
        self.stage_events = None
        self.graph = None
        self.overlap = not args.no_overlap
        # several GPUs: launching the assignment one step ahead takes its count exchange off the critical path (+1.6 % at
        # N = 2); one GPU: no gain (profiles/r02_schedule.md), so it stays opt-in there
        self.pipeline = bool(args.pipeline) or world > 1

    def reseed(self, sigma, seed_base=1000):
        """Same labels, new head outputs with another background spread (in place of the old ones)."""
        import torch
        from retinanet_tensorflow_b200 import synthetic
        self.path.join()
        torch.cuda.synchronize()
        self.logits = self.regs = None
        self.logits, self.regs = synthetic.head_outputs(self.path.label, self.path.deltas, self.wl.num_classes,
                                                         seed=seed_base + self.rank, device=self.dev, logit_sigma=sigma)
        if self.args.logits_dtype == 'bf16':
            self.logits = self.logits.bfloat16()

    def hook_stages(self):
        import torch
        self.stage_events = {k: [] for k in ('assign', 'cls_sums', 'cls_loss', 'select', 'nms', 'tail')}
        self.recording = False

        def timed(name, fn):
            if not self.recording:
                return fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            self.stage_events[name].append((a, b))
            return out
        self.path.stage_hook = timed

    def one_step(self):
        if self.graph is not None:
            self.graph.replay()
            return
        self.path.step(*self.gt, self.logits, self.regs, overlap_detect=self.overlap, with_flip=self.args.flip)
        if self.pipeline:   # the NEXT step's assignment starts now, beside this step's loss kernel
            self.path.prefetch_assign(*self.gt, with_flip=self.args.flip)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(self, steps, warmup):
        """W untimed + K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks -> ms."""
        import torch
        import torch.distributed as dist
        for _ in range(warmup):
            self.one_step()
        self.barrier()
        if self.stage_events is not None:
            for v in self.stage_events.values():
                del v[:]
            self.recording = True
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            self.one_step()
        self.path.join()          # the timed region ends when the last step's detections are done too
        end.record()
        self.barrier()
        self.recording = False
        t = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def stage_ms(self):
        return {k: (sum(a.elapsed_time(b) for a, b in v) / len(v) if v else None) for k, v in self.stage_events.items()}

    def cls_bytes(self):
        esz = 2 if self.args.logits_dtype == 'bf16' else 4
        return self.batch * self.path.anchors * (2 * esz * self.wl.num_classes + 4)


def sharded_equals_single(args, wl, rank, world, dev):
    """Driver-run multi-GPU correctness: every rank runs a small shard (4 images) with the statistics exchange;
    rank 0 also runs the whole 4 x N batch alone.  Global normalisers => same losses, and for rank 0's images
    bit-identical gradients and detections."""
    import torch
    import torch.distributed as dist
    from retinanet_tensorflow_b200 import synthetic
    b = 4
    sub = argparse.Namespace(**vars(args))
    sub.flip = False
    h = Harness(sub, wl, b, rank, world, dev, sigma=1.0, group=None, seed_base=2000)
    h.pipeline = False
    h.one_step()
    h.path.join()
    torch.cuda.synchronize()
    out = {'images_per_gpu': b}
    if rank == 0:
        names = h.path.names
        lg, rg, gts = {pn: [] for pn in names}, {pn: [] for pn in names}, [[], [], []]
        for r in range(world):
            hr = h if r == 0 else Harness(sub, wl, b, r, 1, dev, sigma=1.0, group=False, seed_base=2000)
            C = wl.num_classes
            pl = hr.path.per_level(hr.logits, (C,))
            pr = hr.path.per_level(hr.regs, (4,))
            for pn in names:
                lg[pn].append(pl[pn].clone())
                rg[pn].append(pr[pn].clone())
            for k in range(3):
                gts[k].append(hr.gt[k])
        whole = Harness(sub, wl, b * world, 0, 1, dev, sigma=1.0, group=False, seed_base=2000)
        whole.pipeline = False
        whole.gt = [torch.cat(g, 0) for g in gts]
        whole.logits = {pn: torch.cat(v, 0) for pn, v in lg.items()}
        whole.regs = {pn: torch.cat(v, 0) for pn, v in rg.items()}
        whole.one_step()
        whole.path.join()
        torch.cuda.synchronize()
        a, w_ = h.path.loss_out.double(), whole.path.loss_out.double()
        rel = float(((a - w_).abs() / w_.abs().clamp(min=1e-30)).max())
        C = wl.num_classes
        gs = h.path.per_level(h.path.dlogits, (C,))
        gw = whole.path.per_level(whole.path.dlogits, (C,))
        grad_equal = all(torch.equal(gs[pn], gw[pn][:b]) for pn in names)
        det_equal = bool(torch.equal(h.path.det_count, whole.path.det_count[:b]))
        n0 = int(h.path.det_count[0])
        det_equal = det_equal and bool(torch.equal(h.path.det_boxes[0, :n0], whole.path.det_boxes[0, :n0]))
        out.update(loss_max_rel_diff=rel, gradients_bit_identical=bool(grad_equal), detections_identical=det_equal,
                   ok=bool(rel <= 1e-6 and grad_equal and det_equal))
    if world > 1:
        dist.barrier()
    return out


def run_modules(args, wl, dev, harness):
    """`--api modules`: the reference's own call chain through the drop-in modules, per image where the reference
    works per image: dataset.build_labels (dataset.py:167) -> batch -> utils.process_labels_and_logits
    (train.py:218) -> losses.loss(mode='focal') (train.py:220) -> boxes_decode per level + nms_classwise per image
    (train.py:68-75).  Dense one-hot labels, masked [M, C] tensors, host round-trips for the data-dependent
    shapes -- what the reference's API shape costs."""
    import ctypes as C
    import torch
    from retinanet_tensorflow_b200 import _lib, dataset, losses, utils
    from retinanet_tensorflow_b200.levels import build_levels
    levels = build_levels()
    names = list(levels.keys())
    B, Cn = harness.batch, wl.num_classes
    boxes_h, classes_h, counts_h = harness.gt_host
    gt = [(torch.from_numpy(boxes_h[i, :counts_h[i]]).to(dev), torch.from_numpy(classes_h[i, :counts_h[i]]).to(dev))
          for i in range(B)]
    logits = harness.path.per_level(harness.logits.float(), (Cn,))
    regs = harness.path.per_level(harness.regs, (4,))
    image = torch.empty((B, wl.image_h, wl.image_w, 0), device=dev)
    size = [wl.image_h, wl.image_w]
    times = {}

    def section(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        times[name] = times.get(name, 0.0) + time.perf_counter() - t0
        return out

    def step(detect_images):
        per_image = section('build_labels', lambda: [dataset.build_labels(size, c, b, levels, Cn) for b, c in gt])
        feats = section('batch', lambda: {
            'image': image,
            'detection': {'classifications': {pn: torch.stack([p[0][pn] for p in per_image]) for pn in names},
                          'regressions': {pn: torch.stack([p[1][pn] for p in per_image]) for pn in names}},
            'trainable_masks': {pn: torch.stack([p[2][pn] for p in per_image]) for pn in names}})
        del per_image
        head = {'detection': {'classifications': logits, 'regressions': regs}}
        lab, log = section('process_labels_and_logits', lambda: utils.process_labels_and_logits(feats, head, levels))

        def loss_fb():
            x = log['detection_trainable'].classification.unscaled.requires_grad_(True)
            r = log['detection_trainable'].regression.requires_grad_(True)
            lt = lab['detection_trainable']
            gl = utils.Detection(utils.Classification(x, None), r, None)
            cl, rl = losses.loss(lt, gl, mode='focal')
            (cl + rl).backward()
            return float(cl.detach()), float(rl.detach()), x, lt.classification.prob
        cl, rl, x, z = section('loss_fwd_bwd', loss_fb)

        def detect():
            det = log['detection']
            n = 0
            for i in range(detect_images):
                dec = [utils.boxes_decode(det.classification.prob[pn][i], det.regression_postprocessed[pn][i]) for pn in names]
                n += int(utils.nms_classwise(utils.merge_boxes_decoded(dec), num_classes=Cn).scores.shape[0])
            return n
        ndet = section('decode_nms', detect)
        return cl, rl, ndet, x, z

    # warm-up: two full steps on a defragmented allocator (the 4 GB masked tensors would otherwise cost a cudaMalloc
    # each in the timed steps, depending on what the earlier legs left in torch's cache); then the median of 3
    torch.cuda.empty_cache()
    step(B)
    step(B)
    steps = 3
    runs = []
    for _ in range(steps):
        times.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cl, rl, ndet, x, z = step(B)
        torch.cuda.synchronize()
        runs.append((time.perf_counter() - t0, dict(times)))
    runs.sort(key=lambda r: r[0])
    dt, med_times = runs[len(runs) // 2]
    times = {k: v * steps for k, v in med_times.items()}   # (reported below as per-step values)
    # the dense-label loss kernel alone (read logits + read labels + write gradient = 12 B per logit)
    M = int(x.shape[0])
    xd, zd = x.detach().contiguous(), z.contiguous()
    g = torch.empty_like(xd)
    stats = torch.zeros((_lib.RN_STATS_HEADER + 3 * Cn,), dtype=torch.float64, device=dev)
    _lib.call('rn_label_stats_dense', M, Cn, _lib.ptr(zd), _lib.ptr(stats), _lib.stream_ptr())
    nb = _lib.lib.rn_loss_workspace_bytes(Cn)
    ws = torch.zeros((nb,), dtype=torch.uint8, device=dev)
    seg = _lib.rows_of([M])

    def kernel():
        _lib.call('rn_cls_loss_fwd_bwd', C.byref(seg), Cn, C.byref(_lib.maps_of([xd])), C.byref(_lib.maps_of([zd])), 1,
                  _lib.RN_CLS_FOCAL, 2.0, 0.25, 1e-7, C.byref(_lib.maps_of([g])), _lib.RN_DTYPE_F32, _lib.ptr(stats),
                  _lib.ptr(ws), nb, _lib.stream_ptr())
    for _ in range(3):
        kernel()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        kernel()
    b.record()
    torch.cuda.synchronize()
    k_ms = a.elapsed_time(b) / 10
    peak, _ = measured_peaks()
    alg = 12.0 * M * Cn
    return {'value': B / dt, 'unit': UNIT, 'ms_per_step': 1e3 * dt, 'steps': steps, 'images': B,
            'timing': 'median of %d steps after 2 warm-up steps, host wall clock with a device synchronize per stage' % steps,
            'stage_ms': {k: 1e3 * v / steps for k, v in times.items()},
            'dense_label_loss_kernel': {'ms': k_ms, 'alg_bytes_per_launch': alg, 'achieved_GBps': alg / (k_ms * 1e-3) / 1e9,
                                        'frac': alg / (k_ms * 1e-3) / 1e9 / peak, 'rows': M,
                                        'bytes_per_logit': 12},
            'frac': alg / (k_ms * 1e-3) / 1e9 / peak,
            'check': {'class_loss': cl, 'regr_loss': rl, 'detections': ndet},
            'calls': 'dataset.build_labels x %d images -> torch.stack -> utils.process_labels_and_logits -> '
                     'losses.loss(mode=focal) + backward -> utils.boxes_decode x %d levels + utils.nms_classwise per image'
                     % (B, len(names))}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from retinanet_tensorflow_b200 import _lib, synthetic

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
    if args.scaling == 'strong' and world > 1:
        batch = max(1, batch // world)

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
    numa_node = bind_to_gpu_numa_node(torch.cuda.current_device())

    # nvidia-smi starts now (it needs seconds to come up on an 8-GPU box); only the samples inside the warm-up +
    # timed window are used
    sampler = ClockSampler('GPU-' + str(torch.cuda.get_device_properties(dev).uuid))
    if rank == 0:
        sampler.start()

    # ---- inputs resident in HBM
    group = None if world > 1 else False
    hz = Harness(args, wl, batch, rank, world, dev, sigma=args.logit_sigma, group=group)
    path = hz.path
    batch = hz.batch

    if args.mode == 'infer':
        # BASELINE configs[3]: inference post-processing only (standalone select -> top-k -> NMS)
        path.topk_per_level = args.topk
        t_load0 = time.time()
        for _ in range(args.warmup):
            path.detect(hz.logits, hz.regs)
        hz.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            path.detect(hz.logits, hz.regs)
        b.record()
        hz.barrier()
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
    hz.hook_stages()
    if args.graph:
        hz.graph = path.capture(*hz.gt, hz.logits, hz.regs, with_flip=args.flip)
        hz.overlap = hz.pipeline = False

    # extra untimed warm-up until the clock sampler delivers (rank 0 decides, every rank runs the same steps)
    warmup_extra = 0
    for _ in range(200):
        ready = torch.tensor([1 if (rank != 0 or sampler.ready()) else 0], device=dev)
        if world > 1:
            dist.broadcast(ready, src=0)
        if int(ready.item()):
            break
        for _ in range(25):
            hz.one_step()
        warmup_extra += 25
        torch.cuda.synchronize()
    t_load0 = time.time()
    for _ in range(max(args.warmup - 1, 0)):
        hz.one_step()
    hz.barrier()
    launches0 = _lib.lib.rn_launch_count()
    ms_total = hz.run(args.steps, 1)
    launches = _lib.lib.rn_launch_count() - launches0
    launches = launches * args.steps // (args.steps + 1)   # run() does one more (untimed) step inside the counter window
    if hz.graph is not None:
        # kernels inside a replayed graph are not visible to the host-side counter or the stage events: count
        # and time them on a few eager steps after the timed region
        l0 = _lib.lib.rn_launch_count()
        hz.recording = True
        for _ in range(10):
            path.step(*hz.gt, hz.logits, hz.regs, with_flip=args.flip)
        torch.cuda.synchronize()
        hz.recording = False
        launches = (_lib.lib.rn_launch_count() - l0) // 10 * args.steps
    clocks = sampler.stop(t_load0, time.time()) if rank == 0 else None
    path.check_status(ignore=_lib.RN_FLAG_CAND_OVERFLOW | _lib.RN_FLAG_DET_OVERFLOW)
    loss_vals = path.loss_out.cpu().tolist()
    det_total = int(path.det_count.sum().item())
    cand_total = int(path.cand_count.sum().item())
    stage_ms = hz.stage_ms()
    cls_bytes = hz.cls_bytes()

    # ---- the same step with a wider background-logit spread (SURVEY 8d's sigma = 1.5): the dominant kernel must
    # not depend on the logits being confident negatives
    sigma15 = None
    if not args.quick and hz.graph is None and abs(args.logit_sigma - 1.5) > 1e-6:
        hz.reseed(1.5)
        n15 = max(10, min(args.steps, 30))
        ms15 = hz.run(n15, 3)
        st15 = hz.stage_ms()
        path.check_status(ignore=_lib.RN_FLAG_CAND_OVERFLOW | _lib.RN_FLAG_DET_OVERFLOW)
        sigma15 = {'ms_per_step': ms15 / n15, 'value': world * batch * n15 / (ms15 * 1e-3), 'steps': n15,
                   'ms_per_launch': st15['cls_loss'],
                   'candidates_per_image': int(path.cand_count.sum().item()) / batch}
        hz.reseed(args.logit_sigma)

    # ---- multi-GPU extras: strong scaling (fixed global batch) and sharded == single
    strong = check_multi = None
    if world > 1 and not args.quick and args.scaling != 'strong':
        sb = max(1, (wl.batch if args.batch is None else args.batch) // world)
        hs = Harness(args, wl, sb, rank, world, dev, sigma=args.logit_sigma, group=None)
        ns = max(20, min(args.steps, 100))
        ms_s = hs.run(ns, 5)
        strong_graph = None
        if hs.batch <= 16:
            # a few images per GPU: the step is launch / latency bound -> replay it as ONE CUDA graph (the statistics
            # exchange is capturable: its epochs live in device memory); every rank must take the same decision
            ok = torch.tensor([1], device=dev)
            try:
                hs.path.join()
                hs.graph = hs.path.capture(*hs.gt, hs.logits, hs.regs, with_flip=args.flip)
            except Exception:
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()):
                ms_g = hs.run(ns, 5)
                strong_graph = {'value': world * hs.batch * ns / (ms_g * 1e-3), 'ms_per_step': ms_g / ns}
        strong = {'global_batch': sb * world * (2 if args.flip else 1), 'images_per_gpu': hs.batch,
                  'value': world * hs.batch * ns / (ms_s * 1e-3), 'ms_per_step': ms_s / ns, 'steps': ns,
                  'as_one_cuda_graph': strong_graph,
                  'note': 'BASELINE configs[2] read as a fixed global batch sharded by image; latency-bound at a few '
                          'images per GPU (launches + one statistics exchange per step)'}
        del hs
        check_multi = sharded_equals_single(args, wl, rank, world, dev)

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region.
    # Two device input sets: the H2D copy of step i+1 (copy stream) overlaps the kernels of step i.
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 10))
    logits, regs = hz.logits, hz.regs
    h_logits = torch.empty(logits.shape, dtype=logits.dtype, pin_memory=True)
    h_regs = torch.empty(regs.shape, dtype=regs.dtype, pin_memory=True)
    h_logits.copy_(logits)
    h_regs.copy_(regs)
    h_gt = [torch.from_numpy(x).pin_memory() for x in hz.gt_host]
    h_out = {
        'loss': torch.empty((4,), dtype=torch.float32, pin_memory=True),
        'det_boxes': torch.empty(path.det_boxes.shape, dtype=torch.float32, pin_memory=True),
        'det_scores': torch.empty(path.det_scores.shape, dtype=torch.float32, pin_memory=True),
        'det_classes': torch.empty(path.det_classes.shape, dtype=torch.int32, pin_memory=True),
        'det_count': torch.empty(path.det_count.shape, dtype=torch.int32, pin_memory=True),
    }
    h2d = sum(x.numel() * x.element_size() for x in [h_logits, h_regs] + h_gt)
    d2h = sum(x.numel() * x.element_size() for x in h_out.values())
    sets = [(logits, regs, [torch.empty_like(g) for g in hz.gt]),
            (torch.empty_like(logits), torch.empty_like(regs), [torch.empty_like(g) for g in hz.gt])]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [None, None]
    main = torch.cuda.current_stream()
    path.join()
    torch.cuda.synchronize()

    def upload(k):
        lg, rg, g = sets[k]
        with torch.cuda.stream(copy_stream):
            if consumed[k] is not None:
                copy_stream.wait_event(consumed[k])
            for dst, src in zip(g, h_gt):
                dst.copy_(src, non_blocking=True)
            lg.copy_(h_logits, non_blocking=True)
            rg.copy_(h_regs, non_blocking=True)
            copied[k].record(copy_stream)

    def e2e_run(n):
        upload(0)
        for i in range(n):
            k = i & 1
            if i + 1 < n:
                upload(k ^ 1)            # next step's inputs travel while this step computes
            main.wait_event(copied[k])
            lg, rg, g = sets[k]
            path.step(*g, lg, rg, with_flip=args.flip)
            consumed[k] = torch.cuda.Event()
            consumed[k].record(main)
            h_out['loss'].copy_(path.loss_out, non_blocking=True)
            h_out['det_boxes'].copy_(path.det_boxes, non_blocking=True)
            h_out['det_scores'].copy_(path.det_scores, non_blocking=True)
            h_out['det_classes'].copy_(path.det_classes, non_blocking=True)
            h_out['det_count'].copy_(path.det_count, non_blocking=True)

    e2e_run(2)
    hz.barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # the host->device ceiling of this box for the same bytes, all ranks copying at the same time
    hz.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        sets[1][0].copy_(h_logits, non_blocking=True)
        sets[1][1].copy_(h_regs, non_blocking=True)
    torch.cuda.synchronize()
    ceil_s = (time.perf_counter() - t0) / 3
    te = torch.tensor([e2e_s, ceil_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, ceil_s = float(te[0]), float(te[1])
    ceiling_gbs = (h_logits.numel() * h_logits.element_size() + h_regs.numel() * 4) / ceil_s / 1e9
    e2e_gbs = h2d * e2e_steps / e2e_s / 1e9
    del sets, h_logits, h_regs

    # ---- the reference's own call chain through the drop-in modules (single GPU)
    api_modules = None
    if world == 1 and rank == 0 and (args.api == 'modules' or (not args.quick and args.api != 'none')) \
            and args.logits_dtype == 'f32' and not args.flip and hz.graph is None:
        try:
            api_modules = run_modules(args, wl, dev, hz)
        except torch.cuda.OutOfMemoryError as err:   # dense [B*A, C] labels + masked copies need ~6x the logits
            api_modules = {'unavailable': 'out of memory: %s' % (str(err).splitlines()[0],)}

    if rank == 0:
        A = path.anchors
        C = wl.num_classes
        peak, peak_src = measured_peaks()
        bytes_per_kernel = {
            'assign': batch * (24 * A + 20 * wl.max_gt),
            'cls_sums': batch * A * (4 * C + 4),
            'cls_loss': cls_bytes + batch * A * 16,   # + the regression gradient the same launch writes
            'select': None if path.fused_select_supported() else batch * A * 4 * C,  # fused: finalize only
            'nms': None,
            'tail': None,   # candidate finalize + class-wise NMS in one launch (rn_detect_tail)
        }
        kernels = {}
        for k, v in stage_ms.items():
            if v:
                gbs = bytes_per_kernel[k] / (v * 1e-3) / 1e9 if bytes_per_kernel.get(k) else None
                kernels[k] = {'ms': v, 'alg_GB': (bytes_per_kernel[k] or 0) / 1e9, 'GBps': gbs,
                              'frac_of_hbm': (gbs / peak if gbs else None)}
        for k, v in ncu_kernels('%s_b%d_%s' % (wl.name, batch, args.cls_mode)).items():
            if k in kernels and args.logits_dtype == 'f32':
                kernels[k]['ncu'] = v
        if hz.pipeline and 'assign' in kernels:
            kernels['assign']['note'] = 'runs one step ahead on its own stream, sharing the SMs with the loss kernel'
        achieved = cls_bytes / (stage_ms['cls_loss'] * 1e-3) / 1e9 if stage_ms['cls_loss'] else None
        value = world * batch * args.steps / (ms_total * 1e-3)
        kname = 'cls_tile_kernel (fused sigmoid-focal fwd+bwd + Huber fwd+bwd + candidate pre-selection + final scalars)'
        roofline = {'bound': 'hbm', 'kernel': kname, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': (achieved / peak) if achieved else None,
                    'traffic': ncu_traffic('%s_b%d_%s' % (wl.name, batch, args.cls_mode)),
                    'alg_bytes_per_launch': cls_bytes, 'ms_per_launch': stage_ms['cls_loss'],
                    'peak_source': peak_src,
                    'note': 'achieved counts the classification bytes only (read logits + write dlogits + read label); '
                            'the same launch also moves the regression loss bytes (20 B per anchor)'}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'warmup_extra': warmup_extra, 'ms_per_step': ms_total / args.steps,
            'higher_is_better': True,
            'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'f32' if args.logits_dtype == 'f32' else 'f32 math on bf16 logits',
            'data': 'synthetic',
            'config': workload_config(wl, args, batch),
            'clocks': clocks,
            'e2e': {'value': world * batch * e2e_steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
                    'h2d_gbs_per_gpu': e2e_gbs, 'h2d_ceiling_gbs': ceiling_gbs,
                    'frac_of_ceiling': e2e_gbs / ceiling_gbs if ceiling_gbs else None, 'numa_node': numa_node,
                    'note': 'pinned host buffers -> H2D of GT + logits + regression preds (copy stream, double-buffered: '
                            'the next step\'s upload overlaps this step\'s kernels), whole path, D2H of losses + '
                            'detections; bound by the host->device link (%.2f GB in per step); h2d_ceiling_gbs = the '
                            'same bytes copied with nothing else running, all ranks at once' % (h2d / 1e9)},
            'gpu_launches': int(launches),
            'roofline': roofline,
            'kernels': kernels,
            'fused_loss_select': bool(path.fused_select_supported()),
            'stats_exchange': ('peer-memory kernel' if path._peer is not None else 'nccl') if world > 1 else None,
            'detect_tail_on_side_stream': bool(hz.overlap and path.fused_select_supported()),
            'assign_one_step_ahead': bool(hz.pipeline and hz.graph is None),
            'cuda_graph': hz.graph is not None,
            'cpu_baseline': cpu_baseline,
            'check': {'class_loss': loss_vals[0], 'regr_loss': loss_vals[1], 'num_fg': loss_vals[2],
                      'candidates': cand_total, 'detections': det_total},
        }
        if sigma15 is not None:
            a15 = cls_bytes / (sigma15['ms_per_launch'] * 1e-3) / 1e9
            line['roofline_sigma15'] = {'bound': 'hbm', 'kernel': kname, 'achieved': a15, 'peak': peak, 'unit': 'GB/s',
                                        'frac': a15 / peak, 'alg_bytes_per_launch': cls_bytes,
                                        'ms_per_launch': sigma15['ms_per_launch'], 'ms_per_step': sigma15['ms_per_step'],
                                        'value': sigma15['value'], 'steps': sigma15['steps'],
                                        'logits': 'N(-4.595, 1.50^2) + N(6,1) on (positive anchor, class): %.0f candidates '
                                                  'per image above score 0.5' % sigma15['candidates_per_image']}
        if strong is not None:
            line['strong_scaling'] = strong
        if check_multi is not None:
            line['check']['sharded_equals_single'] = check_multi.get('ok')
            line['check']['sharded_detail'] = check_multi
        if api_modules is not None:
            line['api_modules'] = api_modules
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
