"""Duration of the loss kernel (CUDA events around its launch) in different surroundings: alone, after the
assignment, in the whole step (serial / overlapped tail)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from retinanet_tensorflow_b200 import synthetic
from retinanet_tensorflow_b200.pipeline import DenseHeadPath

wl = synthetic.WORKLOADS['coco']
B = wl.batch
boxes, classes, counts = synthetic.gt_batch(wl, B, 0)
gt = [torch.from_numpy(x).cuda() for x in (boxes, classes, counts)]
path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, B, wl.max_gt, process_group=False, capacity=8192)
path.assign(*gt)
logits, regs = synthetic.head_outputs(path.label, path.deltas, wl.num_classes, seed=1000, device='cuda')
ev = {}


def hook(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    ev.setdefault(name, []).append((a, b))
    return out


def report(tag, n_skip=5):
    torch.cuda.synchronize()
    print(tag, {k: round(sum(a.elapsed_time(b) for a, b in v[n_skip:]) / len(v[n_skip:]), 4) for k, v in ev.items()}, flush=True)
    ev.clear()


path.stage_hook = hook
N = 60


def k2_only():
    for _ in range(N):
        path.loss_and_select(logits, regs, finalize=False)


def k1_k2():
    for _ in range(N):
        path.assign(*gt)
        path.loss_and_select(logits, regs, finalize=False)


def k2_tail():
    for _ in range(N):
        path.loss_and_select(logits, regs, finalize=False)
        path.detect_tail(regs)


def serial():
    for _ in range(N):
        path.step(*gt, logits, regs)


def overlapped():
    for _ in range(N):
        path.step(*gt, logits, regs, overlap_detect=True)
    path.join()


def k1_gap_k2():
    for _ in range(N):
        path.assign(*gt)
        torch.cuda._sleep(200000)   # ~0.1 ms idle gap between the assignment and the loss kernel
        path.loss_and_select(logits, regs, finalize=False)


modes = [k2_only, k1_k2, k2_tail, serial, overlapped, k1_gap_k2]
import subprocess
for rnd in range(3):
    for fn in (modes if rnd % 2 == 0 else modes[::-1]):
        fn()
        clk = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader'],
                             capture_output=True, text=True).stdout.strip()
        report('%d %-12s [%s]' % (rnd, fn.__name__, clk))
