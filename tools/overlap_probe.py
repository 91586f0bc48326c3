"""Does the assignment kernel (K1) really run beside the streaming loss kernel (K2)?  Event timestamps of both
streams relative to one origin, for a few pipelined steps.   python tools/overlap_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from retinanet_tensorflow_b200 import synthetic
from retinanet_tensorflow_b200.pipeline import DenseHeadPath

wl = synthetic.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'coco']
B = wl.batch
dev = torch.device('cuda', 0)
path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, B, wl.max_gt, capacity=8192, process_group=False)
gt = [torch.from_numpy(x).to(dev) for x in synthetic.gt_batch(wl, B)]
path.assign(*gt)
logits, regs = synthetic.head_outputs(path.label, path.deltas, wl.num_classes, seed=1, device=dev)
marks = []


def hook(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    marks.append((name, a, b))
    return out


for _ in range(5):
    path.step(*gt, logits, regs, overlap_detect=True)
    path.prefetch_assign(*gt)
torch.cuda.synchronize()
path.stage_hook = hook
origin = torch.cuda.Event(enable_timing=True)
origin.record()
for _ in range(4):
    path.step(*gt, logits, regs, overlap_detect=True)
    path.prefetch_assign(*gt)
path.join()
torch.cuda.synchronize()
for name, a, b in marks:
    print('%-10s start %8.3f ms  end %8.3f ms  (%.3f)' % (name, origin.elapsed_time(a), origin.elapsed_time(b), a.elapsed_time(b)))
