"""Time the assign kernel alone (CUDA events around N back-to-back launches) for a workload and patch shapes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from retinanet_tensorflow_b200 import synthetic
from retinanet_tensorflow_b200.pipeline import DenseHeadPath

names = sys.argv[1].split(',') if len(sys.argv) > 1 else ['coco']
patches = [tuple(int(x) for x in p.split('x')) for p in (sys.argv[2].split(',') if len(sys.argv) > 2 else ['0x0'])]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
for name in names:
    wl = synthetic.WORKLOADS[name]
    B = wl.batch
    boxes, classes, counts = synthetic.gt_batch(wl, B, 0)
    gb, gc, gn = (torch.from_numpy(x).cuda() for x in (boxes, classes, counts))
    for patch in patches:
        path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, B, wl.max_gt, process_group=False,
                             assign_patch=patch)
        for _ in range(5):
            path.assign(gb, gc, gn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            path.assign(gb, gc, gn)
        e1.record()
        torch.cuda.synchronize()
        print('%s patch %s: %.4f ms per assign' % (name, patch, e0.elapsed_time(e1) / reps), flush=True)
        del path
