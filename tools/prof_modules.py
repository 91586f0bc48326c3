"""Host-overhead profile of the per-image drop-in calls (dataset.build_labels, utils.boxes_decode, utils.nms_classwise)."""
import cProfile
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from retinanet_tensorflow_b200 import dataset, synthetic, utils
from retinanet_tensorflow_b200.levels import build_levels
from retinanet_tensorflow_b200.pipeline import DenseHeadPath

wl = synthetic.WORKLOADS['coco']
B = 8
levels = build_levels()
names = list(levels)
boxes, classes, counts = synthetic.gt_batch(wl, B, 0)
gt = [(torch.from_numpy(boxes[i, :counts[i]]).cuda(), torch.from_numpy(classes[i, :counts[i]]).cuda()) for i in range(B)]
path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, B, wl.max_gt, process_group=False)
path.assign(*(torch.from_numpy(x).cuda() for x in (boxes, classes, counts)))
logits, regs = synthetic.head_outputs(path.label, path.deltas, wl.num_classes, seed=1, device='cuda')
probs = path.per_level(utils.sigmoid(logits), (wl.num_classes,))
dec = path.per_level(regs, (4,))
size = [wl.image_h, wl.image_w]


def labels():
    return [dataset.build_labels(size, c, b, levels, wl.num_classes) for b, c in gt]


def detect():
    n = 0
    for i in range(B):
        d = [utils.boxes_decode(probs[pn][i], dec[pn][i]) for pn in names]
        n += int(utils.nms_classwise(utils.merge_boxes_decoded(d), num_classes=wl.num_classes).scores.shape[0])
    return n


for fn in (labels, detect):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    print('%s: %.3f ms per image' % (fn.__name__, 1e3 * (time.perf_counter() - t0) / 5 / B))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(14)
