"""Small end-to-end run for compute-sanitizer (not collected by pytest):

    compute-sanitizer --tool memcheck  python tests/sanitize_check.py
    compute-sanitizer --tool racecheck python tests/sanitize_check.py

Exercises every kernel family on small shapes: assign, fused/unfused losses in all modes, dice
sums, Huber, select (+fused), finalize with dedup and top-k, NMS, flip, box-map ops, dense labels,
elementwise losses."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from retinanet_tensorflow_b200 import augmentation, dataset, losses, synthetic, utils
from retinanet_tensorflow_b200.levels import build_levels
from retinanet_tensorflow_b200.pipeline import DenseHeadPath


def main():
    for name, classes in (('shapes', 3), ('voc', 20)):
        wl = synthetic.WORKLOADS[name]
        wl = synthetic.Workload(wl.name, wl.config, 192, 256, classes, wl.max_gt, 2)
        boxes, cls, counts = synthetic.gt_batch(wl, 2)
        gt = [torch.from_numpy(x).cuda() for x in (boxes, cls, counts)]
        for mode in ('focal', 'bce_dice', 'balanced_bce', 'focal_dice'):
            path = DenseHeadPath((wl.image_h, wl.image_w), classes, 2, wl.max_gt, cls_mode=mode, capacity=4096,
                                 topk_per_level=50, keep_iou=True, process_group=False)
            path.assign(*gt)
            logits, regs = synthetic.head_outputs(path.label, path.deltas, classes, seed=1, device='cuda')
            logits.view(-1, classes)[3, :min(3, classes)] = 3.0
            path.step(*gt, logits, regs)
            path.step(*gt, logits, regs, overlap_detect=True)
            path.join()
            path.loss(logits, regs)
            path.detect(logits, regs)
            augmentation.flip_compact(path.pyr, 2, path.label, path.deltas)
            torch.cuda.synchronize()
            path.check_status(ignore=12)
    # bfloat16 logits in every kernel that takes them (fused loss + select, dice sums, standalone select), and dense
    # firing into a small candidate list (warp-aggregated appends, full lists, overflow flag)
    wl = synthetic.WORKLOADS['voc']
    wl = synthetic.Workload('voc24', wl.config, 160, 192, 24, wl.max_gt, 2)
    boxes, cls, counts = synthetic.gt_batch(wl, 2)
    gt = [torch.from_numpy(x).cuda() for x in (boxes, cls, counts)]
    for mode in ('focal', 'bce_dice'):
        path = DenseHeadPath((wl.image_h, wl.image_w), 24, 2, wl.max_gt, cls_mode=mode, capacity=64, score_thr=0.05,
                             process_group=False)
        path.assign(*gt)
        logits, regs = synthetic.head_outputs(path.label, path.deltas, 24, seed=2, device='cuda')
        l16 = (logits + 3.0).bfloat16()
        path.step(*gt, l16, regs)
        path.detect(l16, regs)
        path.step(*gt, logits + 3.0, regs)
        torch.cuda.synchronize()
        path.check_status(ignore=12)
    # round 2 (later): assignment with shared CTA patches (ragged decompositions, table and per-anchor encoding) and the
    # one-launch detection tail (dedup, several classes per anchor, per-class limits)
    wl = synthetic.Workload('voc16', 2, 200, 264, 16, 70, 3)
    boxes, cls, counts = synthetic.gt_batch(wl, 3)
    gt = [torch.from_numpy(v).cuda() for v in (boxes, cls, counts)]
    for patch, cap, per_class in (((2, 2), 4096, 1000), ((3, 2), 1024, 2), ((1, 1), 8192, 1000)):
        path = DenseHeadPath((wl.image_h, wl.image_w), 16, 3, wl.max_gt, capacity=cap, max_output_per_class=per_class,
                             keep_iou=True, process_group=False, assign_patch=patch)
        path.assign(*gt)
        logits, regs = synthetic.head_outputs(path.label, path.deltas, 16, seed=6, device='cuda', logit_sigma=1.6)
        logits.view(-1, 16)[7, 0:9] = 3.0
        assert path.tail_fused_supported()
        path.step(*gt, logits, regs)
        path.step(*gt, logits, regs, overlap_detect=True)
        path.join()
        torch.cuda.synchronize()
        path.check_status(ignore=12)
    b, c = synthetic.gt_for_image(synthetic.WORKLOADS['shapes'], 0)
    dataset.build_labels([128, 160], c, b, build_levels(), 3)
    rng = np.random.default_rng(0)
    z = (rng.uniform(size=(64, 5)) > 0.9).astype(np.float32)
    x = torch.tensor(rng.normal(-2, 2, (64, 5)).astype(np.float32), device='cuda', requires_grad=True)
    total = (losses.focal_sigmoid_cross_entropy_with_logits(z, x).sum() + losses.dice_loss(z, x, axis=0).sum() +
             losses.classification_loss(z, x, z.max(-1) > 0.5) + losses.focal_softmax_cross_entropy_with_logits(z, x).sum())
    total.backward()
    probs = rng.uniform(0, 0.7, (6, 7, 9, 4)).astype(np.float32)
    bx = np.concatenate([rng.uniform(0, .4, (6, 7, 9, 2)), rng.uniform(.5, 1, (6, 7, 9, 2))], -1).astype(np.float32)
    utils.nms_classwise(utils.boxes_decode(probs, bx), 4)
    utils.iou(bx, bx)
    # round 2: dense labels through the streaming kernel (C % 4 == 0), trainable masking by the compaction kernels,
    # an assignment launched one step ahead, the Shapes rasteriser and the box drawing
    z8 = (rng.uniform(size=(700, 8)) > 0.97).astype(np.float32)
    x8 = torch.tensor(rng.normal(-3, 2, (700, 8)).astype(np.float32), device='cuda', requires_grad=True)
    (losses.classification_loss(z8, x8, z8.max(-1) > 0.5, mode='focal') +
     losses.classification_loss(z8, x8, z8.max(-1) > 0.5, mode='bce_dice')).backward()
    wl = synthetic.Workload('voc8', 2, 160, 192, 8, 20, 2)
    boxes, cls, counts = synthetic.gt_batch(wl, 2)
    gt = [torch.from_numpy(v).cuda() for v in (boxes, cls, counts)]
    path = DenseHeadPath((wl.image_h, wl.image_w), 8, 2, wl.max_gt, capacity=512, process_group=False)
    path.assign(*gt)
    logits, regs = synthetic.head_outputs(path.label, path.deltas, 8, seed=4, device='cuda')
    for _ in range(3):
        path.step(*gt, logits, regs, overlap_detect=True)
        path.prefetch_assign(*gt)
    path.join()
    levels = build_levels()
    per = [dataset.build_labels([wl.image_h, wl.image_w], torch.from_numpy(cls[i, :counts[i]]),
                                torch.from_numpy(boxes[i, :counts[i]]), levels, 8) for i in range(2)]
    feats = {'image': torch.empty((2, wl.image_h, wl.image_w, 0), device='cuda'),
             'detection': {'classifications': {k: torch.stack([p[0][k] for p in per]) for k in levels},
                           'regressions': {k: torch.stack([p[1][k] for p in per]) for k in levels}},
             'trainable_masks': {k: torch.stack([p[2][k] for p in per]) for k in levels}}
    head = {'detection': {'classifications': path.per_level(logits, (8,)), 'regressions': path.per_level(regs, (4,))}}
    lab, log = utils.process_labels_and_logits(feats, head, levels)
    losses.loss(lab['detection_trainable'], log['detection_trainable'], mode='focal')
    log['detection_trainable'].classification.prob
    from retinanet_tensorflow_b200 import synthetic_shapes
    data = synthetic_shapes.Shapes(image_size=(120, 160)).batch(3)
    synthetic_shapes.draw_bounding_boxes(data['image'], data['gt_boxes'], data['gt_classes'].clamp(min=0), data['gt_counts'],
                                         num_classes=3)
    torch.cuda.synchronize()
    print('SANITIZE_RUN_DONE')


if __name__ == '__main__':
    main()
