"""The reference's dormant training metrics (train.py:137-161; `build_metrics` is defined but its
call is commented out, train.py:228-235), composed from the drop-in functions.

`regr_iou`: decode labels and predictions with the GROUND-TRUE classification
(`utils.boxes_decode(labels.classifications, x.regressions_postprocessed)`), elementwise IoU of
the two box lists, mean (train.py:138-143,154-155).  `class_iou`: tf.metrics.mean_iou over the
binarised [M, C] maps with num_classes=2 (train.py:150-153)."""
from __future__ import annotations

import torch

from . import utils
from ._tensor import as_cuda


def regr_iou(label_classifications, label_boxes, predicted_boxes):
    """Mean IoU between the label boxes and the predicted boxes at the anchors the labels mark as
    foreground.  All arguments are masked `[M, .]` tensors (`detection_trainable`)."""
    a = utils.boxes_decode(label_classifications, label_boxes)
    b = utils.boxes_decode(label_classifications, predicted_boxes)
    if a.boxes.shape[0] == 0:
        return torch.zeros((), device=a.boxes.device)
    return utils.iou(a.boxes, b.boxes).mean()


def class_iou(label_classifications, logits):
    """tf.metrics.mean_iou(labels, sigmoid(logits) > 0.5, num_classes=2) for one batch: the mean
    over the two classes {0, 1} of TP / (TP + FP + FN) on the flattened maps."""
    z = as_cuda(label_classifications) > 0.5
    p = utils.sigmoid(logits) > 0.5
    ious = []
    for cls in (False, True):
        zz, pp = (z == cls), (p == cls)
        inter = (zz & pp).sum().double()
        union = (zz | pp).sum().double()
        if float(union) > 0:
            ious.append(inter / union)
    return torch.stack(ious).mean().float() if ious else torch.zeros((), device=z.device)
