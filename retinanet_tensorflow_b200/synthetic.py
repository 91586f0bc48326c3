"""Seeded synthetic inputs for the dense-head path.

Ground truth is generated on the host with `numpy.random.Generator(PCG64(seed))`,
seed = 1000 * config + image_index, so the oracle, the CUDA path and the bench see the
same boxes.  Logits / regression predictions are generated with torch on whatever device
the caller asks for (the bench generates them in HBM; parity tests on the CPU and upload).

Box layout is the reference's: [ymin, xmin, ymax, xmax] normalised to [0, 1] by the image
size (dataset.py:161).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

PRIOR_LOGIT = -math.log(99.0)  # retinanet.py:52-53 (pi = 0.01)


@dataclass(frozen=True)
class Workload:
    """One row of the SURVEY section-8 config table."""
    name: str
    config: int
    image_h: int
    image_w: int
    num_classes: int
    max_gt: int
    batch: int
    crowded: bool = False


WORKLOADS = {
    'shapes': Workload('shapes', 1, 500, 500, 3, 4, 1),
    'voc': Workload('voc', 2, 500, 667, 20, 50, 32),
    'coco': Workload('coco', 3, 800, 1333, 80, 100, 64),
    'crowded': Workload('crowded', 5, 1024, 1024, 80, 300, 256, crowded=True),
}


def gt_for_image(wl: Workload, image_index: int):
    """-> (boxes f32 [n,4], class_ids i32 [n]) for one image; n in 1..max_gt."""
    rng = np.random.Generator(np.random.PCG64(1000 * wl.config + image_index))
    h, w = wl.image_h, wl.image_w
    if wl.crowded:
        n = wl.max_gt
        centres = rng.uniform(0.2, 0.8, (5, 2))
        which = rng.integers(0, 5, n)
        ctr = centres[which] + rng.normal(0.0, 0.05, (n, 2))
        base = rng.uniform(80.0, 160.0, (5,))[which] * np.exp2(rng.normal(0.0, 0.15, n))
        aspect = np.exp2(rng.normal(0.0, 0.2, n))
    else:
        n = int(rng.integers(1, wl.max_gt + 1))
        ctr = rng.uniform(0.05, 0.95, (n, 2))
        base = np.exp2(rng.uniform(4.0, math.log2(0.6 * min(h, w)), n))
        aspect = np.exp2(rng.uniform(-1.5, 1.5, n))
    bh = base * np.sqrt(aspect) / h
    bw = base / np.sqrt(aspect) / w
    tl = np.stack([ctr[:, 0] - bh / 2, ctr[:, 1] - bw / 2], -1)
    br = np.stack([ctr[:, 0] + bh / 2, ctr[:, 1] + bw / 2], -1)
    tl = np.clip(tl, 0.0, 1.0)
    br = np.clip(br, 0.0, 1.0)
    # the reference requires tl < br strictly (data_loaders/inferred.py:32): keep >= 1 px
    min_sz = np.array([1.0 / h, 1.0 / w])
    br = np.maximum(br, tl + min_sz)
    over = np.maximum(br - 1.0, 0.0)
    tl, br = tl - over, br - over
    boxes = np.concatenate([tl, br], -1).astype(np.float32)
    # float32 rounding must not collapse a box
    assert np.all(boxes[:, :2] < boxes[:, 2:])
    classes = rng.integers(0, wl.num_classes, n).astype(np.int32)
    return boxes, classes


def gt_batch(wl: Workload, batch: int | None = None, first_image: int = 0):
    """Padded batch: boxes f32 [B, max_gt, 4] (padding rows are zeros), class_ids i32
    [B, max_gt] (padding -1), counts i32 [B]."""
    b = wl.batch if batch is None else batch
    boxes = np.zeros((b, wl.max_gt, 4), np.float32)
    classes = np.full((b, wl.max_gt), -1, np.int32)
    counts = np.zeros((b,), np.int32)
    for i in range(b):
        bx, cl = gt_for_image(wl, first_image + i)
        counts[i] = bx.shape[0]
        boxes[i, :bx.shape[0]] = bx
        classes[i, :bx.shape[0]] = cl
    return boxes, classes, counts


def head_outputs(labels, deltas, num_classes, *, seed, device, logit_sigma=1.0):
    """Synthetic head outputs in the level-major merged layout.

    labels: int32 [R] compact labels from the assign kernel (class >= 0 positive, -1
    background, -2 ignored); deltas: f32 [R, 4] encoded targets.  Returns
    (logits f32 [R, C], regression f32 [R, 4]):
      logits ~ N(prior, logit_sigma^2), + N(6, 1) on (positive anchor, matched class);
      regression = target + N(0, 0.1^2) on positives, N(0, 0.3^2) elsewhere.
    """
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    labels = torch.as_tensor(labels, device=device)
    deltas = torch.as_tensor(deltas, device=device)
    r = labels.shape[0]
    logits = torch.empty((r, num_classes), dtype=torch.float32, device=device)
    logits.normal_(PRIOR_LOGIT, logit_sigma, generator=g)
    pos = torch.nonzero(labels >= 0).squeeze(1)
    boost = torch.empty((pos.shape[0],), dtype=torch.float32, device=device).normal_(6.0, 1.0, generator=g)
    logits[pos, labels[pos].long()] += boost
    regression = torch.empty((r, 4), dtype=torch.float32, device=device).normal_(0.0, 0.3, generator=g)
    noise = torch.empty((pos.shape[0], 4), dtype=torch.float32, device=device).normal_(0.0, 0.1, generator=g)
    regression[pos] = deltas[pos] + noise
    return logits, regression
