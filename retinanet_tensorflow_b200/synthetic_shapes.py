"""Synthetic Shapes data on the device -- the reference's data_loaders/shapes.py (:10-53, :113-176) re-designed
and of utils.draw_bounding_boxes (utils.py:120-148), SURVEY section 8f-4.

The reference draws every image with OpenCV on the host, writes it to disk as PNG and reads it back through
tf.data; the boxes travel as NumPy arrays.  Here the shape SPECIFICATIONS are still drawn from Python's `random`
exactly as the reference does (same calls in the same order, so `random.seed(s)` reproduces the reference's images),
but the pixels are rasterised by `rn_shapes_render` straight into an HBM image batch and the ground truth is laid
out as the padded tensors `DenseHeadPath.assign` takes -- no disk, no decode, no host->device copy of pixels.
`draw_bounding_boxes` outlines device-side detection lists on an image batch (no text labels)."""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _lib

CLASS_NAMES = ['square', 'triangle', 'circle']   # shapes.py:16
MAX_SHAPES = 4                                   # random_image draws 1..4 shapes (shapes.py:165)


def _thin(boxes, threshold=0.3):
    """The reference thins the drawn shapes with its NumPy NMS, scores = arange(n) (shapes.py:170-173): visit the
    boxes from the LAST drawn to the first, keep one unless its IoU with an already kept box exceeds the threshold.
    Returns the kept indices (any order; the caller filters in drawing order)."""
    b = np.asarray(boxes, np.float32)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    kept = []
    for i in reversed(range(len(b))):
        ok = True
        for j in kept:
            ih = max(min(b[i, 2], b[j, 2]) - max(b[i, 0], b[j, 0]), np.float32(0))
            iw = max(min(b[i, 3], b[j, 3]) - max(b[i, 1], b[j, 1]), np.float32(0))
            inter = ih * iw
            if inter / (area[j] + area[i] - inter) > threshold:
                ok = False
                break
        if ok:
            kept.append(i)
    return kept


def random_image(image_size):
    """One image specification drawn from Python's `random` with the reference's calls in the reference's order
    (shapes.py:133-176): background colour (3 x randint(0, 255)); n = randint(1, 4); per shape: choice of
    [square, circle, triangle], colour (3 x randint), y then x in [20, size - 21], s in [20, H // 4]; then the NMS
    thinning.  -> (bg_color int array [3], [(shape name, (r, g, b), (x, y, s)), ...]) like the reference."""
    h, w = int(image_size[0]), int(image_size[1])
    bg_color = np.array([random.randint(0, 255) for _ in range(3)])
    drawn = []
    for _ in range(random.randint(1, 4)):
        kind = random.choice(['square', 'circle', 'triangle'])
        colour = tuple(random.randint(0, 255) for _ in range(3))
        y = random.randint(20, h - 21)
        x = random.randint(20, w - 21)
        s = random.randint(20, h // 4)
        drawn.append((kind, colour, (x, y, s)))
    keep = set(_thin([[y - s, x - s, y + s, x + s] for _, _, (x, y, s) in drawn]))
    return bg_color, [d for i, d in enumerate(drawn) if i in keep]


class Shapes(object):
    """The reference's `Shapes(path, num_samples, image_size)` loader (shapes.py:10-53) without the disk: iterating
    yields {'image': uint8 CUDA tensor [H, W, 3], 'class_ids', 'boxes'} (pixel boxes [y - s, x - s, y + s, x + s], as
    the reference yields them); `batch(n)` renders n images with one launch and returns the padded, normalised
    ground-truth tensors of the dense-head path."""

    def __init__(self, path=None, num_samples=1, image_size=(500, 500), device=None):
        self._num_samples = int(num_samples)
        self._image_size = (int(image_size[0]), int(image_size[1]))
        self._class_names = list(CLASS_NAMES)
        self.device = _lib.require_cuda(device)

    @property
    def class_names(self):
        return self._class_names

    @property
    def num_classes(self):
        return len(self._class_names)

    def _specs(self, n):
        specs = np.zeros((n, MAX_SHAPES, 8), np.int32)
        counts = np.zeros((n,), np.int32)
        bg = np.zeros((n, 3), np.uint8)
        boxes, classes = [], []
        for i in range(n):
            bg_color, shapes = random_image(self._image_size)
            bg[i] = bg_color.astype(np.uint8)
            counts[i] = len(shapes)
            bx, cl = [], []
            for j, (shape, color, (x, y, s)) in enumerate(shapes):
                specs[i, j] = (self._class_names.index(shape), x, y, s, color[0], color[1], color[2], 0)
                bx.append([y - s, x - s, y + s, x + s])
                cl.append(self._class_names.index(shape))
            boxes.append(np.array(bx))
            classes.append(np.array(cl))
        return specs, counts, bg, boxes, classes

    def render(self, specs, counts, bg):
        n = int(counts.shape[0])
        h, w = self._image_size
        image = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
        d_specs = torch.from_numpy(specs).to(self.device)
        d_counts = torch.from_numpy(counts).to(self.device)
        d_bg = torch.from_numpy(bg).to(self.device)
        _lib.call('rn_shapes_render', n, h, w, MAX_SHAPES, _lib.ptr(d_specs), _lib.ptr(d_counts), _lib.ptr(d_bg),
                  _lib.ptr(image), _lib.stream_ptr())
        return image

    def __iter__(self):
        for _ in range(self._num_samples):
            specs, counts, bg, boxes, classes = self._specs(1)
            yield {'image': self.render(specs, counts, bg)[0], 'class_ids': classes[0], 'boxes': boxes[0]}

    def batch(self, n):
        """-> dict(image u8 [n, H, W, 3] (device), gt_boxes f32 [n, 4, 4] normalised [ymin, xmin, ymax, xmax] by the
        image size (dataset.py:161), gt_classes i32 [n, 4] (-1 padding), gt_counts i32 [n]) -- device tensors ready
        for `DenseHeadPath.assign`."""
        specs, counts, bg, boxes, classes = self._specs(n)
        h, w = self._image_size
        gt_boxes = np.zeros((n, MAX_SHAPES, 4), np.float32)
        gt_classes = np.full((n, MAX_SHAPES), -1, np.int32)
        for i in range(n):
            k = int(counts[i])
            gt_boxes[i, :k] = (boxes[i] / np.array([h, w, h, w], np.float64)).astype(np.float32)
            gt_classes[i, :k] = classes[i]
        return {'image': self.render(specs, counts, bg),
                'gt_boxes': torch.from_numpy(gt_boxes).to(self.device),
                'gt_classes': torch.from_numpy(gt_classes).to(self.device),
                'gt_counts': torch.from_numpy(counts).to(self.device)}


def class_palette(num_classes):
    """The class colours of utils.draw_bounding_boxes (utils.py:121-122): RandomState(42), three uniforms per class,
    truncated to uint8 the way cv2 takes a float colour."""
    rng = np.random.RandomState(42)
    colors = [(rng.uniform(0, 255), rng.uniform(0, 255), rng.uniform(0, 255)) for _ in range(num_classes)]
    return np.array(colors, np.float64)


def draw_bounding_boxes(image, boxes, class_ids, counts=None, num_classes=None, palette=None):
    """Outline detections on an image batch, in place, on the device.  image u8 [B, H, W, 3]; boxes f32
    [B, K, 4] normalised; class_ids i32 [B, K]; counts i32 [B] (default K everywhere) -- e.g. `path.det_boxes,
    path.det_classes, path.det_count` straight from `DenseHeadPath`.  Colours follow utils.py:121-122."""
    if image.dim() != 4 or image.shape[-1] != 3 or image.dtype != torch.uint8 or not image.is_cuda:
        raise ValueError('image must be a uint8 CUDA tensor [B, H, W, 3]')
    b, h, w, _ = image.shape
    boxes = boxes.contiguous().float()
    class_ids = class_ids.contiguous().to(torch.int32)
    k = int(boxes.shape[1])
    if counts is None:
        counts = torch.full((b,), k, dtype=torch.int32, device=image.device)
    n_cls = int(num_classes if num_classes is not None else (int(class_ids.max().item()) + 1 if class_ids.numel() else 1))
    pal = class_palette(n_cls) if palette is None else np.asarray(palette, np.float64)
    d_pal = torch.from_numpy(np.clip(np.rint(pal), 0, 255).astype(np.uint8)).to(image.device)
    if not image.is_contiguous():
        raise ValueError('image must be contiguous')
    _lib.call('rn_draw_boxes', b, h, w, k, _lib.ptr(boxes), _lib.ptr(class_ids), _lib.ptr(counts.contiguous().to(torch.int32)),
              _lib.ptr(d_pal), n_cls, _lib.ptr(image), _lib.stream_ptr())
    return image
