"""Label construction -- host mirror of the reference's dataset.py label functions
(dataset.py:10-142) on top of librndense.so.  Same names, argument order and return
structure; tensors are torch CUDA tensors (anything array-like / DLPack-capable is accepted).

The reference builds labels one image at a time on tf.data CPU threads (dataset.py:167,231);
`build_labels_batch` is the batched form the GPU wants (one launch for B images, compact
outputs), `build_labels` / `level_labels` keep the reference's dense per-image return values.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._pyramid import level_shapes, pyramid_for, pyramid_from_grids, split_merged
from ._tensor import as_cuda, boxmap_op, flat_pyramid, host_ints, new_status
from .pipeline import raise_for_status

NEG_IOU_THRESHOLD = 0.4  # dataset.py:10
POS_IOU_THRESHOLD = 0.5  # dataset.py:11
MEAN = [0.46618041, 0.44669811, 0.40252436]  # dataset.py:12
STD = [0.27940595, 0.27489075, 0.28920765]  # dataset.py:13


def position_grid(size):
    """Cell centres (cy, cx) of a [H, W] grid in normalised image coordinates, float32
    [H, W, 2]: (i + 0.5) / n evaluated as the reference's float32 linspace (dataset.py:16-25)."""
    h, w = host_ints(size)
    pyr = pyramid_from_grids([(h, w)], np.zeros((1, 1, 2), np.float32))
    centres = torch.empty((h * w, 4), dtype=torch.float32, device=_lib.require_cuda())
    _lib.call('rn_anchor_generate', C.byref(pyr), None, _lib.ptr(centres), _lib.stream_ptr())
    return centres[:, :2].reshape(h, w, 2)


def _flat_box_op(box, op):
    box = as_cuda(box)
    if box.shape[-1] != 4:
        raise ValueError('boxes must have 4 values in the last dimension')
    n = box.numel() // 4
    if n == 0:
        return box.clone()
    return boxmap_op(flat_pyramid(n), 1, op, box.reshape(n, 4)).reshape(box.shape)


def to_center_box(box):
    """[ymin, xmin, ymax, xmax] -> [cy, cx, h, w] as `a + (b - a) / 2`, `b - a` (dataset.py:28-32)."""
    return _flat_box_op(box, _lib.RN_BOXOP_CORNER_TO_CENTRE)


def from_center_box(box):
    """[cy, cx, h, w] -> corners `pos -/+ size / 2` (dataset.py:35-39)."""
    return _flat_box_op(box, _lib.RN_BOXOP_CENTRE_TO_CORNER)


def build_labels_batch(image_size, class_ids, boxes, counts, levels, num_classes, *, names=None, strides=None,
                       anchor_size_mode='tf1_int_trunc', neg_iou=NEG_IOU_THRESHOLD, pos_iou=POS_IOU_THRESHOLD,
                       return_iou=False, check=True):
    """Batched assignment + encoding, compact outputs in the merged level-major layout.

    class_ids int32 [B, N], boxes float32 [B, N, 4] (padded), counts int32 [B].
    Returns dict(pyramid, matched int32 [R], label int32 [R] (class | -1 background | -2
    ignored), deltas float32 [R, 4], iou float32 [R] or None)."""
    boxes = as_cuda(boxes)
    class_ids = as_cuda(class_ids, torch.int32)
    counts = as_cuda(counts, torch.int32)
    if boxes.dim() != 3 or boxes.shape[-1] != 4:
        raise ValueError('boxes must be [B, N, 4]')
    b, n = int(boxes.shape[0]), int(boxes.shape[1])
    if tuple(class_ids.shape) != (b, n) or tuple(counts.shape) != (b,):
        raise ValueError('class_ids must be [B, N] and counts [B]')
    pyr = pyramid_for(host_ints(image_size), levels, anchor_size_mode, names, strides)
    r = b * int(pyr.anchors_per_image)
    dev = boxes.device
    matched = torch.empty((r,), dtype=torch.int32, device=dev)
    label = torch.empty((r,), dtype=torch.int32, device=dev)
    deltas = torch.empty((r, 4), dtype=torch.float32, device=dev)
    iou = torch.empty((r,), dtype=torch.float32, device=dev) if return_iou else None
    status = new_status()

    def merged(t, nbytes):
        m = _lib.Maps()
        if t is not None:
            _lib.call('rn_maps_from_merged', C.byref(pyr), b, _lib.ptr(t), nbytes, C.byref(m))
        return m

    _lib.call('rn_assign_encode', C.byref(pyr), b, _lib.ptr(boxes), _lib.ptr(class_ids), _lib.ptr(counts), n,
              int(num_classes), float(neg_iou), float(pos_iou), C.byref(merged(matched, 4)),
              C.byref(merged(label, 4)), C.byref(merged(deltas, 16)), C.byref(merged(iou, 4)), None,
              _lib.ptr(status), _lib.stream_ptr())
    if check:
        raise_for_status(int(status.item()))  # the reference's tf.assert_* inside utils.iou
    return {'pyramid': pyr, 'batch': b, 'matched': matched, 'label': label, 'deltas': deltas, 'iou': iou}


def _dense_from_compact(res, num_classes):
    """Compact labels -> the reference's dense per-level tensors (dataset.py:87-100)."""
    pyr, b = res['pyramid'], res['batch']
    r = res['label'].shape[0]
    dev = res['label'].device
    onehot = torch.empty((r, num_classes), dtype=torch.float32, device=dev)
    mask = torch.empty((r,), dtype=torch.uint8, device=dev)
    rows = _lib.Rows()
    _lib.call('rn_rows_from_pyramid', C.byref(pyr), b, C.byref(rows))

    def merged(t, nbytes):
        m = _lib.Maps()
        _lib.call('rn_maps_from_merged', C.byref(pyr), b, _lib.ptr(t), nbytes, C.byref(m))
        return m

    _lib.call('rn_labels_dense', C.byref(rows), int(num_classes), C.byref(merged(res['label'], 4)),
              C.byref(merged(onehot, 4 * num_classes)), C.byref(merged(mask, 1)), _lib.stream_ptr())
    return (split_merged(pyr, b, onehot, (num_classes,)), split_merged(pyr, b, res['deltas'], (4,)),
            split_merged(pyr, b, mask.view(torch.bool)))


def level_labels(image_size, class_id, true_box, level, factor, num_classes, *, anchor_size_mode='tf1_int_trunc'):
    """Labels of one image for one pyramid level (dataset.py:43-123).

    Returns (classification float32 [H, W, A, C] one-hot, zero on background / ignored anchors;
    regression float32 [H, W, A, 4] = (dy, dx, log dh, log dw) against the arg-max IoU box for
    EVERY anchor; trainable_mask bool [H, W, A] = IoU < 0.4 or IoU >= 0.5)."""
    true_box = as_cuda(true_box)
    class_id = as_cuda(class_id, torch.int32)
    n = int(true_box.shape[0])
    if n < 1:
        raise AssertionError('level_labels needs at least one box (data_loaders/inferred.py:30)')
    res = build_labels_batch(image_size, class_id.reshape(1, n), true_box.reshape(1, n, 4),
                             torch.tensor([n], dtype=torch.int32), {'L': level}, num_classes, names=['L'],
                             strides=[int(factor)], anchor_size_mode=anchor_size_mode)
    cls, reg, mask = _dense_from_compact(res, int(num_classes))
    return cls[0][0], reg[0][0], mask[0][0]


def build_labels(image_size, class_ids, boxes, levels, num_classes, *, anchor_size_mode='tf1_int_trunc'):
    """Labels of one image for every level (dataset.py:126-142): three dicts keyed like `levels`
    (classifications, regressions, trainable_masks), `factor = 2 ** int(name[-1])`."""
    boxes = as_cuda(boxes)
    class_ids = as_cuda(class_ids, torch.int32)
    n = int(boxes.shape[0])
    if n < 1:
        raise AssertionError('build_labels needs at least one box (data_loaders/inferred.py:30)')
    names = list(levels)
    res = build_labels_batch(image_size, class_ids.reshape(1, n), boxes.reshape(1, n, 4),
                             torch.tensor([n], dtype=torch.int32), levels, num_classes, names=names,
                             anchor_size_mode=anchor_size_mode)
    cls, reg, mask = _dense_from_compact(res, int(num_classes))
    return ({pn: cls[i][0] for i, pn in enumerate(names)},
            {pn: reg[i][0] for i, pn in enumerate(names)},
            {pn: mask[i][0] for i, pn in enumerate(names)})
