"""Geometry, decode, masking and NMS -- host mirror of the reference's utils.py hot-path
functions (utils.py:9-13, 22-117, 151-284) on top of librndense.so.  Same names, argument order,
keyword defaults and return types (the namedtuples of utils.py:10-13); tensors are torch CUDA
tensors (array-likes and DLPack exporters are accepted).

Data-dependent errors: the reference raises tf.errors.InvalidArgumentError from tf.assert_*
inside `iou` (utils.py:65-68,94); here the kernels set a device status word and the wrapper
raises AssertionError after reading it.
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple
from typing import List

import numpy as np
import torch

from . import _lib
from ._pyramid import pyramid_from_grids
from ._tensor import as_cuda, boxmap_op, flat_pyramid, host_ints, new_status
from .pipeline import raise_for_status

NMS_MAX_OUTPUT_SIZE = 1000  # utils.py:9
BoxesDecoded = namedtuple('BoxesDecoded', ['boxes', 'scores', 'class_ids'])
ClassmapDecoded = namedtuple('ClassmapDecoded', ['fg_mask'])
Detection = namedtuple('Detection', ['classification', 'regression', 'regression_postprocessed'])
Classification = namedtuple('Classification', ['unscaled', 'prob'])


# ---- geometry (utils.py:22-117) -----------------------------------------------------------------
def _boxmap5(regression):
    regression = as_cuda(regression)
    if regression.dim() != 5 or regression.shape[-1] != 4:
        raise ValueError('expected a [B, H, W, A, 4] box map, got %s' % (tuple(regression.shape),))
    return regression


def boxmap_anchor_relative_to_image_relative(regression, name='boxmap_anchor_relative_to_image_relative'):
    """Add every cell's centre ((i + 0.5) / n, float32 linspace) to the (y, x) part of a
    [B, H, W, A, 4] map (utils.py:22-36)."""
    regression = _boxmap5(regression)
    b, h, w, a, _ = regression.shape
    if regression.numel() == 0:
        return regression.clone()
    pyr = pyramid_from_grids([(h, w)], np.zeros((1, a, 2), np.float32))
    return boxmap_op(pyr, b, _lib.RN_BOXOP_ADD_CENTRE, regression)


def boxmap_center_relative_to_corner_relative(regression, name='boxmap_center_relative_to_corner_relative'):
    """[cy, cx, h, w] -> [cy - h/2, cx - w/2, cy + h/2, cx + w/2] (utils.py:39-44)."""
    regression = as_cuda(regression)
    n = regression.numel() // 4
    if n == 0:
        return regression.clone()
    return boxmap_op(flat_pyramid(n), 1, _lib.RN_BOXOP_CENTRE_TO_CORNER, regression.reshape(n, 4)).reshape(
        regression.shape)


def anchor_boxmap(grid_size, anchor_boxes, name='anchor_boxmap'):
    """Anchors of a [H, W] grid as corner boxes [1, H, W, A, 4]; `anchor_boxes` is [A, 2]
    normalised (h, w) (utils.py:47-58)."""
    h, w = host_ints(grid_size)
    sizes = as_cuda(anchor_boxes).cpu().numpy().reshape(1, -1, 2)
    pyr = pyramid_from_grids([(h, w)], sizes)
    boxes = torch.empty((h * w * sizes.shape[1], 4), dtype=torch.float32, device=_lib.require_cuda())
    _lib.call('rn_anchor_generate', C.byref(pyr), _lib.ptr(boxes), None, _lib.stream_ptr())
    return boxes.reshape(1, h, w, sizes.shape[1], 4)


def iou(a, b, name='iou'):
    """Broadcasting IoU of corner boxes, 0 where the boxes are disjoint (utils.py:62-97).
    Raises AssertionError when a box has min > max or an IoU is NaN / outside [0, 1]."""
    a, b = as_cuda(a), as_cuda(b)
    a, b = torch.broadcast_tensors(a, b)
    shape = a.shape[:-1]
    a, b = a.contiguous().reshape(-1, 4), b.contiguous().reshape(-1, 4)
    out = torch.empty((a.shape[0],), dtype=torch.float32, device=a.device)
    status = new_status()
    _lib.call('rn_iou_pairwise', a.shape[0], _lib.ptr(a), _lib.ptr(b), _lib.ptr(out), _lib.ptr(status),
              _lib.stream_ptr())
    raise_for_status(int(status.item()))
    return out.reshape(shape)


def _anchor_sizes(anchor_boxes):
    return as_cuda(anchor_boxes).cpu().numpy().reshape(1, -1, 2)


def scale_regression(regression, anchor_boxes, name='scale_regression'):
    """regression [.., A, 4] * [ah, aw, ah, aw] (utils.py:100-105)."""
    regression = as_cuda(regression)
    sizes = _anchor_sizes(anchor_boxes)
    a = sizes.shape[1]
    if regression.shape[-1] != 4 or regression.shape[-2] != a:
        raise ValueError('regression must be [.., %d, 4]' % a)
    cells = regression.numel() // (4 * a)
    if cells == 0:
        return regression.clone()
    pyr = pyramid_from_grids([(cells, 1)], sizes)
    return boxmap_op(pyr, 1, _lib.RN_BOXOP_SCALE, regression.reshape(cells, 1, a, 4)).reshape(regression.shape)


def regression_postprocess(regression, anchor_boxes, name='regression_postprocess'):
    """Decode a [B, H, W, A, 4] delta map to corner boxes: [dy, dx, exp(dh), exp(dw)] *
    [ah, aw, ah, aw], + cell centre, centre -> corners (utils.py:108-117)."""
    regression = _boxmap5(regression)
    b, h, w, a, _ = regression.shape
    sizes = _anchor_sizes(anchor_boxes)
    if sizes.shape[1] != a:
        raise ValueError('anchor_boxes must be [%d, 2]' % a)
    if regression.numel() == 0:
        return regression.clone()
    return boxmap_op(pyramid_from_grids([(h, w)], sizes), b, _lib.RN_BOXOP_DECODE, regression)


# ---- dict helpers (utils.py:151-167, 230-237) ---------------------------------------------------
def merge_outputs(dict, name='merge_outputs'):
    """Concatenate the dict's values along axis 0 in dict order (utils.py:151-153)."""
    return torch.cat([as_cuda(v, None) for v in dict.values()], 0)


def all_same(items):
    return all(item == items[0] for item in items)


def dict_map(f, dict):
    return {key: f(value) for key, value in dict.items()}


def dict_starmap(f, dicts):
    assert all_same([d.keys() for d in dicts])  # utils.py:165
    return {key: f(*[d[key] for d in dicts]) for key in dicts[0].keys()}


def dict_update(dict, keys, f):
    if len(keys) == 0:
        return f(dict)
    head, rest = keys[0], keys[1:]
    return {**dict, head: dict_update(dict[head], rest, f)}


# ---- decode + NMS (utils.py:171-227) --------------------------------------------------------------
def classmap_decode(classmap, name='classmap_decoder'):
    """fg_mask = max over classes > 0.5 (utils.py:171-179)."""
    classmap = as_cuda(classmap)
    c = int(classmap.shape[-1])
    rows = classmap.numel() // max(c, 1)
    out = torch.empty((rows,), dtype=torch.uint8, device=classmap.device)
    _lib.call('rn_classmap_fg', rows, c, _lib.ptr(classmap), 0.5, _lib.ptr(out), _lib.stream_ptr())
    return ClassmapDecoded(fg_mask=out.view(torch.bool).reshape(classmap.shape[:-1]))


def boxes_decode(classifications, regressions, name='boxes_decode', *, score_threshold=0.5, topk=None):
    """One candidate per anchor: score = max_c prob, class = argmax_c prob (first maximum,
    int64), keep score > 0.5, row-major order (utils.py:183-195).  `classifications` are
    probabilities [.., C], `regressions` already-decoded boxes [.., 4].
    `topk` (extension, not in the reference): keep only the k best candidates."""
    classifications = as_cuda(classifications)
    regressions = as_cuda(regressions)
    c = int(classifications.shape[-1])
    rows = classifications.numel() // max(c, 1)
    if regressions.numel() != rows * 4:
        raise ValueError('classifications [.., C] and regressions [.., 4] must share their leading shape')
    dev = classifications.device
    if rows == 0:
        return BoxesDecoded(torch.empty((0, 4), device=dev), torch.empty((0,), device=dev),
                            torch.empty((0,), dtype=torch.int64, device=dev))
    pyr = flat_pyramid(rows)
    cap = rows
    # one allocation for the four output arrays and the workspace, one zero-fill for (count, status)
    wsb = _lib.lib.rn_select_workspace_bytes(1, cap)
    seg = (16 * cap + 255) & ~255, (4 * cap + 255) & ~255
    buf = torch.empty((seg[0] + 3 * seg[1] + wsb,), dtype=torch.uint8, device=dev)
    boxes = buf[:16 * cap].view(torch.float32).view(1, cap, 4)
    scores = buf[seg[0]:seg[0] + 4 * cap].view(torch.float32).view(1, cap)
    classes = buf[seg[0] + seg[1]:seg[0] + seg[1] + 4 * cap].view(torch.int32).view(1, cap)
    anchor = buf[seg[0] + 2 * seg[1]:seg[0] + 2 * seg[1] + 4 * cap].view(torch.int32)
    ws = buf[seg[0] + 3 * seg[1]:]
    count_status = torch.zeros((2,), dtype=torch.int32, device=dev)
    _lib.call('rn_decode_select', C.byref(pyr), 1, c, C.byref(_lib.maps_of([classifications.reshape(rows, c)])), 0,
              C.byref(_lib.maps_of([regressions.reshape(rows, 4)])), 0, float(score_threshold), int(topk or 0), cap,
              _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(classes), _lib.ptr(anchor), _lib.ptr(count_status),
              C.c_void_p(count_status.data_ptr() + 4), _lib.ptr(ws), wsb, _lib.stream_ptr())
    k = int(count_status[0].item())   # data-dependent shape, as in the reference: one host read per call
    return BoxesDecoded(boxes=boxes[0, :k], scores=scores[0, :k], class_ids=classes[0, :k].to(torch.int64))


def merge_boxes_decoded(decoded: List[BoxesDecoded]):
    """Concatenate candidate lists in list order (utils.py:223-227)."""
    return BoxesDecoded(boxes=torch.cat([d.boxes for d in decoded], 0),
                        scores=torch.cat([d.scores for d in decoded], 0),
                        class_ids=torch.cat([d.class_ids for d in decoded], 0))


def _nms_run(decoded, num_classes, classes_i32, max_output_size, iou_threshold):
    boxes = as_cuda(decoded.boxes).reshape(-1, 4)
    scores = as_cuda(decoded.scores).reshape(-1)
    n = int(boxes.shape[0])
    dev = boxes.device
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=dev)
    ocap = n
    wsb = _lib.lib.rn_nms_workspace_bytes(1, num_classes, n)
    seg = (16 * ocap + 255) & ~255, (4 * ocap + 255) & ~255
    buf = torch.empty((seg[0] + 3 * seg[1] + wsb,), dtype=torch.uint8, device=dev)   # outputs + workspace
    det_boxes = buf[:16 * ocap].view(torch.float32)
    det_scores = buf[seg[0]:seg[0] + 4 * ocap].view(torch.float32)
    det_classes = buf[seg[0] + seg[1]:seg[0] + seg[1] + 4 * ocap].view(torch.int32)
    det_index = buf[seg[0] + 2 * seg[1]:seg[0] + 2 * seg[1] + 4 * ocap].view(torch.int32)
    ws = buf[seg[0] + 3 * seg[1]:]
    small = torch.zeros((3,), dtype=torch.int32, device=dev)   # [candidate count, detection count, status]
    small[0] = n
    _lib.call('rn_nms_classwise', 1, num_classes, n, _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(classes_i32),
              _lib.ptr(small), float(iou_threshold), int(max_output_size), ocap, _lib.ptr(det_boxes),
              _lib.ptr(det_scores), _lib.ptr(det_classes), _lib.ptr(det_index), C.c_void_p(small.data_ptr() + 4),
              C.c_void_p(small.data_ptr() + 8), _lib.ptr(ws), wsb, _lib.stream_ptr())
    m = int(small[1].item())
    return det_index[:m].to(torch.int64)


def nms(decoded: BoxesDecoded, max_output_size=NMS_MAX_OUTPUT_SIZE, name='nms', *, iou_threshold=0.5):
    """tf.image.non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5) + gather
    (utils.py:213-220): greedy, descending score, ties -> lower index, suppress iff IoU > 0.5."""
    n = int(as_cuda(decoded.boxes).reshape(-1, 4).shape[0])
    zeros = torch.zeros((max(n, 1),), dtype=torch.int32, device=_lib.require_cuda())
    idx = _nms_run(decoded, 1, zeros, max_output_size, iou_threshold)
    return BoxesDecoded(boxes=as_cuda(decoded.boxes).reshape(-1, 4)[idx], scores=as_cuda(decoded.scores)[idx],
                        class_ids=as_cuda(decoded.class_ids, None)[idx])


def nms_classwise(decoded: BoxesDecoded, num_classes, name='nms_classwise', *,
                  max_output_size=NMS_MAX_OUTPUT_SIZE, iou_threshold=0.5):
    """For c in 0..num_classes-1: candidates with class_ids == c -> nms; results concatenated
    class-major, selection order within a class (utils.py:198-210)."""
    class_ids = as_cuda(decoded.class_ids, None)
    cls32 = class_ids.to(torch.int32)
    # class ids outside [0, num_classes) never match `tf.equal(class_ids, c)`: the kernel sorts them behind the last
    # class segment and never visits them (tests/test_gpu_decode.py), so no host-side filter (and no extra sync)
    idx = _nms_run(decoded, int(num_classes), cls32.contiguous(), max_output_size, iou_threshold)
    return BoxesDecoded(boxes=as_cuda(decoded.boxes).reshape(-1, 4)[idx], scores=as_cuda(decoded.scores)[idx],
                        class_ids=class_ids[idx])


# ---- train-time decode + masking (utils.py:240-284) -------------------------------------------------
def sigmoid(x):
    """tf.nn.sigmoid (utils.py:249) with the device function the fused kernels use."""
    x = as_cuda(x)
    out = torch.empty_like(x)
    _lib.call('rn_sigmoid', x.numel(), _lib.ptr(x), _lib.ptr(out), _lib.stream_ptr())
    return out


def _anchor_sizes_for(level, image_size, anchor_size_mode='tf1_int_trunc'):
    """`tf.to_float(l.anchor_sizes / image_size)` (utils.py:264, dataset.py:53) through the same
    host routine the kernels' pyramid uses."""
    from ._pyramid import pyramid_for
    pyr = pyramid_for(image_size, {'L': level}, anchor_size_mode, ['L'], [1])
    lv = pyr.level[0]
    return np.stack([np.array(lv.anchor_h[:lv.num_anchors], np.float32),
                     np.array(lv.anchor_w[:lv.num_anchors], np.float32)], -1)


class _MaskPlan(object):
    """One scan of the trainable masks (rn_mask_scan), shared by every tensor that is masked with them."""

    def __init__(self, trainable_masks):
        self.keys = list(trainable_masks.keys())
        self.masks = [as_cuda(trainable_masks[k], None) for k in self.keys]
        self.masks = [(m if m.dtype in (torch.bool, torch.uint8) else (m != 0)).contiguous() for m in self.masks]
        self.rows = _lib.rows_of([int(m.numel()) for m in self.masks])
        self.mask_maps = _lib.maps_of([m.view(torch.uint8).reshape(-1) for m in self.masks])
        dev = self.masks[0].device if self.masks else _lib.require_cuda()
        nbytes = max(int(_lib.lib.rn_mask_workspace_bytes(C.byref(self.rows))), 4)
        self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        count = torch.zeros((1,), dtype=torch.int64, device=dev)
        _lib.call('rn_mask_scan', C.byref(self.rows), C.byref(self.mask_maps), _lib.ptr(self.ws), nbytes, _lib.ptr(count),
                  _lib.stream_ptr())
        self.count = int(count.item())   # M: the masked tensors have a data-dependent shape, as in the reference

    def gather(self, per_level):
        """{level: [.., K]} (leading shape = the mask's) -> [M, K], levels concatenated in dict order."""
        if list(per_level.keys()) != self.keys:
            raise AssertionError('dict keys differ')   # utils.dict_starmap's assert (utils.py:165)
        ts = [as_cuda(per_level[k], None) for k in self.keys]
        tail = tuple(ts[0].shape[self.masks[0].dim():])
        k = 1
        for d in tail:
            k *= int(d)
        for t, m in zip(ts, self.masks):
            if tuple(t.shape[:m.dim()]) != tuple(m.shape) or tuple(t.shape[m.dim():]) != tail or t.dtype != ts[0].dtype:
                raise ValueError('tensor %s does not match its mask %s' % (tuple(t.shape), tuple(m.shape)))
        out = torch.empty((self.count,) + tail, dtype=ts[0].dtype, device=ts[0].device)
        if self.count:
            _lib.call('rn_mask_gather', C.byref(self.rows), k * ts[0].element_size(), C.byref(_lib.maps_of(ts)),
                      C.byref(self.mask_maps), _lib.ptr(self.ws), _lib.ptr(out), _lib.stream_ptr())
        return out


class _LazyClassification(Classification):
    """`Classification(unscaled, prob)` whose `prob` is computed on first access.  The reference builds
    sigmoid(logits) for every level and its masked copy eagerly (utils.py:249,270) although its loss only feeds
    them to tf.summary.histogram (losses.py:160-163); here they cost nothing until somebody reads them."""

    def __new__(cls, unscaled, make_prob):
        self = super().__new__(cls, unscaled, None)
        self._make_prob, self._prob = make_prob, None
        return self

    @property
    def prob(self):
        if self._prob is None:
            self._prob = self._make_prob()
        return self._prob

    def __iter__(self):
        return iter((self.unscaled, self.prob))

    def __getitem__(self, i):
        return (self.unscaled, self.prob)[i]


class _LazyDict(dict):
    """{level: f(source[level])}, each value computed on first access."""

    def __init__(self, f, source):
        super().__init__({k: None for k in source})
        self._f, self._src = f, source

    def __getitem__(self, k):
        v = super().__getitem__(k)
        if v is None:
            v = self._f(self._src[k])
            super().__setitem__(k, v)
        return v

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def get(self, k, default=None):
        return self[k] if k in self else default


def postprocess_and_mask(input, trainable_masks, image_size, levels, name='postprocess_and_mask', *, _plan=None):
    """Decode the regressions of every level, then drop the ignored anchors of every level
    (boolean mask over [B, H, W, A], rows ordered b, y, x, a) and concatenate the levels
    (utils.py:258-284).  Compatibility form that materialises the masked tensors exactly like
    the reference (one order-preserving compaction launch per tensor, rn_mask_gather);
    `pipeline.DenseHeadPath.loss` is the fused form that never does."""
    image_size = host_ints(image_size)
    detection = Detection(
        classification=input['detection']['classifications'],
        regression=input['detection']['regressions'],
        regression_postprocessed=dict_starmap(
            lambda r, l: regression_postprocess(r, _anchor_sizes_for(l, image_size)),
            (input['detection']['regressions'], {k: levels[k] for k in input['detection']['regressions']})))
    plan = _MaskPlan(trainable_masks) if _plan is None else _plan
    cls = detection.classification
    if isinstance(cls, _LazyClassification):
        masked_cls = _LazyClassification(plan.gather(cls.unscaled) if cls.unscaled else None, lambda: plan.gather(cls.prob))
    else:
        masked_cls = Classification(unscaled=plan.gather(cls.unscaled) if cls.unscaled else None,
                                    prob=plan.gather(cls.prob) if cls.prob else None)
    trainable = Detection(
        classification=masked_cls,
        regression=plan.gather(detection.regression),
        regression_postprocessed=plan.gather(detection.regression_postprocessed))
    return {**input, 'detection': detection, 'detection_trainable': trainable}


def process_labels_and_logits(labels, logits, levels, name='process_labels_and_logits'):
    """Wrap label maps as probabilities and logits as (unscaled, sigmoid), decode and mask both
    with the labels' trainable masks (utils.py:240-255).  The sigmoid maps are evaluated lazily."""
    labels = dict_update(labels, ['detection', 'classifications'], lambda c: Classification(unscaled=None, prob=c))
    logits = dict_update(logits, ['detection', 'classifications'],
                         lambda c: _LazyClassification(c, lambda: _LazyDict(sigmoid, c)))
    image_size = list(as_cuda(labels['image'], None).shape[1:3])
    plan = _MaskPlan(labels['trainable_masks'])
    labels = postprocess_and_mask(labels, labels['trainable_masks'], image_size=image_size, levels=levels, _plan=plan)
    logits = postprocess_and_mask(logits, labels['trainable_masks'], image_size=image_size, levels=levels, _plan=plan)
    return labels, logits
