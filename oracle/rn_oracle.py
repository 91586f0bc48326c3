"""CPU oracle for the RetinaNet dense-head data path.  TEST INFRASTRUCTURE ONLY.

This file is a literal float32 NumPy restatement of the reference's TensorFlow-1.x
graph code for the path named in BASELINE.json (`north_star`).  It keeps the
reference's dense `[N, H, W, A]` formulation and its order of operations so that
every intermediate is rounded to float32 exactly where TF rounds it (TF runs each
op as its own kernel: no FMA contraction).

Who may import this: `tests/`, `__graft_entry__.smoke()`, and `bench.py`'s
`cpu_baseline` / `--impl reference` legs -- as the checker or the timed CPU
baseline, never as the shipped product path.  The product path
(`retinanet_tensorflow_b200`) never imports it and fails loudly without its CUDA
library.

Parity status:
  * PINNED by the reference's own tests / importable modules: anchor sizes
    (levels_test.py, and levels.py imported directly), grid-centre add
    (utils_test.py:7-42), anchor_boxmap (utils_test.py:44-74), centre->corner
    (utils_test.py:76-97, bit exact), iou (utils_test.py:99-118), fg mask pattern
    (utils_test.py:120-138), merge order (losses_test.py:7-15), Huber == 2.0
    (losses_test.py:17-27), assignment pattern (dataset_test.py:9-28),
    scale_regression (retinanet_old_test.py:15-37), greedy NMS vs the reference's
    own NumPy NMS (data_loaders/shapes.py:76-110).  See tests/test_oracle_golden.py.
  * PARITY UNPINNED (TensorFlow is not installed and not installable here, and
    the reference tests hold no vectors): focal / BCE / balanced BCE / dice /
    jaccard / fixed-IoU values and gradients, `tf.image.non_max_suppression`
    tie order, `tf.argmax` tie order, `tf.linspace` rounding, and the implicit
    ndarray -> int32 truncation of anchor sizes (dataset.py:53).  For those the
    rule written in SURVEY.md Appendix A *is* the contract; each is a switch here.

Every function cites the reference file:line it follows (paths are relative to
the reference checkout).
"""
from __future__ import annotations

from collections import namedtuple
from itertools import product

import numpy as np

F32 = np.float32

NEG_IOU_THRESHOLD = 0.4  # dataset.py:10
POS_IOU_THRESHOLD = 0.5  # dataset.py:11
NMS_MAX_OUTPUT_SIZE = 1000  # utils.py:9

# utils.py:10-13
BoxesDecoded = namedtuple('BoxesDecoded', ['boxes', 'scores', 'class_ids'])
ClassmapDecoded = namedtuple('ClassmapDecoded', ['fg_mask'])
Detection = namedtuple('Detection', ['classification', 'regression', 'regression_postprocessed'])
Classification = namedtuple('Classification', ['unscaled', 'prob'])


# --------------------------------------------------------------------------------------
# levels.py
# --------------------------------------------------------------------------------------
def compute_box_size(base_size, aspect_ratio, scale_ratio):
    """levels.py:49-52 (float64 NumPy in the reference too)."""
    aspect_ratio = np.array(aspect_ratio)
    return np.sqrt(base_size ** 2 / aspect_ratio.prod()) * aspect_ratio * scale_ratio


class Level(object):
    """levels.py:32-44."""

    def __init__(self, anchor_size, anchor_aspect_ratios, anchor_scale_ratios):
        self._anchor_size = anchor_size
        self._anchor_aspect_ratios = anchor_aspect_ratios
        self._anchor_scale_ratios = anchor_scale_ratios

    @property
    def anchor_sizes(self):
        return np.stack([
            compute_box_size(self._anchor_size, a, s)
            for a, s in product(self._anchor_aspect_ratios, self._anchor_scale_ratios)
        ], 0)


class Levels(object):
    """levels.py:5-29 (P3..P7, base sizes 32..512)."""

    def __init__(self, anchor_aspect_ratios, anchor_scale_ratios):
        self._anchor_aspect_ratios = anchor_aspect_ratios
        self._anchor_scale_ratios = anchor_scale_ratios
        self._levels = {
            'P%d' % (i + 3): Level(32 * 2 ** i, anchor_aspect_ratios, anchor_scale_ratios)
            for i in range(5)
        }

    @property
    def num_anchors(self):
        return len(self._anchor_aspect_ratios) * len(self._anchor_scale_ratios)

    def keys(self):
        return self._levels.keys()

    def __getitem__(self, item):
        return self._levels[item]

    def __iter__(self):
        return iter(self.keys())


def build_levels():
    """levels.py:55-59."""
    return Levels([(1, 2), (1, 1), (2, 1)], [2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3)])


# --------------------------------------------------------------------------------------
# TF1 runtime semantics restated (SURVEY Appendix A, items marked "recalled")
# --------------------------------------------------------------------------------------
def tf_linspace(start, stop, num):
    """TF1 CPU `LinSpace` kernel: step=(stop-start)/(num-1); v_i = start + step*i, all in
    float32 with a separate multiply and add (dataset.py:19-20, utils.py:27-28)."""
    start, stop, num = F32(start), F32(stop), int(num)
    if num == 1:
        return np.array([start], F32)
    step = F32(F32(stop - start) / F32(num - 1))
    i = np.arange(num, dtype=F32)
    return (start + (step * i).astype(F32)).astype(F32)


def anchor_size_normalised(anchor_sizes_px, image_size, mode='tf1_int_trunc'):
    """`tf.to_float(level.anchor_sizes / image_size)` (dataset.py:53, utils.py:264).

    mode 'tf1_int_trunc' (default): `image_size` is an int32 tensor, so TF1 converts the
    float64 ndarray to int32 (`astype`: truncation toward zero), true-divides in float64,
    then casts to float32.  mode 'float': no truncation (float64 divide, float32 cast).
    """
    px = np.asarray(anchor_sizes_px, np.float64)
    size = np.asarray(image_size, np.int64).astype(np.float64)
    if mode == 'tf1_int_trunc':
        px = np.trunc(px)
    elif mode != 'float':
        raise ValueError('anchor_size_mode must be tf1_int_trunc or float')
    return (px / size).astype(F32)


def grid_size_for(image_size, factor):
    """`tf.to_int32(tf.ceil(image_size / factor))` (dataset.py:55)."""
    return np.ceil(np.asarray(image_size, np.int64).astype(np.float64) / factor).astype(np.int32)


# --------------------------------------------------------------------------------------
# dataset.py:16-39
# --------------------------------------------------------------------------------------
def _grid_axes(size):
    """Shared by dataset.position_grid (dataset.py:16-20) and
    utils.boxmap_anchor_relative_to_image_relative (utils.py:24-28)."""
    size = np.asarray(size, np.int64)
    cell = (1.0 / size.astype(np.float64)).astype(F32)  # tf.to_float(1 / size)
    two = F32(2)
    y = tf_linspace(cell[0] / two, F32(1) - cell[0] / two, size[0])
    x = tf_linspace(cell[1] / two, F32(1) - cell[1] / two, size[1])
    return y, x


def position_grid(size):
    """dataset.py:16-25 -> [H, W, 2] float32 (cy, cx)."""
    y, x = _grid_axes(size)
    xg, yg = np.meshgrid(x, y)  # 'xy' indexing like tf.meshgrid
    return np.stack([yg, xg], -1).astype(F32)


def to_center_box(box):
    """dataset.py:28-32."""
    box = np.asarray(box, F32)
    a, b = box[..., :2], box[..., 2:]
    size = (b - a).astype(F32)
    return np.concatenate([(a + size / F32(2)).astype(F32), size], -1)


def from_center_box(box):
    """dataset.py:35-39."""
    box = np.asarray(box, F32)
    pos, size = box[..., :2], box[..., 2:]
    half = (size / F32(2)).astype(F32)
    return np.concatenate([(pos - half).astype(F32), (pos + half).astype(F32)], -1)


# --------------------------------------------------------------------------------------
# utils.py:22-117
# --------------------------------------------------------------------------------------
def boxmap_anchor_relative_to_image_relative(regression):
    """utils.py:22-36. regression [B, H, W, A, 4]."""
    regression = np.asarray(regression, F32)
    y, x = _grid_axes(regression.shape[1:3])
    xg, yg = np.meshgrid(x, y)
    grid_pos = np.stack([yg, xg], -1)[:, :, None, :]  # [H, W, 1, 2]
    pos, size = regression[..., :2], regression[..., 2:]
    return np.concatenate([(pos + grid_pos).astype(F32), size], -1)


def boxmap_center_relative_to_corner_relative(regression):
    """utils.py:39-44."""
    regression = np.asarray(regression, F32)
    pos = regression[..., :2]
    half = (regression[..., 2:] / F32(2)).astype(F32)
    return np.concatenate([(pos - half).astype(F32), (pos + half).astype(F32)], -1)


def anchor_boxmap(grid_size, anchor_boxes):
    """utils.py:47-58 -> [1, H, W, A, 4] corner boxes."""
    anchor_boxes = np.asarray(anchor_boxes, F32)
    n = anchor_boxes.shape[0]
    ab = np.concatenate([np.zeros_like(anchor_boxes), anchor_boxes], -1).reshape(1, 1, 1, n, 4)
    ab = np.tile(ab, (1, int(grid_size[0]), int(grid_size[1]), 1, 1))
    return boxmap_center_relative_to_corner_relative(boxmap_anchor_relative_to_image_relative(ab))


class OracleAssertionError(AssertionError):
    """Stands in for tf.errors.InvalidArgumentError raised by tf.assert_* in utils.iou."""


def iou(a, b, check=True):
    """utils.py:62-97.  Broadcasting IoU in float32, `where(invalid, 0, iou)`, with the
    reference's four runtime asserts."""
    a = np.asarray(a, F32)
    b = np.asarray(b, F32)
    if check:
        if not (np.all(a[..., :2] <= a[..., 2:]) and np.all(b[..., :2] <= b[..., 2:])):
            raise OracleAssertionError('assert_less_equal(box[:2], box[2:]) failed')  # utils.py:66-67
    y_top = np.maximum(a[..., 0], b[..., 0])
    x_left = np.maximum(a[..., 1], b[..., 1])
    y_bottom = np.minimum(a[..., 2], b[..., 2])
    x_right = np.minimum(a[..., 3], b[..., 3])
    invalid = np.logical_or(y_bottom < y_top, x_right < x_left)
    inter = ((y_bottom - y_top) * (x_right - x_left)).astype(F32)
    area_a = ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])).astype(F32)
    area_b = ((b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])).astype(F32)
    with np.errstate(divide='ignore', invalid='ignore'):
        out = (inter / ((area_a + area_b).astype(F32) - inter).astype(F32)).astype(F32)
    out = np.where(invalid, F32(0), out).astype(F32)
    if check:
        if not (np.all(out >= 0.0) and np.all(out <= 1.0)):  # NaN fails both, like TF
            raise OracleAssertionError('assert 0 <= iou <= 1 failed')  # utils.py:94
    return out


def scale_regression(regression, anchor_boxes):
    """utils.py:100-105."""
    regression = np.asarray(regression, F32)
    anchor_boxes = np.tile(np.asarray(anchor_boxes, F32), (1, 2))
    anchor_boxes = anchor_boxes.reshape(1, 1, 1, anchor_boxes.shape[0], anchor_boxes.shape[1])
    return (regression * anchor_boxes).astype(F32)


def regression_postprocess(regression, anchor_boxes):
    """utils.py:108-117: [dy,dx,exp(dh),exp(dw)] * [ah,aw,ah,aw] + cell centre -> corners."""
    regression = np.asarray(regression, F32)
    shifts, scales = regression[..., :2], regression[..., 2:]
    regression = np.concatenate([shifts, np.exp(scales).astype(F32)], -1)
    regression = scale_regression(regression, anchor_boxes)
    regression = boxmap_anchor_relative_to_image_relative(regression)
    return boxmap_center_relative_to_corner_relative(regression)


# --------------------------------------------------------------------------------------
# dataset.py:43-142  (assignment + encoding, dense formulation kept on purpose)
# --------------------------------------------------------------------------------------
def level_labels(image_size, class_id, true_box, level, factor, num_classes,
                 anchor_size_mode='tf1_int_trunc', return_aux=False):
    """dataset.py:43-123.  Returns (classification [H,W,A,C] f32, regression [H,W,A,4] f32,
    trainable_mask [H,W,A] bool); with return_aux also (iou_index int64, iou_value f32)."""
    class_id = np.asarray(class_id, np.int64)
    true_box = np.asarray(true_box, F32)
    num_objects = true_box.shape[0]
    sizes_px = level.anchor_sizes
    num_anchors = sizes_px.shape[0]

    true_box = to_center_box(true_box).reshape(num_objects, 1, 1, 1, 4)  # dataset.py:48-50
    anchor_size = anchor_size_normalised(sizes_px, image_size, anchor_size_mode)  # dataset.py:53
    grid_size = grid_size_for(image_size, factor)  # dataset.py:55
    anchor_position = position_grid(grid_size)  # dataset.py:57
    h, w = int(grid_size[0]), int(grid_size[1])
    anchor_position = np.tile(anchor_position.reshape(1, h, w, 1, 2), (1, 1, 1, num_anchors, 1))
    anchor_size = np.tile(anchor_size.reshape(1, 1, 1, num_anchors, 2), (1, h, w, 1, 1))
    anchor = np.concatenate([anchor_position, anchor_size], -1)  # dataset.py:69

    iou_all = iou(from_center_box(anchor), from_center_box(true_box))  # [N,H,W,A] dataset.py:74
    iou_index = np.argmax(iou_all, 0)  # first maximum wins; all-zero column -> 0 (dataset.py:76)
    iou_value = np.max(iou_all, 0)  # dataset.py:78

    bg_mask = iou_value < F32(POS_IOU_THRESHOLD)  # dataset.py:83
    trainable_mask = np.logical_or(iou_value < F32(NEG_IOU_THRESHOLD),
                                   iou_value >= F32(POS_IOU_THRESHOLD))  # dataset.py:87

    classification = class_id[iou_index]  # tf.gather, dataset.py:91
    onehot = np.zeros((h, w, num_anchors, num_classes), F32)  # tf.one_hot, dataset.py:93
    valid = (classification >= 0) & (classification < num_classes)  # out-of-range -> all-zero row
    yy, xx, aa = np.nonzero(valid)
    onehot[yy, xx, aa, classification[yy, xx, aa]] = 1
    onehot[bg_mask] = 0  # tf.where(bg_mask_expanded, zeros, classification) dataset.py:96-100

    true_position, true_size = true_box[..., :2], true_box[..., 2:]  # dataset.py:105
    with np.errstate(divide='ignore', invalid='ignore'):
        shifts = ((true_position - anchor_position).astype(F32) / anchor_size).astype(F32)  # :108
        scales = (true_size / anchor_size).astype(F32)  # :110
        regression = np.concatenate([shifts, np.log(scales).astype(F32)], -1)  # [N,H,W,A,4] :112
        sel = np.zeros((num_objects, h, w, num_anchors, 1), F32)  # one_hot(iou_index, axis=0) :118
        yy, xx, aa = np.indices((h, w, num_anchors))
        sel[iou_index, yy, xx, aa, 0] = 1
        regression = np.sum((regression * sel).astype(F32), 0, dtype=F32)  # :121
    if return_aux:
        return onehot, regression, trainable_mask, iou_index, iou_value
    return onehot, regression, trainable_mask


def build_labels(image_size, class_ids, boxes, levels, num_classes, anchor_size_mode='tf1_int_trunc'):
    """dataset.py:126-142."""
    labels = {
        pn: level_labels(image_size, class_ids, boxes, level=levels[pn], factor=2 ** int(pn[-1]),
                         num_classes=num_classes, anchor_size_mode=anchor_size_mode)
        for pn in levels
    }
    return ({pn: labels[pn][0] for pn in labels},
            {pn: labels[pn][1] for pn in labels},
            {pn: labels[pn][2] for pn in labels})


def flip(sample):
    """augmentation.py:5-22: reverse the W axis of the per-level maps and negate dx."""
    def rev(x):
        return np.flip(x, 1)
    regs = {}
    for pn, r in sample['detection']['regressions'].items():
        r = rev(r).copy()
        r[..., 1] = -r[..., 1]
        regs[pn] = r
    return {
        'image': rev(sample['image']) if 'image' in sample else None,
        'detection': {
            'classifications': {k: rev(v) for k, v in sample['detection']['classifications'].items()},
            'regressions': regs,
        },
        'trainable_masks': {k: rev(v) for k, v in sample['trainable_masks'].items()},
    }


# --------------------------------------------------------------------------------------
# utils.py:151-284  (masking, merge, decode, NMS)
# --------------------------------------------------------------------------------------
def sigmoid(x):
    """tf.nn.sigmoid (utils.py:249, losses.py:10): 1 / (1 + exp(-x)) in float32."""
    x = np.asarray(x, F32)
    with np.errstate(over='ignore'):
        return (F32(1) / (F32(1) + np.exp(-x).astype(F32)).astype(F32)).astype(F32)


def merge_outputs(d):
    """utils.py:151-153: concat dict values in dict order along axis 0."""
    return np.concatenate(list(d.values()), 0)


def classmap_decode(classmap):
    """utils.py:171-179."""
    return ClassmapDecoded(fg_mask=np.max(np.asarray(classmap, F32), -1) > F32(0.5))


def boxes_decode(classifications, regressions):
    """utils.py:183-195: one candidate per anchor, strict `> 0.5`, row-major compaction."""
    classifications = np.asarray(classifications, F32)
    regressions = np.asarray(regressions, F32)
    cmax = np.max(classifications, -1)
    class_ids = np.argmax(classifications, -1).astype(np.int64)  # first maximum wins
    fg = cmax > F32(0.5)
    return BoxesDecoded(boxes=regressions[fg], scores=cmax[fg], class_ids=class_ids[fg])


def merge_boxes_decoded(decoded):
    """utils.py:223-227."""
    return BoxesDecoded(boxes=np.concatenate([d.boxes for d in decoded], 0),
                        scores=np.concatenate([d.scores for d in decoded], 0),
                        class_ids=np.concatenate([d.class_ids for d in decoded], 0))


def tf_nms_iou(box_i, boxes_j):
    """IoU as computed by TF's CPU NonMaxSuppression kernel: corners
    normalised by min/max, `area <= 0 -> 0`, inter = max(dy,0)*max(dx,0)."""
    box_i = np.asarray(box_i, F32)
    boxes_j = np.asarray(boxes_j, F32).reshape(-1, 4)
    ymin_i, xmin_i = min(box_i[0], box_i[2]), min(box_i[1], box_i[3])
    ymax_i, xmax_i = max(box_i[0], box_i[2]), max(box_i[1], box_i[3])
    ymin_j = np.minimum(boxes_j[:, 0], boxes_j[:, 2])
    xmin_j = np.minimum(boxes_j[:, 1], boxes_j[:, 3])
    ymax_j = np.maximum(boxes_j[:, 0], boxes_j[:, 2])
    xmax_j = np.maximum(boxes_j[:, 1], boxes_j[:, 3])
    area_i = F32(F32(ymax_i - ymin_i) * F32(xmax_i - xmin_i))
    area_j = ((ymax_j - ymin_j).astype(F32) * (xmax_j - xmin_j).astype(F32)).astype(F32)
    iy0 = np.maximum(ymin_i, ymin_j)
    ix0 = np.maximum(xmin_i, xmin_j)
    iy1 = np.minimum(ymax_i, ymax_j)
    ix1 = np.minimum(xmax_i, xmax_j)
    inter = (np.maximum((iy1 - iy0).astype(F32), F32(0)) *
             np.maximum((ix1 - ix0).astype(F32), F32(0))).astype(F32)
    with np.errstate(divide='ignore', invalid='ignore'):
        out = (inter / ((area_i + area_j).astype(F32) - inter).astype(F32)).astype(F32)
    return np.where((area_i <= 0) | (area_j <= 0), F32(0), out).astype(F32)


def non_max_suppression(boxes, scores, max_output_size=NMS_MAX_OUTPUT_SIZE, iou_threshold=0.5):
    """`tf.image.non_max_suppression` (utils.py:215) restated: visit candidates in descending
    score, ties -> lower index first (stable; the contract chosen in SURVEY A.4); keep a
    candidate iff its IoU with every kept box is NOT > threshold; stop at max_output_size.
    Returns int32 indices in selection order."""
    boxes = np.asarray(boxes, F32).reshape(-1, 4)
    scores = np.asarray(scores, F32).reshape(-1)
    order = np.argsort(-scores, kind='stable')
    thr = F32(iou_threshold)
    kept = []
    for i in order:
        if len(kept) >= max_output_size:
            break
        if kept:
            if np.any(tf_nms_iou(boxes[i], boxes[np.array(kept)]) > thr):
                continue
        kept.append(int(i))
    return np.array(kept, np.int32)


def nms(decoded, max_output_size=NMS_MAX_OUTPUT_SIZE, iou_threshold=0.5):
    """utils.py:213-220."""
    idx = non_max_suppression(decoded.boxes, decoded.scores, max_output_size, iou_threshold)
    return BoxesDecoded(boxes=decoded.boxes[idx], scores=decoded.scores[idx], class_ids=decoded.class_ids[idx])


def nms_classwise(decoded, num_classes, max_output_size=NMS_MAX_OUTPUT_SIZE, iou_threshold=0.5,
                  return_indices=False):
    """utils.py:198-210: per class c (mask `class_ids == c`, order preserved) -> nms -> concat
    class-major.  With return_indices also returns the indices into `decoded`."""
    out, out_idx = [], []
    for c in range(num_classes):
        m = np.nonzero(decoded.class_ids == c)[0]
        sub = BoxesDecoded(decoded.boxes[m], decoded.scores[m], decoded.class_ids[m])
        idx = non_max_suppression(sub.boxes, sub.scores, max_output_size, iou_threshold)
        out.append(BoxesDecoded(sub.boxes[idx], sub.scores[idx], sub.class_ids[idx]))
        out_idx.append(m[idx])
    merged = merge_boxes_decoded(out)
    if return_indices:
        return merged, np.concatenate(out_idx).astype(np.int64)
    return merged


def topk_prefilter(decoded_level, k):
    """NOT in the reference (BASELINE config 4 extension): keep the k
    highest scores of one (image, level) candidate list (ties -> lower index), then restore
    ascending index order.  Identity when k >= #candidates."""
    n = decoded_level.scores.shape[0]
    if k is None or k >= n:
        return decoded_level
    order = np.argsort(-decoded_level.scores, kind='stable')[:k]
    order = np.sort(order)
    return BoxesDecoded(decoded_level.boxes[order], decoded_level.scores[order], decoded_level.class_ids[order])


def postprocess_and_mask(inp, trainable_masks, image_size, levels, anchor_size_mode='tf1_int_trunc'):
    """utils.py:258-284.  inp['detection'] = {'classifications': Classification(dict,dict),
    'regressions': dict}; per-level tensors are [B,H,W,A,.]."""
    cls = inp['detection']['classifications']
    regs = inp['detection']['regressions']
    post = {pn: regression_postprocess(regs[pn], anchor_size_normalised(levels[pn].anchor_sizes, image_size,
                                                                        anchor_size_mode))
            for pn in regs}
    detection = Detection(classification=cls, regression=regs, regression_postprocessed=post)

    def masked(d):
        return merge_outputs({pn: d[pn][trainable_masks[pn]] for pn in d})

    trainable = Detection(
        classification=Classification(
            unscaled=masked(cls.unscaled) if cls.unscaled else None,
            prob=masked(cls.prob) if cls.prob else None),
        regression=masked(regs),
        regression_postprocessed=masked(post))
    return {**inp, 'detection': detection, 'detection_trainable': trainable}


def process_labels_and_logits(labels, logits, levels, anchor_size_mode='tf1_int_trunc'):
    """utils.py:240-255."""
    labels = {**labels, 'detection': {**labels['detection'], 'classifications': Classification(
        unscaled=None, prob=labels['detection']['classifications'])}}
    c = logits['detection']['classifications']
    logits = {**logits, 'detection': {**logits['detection'], 'classifications': Classification(
        unscaled=c, prob={k: sigmoid(v) for k, v in c.items()})}}
    image_size = np.asarray(labels['image'].shape[1:3], np.int32)
    labels = postprocess_and_mask(labels, labels['trainable_masks'], image_size, levels, anchor_size_mode)
    logits = postprocess_and_mask(logits, labels['trainable_masks'], image_size, levels, anchor_size_mode)
    return labels, logits


# --------------------------------------------------------------------------------------
# losses.py
# --------------------------------------------------------------------------------------
def _rsum(x, axis=None):
    """tf.reduce_sum over float32 data.  TF's CPU reduction is an Eigen tree whose order is not
    knowable here, and NumPy's axis-0 reduction is a naive sequential add whose float32 error
    grows like sqrt(M) ulp (1e-5 relative at M ~ 1e5).  The oracle therefore accumulates in
    float64 and rounds ONCE to float32: the correctly rounded value any good float32 reduction
    is close to."""
    return np.sum(np.asarray(x), axis=axis, dtype=np.float64).astype(F32)


def _rmean(x):
    x = np.asarray(x)
    return F32(np.sum(x, dtype=np.float64) / max(x.size, 1)) if x.size else F32(np.nan)


def bce_with_logits(labels, logits):
    """tf.nn.sigmoid_cross_entropy_with_logits (losses.py:107,124):
    max(x,0) - x*z + log1p(exp(-|x|)), float32."""
    z = np.asarray(labels, F32)
    x = np.asarray(logits, F32)
    return (np.maximum(x, F32(0)) - (x * z).astype(F32) +
            np.log1p(np.exp(-np.abs(x)).astype(F32)).astype(F32)).astype(F32)


def focal_sigmoid_cross_entropy_with_logits(labels, logits, focus=2.0, alpha=0.25, eps=1e-7):
    """losses.py:6-15 (elementwise, unreduced; naive `1 - p` and eps inside the log)."""
    z = np.asarray(labels, F32)
    x = np.asarray(logits, F32)
    a = np.full(z.shape, alpha, F32)
    prob = sigmoid(x)
    is1 = z == 1
    prob_true = np.where(is1, prob, (F32(1) - prob).astype(F32))
    a = np.where(is1, a, (F32(1) - a).astype(F32))
    with np.errstate(divide='ignore'):
        mod = np.power((F32(1) - prob_true).astype(F32), F32(focus)).astype(F32)
        return ((-a * mod).astype(F32) * np.log((prob_true + F32(eps)).astype(F32)).astype(F32)).astype(F32)


def focal_softmax_cross_entropy_with_logits(labels, logits, focus=2.0, alpha=0.25, eps=1e-7):
    """losses.py:19-34."""
    z = np.asarray(labels, F32)
    x = np.asarray(logits, F32)
    a = (np.ones_like(z) * F32(alpha)).astype(F32)
    e = np.exp((x - x.max(-1, keepdims=True)).astype(F32)).astype(F32)
    prob = (e / e.sum(-1, keepdims=True, dtype=F32)).astype(F32)
    is1 = z == 1
    a_bal = np.where(is1, a, (F32(1) - a).astype(F32))
    prob_true = np.where(is1, prob, (F32(1) - prob).astype(F32))
    mod = np.power((F32(1) - prob_true).astype(F32), F32(focus)).astype(F32)
    log_prob = np.log((prob + F32(eps)).astype(F32)).astype(F32)
    return (-np.sum((a_bal * mod * z * log_prob).astype(F32), -1, dtype=F32)).astype(F32)


def _sums(labels, logits, axis):
    z = np.asarray(labels, F32)
    p = sigmoid(logits)
    inter = _rsum((z * p).astype(F32), axis)
    return z, p, inter


def jaccard_loss(labels, logits, smooth=1., axis=None):
    """losses.py:37-47."""
    z, p, inter = _sums(labels, logits, axis)
    union = _rsum(z, axis) + _rsum(p, axis)
    with np.errstate(divide='ignore', invalid='ignore'):
        jac = (inter + F32(smooth)) / (union - inter + F32(smooth))
        return ((F32(1) - jac) * F32(smooth)).astype(F32)


def dice_loss(labels, logits, smooth=1., axis=None):
    """losses.py:50-60."""
    z, p, inter = _sums(labels, logits, axis)
    union = _rsum(z, axis) + _rsum(p, axis)
    with np.errstate(divide='ignore', invalid='ignore'):
        coef = (F32(2) * inter + F32(smooth)) / (union + F32(smooth))
        return (F32(1) - coef).astype(F32)


def fixed_iou_loss(labels, logits, smooth=1., axis=0):
    """losses.py:63-73."""
    z, p, inter = _sums(labels, logits, axis)
    union = _rsum(z, axis) + _rsum(((F32(1) - z) * p).astype(F32), axis)
    with np.errstate(divide='ignore', invalid='ignore'):
        return (F32(1) - (inter + F32(smooth)) / (union + F32(smooth))).astype(F32)


def balanced_sigmoid_cross_entropy_with_logits(labels, logits, axis=None):
    """losses.py:96-110 (weights broadcast over `axis`; no keepdims, as in the reference)."""
    z = np.asarray(labels, F32)
    n_pos = _rsum(z, axis)
    n_neg = _rsum((F32(1) - z).astype(F32), axis)
    with np.errstate(divide='ignore', invalid='ignore'):
        w_pos = (n_neg / (n_pos + n_neg)).astype(F32)
        w_neg = (n_pos / (n_pos + n_neg)).astype(F32)
    ones = np.ones_like(z)
    weight = np.where(z == 1, ones * w_pos, ones * w_neg).astype(F32)
    return (bce_with_logits(z, logits) * weight).astype(F32)


CLS_MODES = ('bce_dice', 'focal', 'bce', 'balanced_bce', 'balanced_bce_dice', 'focal_dice')


def classification_loss(labels, logits, fg_mask, mode='bce_dice'):
    """losses.py:115-141.  mode 'bce_dice' is the code as committed (mean BCE + mean dice,
    smooth=0, axis=0); the other modes enable the commented-out alternatives the way the
    comments spell them (focal: sum / max(num_fg, 1), losses.py:119-122; balanced BCE with
    axis=0, losses.py:127-128)."""
    labels = np.asarray(labels, F32)
    logits = np.asarray(logits, F32)
    total = F32(0)
    if mode in ('focal', 'focal_dice'):
        focal = focal_sigmoid_cross_entropy_with_logits(labels, logits)
        num_fg = _rsum(np.asarray(fg_mask).astype(F32))
        total = total + _rsum(focal) / np.maximum(num_fg, F32(1.0))
    if mode in ('bce', 'bce_dice'):
        total = total + _rmean(bce_with_logits(labels, logits))
    if mode in ('balanced_bce', 'balanced_bce_dice'):
        total = total + _rmean(balanced_sigmoid_cross_entropy_with_logits(labels, logits, axis=0))
    if mode in ('bce_dice', 'balanced_bce_dice', 'focal_dice'):
        total = total + _rmean(dice_loss(labels, logits, smooth=0., axis=0))
    if mode not in CLS_MODES:
        raise ValueError(mode)
    return F32(total)


def huber(labels, predictions, delta=1.0):
    """tf.losses.huber_loss elementwise part (losses.py:146)."""
    e = (np.asarray(predictions, F32) - np.asarray(labels, F32)).astype(F32)
    ae = np.abs(e)
    q = np.minimum(ae, F32(delta))
    lin = (ae - q).astype(F32)
    return (F32(0.5) * (q * q).astype(F32) + F32(delta) * lin).astype(F32)


def regression_loss(labels, logits, fg_mask):
    """losses.py:144-152: Huber, weights fg[:,None], SUM_BY_NONZERO_WEIGHTS -> sum / (#non-zero
    broadcast weights), 0 when there is none."""
    h = huber(labels, logits)
    w = np.asarray(fg_mask).astype(F32).reshape(-1, 1)
    w = np.broadcast_to(w, h.shape)
    num_present = np.sum(w != 0, dtype=np.float64)
    s = _rsum((h * w).astype(F32))
    return F32(s / F32(num_present)) if num_present > 0 else F32(0)


def loss(labels, logits, mode='bce_dice'):
    """losses.py:155-175.  labels/logits are `Detection`s of [M,.] arrays."""
    fg_mask = classmap_decode(labels.classification.prob).fg_mask
    class_loss = classification_loss(labels.classification.prob, logits.classification.unscaled, fg_mask, mode)
    regr_loss = regression_loss(labels.regression, logits.regression, fg_mask)
    return class_loss, regr_loss


# --------------------------------------------------------------------------------------
# Analytic gradients (the reference relies on TF autodiff; formulas from SURVEY A.3, and
# checked against torch.autograd in float64 by tests/test_oracle_golden.py).
# `precise=True` evaluates in float64 (used to bound conditioning in tests).
# --------------------------------------------------------------------------------------
def focal_grad(labels, logits, focus=2.0, alpha=0.25, eps=1e-7, dtype=F32):
    """d focal / d logits, elementwise, following the autodiff chain of losses.py:9-13 with the
    same float32 intermediates (p, 1-p, prob_true, 1-prob_true)."""
    T = dtype
    z = np.asarray(labels, T)
    x = np.asarray(logits, T)
    p = (T(1) / (T(1) + np.exp(-x))).astype(T)
    is1 = z == 1
    q = (T(1) - p).astype(T)
    pt = np.where(is1, p, q)
    at = np.where(is1, T(alpha), T(1) - T(alpha)).astype(T)
    w = (T(1) - pt).astype(T)
    with np.errstate(divide='ignore', invalid='ignore'):
        lg = np.log((pt + T(eps)).astype(T)).astype(T)
        # d/dpt [ w^f * log(pt+eps) ] = -f w^(f-1) log(pt+eps) + w^f / (pt+eps)
        d_pt = (-T(focus) * np.power(w, T(focus - 1)) * lg + np.power(w, T(focus)) / (pt + T(eps))).astype(T)
    d_pt = (-at * d_pt).astype(T)
    sgn = np.where(is1, T(1), T(-1)).astype(T)  # d pt / d p
    return (d_pt * sgn * p * q).astype(T)  # sigmoid grad = p (1 - p)


def classification_loss_grad(labels, logits, fg_mask, mode='bce_dice', dtype=F32):
    """d classification_loss / d logits  [M, C]."""
    T = dtype
    z = np.asarray(labels, T)
    x = np.asarray(logits, T)
    m, c = z.shape
    p = (T(1) / (T(1) + np.exp(-x))).astype(T)
    # d BCE / dx exactly as TF autodiff composes it from max(x,0) - x z + log1p(exp(-|x|)):
    # relu'(x) - z - sign(x) e/(1+e), e = exp(-|x|)  (== p - z, but without the cancellation
    # of `p - 1` when z = 1 and p -> 1)
    e = np.exp(-np.abs(x)).astype(T)
    sfrac = (e / (T(1) + e)).astype(T)
    dbce = np.where(x >= 0, (T(1) - z) - sfrac, sfrac - z).astype(T)
    g = np.zeros_like(x, dtype=T)
    if mode in ('focal', 'focal_dice'):
        num_fg = max(float(np.sum(np.asarray(fg_mask))), 1.0)
        g = g + focal_grad(z, x, dtype=T) / T(num_fg)
    if mode in ('bce', 'bce_dice'):
        g = g + dbce / T(m * c)
    if mode in ('balanced_bce', 'balanced_bce_dice'):
        n_pos = _rsum(z, 0).astype(T)
        n_neg = _rsum(T(1) - z, 0).astype(T)
        with np.errstate(divide='ignore', invalid='ignore'):
            w = np.where(z == 1, n_neg / (n_pos + n_neg), n_pos / (n_pos + n_neg)).astype(T)
        g = g + w * dbce / T(m * c)
    if mode in ('bce_dice', 'balanced_bce_dice', 'focal_dice'):
        inter = np.sum(z * p, 0, dtype=np.float64).astype(T)
        union = (np.sum(z, 0, dtype=np.float64) + np.sum(p, 0, dtype=np.float64)).astype(T)
        with np.errstate(divide='ignore', invalid='ignore'):
            # d/dp [1 - 2I/U] = -2 (z U - I) / U^2 ; mean over C
            g = g + (-(T(2) * (z * union - inter) / (union * union)) * p * (T(1) - p)) / T(c)
    return g.astype(T)


def regression_loss_grad(labels, logits, fg_mask, dtype=F32):
    """d regression_loss / d logits  [M, 4]: clip(e, -1, 1) * fg / (4 num_fg)."""
    T = dtype
    e = np.asarray(logits, T) - np.asarray(labels, T)
    fg = np.asarray(fg_mask).astype(T).reshape(-1, 1)
    n = float(np.sum(np.asarray(fg_mask))) * e.shape[1]
    if n == 0:
        return np.zeros_like(e)
    return (np.clip(e, T(-1), T(1)) * fg / T(n)).astype(T)
