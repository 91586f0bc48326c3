"""The reference-facing modules (levels / dataset / utils / losses) on the GPU.  The first block
restates the reference's own tests (same inputs, same expectations; file:line cited) against
the drop-in functions; the second block compares the remaining signatures with the oracle."""
import os

import numpy as np
import pytest

import helpers
import rn_oracle as oracle

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def ref_tests():
    return np.load(os.path.join(G, 'reference_tests.npz'))


def _np(t):
    return t.detach().cpu().numpy()


# ---- the reference's own tests, re-expressed -------------------------------------------------------
def test_boxmap_anchor_relative_to_image_relative(ref_tests):  # utils_test.py:7-42
    from retinanet_tensorflow_b200 import utils
    a = _np(utils.boxmap_anchor_relative_to_image_relative(ref_tests['g2_in']))
    assert np.allclose(a, ref_tests['g2_out']) and a.shape == (1, 3, 4, 1, 4)
    np.testing.assert_array_equal(a, oracle.boxmap_anchor_relative_to_image_relative(ref_tests['g2_in']))


def test_anchor_boxmap(ref_tests):  # utils_test.py:44-74
    from retinanet_tensorflow_b200 import utils
    a = _np(utils.anchor_boxmap(ref_tests['g3_grid'], ref_tests['g3_sizes']))
    assert np.allclose(a, ref_tests['g3_out']) and a.shape == (1, 3, 4, 1, 4)
    np.testing.assert_array_equal(a, oracle.anchor_boxmap(ref_tests['g3_grid'], ref_tests['g3_sizes']))


def test_boxmap_center_relative_to_corner_relative(ref_tests):  # utils_test.py:76-97
    from retinanet_tensorflow_b200 import utils
    a = _np(utils.boxmap_center_relative_to_corner_relative(ref_tests['g4_in']))
    assert np.array_equal(a, ref_tests['g4_out']) and a.shape == (1, 3, 4, 1, 4)


def test_iou(ref_tests):  # utils_test.py:99-118
    from retinanet_tensorflow_b200 import utils
    a = _np(utils.iou(ref_tests['g5_a'], ref_tests['g5_b']))
    assert np.allclose(a, ref_tests['g5_out']) and a.shape == (4,)
    np.testing.assert_array_equal(a, oracle.iou(ref_tests['g5_a'], ref_tests['g5_b']))


def test_classmap_decode(ref_tests):  # utils_test.py:120-138 (fg pattern)
    from retinanet_tensorflow_b200 import utils
    np.testing.assert_array_equal(_np(utils.classmap_decode(ref_tests['g6_in']).fg_mask), ref_tests['g6_fg'])


def test_merge_outputs(ref_tests):  # losses_test.py:7-15
    import torch
    from retinanet_tensorflow_b200 import utils
    outputs = {'a': torch.tensor(ref_tests['g7_a']).cuda(), 'b': torch.tensor(ref_tests['g7_b']).cuda()}
    masks = {'a': torch.tensor(ref_tests['g7_mask_a']).cuda(), 'b': torch.tensor(ref_tests['g7_mask_b']).cuda()}
    merged = utils.merge_outputs(utils.dict_starmap(lambda x, m: x[m], (outputs, masks)))
    np.testing.assert_array_equal(_np(merged), ref_tests['g7_out'])


def test_regression_loss(ref_tests):  # losses_test.py:17-27
    from retinanet_tensorflow_b200 import losses
    loss = losses.regression_loss(labels=ref_tests['g8_labels'], logits=ref_tests['g8_logits'],
                                  fg_mask=ref_tests['g8_fg'])
    assert np.array_equal(_np(loss), ref_tests['g8_out'])


def test_level_labels(ref_tests):  # dataset_test.py:8-28 (5-arg call + id map are stale; pattern valid)
    from retinanet_tensorflow_b200 import dataset
    from retinanet_tensorflow_b200.levels import Level
    ids = ref_tests['g9_class_ids']
    cls, reg, mask = dataset.level_labels(np.array([32, 32], np.int32), np.arange(4), ref_tests['g9_boxes'],
                                          Level(16, [(1, 1)], [1, 1.5]), 2 ** 4, 4)
    cls, reg, mask = _np(cls), _np(reg), _np(mask)
    assert cls.shape == (2, 2, 2, 4) and reg.shape == (2, 2, 2, 4) and mask.shape == (2, 2, 2)
    class_map = np.where(cls.max(-1) > 0, ids[cls.argmax(-1)], 0)
    np.testing.assert_array_equal(class_map, ref_tests['g9_class_map'])
    o_cls, o_reg, o_mask = oracle.level_labels(np.array([32, 32], np.int32), np.arange(4), ref_tests['g9_boxes'],
                                               oracle.Level(16, [(1, 1)], [1, 1.5]), 2 ** 4, 4)
    np.testing.assert_array_equal(cls, o_cls)
    np.testing.assert_array_equal(mask, o_mask)
    helpers.assert_close(reg, o_reg, rtol=1e-5, atol=1e-7, what='regression')


def test_scale_regression(ref_tests):  # retinanet_old_test.py:15-37
    from retinanet_tensorflow_b200 import utils
    a = _np(utils.scale_regression(ref_tests['g10_in'], ref_tests['g10_sizes']))
    assert np.array_equal(a, ref_tests['g10_out']) and a.shape == (1, 1, 1, 2, 4)


def test_levels_module():  # levels_test.py:5-14
    from retinanet_tensorflow_b200.levels import Level, compute_box_size
    box_size = compute_box_size(32, (1, 2), 1)
    assert len(box_size) == 2 and np.isclose(box_size.prod(), 32 ** 2) and box_size[1] / box_size[0] == 2
    assert np.array_equal(Level(32, [(1, 4)], [2 ** 0, 2 ** 1]).anchor_sizes, [[16, 64], [32, 128]])


# ---- the remaining signatures vs the oracle ---------------------------------------------------------
def test_position_grid_and_box_forms():
    from retinanet_tensorflow_b200 import dataset
    for size in ([63, 84], [1, 1], [7, 11]):
        np.testing.assert_array_equal(_np(dataset.position_grid(size)), oracle.position_grid(size))
    rng = np.random.default_rng(1)
    tl = rng.uniform(0, 0.5, (37, 2))
    box = np.concatenate([tl, tl + rng.uniform(0.01, 0.5, (37, 2))], -1).astype(np.float32)
    c = _np(dataset.to_center_box(box))
    np.testing.assert_array_equal(c, oracle.to_center_box(box))
    np.testing.assert_array_equal(_np(dataset.from_center_box(c)), oracle.from_center_box(c))


def test_build_labels_dicts_match_oracle():
    from retinanet_tensorflow_b200 import dataset, synthetic
    from retinanet_tensorflow_b200.levels import build_levels
    wl = synthetic.WORKLOADS['shapes']
    boxes, classes = synthetic.gt_for_image(wl, 5)
    cls, reg, mask = dataset.build_labels(np.array([500, 500], np.int32), classes, boxes, build_levels(), 3)
    o_cls, o_reg, o_mask = oracle.build_labels(np.array([500, 500], np.int32), classes, boxes, oracle.build_levels(), 3)
    assert list(cls) == list(o_cls) == ['P3', 'P4', 'P5', 'P6', 'P7']
    for pn in o_cls:
        np.testing.assert_array_equal(_np(cls[pn]), o_cls[pn])
        np.testing.assert_array_equal(_np(mask[pn]), o_mask[pn])
        helpers.assert_close(_np(reg[pn]), o_reg[pn], rtol=1e-5, atol=1e-7, what='regression ' + pn)


def test_build_labels_rejects_invalid_boxes():
    from retinanet_tensorflow_b200 import dataset
    from retinanet_tensorflow_b200.levels import build_levels
    with pytest.raises(AssertionError):  # utils.py:66-67
        dataset.build_labels([128, 128], [0], np.array([[0.6, 0.2, 0.2, 0.5]], np.float32), build_levels(), 3)
    with pytest.raises(AssertionError):  # inferred.py:30
        dataset.build_labels([128, 128], np.zeros((0,), np.int32), np.zeros((0, 4), np.float32), build_levels(), 3)


def test_iou_broadcasts_and_asserts():
    from retinanet_tensorflow_b200 import utils
    rng = np.random.default_rng(2)
    tl = rng.uniform(0, 0.5, (6, 1, 2))
    a = np.concatenate([tl, tl + rng.uniform(0.01, 0.5, (6, 1, 2))], -1).astype(np.float32)
    tl = rng.uniform(0, 0.5, (1, 9, 2))
    b = np.concatenate([tl, tl + rng.uniform(0.01, 0.5, (1, 9, 2))], -1).astype(np.float32)
    np.testing.assert_array_equal(_np(utils.iou(a, b)), oracle.iou(a, b))
    with pytest.raises(AssertionError):
        utils.iou(np.array([[0.5, 0.1, 0.2, 0.3]], np.float32), np.array([[0.1, 0.1, 0.3, 0.3]], np.float32))
    with pytest.raises(AssertionError):
        utils.iou(np.ones((1, 4), np.float32), np.ones((1, 4), np.float32))  # 0/0


def test_regression_postprocess_matches_oracle():
    from retinanet_tensorflow_b200 import utils
    rng = np.random.default_rng(3)
    reg = rng.normal(0, 0.5, (2, 7, 5, 3, 4)).astype(np.float32)
    sizes = rng.uniform(0.05, 0.5, (3, 2)).astype(np.float32)
    helpers.assert_close(_np(utils.regression_postprocess(reg, sizes)), oracle.regression_postprocess(reg, sizes),
                         rtol=1e-5, atol=1e-6, what='regression_postprocess')


def test_boxes_decode_nms_and_classwise_match_oracle():
    import torch
    from retinanet_tensorflow_b200 import utils
    rng = np.random.default_rng(4)
    probs = rng.uniform(0, 0.62, (9, 11, 9, 5)).astype(np.float32)
    ctr = rng.uniform(0.2, 0.8, (9, 11, 9, 2))
    sz = rng.uniform(0.05, 0.3, (9, 11, 9, 2))
    boxes = np.concatenate([ctr - sz / 2, ctr + sz / 2], -1).astype(np.float32)
    got = utils.boxes_decode(probs, boxes)
    want = oracle.boxes_decode(probs, boxes)
    assert got.class_ids.dtype == torch.int64 and want.scores.shape[0] > 50
    np.testing.assert_array_equal(_np(got.boxes), want.boxes)
    np.testing.assert_array_equal(_np(got.scores), want.scores)
    np.testing.assert_array_equal(_np(got.class_ids), want.class_ids)
    merged = utils.merge_boxes_decoded([got, got])
    assert merged.scores.shape[0] == 2 * want.scores.shape[0]
    g1, w1 = utils.nms(got), oracle.nms(want)
    np.testing.assert_array_equal(_np(g1.boxes), w1.boxes)
    np.testing.assert_array_equal(_np(g1.scores), w1.scores)
    np.testing.assert_array_equal(_np(g1.class_ids), w1.class_ids)
    g7 = utils.nms(got, max_output_size=7)
    np.testing.assert_array_equal(_np(g7.scores), oracle.nms(want, max_output_size=7).scores)
    g2, w2 = utils.nms_classwise(got, num_classes=5), oracle.nms_classwise(want, 5)
    np.testing.assert_array_equal(_np(g2.boxes), w2.boxes)
    np.testing.assert_array_equal(_np(g2.scores), w2.scores)
    np.testing.assert_array_equal(_np(g2.class_ids), w2.class_ids)
    empty = utils.boxes_decode(np.zeros((4, 4, 2, 5), np.float32), boxes[:4, :4, :2])
    assert empty.scores.shape[0] == 0 and utils.nms_classwise(empty, 5).boxes.shape == (0, 4)
    # class ids outside [0, num_classes) never equal a class in utils.py:201-206: those candidates are dropped
    odd = want.class_ids.copy()
    odd[::7] = -1
    odd[3::11] = 5
    odd[5::13] = 77
    g3 = utils.nms_classwise(utils.BoxesDecoded(got.boxes, got.scores, torch.from_numpy(odd).cuda()), num_classes=5)
    w3 = oracle.nms_classwise(oracle.BoxesDecoded(want.boxes, want.scores, odd), 5)
    np.testing.assert_array_equal(_np(g3.boxes), w3.boxes)
    np.testing.assert_array_equal(_np(g3.scores), w3.scores)
    np.testing.assert_array_equal(_np(g3.class_ids), w3.class_ids)
    assert w3.scores.shape[0] > 10 and set(np.unique(w3.class_ids)) <= set(range(5))


def test_dict_helpers():
    from retinanet_tensorflow_b200 import utils
    assert utils.dict_map(lambda v: v + 1, {'a': 1, 'b': 2}) == {'a': 2, 'b': 3}
    assert utils.dict_starmap(lambda x, y: x * y, [{'a': 2, 'b': 3}, {'a': 5, 'b': 7}]) == {'a': 10, 'b': 21}
    with pytest.raises(AssertionError):  # utils.py:165
        utils.dict_starmap(lambda x, y: x, [{'a': 1}, {'b': 1}])
    assert utils.dict_update({'x': {'y': 1, 'z': 2}}, ['x', 'y'], lambda v: v + 10) == {'x': {'y': 11, 'z': 2}}
    assert utils.all_same([1, 1, 1]) and not utils.all_same([1, 2])


def _loss_inputs(m=300, c=6, seed=5):
    rng = np.random.default_rng(seed)
    z = np.zeros((m, c), np.float32)
    pos = rng.choice(m, m // 10, replace=False)
    z[pos, rng.integers(0, c, pos.shape[0])] = 1
    x = rng.normal(-3, 2, (m, c)).astype(np.float32)
    return z, x


def test_elementwise_losses_and_autograd():
    import torch
    from retinanet_tensorflow_b200 import losses
    z, x = _loss_inputs()
    xt = torch.tensor(x, device='cuda', requires_grad=True)
    out = losses.focal_sigmoid_cross_entropy_with_logits(z, xt)
    helpers.assert_close(_np(out), oracle.focal_sigmoid_cross_entropy_with_logits(z, x), rtol=1e-5, atol=1e-9,
                         what='focal')
    out.sum().backward()
    g_ref = oracle.focal_grad(z, x)
    helpers.assert_close(_np(xt.grad), g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max(), what='focal grad')
    for axis in (None, 0):
        out = losses.balanced_sigmoid_cross_entropy_with_logits(z, x, axis=axis)
        helpers.assert_close(_np(out), oracle.balanced_sigmoid_cross_entropy_with_logits(z, x, axis=axis), rtol=1e-5,
                             atol=1e-9, what='balanced axis=%s' % axis)
    out = losses.focal_sigmoid_cross_entropy_with_logits(z, x, focus=1.5, alpha=0.4, eps=1e-6)
    helpers.assert_close(_np(out), oracle.focal_sigmoid_cross_entropy_with_logits(z, x, 1.5, 0.4, 1e-6), rtol=2e-5,
                         atol=1e-9, what='focal 1.5')


@pytest.mark.parametrize('fn', ['dice_loss', 'jaccard_loss', 'fixed_iou_loss'])
@pytest.mark.parametrize('axis', [None, 0])
def test_column_sum_losses(fn, axis):
    import torch
    from retinanet_tensorflow_b200 import losses
    z, x = _loss_inputs()
    for smooth in (1.0, 1e-7):
        xt = torch.tensor(x, device='cuda', requires_grad=True)
        got = getattr(losses, fn)(z, xt, smooth=smooth, axis=axis)
        want = getattr(oracle, fn)(z, x, smooth=smooth, axis=axis)
        helpers.assert_close(_np(got), want, rtol=1e-5, atol=1e-7, what=fn)
        got.sum().backward()
        # float64 autograd reference for the gradient
        x64 = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        z64 = torch.tensor(z, dtype=torch.float64)
        p = torch.sigmoid(x64)
        dim = 0 if axis == 0 else None
        red = (lambda t: t.sum(0)) if dim == 0 else (lambda t: t.sum())
        i, zz, pp = red(z64 * p), red(z64), red(p)
        val = {'dice_loss': 1 - (2 * i + smooth) / (zz + pp + smooth),
               'jaccard_loss': (1 - (i + smooth) / (zz + pp - i + smooth)) * smooth,
               'fixed_iou_loss': 1 - (i + smooth) / (zz + red((1 - z64) * p) + smooth)}[fn]
        (g,) = torch.autograd.grad(val.sum(), x64)
        helpers.assert_close(_np(xt.grad), g.numpy(), rtol=2e-5, atol=1e-6 * np.abs(g.numpy()).max(),
                             what=fn + ' grad')


@pytest.mark.parametrize('mode', [None, 'focal', 'balanced_bce'])
def test_loss_on_detections_matches_oracle(mode):
    """losses.loss on masked Detections, the call train.py:220 makes."""
    import torch
    from retinanet_tensorflow_b200 import losses, utils
    z, x = _loss_inputs(500, 4, seed=8)
    rng = np.random.default_rng(9)
    t = rng.normal(0, 1, (500, 4)).astype(np.float32)
    r = (t + rng.normal(0, 0.7, (500, 4))).astype(np.float32)
    xt = torch.tensor(x, device='cuda', requires_grad=True)
    rt = torch.tensor(r, device='cuda', requires_grad=True)
    labels = utils.Detection(utils.Classification(None, torch.tensor(z).cuda()), torch.tensor(t).cuda(), None)
    logits = utils.Detection(utils.Classification(xt, utils.sigmoid(xt.detach())), rt, None)
    class_loss, regr_loss = losses.loss(labels, logits, mode=mode)
    omode = 'bce_dice' if mode is None else mode
    o_labels = oracle.Detection(oracle.Classification(None, z), t, None)
    o_logits = oracle.Detection(oracle.Classification(x, oracle.sigmoid(x)), r, None)
    oc, orr = oracle.loss(o_labels, o_logits, mode=omode)
    np.testing.assert_allclose(_np(class_loss), oc, rtol=1e-5)
    np.testing.assert_allclose(_np(regr_loss), orr, rtol=1e-5)
    (class_loss + regr_loss).backward()
    fg = z.max(-1) > 0.5
    g_ref = oracle.classification_loss_grad(z, x, fg, mode=omode)
    helpers.assert_close(_np(xt.grad), g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max(), what='dlogits')
    gr_ref = oracle.regression_loss_grad(t, r, fg)
    helpers.assert_close(_np(rt.grad), gr_ref, rtol=1e-5, atol=1e-6 * np.abs(gr_ref).max(), what='dreg')


def test_process_labels_and_logits_matches_oracle():
    """utils.process_labels_and_logits on the reference's dict structure (train.py:218)."""
    import torch
    from retinanet_tensorflow_b200 import synthetic, utils
    from retinanet_tensorflow_b200.levels import build_levels
    wl = synthetic.WORKLOADS['shapes']
    ref = helpers.oracle_labels_batch(wl, 2)
    rng = np.random.default_rng(10)
    logit_cls = {pn: rng.normal(-3, 2, v.shape).astype(np.float32) for pn, v in ref['cls'].items()}
    logit_reg = {pn: rng.normal(0, 0.3, v.shape).astype(np.float32) for pn, v in ref['reg'].items()}
    image = np.zeros((2, wl.image_h, wl.image_w, 3), np.float32)

    def cu(d, dtype=None):
        return {k: torch.tensor(v).cuda() for k, v in d.items()}

    labels = {'image': torch.tensor(image).cuda(), 'trainable_masks': cu(ref['mask']),
              'detection': {'classifications': cu(ref['cls']), 'regressions': cu(ref['reg'])}}
    logits = {'detection': {'classifications': cu(logit_cls), 'regressions': cu(logit_reg)}}
    got_l, got_p = utils.process_labels_and_logits(labels, logits, build_levels())
    o_labels = {'image': image, 'trainable_masks': ref['mask'],
                'detection': {'classifications': ref['cls'], 'regressions': ref['reg']}}
    o_logits = {'detection': {'classifications': logit_cls, 'regressions': logit_reg}}
    want_l, want_p = oracle.process_labels_and_logits(o_labels, o_logits, oracle.build_levels())
    for got, want in ((got_l, want_l), (got_p, want_p)):
        gt, wt = got['detection_trainable'], want['detection_trainable']
        np.testing.assert_array_equal(_np(gt.regression), wt.regression)
        helpers.assert_close(_np(gt.regression_postprocessed), wt.regression_postprocessed, rtol=1e-5, atol=1e-6,
                             what='regression_postprocessed')
        helpers.assert_close(_np(gt.classification.prob), wt.classification.prob, rtol=1e-6, what='prob')
        if wt.classification.unscaled is not None:
            np.testing.assert_array_equal(_np(gt.classification.unscaled), wt.classification.unscaled)
        else:
            assert gt.classification.unscaled is None


def test_focal_softmax_matches_oracle_and_autograd():
    import torch
    from retinanet_tensorflow_b200 import losses
    z, x = _loss_inputs(200, 7, seed=12)
    xt = torch.tensor(x, device='cuda', requires_grad=True)
    got = losses.focal_softmax_cross_entropy_with_logits(z, xt)
    want = oracle.focal_softmax_cross_entropy_with_logits(z, x)
    assert got.shape == (200,)
    helpers.assert_close(_np(got), want, rtol=1e-5, atol=1e-8, what='focal softmax')
    got.sum().backward()
    x64 = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    z64 = torch.tensor(z, dtype=torch.float64)
    p = torch.softmax(x64, -1)
    a = torch.where(z64 == 1, torch.tensor(0.25, dtype=torch.float64), torch.tensor(0.75, dtype=torch.float64))
    pt = torch.where(z64 == 1, p, 1 - p)
    val = -(a * (1 - pt) ** 2.0 * z64 * torch.log(p + 1e-7)).sum()
    (g,) = torch.autograd.grad(val, x64)
    helpers.assert_close(_np(xt.grad), g.numpy(), rtol=2e-5, atol=1e-6 * np.abs(g.numpy()).max(), what='grad')


def test_dormant_metrics():
    """train.py:137-161 (regr_iou / class_iou), composed from the drop-in functions."""
    import torch
    from retinanet_tensorflow_b200 import metrics
    rng = np.random.default_rng(13)
    z, x = _loss_inputs(400, 5, seed=14)
    tl = rng.uniform(0, 0.5, (400, 2))
    lab_boxes = np.concatenate([tl, tl + rng.uniform(0.05, 0.4, (400, 2))], -1).astype(np.float32)
    pred_boxes = (lab_boxes + rng.normal(0, 0.01, lab_boxes.shape)).astype(np.float32)
    pred_boxes[:, 2:] = np.maximum(pred_boxes[:, 2:], pred_boxes[:, :2] + 1e-3)
    got = float(metrics.regr_iou(z, lab_boxes, pred_boxes))
    a = oracle.boxes_decode(z, lab_boxes)
    b = oracle.boxes_decode(z, pred_boxes)
    want = float(oracle.iou(a.boxes, b.boxes).astype(np.float64).mean())
    assert abs(got - want) < 1e-6 and 0.5 < got <= 1.0
    ci = float(metrics.class_iou(z, x))
    pz, zz = oracle.sigmoid(x) > 0.5, z > 0.5
    want_ci = np.mean([((zz == c) & (pz == c)).sum() / max(((zz == c) | (pz == c)).sum(), 1) for c in (False, True)])
    assert abs(ci - want_ci) < 1e-6


def test_label_path_is_reentrant_from_four_threads():
    """dataset.py:231 maps build_labels over <= 4 tf.data threads: the library keeps no mutable global
    state, so concurrent calls (ctypes releases the GIL) give the same labels as sequential ones."""
    import threading
    import torch
    from retinanet_tensorflow_b200 import dataset, synthetic
    from retinanet_tensorflow_b200.levels import build_levels
    wl = synthetic.WORKLOADS['voc']
    levels = build_levels()
    images = [synthetic.gt_for_image(wl, i) for i in range(12)]
    size = [wl.image_h, wl.image_w]

    def labels(i):
        b, c = images[i]
        cls, reg, mask = dataset.build_labels(size, c, b, levels, wl.num_classes)
        return [x[pn].clone() for x in (cls, reg, mask) for pn in levels]

    want = [labels(i) for i in range(len(images))]
    got = [None] * len(images)
    errors = []

    def worker(k):
        try:
            torch.cuda.set_device(0)
            for i in range(k, len(images), 4):
                got[i] = labels(i)
            torch.cuda.synchronize()
        except Exception as err:   # surfaced below: a thread must not fail silently
            errors.append(err)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for w, g in zip(want, got):
        assert all(torch.equal(a, b) for a, b in zip(w, g))
