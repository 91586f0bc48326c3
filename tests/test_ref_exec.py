"""Pins oracle/rn_oracle.py to the reference's own source EXECUTED (unmodified dataset.py / utils.py / losses.py /
augmentation.py on the NumPy TensorFlow stand-in, oracle/ref_exec.py).  The committed fixtures
tests/golden/ref_exec_*.npz are what the reference returned (tests/golden/make_ref_fixtures.py); the tests that need
/root/reference itself (stub validation on the reference's own unit tests, fixture freshness) skip where it is absent.
CPU only."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, 'oracle'), HERE, os.path.join(HERE, 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)

import rn_oracle as oracle  # noqa: E402
import ref_exec  # noqa: E402
from helpers import assert_close  # noqa: E402

CASES = ['shapes', 'voc_small', 'wide']
needs_reference = pytest.mark.skipif(not ref_exec.available(), reason='reference checkout not present (GPU box)')


def fixture(case):
    return np.load(os.path.join(HERE, 'golden', 'ref_exec_%s.npz' % case))


# ---- the stub itself: the reference's own unit tests must pass on it ---------------------------------------------
@needs_reference
def test_reference_unit_tests_pass_on_the_stub():
    res = ref_exec.run_reference_tests()
    must_pass = ['utils_test.UtilsTest.test_anchor_boxmap',
                 'utils_test.UtilsTest.test_boxmap_anchor_relative_to_image_relative',
                 'utils_test.UtilsTest.test_boxmap_center_relative_to_corner_relative',
                 'utils_test.UtilsTest.test_iou',
                 'losses_test.ObjectivesTest.test_regression_loss']
    for t in must_pass:
        assert res[t] == 'ok', (t, res[t])
    # stale in the reference itself (they call signatures the code no longer has): the failure must be THAT one
    assert 'num_classes' in res['dataset_test.DatasetTest.test_level_labels']
    assert 'positional argument' in res['augmentation_test.AugmentationTest.test_flip']
    assert 'merge_outputs' in res['losses_test.ObjectivesTest.test_merge_outputs']
    assert res['utils_test.UtilsTest.test_classmap_decode'].startswith('fail')   # returns a namedtuple now


@needs_reference
def test_fixture_is_what_the_reference_returns_now():
    import make_ref_fixtures
    out = make_ref_fixtures.run_case(ref_exec.load(), 'wide')
    fx = fixture('wide')
    assert set(out) == set(fx.files)
    for k in fx.files:
        assert np.array_equal(np.asarray(out[k]), fx[k], equal_nan=True) if fx[k].dtype.kind == 'f' else \
            np.array_equal(np.asarray(out[k]), fx[k]), k


def test_stub_gradients_match_torch_autograd():
    """tf.gradients of the stand-in (TF's registered gradient formulas) against torch.autograd on the same
    composite expressions, in float64 where rounding cannot hide a wrong formula."""
    import torch
    sys.path.insert(0, ref_exec.STUB)
    saved = sys.modules.pop('tensorflow', None)
    try:
        import tensorflow as tf
        rng = np.random.default_rng(0)
        x = rng.normal(0, 2, (50, 6))
        z = (rng.random((50, 6)) < 0.2).astype(np.float64)
        w = rng.random(50) < 0.5
        tx = tf.convert_to_tensor(x)
        tz = tf.convert_to_tensor(z)
        p = tf.nn.sigmoid(tx)
        pt = tf.where(tf.equal(tz, 1), p, 1 - p)
        val = tf.reduce_sum(-(1 - pt) ** 2.0 * tf.log(pt + 1e-7)) / tf.maximum(tf.reduce_sum(tz), 1.0) \
            + tf.reduce_mean(tf.nn.sigmoid_cross_entropy_with_logits(labels=tz, logits=tx)) \
            + tf.reduce_mean(1 - 2 * tf.reduce_sum(tz * p, axis=0) / (tf.reduce_sum(tz, axis=0) + tf.reduce_sum(p, axis=0))) \
            + tf.reduce_sum(tf.boolean_mask(tf.abs(tx), w)) + tf.reduce_sum(tf.minimum(tf.abs(tx), 1.0) * tf.exp(tx[:, :1]))
        g = tf.gradients(val, tx)[0].numpy()
        X = torch.tensor(x, requires_grad=True)
        Z = torch.tensor(z)
        P = torch.sigmoid(X)
        PT = torch.where(Z == 1, P, 1 - P)
        V = (-(1 - PT) ** 2.0 * torch.log(PT + 1e-7)).sum() / Z.sum().clamp(min=1.0) \
            + torch.nn.functional.binary_cross_entropy_with_logits(X, Z) \
            + (1 - 2 * (Z * P).sum(0) / (Z.sum(0) + P.sum(0))).mean() \
            + X.abs()[torch.tensor(w)].sum() + (X.abs().clamp(max=1.0) * torch.exp(X[:, :1])).sum()
        V.backward()
        assert abs(float(val.numpy()) - float(V.detach())) < 1e-9 * abs(float(V.detach()))
        np.testing.assert_allclose(g, X.grad.numpy(), rtol=1e-9, atol=1e-12)
    finally:
        sys.modules.pop('tensorflow', None)
        if saved is not None:
            sys.modules['tensorflow'] = saved
        sys.path.remove(ref_exec.STUB)


# ---- the oracle against the executed reference ----------------------------------------------------------------------
def _levels():
    return oracle.build_levels()


@pytest.mark.parametrize('case', CASES)
def test_oracle_labels_equal_reference(case):
    """dataset.build_labels: one-hot classifications and trainable masks bit-exact; encoded deltas equal where they
    are finite on both sides (the reference's `regression * one_hot` sum is the matched row exactly, up to -0)."""
    fx = fixture(case)
    levels = _levels()
    size = fx['image_size']
    C = int(fx['num_classes'])
    for i in range(int(fx['batch'])):
        cls, reg, mask = oracle.build_labels(size, fx['classes_%d' % i], fx['boxes_%d' % i], levels, C)
        for pn in levels:
            assert np.array_equal(cls[pn].astype(np.uint8), fx['cls_' + pn][i]), (case, pn, 'classification')
            assert np.array_equal(mask[pn], fx['mask_' + pn][i]), (case, pn, 'trainable mask')
            assert_close(reg[pn], fx['reg_' + pn][i], rtol=1e-6, atol=1e-7, what='%s %s deltas' % (case, pn))


def _detections(fx, levels):
    names = list(levels.keys())
    B = int(fx['batch'])
    h, w = [int(v) for v in fx['image_size']]
    labels = {'image': np.zeros((B, h, w, 3), np.float32),
              'detection': {'classifications': {pn: fx['cls_' + pn].astype(np.float32) for pn in names},
                            'regressions': {pn: fx['reg_' + pn] for pn in names}},
              'trainable_masks': {pn: fx['mask_' + pn] for pn in names}}
    logits = {'detection': {'classifications': {pn: fx['logits_' + pn] for pn in names},
                            'regressions': {pn: fx['regs_' + pn] for pn in names}}}
    return oracle.process_labels_and_logits(labels, logits, levels)


@pytest.mark.parametrize('case', CASES)
def test_oracle_masking_equals_reference(case):
    """utils.process_labels_and_logits: row order (b, y, x, a), level concatenation P3..P7, decoded boxes."""
    fx = fixture(case)
    lab, log = _detections(fx, _levels())
    lt, gt = lab['detection_trainable'], log['detection_trainable']
    assert np.array_equal(lt.classification.prob, fx['lab_prob'])
    assert np.array_equal(gt.classification.unscaled, fx['log_unscaled'])
    assert np.array_equal(lt.regression, fx['lab_reg'], equal_nan=True)
    assert np.array_equal(gt.regression, fx['log_reg'])
    assert_close(gt.classification.prob, fx['log_prob'], rtol=2e-7, what='sigmoid')
    assert_close(gt.regression_postprocessed, fx['log_regpp'], rtol=1e-6, atol=1e-7, what='decoded boxes (logits)')
    ok = np.isfinite(fx['lab_regpp']).all(-1)
    assert_close(lt.regression_postprocessed[ok], fx['lab_regpp'][ok], rtol=1e-6, atol=1e-7, what='decoded boxes (labels)')


@pytest.mark.parametrize('case', CASES)
def test_oracle_losses_equal_reference(case):
    """Every loss of losses.py: value within 1e-5 relative (north_star), gradients within 1e-5 relative plus the
    absolute floor DESIGN.md section 2 derives (1e-6 of the largest gradient element: the reference's own
    `1 - (1 - p)` quantises)."""
    fx = fixture(case)
    z, x, fg = fx['lab_prob'], fx['log_unscaled'], fx['fg_mask']
    M, C = z.shape
    assert np.array_equal(oracle.classmap_decode(z).fg_mask, fg)
    # committed classification_loss + regression_loss through losses.loss
    lab, log = _detections(fx, _levels())
    cl, rl = oracle.loss(lab['detection_trainable'], log['detection_trainable'])
    assert_close(cl, fx['loss_class'], rtol=1e-5, what='class loss')
    assert_close(rl, fx['loss_regr'], rtol=1e-5, what='regression loss')
    assert_close(oracle.focal_sigmoid_cross_entropy_with_logits(z, x), fx['elem_focal'], rtol=1e-5, atol=1e-12,
                 what='elementwise focal')

    def check(name, value, grad):
        assert_close(value, fx['val_' + name], rtol=1e-5, what=name + ' value')
        ref = fx['dx_' + name]
        assert_close(grad, ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max(), what=name + ' gradient')

    check('focal', oracle.classification_loss(z, x, fg, 'focal'), oracle.classification_loss_grad(z, x, fg, 'focal'))
    check('bce', oracle.classification_loss(z, x, fg, 'bce'), oracle.classification_loss_grad(z, x, fg, 'bce'))
    check('balanced_bce', oracle.classification_loss(z, x, fg, 'balanced_bce'),
          oracle.classification_loss_grad(z, x, fg, 'balanced_bce'))
    check('bce_dice', oracle.classification_loss(z, x, fg, 'bce_dice'), oracle.classification_loss_grad(z, x, fg, 'bce_dice'))
    dice = oracle._rmean(oracle.dice_loss(z, x, smooth=0., axis=0))
    gd = oracle.classification_loss_grad(z, x, fg, 'bce_dice') - oracle.classification_loss_grad(z, x, fg, 'bce')
    assert_close(dice, fx['val_dice'], rtol=1e-5, what='dice value')
    assert_close(gd, fx['dx_dice'], rtol=1e-4, atol=1e-6 * np.abs(fx['dx_dice']).max(), what='dice gradient')
    assert_close(oracle._rmean(oracle.jaccard_loss(z, x, axis=0)), fx['val_jaccard'], rtol=1e-5, what='jaccard')
    assert_close(oracle._rmean(oracle.fixed_iou_loss(z, x, axis=0, smooth=1e-7)), fx['val_fixed_iou'], rtol=1e-5,
                 what='fixed iou')
    assert_close(np.sum(oracle.focal_softmax_cross_entropy_with_logits(z, x), dtype=np.float64), fx['val_focal_softmax'],
                 rtol=1e-5, what='softmax focal')
    assert_close(oracle.regression_loss(fx['lab_reg'], fx['log_reg'], fg), fx['val_huber'], rtol=1e-5, what='huber')
    # TF's autodiff of 0.5 q^2 + (|e| - q), q = min(|e|, 1), sums (q g - g) + g: exact in real numbers, but in
    # float32 the cancellation leaves an absolute error of ~1 ulp of g = 1/(4 num_fg) on every element -- the
    # reference's own gradient is quantised at that step, so small elements cannot agree to 1e-5 relative
    gmax = np.abs(fx['dx_huber']).max()
    assert_close(oracle.regression_loss_grad(fx['lab_reg'], fx['log_reg'], fg), fx['dx_huber'], rtol=1e-5,
                 atol=2.4e-7 * gmax, what='huber gradient')


@pytest.mark.parametrize('case', CASES)
def test_oracle_full_chain_gradients_equal_reference(case):
    """tf.gradients of class_loss + regr_loss w.r.t. the per-level head outputs (through boolean_mask and concat):
    zero on ignored anchors, the masked gradient elsewhere."""
    fx = fixture(case)
    levels = _levels()
    z, x, fg = fx['lab_prob'], fx['log_unscaled'], fx['fg_mask']
    g = oracle.classification_loss_grad(z, x, fg, 'bce_dice')
    gr = oracle.regression_loss_grad(fx['lab_reg'], fx['log_reg'], fg)
    off = 0
    for pn in levels:
        m = fx['mask_' + pn]
        n = int(m.sum())
        full = np.zeros(fx['logits_' + pn].shape, np.float32)
        full[m] = g[off:off + n]
        ref = fx['grad_logits_' + pn]
        assert_close(full, ref, rtol=1e-5, atol=1e-6 * max(np.abs(ref).max(), 1e-30), what='%s d/dlogits' % pn)
        fullr = np.zeros(fx['regs_' + pn].shape, np.float32)
        fullr[m] = gr[off:off + n]
        assert_close(fullr, fx['grad_regs_' + pn], rtol=1e-5, atol=2.4e-7 * max(np.abs(gr).max(), 1e-30),
                     what='%s d/dregs' % pn)   # see the Huber note in test_oracle_losses_equal_reference
        off += n
    assert off == z.shape[0]


@pytest.mark.parametrize('case', CASES)
def test_oracle_decode_and_nms_equal_reference(case):
    """train.py:68-75's chain: candidate order, classes and kept detections bit-exact; boxes within 1e-6."""
    fx = fixture(case)
    levels = _levels()
    names = list(levels.keys())
    _, log = _detections(fx, levels)
    det = log['detection']
    C = int(fx['num_classes'])
    for i in range(int(fx['batch'])):
        decoded = [oracle.boxes_decode(det.classification.prob[pn][i], det.regression_postprocessed[pn][i]) for pn in names]
        merged = oracle.merge_boxes_decoded(decoded)
        assert np.array_equal(merged.class_ids, fx['cand_classes_%d' % i])
        assert_close(merged.scores, fx['cand_scores_%d' % i], rtol=2e-7, what='candidate scores')
        assert_close(merged.boxes, fx['cand_boxes_%d' % i], rtol=1e-6, atol=1e-7, what='candidate boxes')
        # NMS on the REFERENCE's candidates, so a 1-ulp score difference cannot reorder anything
        ref_cand = oracle.BoxesDecoded(fx['cand_boxes_%d' % i], fx['cand_scores_%d' % i], fx['cand_classes_%d' % i])
        kept = oracle.nms_classwise(ref_cand, C)
        assert np.array_equal(kept.class_ids, fx['det_classes_%d' % i])
        assert np.array_equal(kept.scores, fx['det_scores_%d' % i])
        assert np.array_equal(kept.boxes, fx['det_boxes_%d' % i])


@pytest.mark.parametrize('case', CASES)
def test_oracle_flip_equals_reference(case):
    fx = fixture(case)
    levels = _levels()
    names = list(levels.keys())
    sample = {'image': np.zeros((4, 4, 3), np.float32),
              'detection': {'classifications': {pn: fx['cls_' + pn][0].astype(np.float32) for pn in names},
                            'regressions': {pn: fx['reg_' + pn][0] for pn in names}},
              'trainable_masks': {pn: fx['mask_' + pn][0] for pn in names}}
    out = oracle.flip(sample)
    for pn in names:
        assert np.array_equal(out['detection']['classifications'][pn].astype(np.uint8), fx['flip_cls_' + pn])
        assert np.array_equal(out['detection']['regressions'][pn], fx['flip_reg_' + pn], equal_nan=True)
        assert np.array_equal(out['trainable_masks'][pn], fx['flip_mask_' + pn])
