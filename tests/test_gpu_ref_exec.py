"""The CUDA path against the reference's own source EXECUTED (fixtures tests/golden/ref_exec_*.npz, written by
tests/golden/make_ref_fixtures.py from unmodified /root/reference code; see oracle/ref_exec.py).  Everything goes
through the C ABI (DenseHeadPath -> ctypes -> librndense.so).  Bit-exact: labels, masks, candidate and detection
indices/classes/order.  Floating point: 1e-5 relative (north_star) with the absolute floors documented in
DESIGN.md section 2 (the reference's own quantisation steps)."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ['shapes', 'voc_small', 'wide']
NAMES = ['P3', 'P4', 'P5', 'P6', 'P7']


def fixture(case):
    return np.load(os.path.join(HERE, 'golden', 'ref_exec_%s.npz' % case))


def _path(fx, mode, **kw):
    import torch
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    B = int(fx['batch'])
    C = int(fx['num_classes'])
    max_gt = max(len(fx['classes_%d' % i]) for i in range(B))
    boxes = np.zeros((B, max_gt, 4), np.float32)
    classes = np.zeros((B, max_gt), np.int32)
    counts = np.zeros((B,), np.int32)
    for i in range(B):
        n = len(fx['classes_%d' % i])
        boxes[i, :n], classes[i, :n], counts[i] = fx['boxes_%d' % i], fx['classes_%d' % i], n
    path = DenseHeadPath(tuple(int(v) for v in fx['image_size']), C, B, max_gt, cls_mode=mode, process_group=False, **kw)
    path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda())
    return path


def _merged(fx, key, trailing):
    return helpers.merge_levels({pn: fx[key + pn] for pn in NAMES}, trailing)


@pytest.mark.parametrize('case', CASES)
def test_assign_equals_executed_reference(case):
    """dataset.build_labels (dataset.py:43-142): labels bit-exact, deltas within 1e-5 relative."""
    fx = fixture(case)
    path = _path(fx, 'focal')
    cls = _merged(fx, 'cls_', 1).astype(np.float32)
    mask = _merged(fx, 'mask_', 0)
    want = helpers.compact_labels(cls, mask)
    got = path.label.cpu().numpy()
    assert np.array_equal(got, want)
    reg = _merged(fx, 'reg_', 1)
    helpers.assert_close(path.deltas.cpu().numpy(), reg, rtol=1e-5, atol=1e-6, what='deltas')
    path.check_status()


def _inputs(fx):
    import torch
    logits = {pn: torch.from_numpy(fx['logits_' + pn]).cuda() for pn in NAMES}
    regs = {pn: torch.from_numpy(fx['regs_' + pn]).cuda() for pn in NAMES}
    return logits, regs


@pytest.mark.parametrize('mode', ['bce_dice', 'focal', 'bce', 'balanced_bce'])
@pytest.mark.parametrize('case', CASES)
def test_loss_equals_executed_reference(case, mode):
    """losses.py on utils.process_labels_and_logits' masked tensors: values and gradients (tf.gradients with TF's
    registered gradient formulas in float32)."""
    import torch
    fx = fixture(case)
    path = _path(fx, mode)
    logits, regs = _inputs(fx)
    out = path.loss(logits, regs).cpu().numpy()
    torch.cuda.synchronize()
    mask = _merged(fx, 'mask_', 0)
    np.testing.assert_allclose(out[0], fx['val_' + mode], rtol=1e-5)
    np.testing.assert_allclose(out[1], fx['val_huber'], rtol=1e-5)
    assert out[2] == fx['fg_mask'].sum() and out[3] == mask.sum() == fx['lab_prob'].shape[0]
    g = path.dlogits.cpu().numpy()
    assert not g[~mask].any()
    ref = fx['dx_' + mode]
    helpers.assert_close(g[mask], ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max(), what='dlogits ' + mode)
    gr = path.dreg.cpu().numpy()
    assert not gr[~mask].any()
    # the reference's Huber gradient is quantised at 1 ulp of 1/(4 num_fg) (tests/test_ref_exec.py explains)
    helpers.assert_close(gr[mask], fx['dx_huber'], rtol=1e-5, atol=2.4e-7 * np.abs(fx['dx_huber']).max(), what='dreg')
    if mode == 'bce_dice':   # the committed losses.loss through the whole chain, per level
        np.testing.assert_allclose(out[0], fx['loss_class'], rtol=1e-5)
        np.testing.assert_allclose(out[1], fx['loss_regr'], rtol=1e-5)
        full = _merged(fx, 'grad_logits_', 1)
        helpers.assert_close(g, full, rtol=1e-5, atol=1e-6 * np.abs(full).max(), what='grad_logits maps')


@pytest.mark.parametrize('case', ['voc_small'])
def test_bf16_loss_close_to_executed_reference(case):
    """bfloat16 logits in / gradients out (float32 math): the loss of the ROUNDED logits is what is computed, so the
    comparison is against the reference evaluated on bf16-rounded inputs is not available; bound the difference by
    bfloat16's own resolution instead (2^-8 relative on the logits -> a few 1e-3 on the loss)."""
    import torch
    fx = fixture(case)
    path = _path(fx, 'focal')
    logits, regs = _inputs(fx)
    lb = {pn: v.to(torch.bfloat16) for pn, v in logits.items()}
    out = path.loss(lb, regs).cpu().numpy()
    np.testing.assert_allclose(out[0], fx['val_focal'], rtol=2e-2)
    np.testing.assert_allclose(out[1], fx['val_huber'], rtol=1e-5)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('case', CASES)
def test_detections_equal_executed_reference(case, fused):
    """train.py:68-75: boxes_decode per level -> merge -> nms_classwise.  Candidates (anchor order, class) and kept
    detections (class-major, selection order) bit-exact; scores within 1 ulp of sigmoid; boxes within 1e-5."""
    import torch
    fx = fixture(case)
    path = _path(fx, 'focal')
    logits, regs = _inputs(fx)
    if fused:
        path.loss_and_select(logits, regs)
    else:
        path.select(logits, regs)
    path.nms()
    torch.cuda.synchronize()
    path.check_status()
    cc, dc = path.cand_count.cpu().numpy(), path.det_count.cpu().numpy()
    for i in range(int(fx['batch'])):
        n = len(fx['cand_scores_%d' % i])
        assert cc[i] == n
        assert np.array_equal(path.cand_classes[i, :n].cpu().numpy(), fx['cand_classes_%d' % i])
        helpers.assert_close(path.cand_scores[i, :n].cpu().numpy(), fx['cand_scores_%d' % i], rtol=3e-7, what='scores')
        helpers.assert_close(path.cand_boxes[i, :n].cpu().numpy(), fx['cand_boxes_%d' % i], rtol=1e-5, atol=1e-6,
                             what='candidate boxes')
        k = len(fx['det_scores_%d' % i])
        assert dc[i] == k
        assert np.array_equal(path.det_classes[i, :k].cpu().numpy(), fx['det_classes_%d' % i])
        helpers.assert_close(path.det_scores[i, :k].cpu().numpy(), fx['det_scores_%d' % i], rtol=3e-7, what='det scores')
        helpers.assert_close(path.det_boxes[i, :k].cpu().numpy(), fx['det_boxes_%d' % i], rtol=1e-5, atol=1e-6,
                             what='detection boxes')


@pytest.mark.parametrize('case', ['shapes', 'wide'])
def test_flip_equals_executed_reference(case):
    """augmentation.flip (augmentation.py:5-22) of image 0's label maps, on the device."""
    import ctypes as C
    import torch
    from retinanet_tensorflow_b200 import _lib
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    fx = fixture(case)
    Cn = int(fx['num_classes'])
    n = len(fx['classes_0'])
    path = DenseHeadPath(tuple(int(v) for v in fx['image_size']), Cn, 2, n, cls_mode='focal', process_group=False)
    boxes = torch.from_numpy(fx['boxes_0'][None]).cuda()
    classes = torch.from_numpy(fx['classes_0'][None]).cuda()
    counts = torch.tensor([n], dtype=torch.int32).cuda()
    path.assign(boxes, classes, counts, with_flip=True)   # image 1 of every level = flipped image 0
    lab = helpers.split_levels(path.label.cpu().numpy(), {pn: np.zeros((2,) + fx['flip_mask_' + pn].shape) for pn in NAMES})
    dl = helpers.split_levels(path.deltas.cpu().numpy(), {pn: np.zeros((2,) + fx['flip_mask_' + pn].shape) for pn in NAMES})
    for pn in NAMES:
        want = helpers.compact_labels(fx['flip_cls_' + pn].reshape(-1, Cn).astype(np.float32), fx['flip_mask_' + pn].reshape(-1))
        assert np.array_equal(lab[pn][1].reshape(-1), want), pn
        helpers.assert_close(dl[pn][1], fx['flip_reg_' + pn], rtol=1e-5, atol=1e-6, what='flipped deltas ' + pn)
    del C, _lib
