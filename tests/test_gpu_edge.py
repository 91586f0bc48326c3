"""Edge cases of the path on the GPU: flip (augmentation.py), ragged / empty ground truth,
capacity overflow, candidate lists larger than the shared-memory sort, class counts that are
not a multiple of four, extreme thresholds."""
import numpy as np
import pytest

import helpers
import rn_oracle as oracle

pytestmark = pytest.mark.gpu


def test_flip_reference_vector():
    """augmentation_test.py:7-45 (positional 4-tuple signature is stale; semantics valid)."""
    import torch
    from retinanet_tensorflow_b200 import augmentation
    sample = {
        'image': np.array([[1, 2], [3, 4]]),
        'detection': {
            'classifications': {'P3': np.array([[[1], [2]], [[3], [4]]])},
            'regressions': {'P3': np.array([[[[0., 0., .5, .5]], [[.25, .25, .5, .5]]],
                                            [[[1., 1., .5, .5]], [[.75, .75, .25, .25]]]], np.float32)},
        },
        'trainable_masks': {'P3': np.array([[[True], [False]], [[False], [True]]])},
    }
    out = augmentation.flip(sample)
    np.testing.assert_array_equal(out['image'].cpu().numpy(), [[2, 1], [4, 3]])
    np.testing.assert_array_equal(out['detection']['classifications']['P3'].cpu().numpy(), [[[2], [1]], [[4], [3]]])
    np.testing.assert_array_equal(out['detection']['regressions']['P3'].cpu().numpy(),
                                  [[[[.25, -.25, .5, .5]], [[0., -0., .5, .5]]],
                                   [[[.75, -.75, .25, .25]], [[1., -1., .5, .5]]]])
    np.testing.assert_array_equal(out['trainable_masks']['P3'].cpu().numpy(), [[[False], [True]], [[True], [False]]])


def test_flip_compact_batch_matches_oracle():
    import torch
    from retinanet_tensorflow_b200 import augmentation, synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS['shapes']
    boxes, classes, counts = synthetic.gt_batch(wl, 2)
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, 2, wl.max_gt, process_group=False)
    path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda())
    lab, dl = augmentation.flip_compact(path.pyr, 2, path.label, path.deltas)
    lab_lv = path.per_level(path.label)
    dl_lv = path.per_level(path.deltas, (4,))
    got_lab = path.per_level(lab)
    got_dl = path.per_level(dl, (4,))
    for pn in path.names:
        for b in range(2):
            sample = {'detection': {'classifications': {pn: lab_lv[pn][b].cpu().numpy()[..., None]},
                                    'regressions': {pn: dl_lv[pn][b].cpu().numpy()}},
                      'trainable_masks': {pn: lab_lv[pn][b].cpu().numpy() != -2}}
            want = oracle.flip(sample)
            np.testing.assert_array_equal(got_lab[pn][b].cpu().numpy(), want['detection']['classifications'][pn][..., 0])
            np.testing.assert_array_equal(got_dl[pn][b].cpu().numpy(), want['detection']['regressions'][pn])
    # flipping twice is the identity
    lab2, dl2 = augmentation.flip_compact(path.pyr, 2, lab, dl)
    assert torch.equal(lab2, path.label) and torch.equal(dl2, path.deltas)


def test_ragged_and_empty_ground_truth():
    """Counts 0 / 1 / max in one batch: padding rows never participate; the empty image is
    all-background and flagged (the reference asserts >= 1 box, inferred.py:30)."""
    import torch
    from retinanet_tensorflow_b200 import _lib, synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS['voc']
    boxes, classes, counts = synthetic.gt_batch(wl, 3)
    counts[:] = [0, 1, wl.max_gt]
    b2, c2 = synthetic.gt_for_image(wl, 2)
    rng = np.random.default_rng(0)
    reps = -(-wl.max_gt // b2.shape[0])
    boxes[2] = np.tile(b2, (reps, 1))[:wl.max_gt]
    classes[2] = np.tile(c2, reps)[:wl.max_gt]
    boxes[0] = rng.uniform(0.1, 0.9, boxes[0].shape)   # garbage in the padding of the empty image
    boxes[1, 1:] = 0.5
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, 3, wl.max_gt, keep_iou=True, process_group=False)
    path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda())
    lab = path.per_level(path.label)
    mat = path.per_level(path.matched)
    status = int(path.status.item())
    assert status & _lib.RN_FLAG_EMPTY_GT
    with pytest.raises(AssertionError):
        path.check_status()
    levels = oracle.build_levels()
    for b in (1, 2):
        n = counts[b]
        o_cls, o_reg, o_mask = oracle.build_labels(helpers.image_size_of(wl), classes[b, :n], boxes[b, :n], levels,
                                                   wl.num_classes)
        for pn in levels:
            want = helpers.compact_labels(o_cls[pn].reshape(-1, wl.num_classes), o_mask[pn].reshape(-1))
            np.testing.assert_array_equal(lab[pn][b].cpu().numpy().reshape(-1), want)
    for pn in levels:
        assert int((lab[pn][0] != -1).sum()) == 0 and int((mat[pn][0] != -1).sum()) == 0


def test_candidate_overflow_flag_and_large_lists():
    """score_thr so low that every anchor fires: (a) capacity overflow is flagged and the output
    is truncated, (b) with enough capacity the > 16384-entry global-memory sort path returns
    exactly the oracle's candidates."""
    import torch
    from retinanet_tensorflow_b200 import _lib, synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS['shapes']
    boxes, classes, counts = synthetic.gt_batch(wl, 1)
    for cap in (1000, 50000):
        path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, 1, wl.max_gt, score_thr=0.001, capacity=cap,
                             process_group=False)
        path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda())
        logits, regs = synthetic.head_outputs(path.label, path.deltas, wl.num_classes, seed=5, device='cuda')
        probs = torch.sigmoid(logits)
        path.select(probs, regs, scores_are_logits=False)
        torch.cuda.synchronize()
        status = int(path.status.item())
        n = int(path.cand_count[0])
        p = probs.cpu().numpy()
        fg = p.max(-1) > np.float32(0.001)
        if cap == 1000:
            assert status & _lib.RN_FLAG_CAND_OVERFLOW and n == cap
        else:
            assert status == 0 and n == fg.sum() and n > 16384
            np.testing.assert_array_equal(path.cand_anchor[0, :n].cpu().numpy(), np.nonzero(fg)[0])
            np.testing.assert_array_equal(path.cand_scores[0, :n].cpu().numpy(), p.max(-1)[fg])
            np.testing.assert_array_equal(path.cand_classes[0, :n].cpu().numpy(), p.argmax(-1)[fg])
            path.nms()
            cand = oracle.BoxesDecoded(path.cand_boxes[0, :n].cpu().numpy(), path.cand_scores[0, :n].cpu().numpy(),
                                       path.cand_classes[0, :n].cpu().numpy().astype(np.int64))
            _, want = oracle.nms_classwise(cand, wl.num_classes, return_indices=True)
            m = int(path.det_count[0])
            np.testing.assert_array_equal(path.det_index[0, :m].cpu().numpy(), want)


@pytest.mark.parametrize('num_classes', [1, 5, 7, 12])
def test_odd_class_counts(num_classes):
    """C not a multiple of 4 takes the scalar kernels; C = 12 the vector ones."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.Workload('odd', 9, 160, 224, num_classes, 6, 2)
    ref = helpers.oracle_labels_batch(wl, 2)
    cls = helpers.merge_levels(ref['cls'], 1)
    mask = helpers.merge_levels(ref['mask'])
    reg = helpers.merge_levels(ref['reg'], 1)
    lab = helpers.compact_labels(cls, mask)
    logits, regs = synthetic.head_outputs(lab, reg, num_classes, seed=3, device='cpu')
    boxes, classes, counts = synthetic.gt_batch(wl, 2)
    for mode in ('focal', 'bce_dice'):
        path = DenseHeadPath((wl.image_h, wl.image_w), num_classes, 2, wl.max_gt, cls_mode=mode, process_group=False)
        path.step(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda(),
                  logits.cuda(), regs.cuda())
        np.testing.assert_array_equal(path.label.cpu().numpy(), lab)
        x, r = logits.numpy(), regs.numpy()
        labels = oracle.Detection(oracle.Classification(None, cls[mask]), reg[mask], None)
        preds = oracle.Detection(oracle.Classification(x[mask], oracle.sigmoid(x[mask])), r[mask], None)
        cl, rl = oracle.loss(labels, preds, mode=mode)
        out = path.loss_out.cpu().numpy()
        np.testing.assert_allclose(out[:2], [cl, rl], rtol=1e-5)
        fg = oracle.classmap_decode(cls[mask]).fg_mask
        g_ref = np.zeros_like(x)
        g_ref[mask] = oracle.classification_loss_grad(cls[mask], x[mask], fg, mode=mode)
        helpers.assert_close(path.dlogits.cpu().numpy(), g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max(),
                             what='dlogits C=%d %s' % (num_classes, mode))
        # detections against the oracle fed with the GPU's own probabilities
        probs = torch.sigmoid(logits.cuda())


def test_zero_batch_and_argument_errors():
    import torch
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    path = DenseHeadPath((128, 128), 4, 1, 4, process_group=False)
    with pytest.raises(ValueError):
        path.assign(torch.zeros((1, 3, 4), device='cuda'), torch.zeros((1, 4), dtype=torch.int32, device='cuda'),
                    torch.ones((1,), dtype=torch.int32, device='cuda'))
    with pytest.raises(ValueError):
        path.loss(torch.zeros((7, 4), device='cuda'), torch.zeros((path.rows_total, 4), device='cuda'))
    with pytest.raises(ValueError):
        DenseHeadPath((128, 128), 4, 1, 4, cls_mode='nope', process_group=False)


def test_non_finite_gt_is_flagged():
    """utils.py:94: a NaN / infinite IoU fails the reference's 0 <= iou <= 1 assert."""
    import torch
    from retinanet_tensorflow_b200 import _lib
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    path = DenseHeadPath((128, 128), 3, 1, 2, process_group=False)
    boxes = torch.tensor([[[0.1, 0.1, float('inf'), 0.5], [0.1, 0.1, 0.6, 0.6]]], dtype=torch.float32).cuda()
    path.assign(boxes, torch.tensor([[0, 1]], dtype=torch.int32).cuda(), torch.tensor([2], dtype=torch.int32).cuda())
    assert int(path.status.item()) & _lib.RN_FLAG_IOU_RANGE
    with pytest.raises(AssertionError):
        path.check_status()


def test_orig_and_flip_batch_on_device():
    """dataset.py:182-204: the reference trains on [orig, flipped] pairs.  assign(with_flip=True) builds
    the pair's label maps on the device; labels, statistics, losses and gradients of the 2-image batch
    must equal the oracle's on [labels, augmentation.flip(labels)] (SURVEY config 1: Shapes, batch 2)."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS['shapes']
    ref = helpers.oracle_labels_batch(wl, 1)
    both = {k: {} for k in ('cls', 'reg', 'mask')}
    for pn in ref['cls']:
        sample = {'detection': {'classifications': {pn: ref['cls'][pn][0]}, 'regressions': {pn: ref['reg'][pn][0]}},
                  'trainable_masks': {pn: ref['mask'][pn][0]}}
        fl = oracle.flip(sample)
        both['cls'][pn] = np.stack([ref['cls'][pn][0], fl['detection']['classifications'][pn]], 0)
        both['reg'][pn] = np.stack([ref['reg'][pn][0], fl['detection']['regressions'][pn]], 0)
        both['mask'][pn] = np.stack([ref['mask'][pn][0], fl['trainable_masks'][pn]], 0)
    cls = helpers.merge_levels(both['cls'], 1)
    mask = helpers.merge_levels(both['mask'])
    reg = helpers.merge_levels(both['reg'], 1)
    boxes, classes, counts = synthetic.gt_batch(wl, 1)
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, 2, wl.max_gt, cls_mode='bce_dice', keep_iou=True,
                         process_group=False)
    path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda(),
                with_flip=True)
    torch.cuda.synchronize()
    path.check_status()
    np.testing.assert_array_equal(path.label.cpu().numpy(), helpers.compact_labels(cls, mask))
    helpers.assert_close(path.deltas.cpu().numpy(), reg, rtol=1e-5, atol=1e-7, what='deltas [orig, flip]')
    # matched index and max IoU travel with their anchors
    m_lv, i_lv = path.per_level(path.matched), path.per_level(path.iou_max)
    for pn in path.names:
        assert torch.equal(m_lv[pn][1], torch.flip(m_lv[pn][0], dims=[1]))
        assert torch.equal(i_lv[pn][1], torch.flip(i_lv[pn][0], dims=[1]))
    logits, regs = synthetic.head_outputs(path.label, path.deltas, wl.num_classes, seed=3, device='cuda')
    out = path.loss(logits, regs).cpu().numpy()
    x, r = logits.cpu().numpy(), regs.cpu().numpy()
    labels = oracle.Detection(oracle.Classification(None, cls[mask]), reg[mask], None)
    preds = oracle.Detection(oracle.Classification(x[mask], oracle.sigmoid(x[mask])), r[mask], None)
    cl, rl = oracle.loss(labels, preds, mode='bce_dice')
    fg = oracle.classmap_decode(cls[mask]).fg_mask
    np.testing.assert_allclose(out[:2], [cl, rl], rtol=1e-5)
    assert out[2] == fg.sum() and out[3] == mask.sum()
    g_ref = np.zeros_like(x)
    g_ref[mask] = oracle.classification_loss_grad(cls[mask], x[mask], fg, mode='bce_dice')
    helpers.assert_close(path.dlogits.cpu().numpy(), g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max(), what='dlogits')


def test_denormal_and_zero_sized_ground_truth_boxes():
    """A box whose height is a float32 denormal (and one of zero height) still goes through the
    reference's arithmetic: IEEE division, log of a denormal / of zero (-inf) in the encoded sizes."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.Workload('tiny', 1, 96, 128, 3, 4, 1)
    tiny = np.float32(1e-40)
    assert 0 < tiny < np.finfo(np.float32).tiny
    cases = [np.array([[0.25, 0.25, np.float32(0.25) + tiny, 0.75], [0.1, 0.2, 0.6, 0.7]], np.float32),
             np.array([[0.5, 0.25, 0.5, 0.75], [0.1, 0.2, 0.6, 0.7]], np.float32)]
    for boxes in cases:
        classes = np.array([1, 2], np.int32)
        gb = np.zeros((1, 4, 4), np.float32)
        gb[0, :2] = boxes
        gc = np.zeros((1, 4), np.int32)
        gc[0, :2] = classes
        path = DenseHeadPath((wl.image_h, wl.image_w), 3, 1, 4, process_group=False)
        path.assign(torch.from_numpy(gb).cuda(), torch.from_numpy(gc).cuda(), torch.tensor([2], dtype=torch.int32).cuda())
        torch.cuda.synchronize()
        with np.errstate(all='ignore'):
            o_cls, o_reg, o_mask = oracle.build_labels(np.array([96, 128], np.int32), classes, boxes, oracle.build_levels(), 3)
        lab = path.per_level(path.label)
        dl = path.per_level(path.deltas, (4,))
        for pn in path.names:
            np.testing.assert_array_equal(lab[pn][0].cpu().numpy(), helpers.compact_labels(o_cls[pn], o_mask[pn]))
            got, want = dl[pn][0].cpu().numpy(), o_reg[pn]
            assert np.array_equal(np.isneginf(got), np.isneginf(want))
            fin = np.isfinite(want)
            helpers.assert_close(got[fin], want[fin], rtol=1e-5, atol=1e-7, what='deltas ' + pn)


def test_prefetched_assignment_equals_inline_steps():
    """DenseHeadPath.prefetch_assign (the next step's K1 + statistics on their own stream, into the other label
    set) must reproduce a plain sequence of steps whose ground truth CHANGES from step to step."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS['voc']
    B = 4
    gts = [[torch.from_numpy(x).cuda() for x in synthetic.gt_batch(wl, B, first_image=f)] for f in (0, 10, 20, 30)]
    ref = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, B, wl.max_gt, process_group=False, capacity=8192)
    ref.assign(*gts[0])
    logits, regs = synthetic.head_outputs(ref.label, ref.deltas, wl.num_classes, seed=3, device='cuda')
    want = []
    for g in gts:
        ref.step(*g, logits, regs)
        torch.cuda.synchronize()
        want.append((ref.loss_out.clone(), ref.det_count.clone(), ref.dlogits.clone(), ref.label.clone()))
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, B, wl.max_gt, process_group=False, capacity=8192)
    path.prefetch_assign(*gts[0])
    for i, g in enumerate(gts):
        path.step(*g, logits, regs, overlap_detect=True)
        if i + 1 < len(gts):
            path.prefetch_assign(*gts[i + 1])
        path.join()
        torch.cuda.synchronize()
        assert torch.equal(path.label, want[i][3]), 'labels of step %d' % i
        assert torch.equal(path.loss_out, want[i][0]), (i, path.loss_out.tolist(), want[i][0].tolist())
        assert torch.equal(path.det_count, want[i][1]) and torch.equal(path.dlogits, want[i][2])
    path.check_status(ignore=12)


def test_head_loss_equals_separate_launches():
    """rn_head_loss (classification + Huber + final scalars in one launch) against rn_cls_loss_fwd_bwd +
    rn_huber_fwd_bwd + rn_loss_finalize: same gradients bit for bit, same scalars (the Huber sum is accumulated in
    fixed point in one and in double in the other: 1e-6)."""
    import ctypes as C
    import torch
    from retinanet_tensorflow_b200 import _lib, synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS['voc']
    B = 3
    gt = [torch.from_numpy(x).cuda() for x in synthetic.gt_batch(wl, B)]
    for mode in ('focal', 'bce_dice', 'balanced_bce'):
        path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, B, wl.max_gt, cls_mode=mode, process_group=False)
        path.assign(*gt)
        logits, regs = synthetic.head_outputs(path.label, path.deltas, wl.num_classes, seed=5, device='cuda')
        a = path.loss(logits, regs).clone()
        ga, gra = path.dlogits.clone(), path.dreg.clone()
        st = _lib.stream_ptr()
        stats = path.stats.clone()
        m_logits, _ = path._maps(logits, wl.num_classes, path._LOGIT_DTYPES)
        m_reg, _ = path._maps(regs, 4)
        gb, grb = torch.empty_like(ga), torch.empty_like(gra)
        out = torch.zeros((4,), dtype=torch.float32, device='cuda')
        ws = torch.zeros_like(path._loss_ws)
        _lib.call('rn_cls_loss_fwd_bwd', C.byref(path.rows), wl.num_classes, C.byref(m_logits), C.byref(path._m_label), 0,
                  path.mode, 2.0, 0.25, 1e-7, C.byref(path._merged_maps(gb, 4 * wl.num_classes)), _lib.RN_DTYPE_F32,
                  _lib.ptr(stats), _lib.ptr(ws), ws.numel(), st)
        _lib.call('rn_huber_fwd_bwd', C.byref(path.rows), C.byref(m_reg), C.byref(path._m_deltas), C.byref(path._m_label),
                  C.byref(_lib.Maps()), C.byref(path._merged_maps(grb, 16)), _lib.ptr(stats), _lib.ptr(ws), ws.numel(), st)
        _lib.call('rn_loss_finalize', wl.num_classes, path.mode, _lib.ptr(stats), _lib.ptr(out), st)
        torch.cuda.synchronize()
        assert torch.equal(ga, gb) and torch.equal(gra, grb), mode
        assert torch.allclose(a, out, rtol=1e-6, atol=0), (mode, a.tolist(), out.tolist())


def test_assign_with_the_largest_supported_gt_capacity():
    """max_gt = 2048 needs 90 KB of dynamic shared memory in the assign kernel (an opt-in above 48 KB): the launch must
    succeed and give the same labels as a tight capacity."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS['voc']
    boxes, classes, counts = synthetic.gt_batch(wl, 2)
    small = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, 2, wl.max_gt, process_group=False)
    small.assign(*(torch.from_numpy(x).cuda() for x in (boxes, classes, counts)))
    big_n = 2048
    bb = np.zeros((2, big_n, 4), np.float32)
    cc = np.full((2, big_n), -1, np.int32)
    bb[:, :wl.max_gt], cc[:, :wl.max_gt] = boxes, classes
    big = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, 2, big_n, process_group=False)
    big.assign(torch.from_numpy(bb).cuda(), torch.from_numpy(cc).cuda(), torch.from_numpy(counts).cuda())
    torch.cuda.synchronize()
    big.check_status()
    assert torch.equal(big.label, small.label) and torch.equal(big.matched, small.matched)
    assert torch.equal(big.deltas, small.deltas)


def test_column_sum_losses_backward_twice():
    """dice / jaccard / fixed-IoU keep only plain tensors for backward: a second backward (retain_graph) works and
    gives the same gradient."""
    import torch
    from retinanet_tensorflow_b200 import losses
    g = torch.Generator().manual_seed(0)
    z = (torch.rand((300, 8), generator=g) > 0.9).float().cuda()
    x = torch.randn((300, 8), generator=g).cuda().requires_grad_(True)
    for fn in (lambda: losses.dice_loss(z, x, smooth=0., axis=0).mean(), lambda: losses.jaccard_loss(z, x, axis=0).mean(),
               lambda: losses.fixed_iou_loss(z, x, axis=0, smooth=1e-7).mean()):
        out = fn()
        (g1,) = torch.autograd.grad(out, x, retain_graph=True)
        (g2,) = torch.autograd.grad(out, x)
        assert torch.equal(g1, g2) and torch.isfinite(g1).all() and g1.abs().sum() > 0
