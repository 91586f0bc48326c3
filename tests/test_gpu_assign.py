"""K1 parity (GPU): rn_assign_encode through the C ABI vs the oracle's dense
dataset.level_labels on the same seeded inputs.  Indices, labels and IoU maxima bit-exact;
encoded deltas within 1e-5 relative."""
import numpy as np
import pytest

import helpers
import rn_oracle as oracle

pytestmark = pytest.mark.gpu


def _run_gpu(wl, batch, first_image=0, levels=None, **kw):
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    boxes, classes, counts = synthetic.gt_batch(wl, batch, first_image)
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, batch, wl.max_gt, levels=levels, keep_iou=True,
                         process_group=False, **kw)
    path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda())
    torch.cuda.synchronize()
    path.check_status()
    return path


@pytest.mark.parametrize('name,batch', [('shapes', 3), ('voc', 2), ('coco', 1), ('crowded', 1)])
def test_assign_encode_matches_oracle(name, batch):
    from retinanet_tensorflow_b200 import synthetic
    wl = synthetic.WORKLOADS[name]
    ref = helpers.oracle_labels_batch(wl, batch)
    path = _run_gpu(wl, batch)
    idx = helpers.merge_levels(ref['idx'])
    iou = helpers.merge_levels(ref['iou'])
    lab = helpers.compact_labels(helpers.merge_levels(ref['cls'], 1), helpers.merge_levels(ref['mask']))
    reg = helpers.merge_levels(ref['reg'], 1)
    assert path.rows_total == idx.shape[0]
    np.testing.assert_array_equal(path.iou_max.cpu().numpy(), iou)       # IEEE-exact op sequence
    np.testing.assert_array_equal(path.matched.cpu().numpy(), idx.astype(np.int32))
    np.testing.assert_array_equal(path.label.cpu().numpy(), lab)
    helpers.assert_close(path.deltas.cpu().numpy(), reg, rtol=1e-5, atol=1e-7, what='deltas')
    # label statistics accumulated by the kernel
    stats = path.stats.cpu().numpy()
    assert stats[0] == (lab != -2).sum() and stats[1] == (lab >= 0).sum()
    npos = np.bincount(lab[lab >= 0], minlength=wl.num_classes)
    np.testing.assert_array_equal(stats[8:8 + wl.num_classes], npos)
    assert (lab >= 0).sum() > 0 and (lab == -2).sum() > 0  # the case exercises all three states


@pytest.mark.parametrize('name,batch', [('voc', 2), ('coco', 1), ('crowded', 1)])
def test_assign_patch_decomposition_does_not_change_results(name, batch):
    """rn_assign_encode_patch: one CTA handles sy x sx sub-patches and shares the GT staging and the log-ratio
    table between them (the table is skipped when too many boxes touch the patch -- crowded).  Every output is
    bit-identical for every decomposition, including ragged ones that overhang the grid."""
    from retinanet_tensorflow_b200 import synthetic
    wl = synthetic.WORKLOADS[name]
    base = _run_gpu(wl, batch, assign_patch=(1, 1))
    want = [t.cpu().numpy() for t in (base.matched, base.label, base.iou_max, base.deltas, base.stats)]
    for patch in [(0, 0), (2, 2), (1, 4), (3, 1), (4, 4), (5, 7), (64, 64)]:
        path = _run_gpu(wl, batch, assign_patch=patch)
        got = [t.cpu().numpy() for t in (path.matched, path.label, path.iou_max, path.deltas, path.stats)]
        for g, w, what in zip(got, want, ('matched', 'label', 'iou_max', 'deltas', 'stats')):
            np.testing.assert_array_equal(g.view(np.int32) if g.dtype == np.float32 else g,
                                          w.view(np.int32) if w.dtype == np.float32 else w,
                                          err_msg='%s with patch %s' % (what, patch))


def test_assign_big_batch_takes_the_shared_patch_path_and_matches_small_batches():
    """A launch that fills the device (automatic 2 x 2 decomposition) gives each image the labels it gets alone."""
    from retinanet_tensorflow_b200 import synthetic
    wl = synthetic.WORKLOADS['voc']
    batch = 48
    big = _run_gpu(wl, batch)
    lab = big.label.cpu().numpy()
    dl = big.deltas.cpu().numpy()
    one = _run_gpu(wl, 1, first_image=batch - 1, assign_patch=(1, 1))
    # merged level-major layout: level l holds [B, cells_l * A] rows
    off_b, off_1 = 0, 0
    sizes = [int(r) // batch for r in list(big.rows.rows)[:big.rows.num_segments]]
    for n in sizes:
        np.testing.assert_array_equal(lab[off_b + (batch - 1) * n: off_b + batch * n], one.label.cpu().numpy()[off_1:off_1 + n])
        np.testing.assert_array_equal(dl[off_b + (batch - 1) * n: off_b + batch * n].view(np.int32),
                                      one.deltas.cpu().numpy()[off_1:off_1 + n].view(np.int32))
        off_b += batch * n
        off_1 += n


@pytest.mark.parametrize('aspects,scales', [([(1, 1)], [1]),                                   # A = 1
                                            ([(1, 2), (1, 1), (2, 1), (1, 3)], [1, 1.26, 1.59, 2.0]),  # A = 16
                                            ([(1, 2), (2, 1)], [1, 1.3, 1.7, 2.1, 2.6])])              # A = 10
@pytest.mark.parametrize('patch', [(0, 0), (2, 2), (3, 1)])
def test_assign_other_anchor_counts(aspects, scales, patch):
    """Anchor sets other than the reference's 3 x 3 (1, 10 and the maximum of 16 shapes per cell): threads per cell,
    the per-CTA shape and log-ratio tables and their fall-back (16 shapes x many boxes) against the oracle."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    levels = {'P3': oracle.Level(32, aspects, scales), 'P5': oracle.Level(128, aspects, scales)}
    wl = synthetic.WORKLOADS['voc']
    batch = 2
    boxes, classes, counts = synthetic.gt_batch(wl, batch)
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, batch, wl.max_gt, levels=levels, keep_iou=True,
                         process_group=False, assign_patch=patch)
    path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda())
    torch.cuda.synchronize()
    path.check_status()
    ref = helpers.oracle_labels_batch(wl, batch, levels=levels)
    lab = helpers.compact_labels(helpers.merge_levels(ref['cls'], 1), helpers.merge_levels(ref['mask']))
    np.testing.assert_array_equal(path.iou_max.cpu().numpy(), helpers.merge_levels(ref['iou']))
    np.testing.assert_array_equal(path.matched.cpu().numpy(), helpers.merge_levels(ref['idx']).astype(np.int32))
    np.testing.assert_array_equal(path.label.cpu().numpy(), lab)
    helpers.assert_close(path.deltas.cpu().numpy(), helpers.merge_levels(ref['reg'], 1), rtol=1e-5, atol=1e-7,
                         what='deltas')


def test_assign_float_anchor_size_mode():
    from retinanet_tensorflow_b200 import synthetic
    wl = synthetic.WORKLOADS['shapes']
    ref = helpers.oracle_labels_batch(wl, 2, anchor_size_mode='float')
    path = _run_gpu(wl, 2, anchor_size_mode='float')
    np.testing.assert_array_equal(path.matched.cpu().numpy(), helpers.merge_levels(ref['idx']).astype(np.int32))
    np.testing.assert_array_equal(path.iou_max.cpu().numpy(), helpers.merge_levels(ref['iou']))


def test_assign_reference_dataset_test_vector():
    """dataset_test.py:9-28 (assignment pattern; its regression expectations are stale)."""
    import torch
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    levels = {'P4': oracle.Level(16, [(1, 1)], [1, 1.5])}
    boxes = (np.array([[0, 0, 16, 16], [8, 8, 24, 24], [16, 16, 32, 32], [-4, -4, 20, 20]], np.float64) / 32)
    boxes = boxes.astype(np.float32)[None]
    classes = np.array([[1, 2, 3, 4]], np.int32)
    path = DenseHeadPath((32, 32), 5, 1, 4, levels=levels, keep_iou=True, process_group=False)
    path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(),
                torch.tensor([4], dtype=torch.int32).cuda())
    lab = path.label.cpu().numpy().reshape(2, 2, 2)
    # expected class-id map [[[100,400],[0,0]],[[0,0],[300,0]]] with ids 1..4; the IoU 0.4444 anchor is ignored
    np.testing.assert_array_equal(lab, np.array([[[1, 4], [-1, -1]], [[-1, -1], [3, -2]]]))
    np.testing.assert_allclose(path.iou_max.cpu().numpy().reshape(2, 2, 2),
                               [[[1, 1], [1 / 7, 9 / 43]], [[1 / 7, 9 / 43], [1, 4 / 9]]], rtol=1e-6)


def test_assign_ties_first_gt_wins_and_empty_rows():
    """Two identical GT boxes: tf.argmax keeps the first.  All-zero IoU columns match GT 0."""
    import torch
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    boxes = np.zeros((1, 4, 4), np.float32)
    boxes[0, 0] = [0.70, 0.70, 0.90, 0.95]
    boxes[0, 1] = [0.10, 0.10, 0.40, 0.35]
    boxes[0, 2] = [0.10, 0.10, 0.40, 0.35]
    classes = np.array([[0, 1, 2, -1]], np.int32)
    path = DenseHeadPath((256, 256), 3, 1, 4, keep_iou=True, process_group=False)
    path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(),
                torch.tensor([3], dtype=torch.int32).cuda())
    matched = path.matched.cpu().numpy()
    iou = path.iou_max.cpu().numpy()
    assert not (matched == 2).any()          # the duplicate never wins
    assert (matched[iou == 0] == 0).all()    # argmax of an all-zero column is 0
    ref = oracle.build_labels(np.array([256, 256], np.int32), classes[0, :3], boxes[0, :3], oracle.build_levels(), 3)
    lab = helpers.compact_labels(helpers.merge_levels({k: v[None] for k, v in ref[0].items()}, 1),
                                 helpers.merge_levels({k: v[None] for k, v in ref[2].items()}))
    np.testing.assert_array_equal(path.label.cpu().numpy(), lab)


def test_assign_flags_bad_boxes_and_class_out_of_range():
    import torch
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    path = DenseHeadPath((128, 128), 3, 1, 2, process_group=False)
    boxes = torch.tensor([[[0.6, 0.2, 0.2, 0.5], [0.1, 0.1, 0.6, 0.6]]], dtype=torch.float32).cuda()  # ymin > ymax
    path.assign(boxes, torch.tensor([[0, 7]], dtype=torch.int32).cuda(), torch.tensor([2], dtype=torch.int32).cuda())
    with pytest.raises(AssertionError):
        path.check_status()
    # class id 7 >= C: one_hot gives an all-zero row -> never a foreground label
    assert int((path.label >= 0).sum()) == 0 or int(path.label.max()) < 3
