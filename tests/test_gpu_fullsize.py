"""Full BASELINE size (COCO shape, 64 images, 12.8 M anchors, 1.03 G logits) on the GPU, checked
through size-independent properties of the domain instead of the (too slow) oracle:
encode -> decode round trip, label statistics, batch-permutation invariance, determinism,
fused == unfused, and the invariants of a greedy per-class NMS result."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def coco():
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip('needs ~20 GB of device memory')
    wl = synthetic.WORKLOADS['coco']
    boxes, classes, counts = synthetic.gt_batch(wl, wl.batch)
    gt = [torch.from_numpy(x).cuda() for x in (boxes, classes, counts)]
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, wl.batch, wl.max_gt, capacity=8192, keep_iou=True,
                         process_group=False)
    path.assign(*gt)
    logits, regs = synthetic.head_outputs(path.label, path.deltas, wl.num_classes, seed=42, device='cuda')
    yield wl, path, gt, logits, regs
    del path, logits, regs
    torch.cuda.empty_cache()


def _gather_gt(path, gt_boxes):
    """GT box of every anchor's match, in the merged layout."""
    import torch
    out = torch.empty((path.rows_total, 4), device='cuda')
    m = path.per_level(path.matched)
    o = path.per_level(out, (4,))
    for pn in path.names:
        idx = m[pn].long().clamp(min=0)                       # [B,H,W,A]
        b = torch.arange(path.batch, device='cuda').view(-1, 1, 1, 1).expand_as(idx)
        o[pn].copy_(gt_boxes[b, idx])
    return out


def test_encode_decode_round_trip(coco):
    """utils.regression_postprocess(dataset-encoded deltas) gives back the matched GT box for
    EVERY anchor (the reference encodes all anchors, dataset.py:121)."""
    import torch
    from retinanet_tensorflow_b200 import _lib
    wl, path, gt, logits, regs = coco
    decoded = torch.empty_like(path.deltas)
    _lib.call('rn_decode_boxes', C.byref(path.pyr), path.batch, C.byref(path._merged_maps(path.deltas, 16)),
              C.byref(path._merged_maps(decoded, 16)), _lib.stream_ptr())
    want = _gather_gt(path, gt[0])
    err = (decoded - want).abs().max().item()
    assert err < 2e-6, err          # coordinates are O(1): 2e-6 absolute ~ 1e-5 relative of a small box side


def test_label_statistics_and_states(coco):
    import torch
    wl, path, gt, logits, regs = coco
    lab, iou = path.label, path.iou_max
    assert int((lab != -2).sum()) == int(path.stats[0].item())
    assert int((lab >= 0).sum()) == int(path.stats[1].item())
    npos = torch.bincount(lab[lab >= 0].long(), minlength=wl.num_classes).double()
    assert torch.equal(npos, path.stats[8:8 + wl.num_classes])
    # the three states are exactly the IoU intervals of dataset.py:83,87
    assert bool(((iou >= 0.5) == (lab >= 0)).all())
    assert bool((((iou >= 0.4) & (iou < 0.5)) == (lab == -2)).all())
    assert bool(((iou < 0.4) == (lab == -1)).all())
    assert 0 <= float(iou.min()) and float(iou.max()) <= 1.0
    # argmax of an all-zero column is 0
    assert bool((path.matched[iou == 0] == 0).all())


def test_batch_permutation_invariance(coco):
    """Images are independent: permuting the batch permutes labels / detections and leaves the
    (globally normalised) losses unchanged up to summation order."""
    import torch
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl, path, gt, logits, regs = coco
    path.step(*gt, logits, regs)
    torch.cuda.synchronize()
    base_loss = path.loss_out.clone()
    base_det = path.det_count.clone()
    base_lab = {k: v.clone() for k, v in path.per_level(path.label).items()}
    perm = torch.randperm(wl.batch, generator=torch.Generator().manual_seed(3)).cuda()
    gt_p = [g[perm].contiguous() for g in gt]
    lg = torch.empty_like(logits)
    rg = torch.empty_like(regs)
    for src, dst, w in ((logits, lg, wl.num_classes), (regs, rg, 4)):
        s_lv, d_lv = path.per_level(src, (w,)), path.per_level(dst, (w,))
        for pn in path.names:
            d_lv[pn].copy_(s_lv[pn][perm])
    other = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, wl.batch, wl.max_gt, capacity=8192,
                          process_group=False)
    other.step(*gt_p, lg, rg)
    torch.cuda.synchronize()
    assert torch.allclose(other.loss_out, base_loss, rtol=1e-6, atol=0)
    assert torch.equal(other.det_count, base_det[perm])
    lab_p = other.per_level(other.label)
    for pn in path.names:
        assert torch.equal(lab_p[pn], base_lab[pn][perm])
    del other, lg, rg


def test_determinism_and_fused_equals_unfused(coco):
    import torch
    wl, path, gt, logits, regs = coco
    path.step(*gt, logits, regs)
    a = (path.loss_out.clone(), path.dlogits.clone(), path.det_count.clone(), path.det_index.clone())
    path.step(*gt, logits, regs)
    assert torch.equal(a[0], path.loss_out) and torch.equal(a[1], path.dlogits)
    assert torch.equal(a[2], path.det_count) and torch.equal(a[3], path.det_index)
    path.step(*gt, logits, regs, fused=False)
    assert torch.equal(a[0], path.loss_out) and torch.equal(a[1], path.dlogits)
    assert torch.equal(a[2], path.det_count) and torch.equal(a[3], path.det_index)
    # gradient sanity: zero on ignored anchors, finite everywhere, loss sums positive
    g = path.dlogits
    assert bool(torch.isfinite(g).all()) and float(g[path.label == -2].abs().max()) == 0.0
    assert float(path.loss_out[0]) > 0 and float(path.loss_out[1]) > 0


def test_nms_result_invariants(coco):
    """Class-major, descending score within a class, every score > 0.5, no two kept boxes of a
    class overlap by more than 0.5, and every candidate that was dropped overlaps a kept box of
    its class with a higher (or equal, earlier) score by more than 0.5."""
    import torch
    wl, path, gt, logits, regs = coco
    path.step(*gt, logits, regs)
    torch.cuda.synchronize()

    def iou(a, b):
        tl = torch.maximum(a[:, None, :2], b[None, :, :2])
        br = torch.minimum(a[:, None, 2:], b[None, :, 2:])
        inter = (br - tl).clamp(min=0).prod(-1)
        aa = (a[:, 2:] - a[:, :2]).prod(-1)
        ab = (b[:, 2:] - b[:, :2]).prod(-1)
        return inter / (aa[:, None] + ab[None, :] - inter)

    checked = 0
    for b in range(0, wl.batch, 7):
        m = int(path.det_count[b])
        n = int(path.cand_count[b])
        cls, sc, bx = path.det_classes[b, :m], path.det_scores[b, :m], path.det_boxes[b, :m]
        assert bool((sc > 0.5).all())
        assert bool((cls[1:] >= cls[:-1]).all())
        same = cls[1:] == cls[:-1]
        assert bool((sc[1:][same] <= sc[:-1][same]).all())
        o = iou(bx, bx)
        o[cls[:, None] != cls[None, :]] = 0
        o.fill_diagonal_(0)
        assert float(o.max()) <= 0.5
        kept = torch.zeros(n, dtype=torch.bool, device='cuda')
        kept[path.det_index[b, :m].long()] = True
        cb, cs, cc = path.cand_boxes[b, :n], path.cand_scores[b, :n], path.cand_classes[b, :n]
        dropped = torch.nonzero(~kept).squeeze(1)
        if dropped.numel():
            od = iou(cb[dropped], bx)
            ok = (od > 0.5) & (cc[dropped][:, None] == cls[None, :]) & (sc[None, :] >= cs[dropped][:, None])
            assert bool(ok.any(1).all())
        checked += m
    assert checked > 100
