"""K4/K5 parity (GPU): decode, candidate selection and class-wise NMS through the C ABI vs the
oracle's utils.regression_postprocess / boxes_decode / merge_boxes_decoded / nms_classwise.
Stage-wise on identical inputs: kept indices, classes and order are bit-exact; decoded boxes
and sigmoid within 1e-5 relative."""
import ctypes as C

import numpy as np
import pytest

import helpers
import rn_oracle as oracle

pytestmark = pytest.mark.gpu


def _case(name, batch, seed=11, **kw):
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS[name]
    boxes, classes, counts = synthetic.gt_batch(wl, batch)
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, batch, wl.max_gt, process_group=False, **kw)
    path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda())
    logits, regs = synthetic.head_outputs(path.label, path.deltas, wl.num_classes, seed=seed, device='cuda')
    return wl, path, logits, regs


def _gpu_sigmoid(x):
    import torch
    from retinanet_tensorflow_b200 import _lib
    out = torch.empty_like(x)
    _lib.call('rn_sigmoid', x.numel(), _lib.ptr(x), _lib.ptr(out), _lib.stream_ptr())
    return out


def _gpu_decode(path, regs):
    import torch
    from retinanet_tensorflow_b200 import _lib
    out = torch.empty_like(regs)
    _lib.call('rn_decode_boxes', C.byref(path.pyr), path.batch, C.byref(path._merged_maps(regs, 16)),
              C.byref(path._merged_maps(out, 16)), _lib.stream_ptr())
    return out


def _oracle_candidates(path, probs, boxes, b, topk=None):
    """utils.boxes_decode per level for image b, merged P3..P7 (train.py:68-73); also the anchor
    index of every candidate within the image."""
    decoded, anchors, off = [], [], 0
    p_lv = path.per_level(probs, (path.num_classes,))
    b_lv = path.per_level(boxes, (4,))
    for pn in path.names:
        p = p_lv[pn][b].cpu().numpy()
        d = oracle.boxes_decode(p, b_lv[pn][b].cpu().numpy())
        fg = (p.max(-1) > np.float32(0.5)).reshape(-1)
        idx = np.nonzero(fg)[0] + off
        if topk is not None and topk < d.scores.shape[0]:
            order = np.sort(np.argsort(-d.scores, kind='stable')[:topk])
            d = oracle.BoxesDecoded(d.boxes[order], d.scores[order], d.class_ids[order])
            idx = idx[order]
        decoded.append(d)
        anchors.append(idx)
        off += fg.shape[0]
    return oracle.merge_boxes_decoded(decoded), np.concatenate(anchors)


def test_sigmoid_and_decode_boxes_match_oracle():
    wl, path, logits, regs = _case('shapes', 2)
    helpers.assert_close(_gpu_sigmoid(logits).cpu().numpy(), oracle.sigmoid(logits.cpu().numpy()), rtol=1e-6,
                         what='sigmoid')
    got = path.per_level(_gpu_decode(path, regs), (4,))
    want_in = path.per_level(regs, (4,))
    for pn in path.names:
        sizes = oracle.anchor_size_normalised(path.levels[pn].anchor_sizes, np.array(path.image_size, np.int32))
        want = oracle.regression_postprocess(want_in[pn].cpu().numpy(), sizes)
        helpers.assert_close(got[pn].cpu().numpy(), want, rtol=1e-5, atol=1e-6, what='decode ' + pn)


@pytest.mark.parametrize('name,batch', [('shapes', 3), ('voc', 2), ('coco', 2)])
def test_select_and_nms_exact_on_identical_inputs(name, batch):
    """Probabilities and decoded boxes are computed once (on the GPU) and handed to both sides:
    every index, class, score, box and the output order must then be identical."""
    import torch
    wl, path, logits, regs = _case(name, batch, capacity=8192)
    probs = _gpu_sigmoid(logits)
    boxes = _gpu_decode(path, regs)
    path.select(probs, boxes, scores_are_logits=False, regs_are_deltas=False)
    path.nms()
    torch.cuda.synchronize()
    path.check_status()
    total = 0
    for b in range(batch):
        cand, anchors = _oracle_candidates(path, probs, boxes, b)
        n = int(path.cand_count[b])
        assert n == cand.scores.shape[0]
        np.testing.assert_array_equal(path.cand_anchor[b, :n].cpu().numpy(), anchors)
        np.testing.assert_array_equal(path.cand_classes[b, :n].cpu().numpy(), cand.class_ids)
        np.testing.assert_array_equal(path.cand_scores[b, :n].cpu().numpy(), cand.scores)
        np.testing.assert_array_equal(path.cand_boxes[b, :n].cpu().numpy(), cand.boxes)
        det, det_idx = oracle.nms_classwise(cand, wl.num_classes, return_indices=True)
        m = int(path.det_count[b])
        assert m == det.scores.shape[0]
        np.testing.assert_array_equal(path.det_index[b, :m].cpu().numpy(), det_idx)
        np.testing.assert_array_equal(path.det_classes[b, :m].cpu().numpy(), det.class_ids)
        np.testing.assert_array_equal(path.det_scores[b, :m].cpu().numpy(), det.scores)
        np.testing.assert_array_equal(path.det_boxes[b, :m].cpu().numpy(), det.boxes)
        total += m
    assert total > 0


@pytest.mark.parametrize('name,batch', [('shapes', 2), ('voc', 2), ('coco', 1), ('crowded', 1)])
def test_fused_select_from_logits_and_deltas(name, batch):
    """The fused path (sigmoid and decode inside the kernels) selects exactly what the
    two-step path selects, and its values agree with the oracle's float32 math."""
    import torch
    wl, path, logits, regs = _case(name, batch, capacity=8192)
    probs = _gpu_sigmoid(logits)
    boxes = _gpu_decode(path, regs)
    path.select(logits, regs)  # logits + deltas
    torch.cuda.synchronize()
    for b in range(batch):
        cand, anchors = _oracle_candidates(path, probs, boxes, b)
        n = int(path.cand_count[b])
        assert n == anchors.shape[0]
        np.testing.assert_array_equal(path.cand_anchor[b, :n].cpu().numpy(), anchors)
        np.testing.assert_array_equal(path.cand_classes[b, :n].cpu().numpy(), cand.class_ids)
        np.testing.assert_array_equal(path.cand_scores[b, :n].cpu().numpy(), cand.scores)
        np.testing.assert_array_equal(path.cand_boxes[b, :n].cpu().numpy(), cand.boxes)
        ref_scores = oracle.sigmoid(logits.cpu().numpy()).max(-1)
        lv_rows = path.rows_total // path.batch
        assert lv_rows == path.anchors and ref_scores.shape[0] == path.rows_total


@pytest.mark.parametrize('mode', ['focal', 'bce_dice', 'balanced_bce', 'focal_dice'])
@pytest.mark.parametrize('name,batch', [('voc', 2), ('coco', 2)])
def test_fused_loss_and_select_equals_separate_passes(name, batch, mode):
    """rn_cls_loss_select (+ rn_select_finalize with dedup) must give exactly the loss, gradients,
    candidates and detections of rn_cls_loss_fwd_bwd followed by rn_decode_select."""
    import torch
    if name == 'coco' and mode != 'focal':
        pytest.skip('one large case is enough')
    wl, path, logits, regs = _case(name, batch, capacity=8192, cls_mode=mode)
    assert path.fused_select_supported()
    # make several classes of the same anchor fire, inside one float4 unit and across units,
    # including exact probability ties (first class must win)
    lv = logits.view(-1, wl.num_classes)
    lv[5, 0:3] = torch.tensor([2.0, 3.0, 2.5], device='cuda')
    lv[77, 1] = 4.0
    lv[77, 9] = 4.0
    lv[77, 17] = 3.0
    lv[123, 6] = 30.0
    lv[123, 2] = 40.0   # both saturate to probability 1.0 -> tie in probability space, class 2 wins
    path.loss(logits, regs)
    path.select(logits, regs)
    path.nms()
    want = {k: getattr(path, k).clone() for k in ('loss_out', 'dlogits', 'dreg', 'cand_count', 'cand_anchor',
                                                  'cand_scores', 'cand_classes', 'cand_boxes', 'det_count',
                                                  'det_index', 'det_boxes')}
    for k in ('dlogits', 'cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes', 'det_index', 'det_boxes'):
        getattr(path, k).zero_()
    path.loss_and_select(logits, regs)
    path.nms()
    torch.cuda.synchronize()
    path.check_status()
    assert torch.equal(path.loss_out, want['loss_out']) and torch.equal(path.dlogits, want['dlogits'])
    assert torch.equal(path.dreg, want['dreg'])
    assert torch.equal(path.cand_count, want['cand_count']) and torch.equal(path.det_count, want['det_count'])
    for b in range(batch):
        n, m = int(want['cand_count'][b]), int(want['det_count'][b])
        for k in ('cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes'):
            assert torch.equal(getattr(path, k)[b, :n], want[k][b, :n]), k
        for k in ('det_index', 'det_boxes'):
            assert torch.equal(getattr(path, k)[b, :m], want[k][b, :m]), k
    cls5 = path.cand_classes[0][path.cand_anchor[0] == 5]
    assert cls5.numel() == 1 and int(cls5) == 1     # argmax of [2.0, 3.0, 2.5]
    b77 = 77 % path.anchors
    assert int(path.cand_classes[77 // path.anchors][path.cand_anchor[77 // path.anchors] == b77]) == 1


def test_overlapped_detection_tail_gives_the_same_results():
    """step(overlap_detect=True) over several steps with changing inputs == the serial step."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    wl, path, logits, regs = _case('voc', 2, capacity=8192)
    boxes, classes, counts = synthetic.gt_batch(wl, 2)
    gt = [torch.from_numpy(x).cuda() for x in (boxes, classes, counts)]
    inputs = [(logits + 0.05 * i, regs * (1.0 + 0.01 * i)) for i in range(4)]
    want = []
    for lg, rg in inputs:
        path.step(*gt, lg, rg)
        torch.cuda.synchronize()
        want.append((path.loss_out.clone(), path.det_count.clone(), path.det_index.clone(), path.det_boxes.clone()))
    got = []
    for lg, rg in inputs:      # back to back, no host synchronisation between steps
        path.step(*gt, lg, rg, overlap_detect=True)
        path.join()            # results of THIS step (still asynchronous w.r.t. the host)
        got.append((path.loss_out.clone(), path.det_count.clone(), path.det_index.clone(), path.det_boxes.clone()))
    torch.cuda.synchronize()
    for w, g in zip(want, got):
        assert torch.equal(w[0], g[0]) and torch.equal(w[1], g[1])
        for b in range(2):
            m = int(w[1][b])
            assert torch.equal(w[2][b, :m], g[2][b, :m]) and torch.equal(w[3][b, :m], g[3][b, :m])
    # and fully pipelined (join only at the end): the last step's detections must be right
    for lg, rg in inputs:
        path.step(*gt, lg, rg, overlap_detect=True)
    path.join()
    torch.cuda.synchronize()
    assert torch.equal(path.det_count, want[-1][1]) and torch.equal(path.loss_out, want[-1][0])


def _tail_outputs(path, batch):
    out = {'cand_count': path.cand_count.clone(), 'det_count': path.det_count.clone()}
    for k in ('cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes'):
        out[k] = [getattr(path, k)[b, :int(path.cand_count[b])].clone() for b in range(batch)]
    for k in ('det_index', 'det_boxes', 'det_scores', 'det_classes'):
        out[k] = [getattr(path, k)[b, :int(path.det_count[b])].clone() for b in range(batch)]
    return out


@pytest.mark.parametrize('name,batch,sigma,kw', [
    ('voc', 3, 1.0, {}), ('coco', 2, 1.0, {}), ('crowded', 1, 1.0, {}),
    ('coco', 2, None, {}),                                  # the longest lists that still fit the capacity
    ('voc', 2, 1.5, {'max_output_per_class': 3}),           # per-class output limit
    ('voc', 2, 1.5, {'out_capacity': 40}),                  # detection overflow
    ('voc', 2, 1.5, {'topk_per_level': 25}),                # per-level top-k (extension): levels above k are ranked
    ('coco', 1, 1.3, {'topk_per_level': 1000}),             # ... BASELINE config 4's k, P3 above it
])
def test_one_launch_tail_equals_separate_kernels(name, batch, sigma, kw):
    """rn_detect_tail (finalize + class-wise NMS in one launch) against rn_select_finalize + rn_nms_classwise on the
    candidates the fused loss pass left in the workspace: every output bit-identical."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200 import _lib
    kw = dict({'capacity': 8192}, **kw)
    wl, path, logits0, regs = _case(name, batch, one_launch_tail=False, **kw)
    assert path.fused_select_supported() and not path.tail_fused_supported()
    for sg in ([sigma] if sigma else [1.45, 1.4, 1.35, 1.3, 1.25]):
        logits = logits0
        if sg != 1.0:
            logits, regs = synthetic.head_outputs(path.label, path.deltas, wl.num_classes, seed=5, device='cuda',
                                                  logit_sigma=sg)
        lv = logits.view(-1, wl.num_classes)   # several classes of one anchor fire (duplicates to drop), with ties
        lv[5, 0:3] = torch.tensor([2.0, 3.0, 2.5], device='cuda')
        if wl.num_classes >= 18:
            lv[77, 1] = 4.0
            lv[77, 9] = 4.0
            lv[77, 17] = 3.0
        path.status.zero_()
        path.loss_and_select(logits, regs, finalize=False)
        path.select_finalize(regs)
        path.nms()
        torch.cuda.synchronize()
        status_want = int(path.status.item())
        if not status_want & _lib.RN_FLAG_CAND_OVERFLOW:   # (an overflowing list keeps an arbitrary subset)
            break
    assert not status_want & _lib.RN_FLAG_CAND_OVERFLOW
    want = _tail_outputs(path, batch)
    if sigma is None:
        assert int(want['cand_count'].max()) > 2048, want['cand_count']
    for k in ('cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes', 'det_index', 'det_boxes', 'det_scores',
              'det_classes', 'cand_count', 'det_count', 'status'):
        getattr(path, k).zero_()
    path.one_launch_tail = True
    assert path.tail_fused_supported()
    path.loss_and_select(logits, regs, finalize=False)
    path.detect_tail(regs)
    torch.cuda.synchronize()
    got = _tail_outputs(path, batch)
    assert int(path.status.item()) == status_want
    assert torch.equal(got['cand_count'], want['cand_count']) and torch.equal(got['det_count'], want['det_count'])
    assert int(want['det_count'].sum()) > 0
    for k in ('cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes', 'det_index', 'det_boxes', 'det_scores',
              'det_classes'):
        for b in range(batch):
            assert torch.equal(got[k][b], want[k][b]), (k, b)


def test_one_launch_tail_after_standalone_select_and_empty_lists():
    """dedup = 0 (candidates of the standalone select pass) and images without any candidate."""
    import torch
    from retinanet_tensorflow_b200 import _lib
    import ctypes as C
    wl, path, logits, regs = _case('voc', 2, capacity=4096)
    logits[: logits.shape[0] // 4] = -20.0      # (merged level-major layout: the first rows are image 0 of P3)
    path.select(logits, regs)
    path.nms()
    torch.cuda.synchronize()
    want = _tail_outputs(path, 2)
    # the same select pass again (leaves the workspace as rn_decode_select's first half does), then the fused tail
    path.select(logits, regs)
    m_reg, _ = path._maps(regs, 4)
    for k in ('det_index', 'det_boxes', 'det_count'):
        getattr(path, k).zero_()
    # rn_decode_select finalised cand_count in place; the workspace still holds the unordered entries of that count
    # only when nothing was dropped, which holds without dedup / top-k
    _lib.call('rn_detect_tail', C.byref(path.pyr), path.batch, path.num_classes, C.byref(m_reg), 1, path.capacity, 0, 0,
              _lib.ptr(path.cand_boxes), _lib.ptr(path.cand_scores), _lib.ptr(path.cand_classes),
              _lib.ptr(path.cand_anchor), _lib.ptr(path.cand_count), path.nms_iou, path.max_output_per_class,
              path.out_capacity, _lib.ptr(path.det_boxes), _lib.ptr(path.det_scores), _lib.ptr(path.det_classes),
              _lib.ptr(path.det_index), _lib.ptr(path.det_count), _lib.ptr(path.status), _lib.ptr(path._sel_ws),
              path._sel_ws_bytes, _lib.stream_ptr())
    torch.cuda.synchronize()
    got = _tail_outputs(path, 2)
    for k in want:
        if k.endswith('count'):
            assert torch.equal(got[k], want[k]), k
        else:
            for b in range(2):
                assert torch.equal(got[k][b], want[k][b]), (k, b)
    # no candidates at all
    path.select(torch.full_like(logits, -20.0), regs)
    _lib.call('rn_detect_tail', C.byref(path.pyr), path.batch, path.num_classes, C.byref(m_reg), 1, path.capacity, 0, 0,
              _lib.ptr(path.cand_boxes), _lib.ptr(path.cand_scores), _lib.ptr(path.cand_classes),
              _lib.ptr(path.cand_anchor), _lib.ptr(path.cand_count), path.nms_iou, path.max_output_per_class,
              path.out_capacity, _lib.ptr(path.det_boxes), _lib.ptr(path.det_scores), _lib.ptr(path.det_classes),
              _lib.ptr(path.det_index), _lib.ptr(path.det_count), _lib.ptr(path.status), _lib.ptr(path._sel_ws),
              path._sel_ws_bytes, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert int(path.cand_count.sum()) == 0 and int(path.det_count.sum()) == 0


@pytest.mark.parametrize('topk', [0, 30, 1000])
def test_detect_one_launch_tail_equals_separate_kernels(topk):
    """DenseHeadPath.detect (inference post-processing): select pass + rn_detect_tail against select + finalize + the
    three NMS launches, with and without the per-level top-k."""
    import torch
    wl, path, logits, regs = _case('voc', 3, capacity=8192, topk_per_level=topk, one_launch_tail=False)
    logits = logits + 1.5     # more candidates than k on the fine levels
    path.detect(logits, regs)
    torch.cuda.synchronize()
    want = _tail_outputs(path, 3)
    for k in ('cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes', 'det_index', 'det_boxes', 'det_scores',
              'det_classes', 'cand_count', 'det_count'):
        getattr(path, k).zero_()
    path.one_launch_tail = True
    assert path.tail_fused_supported()
    path.detect(logits, regs)
    torch.cuda.synchronize()
    path.check_status()
    got = _tail_outputs(path, 3)
    assert torch.equal(got['cand_count'], want['cand_count']) and torch.equal(got['det_count'], want['det_count'])
    assert int(want['det_count'].sum()) > 0
    if topk == 30:
        assert int(want['cand_count'].max()) <= 5 * 30
    for k in ('cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes', 'det_index', 'det_boxes', 'det_scores',
              'det_classes'):
        for b in range(3):
            assert torch.equal(got[k][b], want[k][b]), (k, b)


def test_topk_per_level_prefilter():
    """Extension (not in the reference): k best per (image, level), identity for large k."""
    import torch
    wl, path, logits, regs = _case('voc', 2, capacity=8192, topk_per_level=25)
    probs = _gpu_sigmoid(logits)
    boxes = _gpu_decode(path, regs)
    path.select(probs, boxes, scores_are_logits=False, regs_are_deltas=False)
    torch.cuda.synchronize()
    clipped = False
    for b in range(2):
        full, _ = _oracle_candidates(path, probs, boxes, b)
        cand, anchors = _oracle_candidates(path, probs, boxes, b, topk=25)
        clipped |= cand.scores.shape[0] < full.scores.shape[0]
        n = int(path.cand_count[b])
        assert n == anchors.shape[0]
        np.testing.assert_array_equal(path.cand_anchor[b, :n].cpu().numpy(), anchors)
        np.testing.assert_array_equal(path.cand_scores[b, :n].cpu().numpy(), cand.scores)
    assert clipped


def _nms_raw(boxes, scores, classes, num_classes, max_out=1000, iou_thr=0.5, capacity=None, out_capacity=None):
    import torch
    from retinanet_tensorflow_b200 import _lib
    _lib.require_cuda()
    n = boxes.shape[0]
    cap = capacity or max(n, 1)
    ocap = out_capacity or cap
    cb = torch.zeros((1, cap, 4), device='cuda')
    cs = torch.zeros((1, cap), device='cuda')
    cc = torch.zeros((1, cap), dtype=torch.int32, device='cuda')
    cb[0, :n] = torch.from_numpy(boxes).cuda()
    cs[0, :n] = torch.from_numpy(scores).cuda()
    cc[0, :n] = torch.from_numpy(classes.astype(np.int32)).cuda()
    cnt = torch.tensor([n], dtype=torch.int32, device='cuda')
    db = torch.zeros((1, ocap, 4), device='cuda')
    ds = torch.zeros((1, ocap), device='cuda')
    dc = torch.zeros((1, ocap), dtype=torch.int32, device='cuda')
    di = torch.zeros((1, ocap), dtype=torch.int32, device='cuda')
    dn = torch.zeros((1,), dtype=torch.int32, device='cuda')
    status = torch.zeros((1,), dtype=torch.int32, device='cuda')
    wsb = _lib.lib.rn_nms_workspace_bytes(1, num_classes, cap)
    ws = torch.empty((wsb,), dtype=torch.uint8, device='cuda')
    _lib.call('rn_nms_classwise', 1, num_classes, cap, _lib.ptr(cb), _lib.ptr(cs), _lib.ptr(cc), _lib.ptr(cnt),
              iou_thr, max_out, ocap, _lib.ptr(db), _lib.ptr(ds), _lib.ptr(dc), _lib.ptr(di), _lib.ptr(dn),
              _lib.ptr(status), _lib.ptr(ws), wsb, _lib.stream_ptr())
    m = int(dn[0])
    return di[0, :m].cpu().numpy(), int(status[0])


def _random_boxes(rng, n, spread=0.25):
    ctr = rng.uniform(0.2, 0.8, (n, 2))
    sz = rng.uniform(0.05, spread, (n, 2))
    return np.concatenate([ctr - sz / 2, ctr + sz / 2], -1).astype(np.float32)


@pytest.mark.parametrize('n,num_classes', [(1, 1), (31, 1), (400, 1), (3000, 3), (20000, 5)])
def test_nms_random_boxes_vs_oracle(n, num_classes):
    rng = np.random.default_rng(n)
    boxes = _random_boxes(rng, n)
    scores = rng.permutation(n).astype(np.float32) / n + 0.001
    classes = rng.integers(0, num_classes, n)
    got, status = _nms_raw(boxes, scores, classes, num_classes)
    cand = oracle.BoxesDecoded(boxes, scores, classes.astype(np.int64))
    _, want = oracle.nms_classwise(cand, num_classes, return_indices=True)
    np.testing.assert_array_equal(got, want)
    assert status == 0


def test_nms_ties_max_output_and_overflow():
    rng = np.random.default_rng(5)
    boxes = _random_boxes(rng, 600, spread=0.08)
    scores = np.repeat(rng.uniform(0.5, 1.0, 60).astype(np.float32), 10)  # 10-way score ties
    rng.shuffle(scores)
    classes = np.zeros(600, np.int64)
    cand = oracle.BoxesDecoded(boxes, scores, classes)
    got, _ = _nms_raw(boxes, scores, classes, 1)
    np.testing.assert_array_equal(got, oracle.nms_classwise(cand, 1, return_indices=True)[1])  # ties -> lower index
    got, _ = _nms_raw(boxes, scores, classes, 1, max_out=7)
    np.testing.assert_array_equal(got, oracle.nms_classwise(cand, 1, max_output_size=7, return_indices=True)[1])
    got, status = _nms_raw(boxes, scores, classes, 1, out_capacity=5)
    assert got.shape[0] == 5 and status & 8
    got, _ = _nms_raw(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), np.zeros((0,), np.int64), 4)
    assert got.shape[0] == 0


def test_nms_degenerate_and_flipped_boxes():
    """tf.image.non_max_suppression normalises corner order and gives zero-area boxes IoU 0."""
    boxes = np.array([[0.1, 0.1, 0.5, 0.5], [0.5, 0.5, 0.1, 0.1], [0.3, 0.3, 0.3, 0.6], [0.3, 0.3, 0.3, 0.6],
                      [0.12, 0.1, 0.5, 0.52]], np.float32)
    scores = np.array([0.9, 0.8, 0.7, 0.6, 0.5], np.float32)
    classes = np.zeros(5, np.int64)
    got, _ = _nms_raw(boxes, scores, classes, 1)
    want = oracle.nms_classwise(oracle.BoxesDecoded(boxes, scores, classes), 1, return_indices=True)[1]
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got, [0, 2, 3])


def test_cuda_graph_replay_matches_eager_steps():
    """DenseHeadPath.capture: the replayed graph follows in-place input updates and reproduces step()."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    wl, path, logits, regs = _case('voc', 2, capacity=8192)
    boxes, classes, counts = synthetic.gt_batch(wl, 2)
    gt = [torch.from_numpy(x).cuda() for x in (boxes, classes, counts)]
    lg, rg = logits.clone(), regs.clone()
    graph = path.capture(*gt, lg, rg)
    for i in range(3):
        cur_l, cur_r = logits + 0.05 * i, regs * (1.0 + 0.01 * i)
        path.step(*gt, cur_l, cur_r)
        torch.cuda.synchronize()
        want = [x.clone() for x in (path.loss_out, path.dlogits, path.dreg, path.det_count, path.det_index,
                                    path.det_boxes)]
        for x in (path.dlogits, path.dreg, path.det_index, path.det_boxes, path.loss_out):
            x.zero_()
        lg.copy_(cur_l)
        rg.copy_(cur_r)
        graph.replay()
        torch.cuda.synchronize()
        path.check_status()
        assert torch.equal(path.loss_out, want[0]) and torch.equal(path.dlogits, want[1])
        assert torch.equal(path.dreg, want[2]) and torch.equal(path.det_count, want[3])
        for b in range(2):
            m = int(want[3][b])
            assert m > 0
            assert torch.equal(path.det_index[b, :m], want[4][b, :m]) and torch.equal(path.det_boxes[b, :m], want[5][b, :m])
