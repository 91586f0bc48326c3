"""K2/K3 parity (GPU): fused classification + regression loss forward/backward through the C ABI
vs the oracle's utils.process_labels_and_logits + losses.loss (and its analytic gradients) on
the same seeded inputs.  Tolerance: 1e-5 relative (north_star), plus an absolute floor of
1e-6 x max|reference| for the near-zero elements whose value is dominated by float32
cancellation in the reference's own `1 - (1 - p)` ("tolerances")."""
import numpy as np
import pytest

import helpers
import rn_oracle as oracle

pytestmark = pytest.mark.gpu

MODES = ['focal', 'bce_dice', 'bce', 'balanced_bce', 'balanced_bce_dice', 'focal_dice']


def _setup(name, batch, mode, generic=False):
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS[name]
    ref = helpers.oracle_labels_batch(wl, batch)
    cls = helpers.merge_levels(ref['cls'], 1)
    mask = helpers.merge_levels(ref['mask'])
    reg = helpers.merge_levels(ref['reg'], 1)
    lab = helpers.compact_labels(cls, mask)
    logits, regs = synthetic.head_outputs(lab, reg, wl.num_classes, seed=7 + wl.config, device='cpu')
    boxes, classes, counts = synthetic.gt_batch(wl, batch)
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, batch, wl.max_gt, cls_mode=mode,
                         process_group=False, generic_kernels=generic)
    path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda())
    return wl, path, cls, mask, reg, logits, regs


@pytest.fixture(params=[1, 0], ids=['streaming', 'generic'])
def focal_kernel(request):
    """Run focal cases through both the streaming and the generic kernel (a per-path choice, no global switch)."""
    return request.param


def _outside_pure_rtol(got, want, rtol=1e-5):
    """Fraction of elements that a PURE relative tolerance (no absolute floor) rejects, among the non-zero ones."""
    nz = want != 0
    bad = np.abs(got.astype(np.float64) - want.astype(np.float64)) > rtol * np.abs(want.astype(np.float64))
    return float((bad & nz).sum()) / max(int(nz.sum()), 1)


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name,batch', [('shapes', 2), ('voc', 2), ('coco', 1), ('crowded', 1)])
def test_fused_loss_matches_oracle(name, batch, mode, focal_kernel):
    import torch
    if focal_kernel == 0 and (mode != 'focal' or name == 'shapes'):
        pytest.skip('the generic kernel is what runs anyway')
    if name in ('coco', 'crowded') and mode not in ('focal', 'bce_dice'):
        pytest.skip('the headline shapes run the two modes the bench reports')
    wl, path, cls, mask, reg, logits, regs = _setup(name, batch, mode, generic=(focal_kernel == 0))
    out = path.loss(logits.cuda(), regs.cuda()).cpu().numpy()
    torch.cuda.synchronize()
    x, r = logits.numpy(), regs.numpy()
    labels = oracle.Detection(oracle.Classification(None, cls[mask]), reg[mask], None)
    preds = oracle.Detection(oracle.Classification(x[mask], oracle.sigmoid(x[mask])), r[mask], None)
    cls_loss, reg_loss = oracle.loss(labels, preds, mode=mode)
    fg = oracle.classmap_decode(cls[mask]).fg_mask
    np.testing.assert_allclose(out[0], cls_loss, rtol=1e-5)
    np.testing.assert_allclose(out[1], reg_loss, rtol=1e-5)
    assert out[2] == fg.sum() and out[3] == mask.sum()
    # gradients w.r.t. the full maps: zero on ignored anchors (boolean_mask's gradient)
    g_ref = np.zeros_like(x)
    g_ref[mask] = oracle.classification_loss_grad(cls[mask], x[mask], fg, mode=mode)
    g = path.dlogits.cpu().numpy()
    helpers.assert_close(g, g_ref, rtol=1e-5, atol=1e-6 * np.abs(g_ref).max(), what='dlogits ' + mode)
    # how many elements need the absolute floor at all (the reference's `1 - (1 - p)` quantisation, DESIGN.md
    # section 2): a small fraction, and the same for the streaming kernel (ex2/rcp.approx) and the libm one
    frac = _outside_pure_rtol(g, g_ref)
    print('PARITY %s batch %d %s kernel=%s: %.4f %% of the non-zero gradient elements outside pure rtol 1e-5'
          % (name, batch, mode, 'generic' if focal_kernel == 0 else 'streaming', 100 * frac))
    assert frac < 0.05
    gr_ref = np.zeros_like(r)
    gr_ref[mask] = oracle.regression_loss_grad(reg[mask], r[mask], fg)
    helpers.assert_close(path.dreg.cpu().numpy(), gr_ref, rtol=1e-5, atol=1e-6 * np.abs(gr_ref).max(), what='dreg')


def test_fast_focal_extreme_logits_stay_finite_and_close():
    """Saturated logits (|x| up to 60) through the fast kernel: finite, and equal to the generic
    kernel within the usual tolerance."""
    import torch
    from retinanet_tensorflow_b200 import _lib
    wl, path, cls, mask, reg, logits, regs = _setup('voc', 1, 'focal')
    g = torch.Generator().manual_seed(0)
    lg = (logits + torch.randn(logits.shape, generator=g) * 12.0).cuda()
    lg[::97, 3] = 60.0
    lg[::89, 5] = -60.0
    rg = regs.cuda()
    a = path.loss(lg, rg).clone()
    ga = path.dlogits.clone()
    path.mode |= _lib.RN_CLS_GENERIC_FLAG
    try:
        b = path.loss(lg, rg).clone()
        gb = path.dlogits.clone()
    finally:
        path.mode &= ~_lib.RN_CLS_GENERIC_FLAG
    assert torch.isfinite(a).all() and torch.isfinite(ga).all()
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5)
    ga, gb, x = ga.cpu().numpy(), gb.cpu().numpy(), lg.cpu().numpy()
    assert np.linalg.norm((ga - gb).astype(np.float64)) <= 1e-5 * np.linalg.norm(gb.astype(np.float64))
    # Elementwise outside saturation.  For x > 9 the reference's own `1 - p` is quantised in steps
    # of 6e-8 that are > 0.05 % of its value (a 1-ulp change of p moves the result by that much on
    # ANY implementation), so only a loose bound is meaningful there.
    sat = x > 9.0
    helpers.assert_close(np.where(sat, 0, ga), np.where(sat, 0, gb), rtol=2e-5, atol=1e-6 * float(np.abs(gb).max()),
                         what='fast vs generic')
    helpers.assert_close(ga[sat], gb[sat], rtol=0.2, atol=1e-6 * float(np.abs(gb).max()), what='saturated')


def test_loss_is_deterministic_and_forward_only_matches():
    import torch
    wl, path, cls, mask, reg, logits, regs = _setup('shapes', 2, 'focal')
    lg, rg = logits.cuda(), regs.cuda()
    a = path.loss(lg, rg).clone()
    g1 = path.dlogits.clone()
    b = path.loss(lg, rg).clone()
    assert torch.equal(a, b) and torch.equal(g1, path.dlogits)
    c = path.loss(lg, rg, want_grad=False).clone()
    assert torch.equal(a, c)


def test_per_level_inputs_equal_merged():
    """Per-level [B,H,W,A,C] tensors (the reference's layout, retinanet.py:68-69) and the merged
    buffer are the same thing to the kernels."""
    import torch
    wl, path, cls, mask, reg, logits, regs = _setup('shapes', 2, 'bce_dice')
    lg, rg = logits.cuda(), regs.cuda()
    a = path.loss(lg, rg).clone()
    lv = {k: v.clone() for k, v in path.per_level(lg, (wl.num_classes,)).items()}
    rv = {k: v.clone() for k, v in path.per_level(rg, (4,)).items()}
    b = path.loss(lv, rv).clone()
    assert torch.equal(a, b)


def test_reference_huber_vector():
    """losses_test.py:17-27: labels [[3],[4],[6]], logits [[1],[2],[3]], fg [T,F,T] -> 2.0.
    The C ABI works on 4-wide rows; replicate the single column so that the mean is unchanged."""
    import ctypes as C
    import torch
    from retinanet_tensorflow_b200 import _lib
    _lib.require_cuda()
    pred = torch.tensor([[1.] * 4, [2.] * 4, [3.] * 4], device='cuda')
    target = torch.tensor([[3.] * 4, [4.] * 4, [6.] * 4], device='cuda')
    fg = torch.tensor([1, 0, 1], dtype=torch.uint8, device='cuda')
    stats = torch.zeros(8 + 3, dtype=torch.float64, device='cuda')
    stats[1] = 2.0
    out = torch.zeros(4, device='cuda')
    dpred = torch.empty_like(pred)
    ws = torch.zeros(_lib.lib.rn_loss_workspace_bytes(1), dtype=torch.uint8, device='cuda')
    rows = _lib.rows_of([3])
    _lib.call('rn_huber_fwd_bwd', C.byref(rows), C.byref(_lib.maps_of([pred])), C.byref(_lib.maps_of([target])),
              C.byref(_lib.Maps()), C.byref(_lib.maps_of([fg])), C.byref(_lib.maps_of([dpred])), _lib.ptr(stats),
              _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.call('rn_loss_finalize', 1, _lib.RN_CLS_BCE, _lib.ptr(stats), _lib.ptr(out), _lib.stream_ptr())
    assert out[1].item() == 2.0
    np.testing.assert_array_equal(dpred.cpu().numpy(), np.array([[-1.] * 4, [0.] * 4, [-1.] * 4]) / 8.0)


@pytest.mark.parametrize('mode', ['focal', 'bce', 'balanced_bce', 'bce_dice', 'focal_dice', 'balanced_bce_dice'])
def test_bf16_logits_in_and_out(mode):
    """bfloat16 logits (bf16x8 accesses) through the fast kernel: float32 math on the upcast values,
    so the loss equals the oracle's on logits.bfloat16().float() and the bf16 gradient equals the
    oracle's gradient rounded to bf16 (up to one bf16 ulp)."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS['voc']          # C = 20 is not a multiple of 8 -> use a 24-class variant
    wl = synthetic.Workload('voc24', wl.config, wl.image_h, wl.image_w, 24, wl.max_gt, 2)
    ref = helpers.oracle_labels_batch(wl, 2)
    cls = helpers.merge_levels(ref['cls'], 1)
    mask = helpers.merge_levels(ref['mask'])
    reg = helpers.merge_levels(ref['reg'], 1)
    lab = helpers.compact_labels(cls, mask)
    logits, regs = synthetic.head_outputs(lab, reg, wl.num_classes, seed=21, device='cpu')
    logits16 = logits.bfloat16()
    boxes, classes, counts = synthetic.gt_batch(wl, 2)
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, 2, wl.max_gt, cls_mode=mode, capacity=8192,
                         process_group=False)
    gt = [torch.from_numpy(x).cuda() for x in (boxes, classes, counts)]
    path.step(*gt, logits16.cuda(), regs.cuda())
    torch.cuda.synchronize()
    assert path.dlogits.dtype == torch.bfloat16
    x = logits16.float().numpy()
    r = regs.numpy()
    labels = oracle.Detection(oracle.Classification(None, cls[mask]), reg[mask], None)
    preds = oracle.Detection(oracle.Classification(x[mask], oracle.sigmoid(x[mask])), r[mask], None)
    cl, rl = oracle.loss(labels, preds, mode=mode)
    out = path.loss_out.cpu().numpy()
    np.testing.assert_allclose(out[:2], [cl, rl], rtol=1e-5)
    fg = oracle.classmap_decode(cls[mask]).fg_mask
    g_ref = np.zeros_like(x)
    g_ref[mask] = oracle.classification_loss_grad(cls[mask], x[mask], fg, mode=mode)
    got = path.dlogits.float().cpu().numpy()
    helpers.assert_close(got, g_ref, rtol=2.0 ** -7, atol=1e-6 * np.abs(g_ref).max(), what='bf16 dlogits ' + mode)
    # detections equal the float32 path's on the same (upcast) logits
    ref_path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, 2, wl.max_gt, cls_mode=mode, capacity=8192,
                             process_group=False)
    ref_path.step(*gt, logits16.float().cuda(), regs.cuda())
    torch.cuda.synchronize()
    assert torch.equal(ref_path.det_count, path.det_count)
    if 'dice' in mode:   # the per-class sums are accumulated in a different order (8 vs 4 classes per thread)
        np.testing.assert_allclose(path.loss_out.cpu().numpy(), ref_path.loss_out.cpu().numpy(), rtol=1e-6)
    else:
        assert torch.equal(ref_path.loss_out, path.loss_out)
    for b in range(2):
        m = int(path.det_count[b])
        assert torch.equal(ref_path.det_index[b, :m], path.det_index[b, :m])
        assert torch.equal(ref_path.det_boxes[b, :m], path.det_boxes[b, :m])


def test_bf16_standalone_select_equals_float32_on_the_same_values():
    """rn_decode_select with RN_SCORES_BF16: float32 math on the upcast values, so candidates and
    detections are bit-identical to the float32 kernel fed logits.bfloat16().float()."""
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    wl = synthetic.WORKLOADS['voc']
    wl = synthetic.Workload('voc24', wl.config, wl.image_h, wl.image_w, 24, wl.max_gt, 2)
    boxes, classes, counts = synthetic.gt_batch(wl, 2)
    gt = [torch.from_numpy(x).cuda() for x in (boxes, classes, counts)]
    paths = [DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, 2, wl.max_gt, capacity=8192, process_group=False)
             for _ in range(2)]
    paths[0].assign(*gt)
    logits, regs = synthetic.head_outputs(paths[0].label, paths[0].deltas, wl.num_classes, seed=5, device='cuda')
    logits16 = logits.bfloat16()
    for scores, is_logits in ((logits16, True), (torch.sigmoid(logits).bfloat16(), False)):
        paths[0].detect(scores, regs, scores_are_logits=is_logits)
        paths[1].detect(scores.float(), regs, scores_are_logits=is_logits)
        torch.cuda.synchronize()
        assert int(paths[0].cand_count.sum()) > 50
        assert torch.equal(paths[0].cand_count, paths[1].cand_count) and torch.equal(paths[0].det_count, paths[1].det_count)
        for b in range(2):
            n, m = int(paths[0].cand_count[b]), int(paths[0].det_count[b])
            for k in ('cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes'):
                assert torch.equal(getattr(paths[0], k)[b, :n], getattr(paths[1], k)[b, :n]), k
            assert torch.equal(paths[0].det_index[b, :m], paths[1].det_index[b, :m])
