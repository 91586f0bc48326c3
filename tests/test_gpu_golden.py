"""GPU path vs the COMMITTED golden fixture (tests/golden/oracle_shapes_case.npz, written by
tests/golden/make_golden.py in the build container): nothing here calls the oracle."""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_committed_case_end_to_end():
    import torch
    from retinanet_tensorflow_b200 import synthetic
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    case = np.load(os.path.join(G, 'oracle_shapes_case.npz'))
    wl = synthetic.WORKLOADS['shapes']
    boxes, classes, counts = synthetic.gt_batch(wl, 2)
    logits = torch.from_numpy(case['logits']).cuda()
    regs = torch.from_numpy(case['regs']).cuda()
    for mode in ('focal', 'bce_dice', 'balanced_bce'):
        path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, 2, wl.max_gt, cls_mode=mode, keep_iou=True,
                             process_group=False)
        path.assign(torch.from_numpy(boxes).cuda(), torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda())
        np.testing.assert_array_equal(path.label.cpu().numpy(), case['label'])
        np.testing.assert_array_equal(path.matched.cpu().numpy(), case['matched'])
        np.testing.assert_array_equal(path.iou_max.cpu().numpy(), case['iou'])
        helpers.assert_close(path.deltas.cpu().numpy(), case['deltas'], rtol=1e-5, atol=1e-7, what='deltas')
        out = path.loss(logits, regs).cpu().numpy()
        np.testing.assert_allclose(out[:2], case['loss_' + mode], rtol=1e-5)
        g = path.dlogits.cpu().numpy()
        want = case['dlogits_' + mode]
        helpers.assert_close(g.reshape(-1)[::7], want, rtol=1e-5, atol=1e-6 * np.abs(want).max(), what='dlogits')
        np.testing.assert_allclose(g.astype(np.float64).sum(), case['dlogits_sum_' + mode], rtol=1e-4, atol=1e-9)
        helpers.assert_close(path.dreg.cpu().numpy(), case['dreg'], rtol=1e-5,
                             atol=1e-6 * np.abs(case['dreg']).max(), what='dreg')
    # detections: the fixture's candidates come from the oracle's sigmoid/exp, the GPU's from its
    # own -> compare values to 1e-5 and require the same candidates / kept set
    path.detect(logits, regs)
    torch.cuda.synchronize()
    for b in range(2):
        n = int(path.cand_count[b])
        assert n == case['cand_scores_%d' % b].shape[0]
        np.testing.assert_array_equal(path.cand_classes[b, :n].cpu().numpy(), case['cand_classes_%d' % b])
        helpers.assert_close(path.cand_scores[b, :n].cpu().numpy(), case['cand_scores_%d' % b], rtol=1e-5,
                             what='scores')
        helpers.assert_close(path.cand_boxes[b, :n].cpu().numpy(), case['cand_boxes_%d' % b], rtol=1e-5, atol=1e-6,
                             what='boxes')
        m = int(path.det_count[b])
        np.testing.assert_array_equal(path.det_index[b, :m].cpu().numpy(), case['det_index_%d' % b])
        helpers.assert_close(path.det_boxes[b, :m].cpu().numpy(), case['det_boxes_%d' % b], rtol=1e-5, atol=1e-6,
                             what='det boxes')
