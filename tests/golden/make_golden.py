"""Regenerates the committed golden fixtures.  Run in the BUILD container (needs
/root/reference); the fixtures travel to the GPU box, the reference does not.

    python tests/golden/make_golden.py

Writes next to this file:
  reference_levels.npz   anchor sizes of the REAL reference levels.py (imported from
                         /root/reference), P3..P7, plus levels_test.py's two vectors.
  reference_shapes.npz   GT produced by the REAL reference data_loaders/shapes.py
                         (random_image, seeded) and its NumPy NMS on seeded random boxes.
  reference_tests.npz    the still-valid vectors transcribed from the reference's own tests
                         (utils_test.py, losses_test.py, dataset_test.py, retinanet_old_test.py).
  oracle_shapes_case.npz a small end-to-end case (shapes workload, 2 images) computed by the
                         oracle: labels, deltas, logits, losses, gradients' checksums, detections.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def reference_levels():
    sys.path.insert(0, REF)
    import levels as ref_levels  # the real reference module (NumPy only)
    lv = ref_levels.build_levels()
    out = {'names': np.array(list(lv.keys())), 'num_anchors': np.int64(lv.num_anchors)}
    for pn in lv:
        out['sizes_' + pn] = lv[pn].anchor_sizes
    out['box_32_1x2_1'] = ref_levels.compute_box_size(32, (1, 2), 1)            # levels_test.py:5-9
    out['level_32_1x4'] = ref_levels.Level(32, [(1, 4)], [2 ** 0, 2 ** 1]).anchor_sizes  # levels_test.py:12-14
    np.savez(os.path.join(HERE, 'reference_levels.npz'), **out)


def reference_shapes():
    sys.path.insert(0, REF)
    from data_loaders import shapes as ref_shapes  # real reference module (NumPy + cv2)
    out = {}
    for seed in range(8):
        random.seed(seed)
        _, shp = ref_shapes.random_image((500, 500))
        names = ['square', 'triangle', 'circle']  # shapes.py:16
        boxes = np.array([[y - s, x - s, y + s, x + s] for _, _, (x, y, s) in shp], np.float64)
        out['boxes_%d' % seed] = boxes
        out['classes_%d' % seed] = np.array([names.index(s[0]) for s in shp], np.int32)
    rng = np.random.default_rng(123)
    for case, n in enumerate((50, 400)):
        ctr = rng.uniform(0.1, 0.9, (n, 2))
        sz = rng.uniform(0.02, 0.3, (n, 2))
        boxes = np.concatenate([ctr - sz / 2, ctr + sz / 2], -1).astype(np.float32)
        scores = rng.permutation(n).astype(np.float32) / n
        out['nms_boxes_%d' % case] = boxes
        out['nms_scores_%d' % case] = scores
        for thr in (0.3, 0.5):
            out['nms_keep_%d_%02d' % (case, int(thr * 10))] = ref_shapes.non_max_suppression(boxes, scores, thr)
    np.savez(os.path.join(HERE, 'reference_shapes.npz'), **out)


def reference_tests():
    c = [0.5, 1.0, 0.25, 0.75]
    out = {
        # utils_test.py:7-42
        'g2_in': np.tile(np.array(c, np.float32), (1, 3, 4, 1, 1)),
        'g2_out': np.array([[[[(2 * i + 1) / 6 + 0.5, (2 * j + 1) / 8 + 1.0, 0.25, 0.75]] for j in range(4)]
                            for i in range(3)], np.float64)[None],
        # utils_test.py:44-74
        'g3_grid': np.array([3, 4]), 'g3_sizes': np.array([[0.2, 0.4]], np.float32),
        'g3_out': np.array([[[[(2 * i + 1) / 6 - 0.1, (2 * j + 1) / 8 - 0.2, (2 * i + 1) / 6 + 0.1,
                               (2 * j + 1) / 8 + 0.2]] for j in range(4)] for i in range(3)], np.float64)[None],
        # utils_test.py:76-97 (np.array_equal in the reference)
        'g4_in': np.tile(np.array([0.5, 1.0, 0.2, 0.4], np.float32), (1, 3, 4, 1, 1)),
        'g4_out': np.tile(np.array([0.4, 0.8, 0.6, 1.2], np.float32), (1, 3, 4, 1, 1)),
        # utils_test.py:99-118
        'g5_a': np.array([[0.1, 0.1, 0.2, 0.2], [100, 100, 200, 200], [0.1, 0.1, 0.2, 0.2], [1., 1., 1., 1.]],
                         np.float32),
        'g5_b': np.array([[0.1, 0.1, 0.3, 0.3], [100, 100, 300, 300], [100, 100, 300, 300], [0., 0., 0., 0.]],
                         np.float32),
        'g5_out': np.array([0.25, 0.25, 0, 0], np.float32),
        # utils_test.py:120-138 (fg pattern only; the int-vector return type is stale)
        'g6_in': np.array([[0.1, 0.9, 0.3, 0.8], [0, 1, 0, 0], [0.1, 0.2, 0.4, 0.3], [0, 0, 0, 0]], np.float32),
        'g6_fg': np.array([True, True, False, False]),
        # losses_test.py:7-15 (merge order semantics)
        'g7_a': np.array([1, 2, 3]), 'g7_b': np.array([4, 5, 6]),
        'g7_mask_a': np.array([False, True, True]), 'g7_mask_b': np.array([True, True, False]),
        'g7_out': np.array([2, 3, 4, 5]),
        # losses_test.py:17-27
        'g8_logits': np.array([[1], [2], [3]], np.float32), 'g8_labels': np.array([[3], [4], [6]], np.float32),
        'g8_fg': np.array([True, False, True]), 'g8_out': np.float32(2.0),
        # dataset_test.py:8-28 (assignment pattern)
        'g9_boxes': (np.array([[0, 0, 16, 16], [8, 8, 24, 24], [16, 16, 32, 32], [-4, -4, 20, 20]], np.float64)
                     / 32).astype(np.float32),
        'g9_class_ids': np.array([100, 200, 300, 400]),
        'g9_class_map': np.array([[[100, 400], [0, 0]], [[0, 0], [300, 0]]]),
        # retinanet_old_test.py:15-37 (np.array_equal in the reference)
        'g10_in': np.array([[0.5, 1.0, 0.5, 1.0], [0.5, 0.5, 0.5, 0.5]], np.float32).reshape(1, 1, 1, 2, 4),
        'g10_sizes': np.array([[0.2, 0.4], [0.4, 0.2]], np.float32),
        'g10_out': np.array([[0.1, 0.4, 0.1, 0.4], [0.2, 0.1, 0.2, 0.1]], np.float32).reshape(1, 1, 1, 2, 4),
    }
    np.savez(os.path.join(HERE, 'reference_tests.npz'), **out)


def oracle_case():
    import helpers
    import rn_oracle as oracle
    from retinanet_tensorflow_b200 import synthetic
    wl = synthetic.WORKLOADS['shapes']
    batch = 2
    ref = helpers.oracle_labels_batch(wl, batch)
    cls = helpers.merge_levels(ref['cls'], 1)
    mask = helpers.merge_levels(ref['mask'])
    reg = helpers.merge_levels(ref['reg'], 1)
    lab = helpers.compact_labels(cls, mask)
    logits, regs = synthetic.head_outputs(lab, reg, wl.num_classes, seed=2024, device='cpu')
    x, r = logits.numpy(), regs.numpy()
    out = {'label': lab, 'matched': helpers.merge_levels(ref['idx']).astype(np.int32),
           'iou': helpers.merge_levels(ref['iou']), 'deltas': reg, 'logits': x, 'regs': r}
    labels = oracle.Detection(oracle.Classification(None, cls[mask]), reg[mask], None)
    preds = oracle.Detection(oracle.Classification(x[mask], oracle.sigmoid(x[mask])), r[mask], None)
    fg = oracle.classmap_decode(cls[mask]).fg_mask
    for mode in ('focal', 'bce_dice', 'balanced_bce'):
        cl, rl = oracle.loss(labels, preds, mode=mode)
        out['loss_' + mode] = np.array([cl, rl], np.float32)
        g = np.zeros_like(x)
        g[mask] = oracle.classification_loss_grad(cls[mask], x[mask], fg, mode=mode)
        out['dlogits_' + mode] = g.reshape(-1)[::7].copy()  # every 7th element keeps the fixture small
        out['dlogits_sum_' + mode] = np.float64(g.astype(np.float64).sum())
    gr = np.zeros_like(r)
    gr[mask] = oracle.regression_loss_grad(reg[mask], r[mask], fg)
    out['dreg'] = gr
    # detections per image from oracle sigmoid / decode
    probs = helpers.split_levels(oracle.sigmoid(x), ref['cls'])
    regl = helpers.split_levels(r, ref['reg'])
    levels = oracle.build_levels()
    for b in range(batch):
        dec = []
        for pn in levels:
            sizes = oracle.anchor_size_normalised(levels[pn].anchor_sizes, helpers.image_size_of(wl))
            boxes = oracle.regression_postprocess(regl[pn][b:b + 1], sizes)[0]
            dec.append(oracle.boxes_decode(probs[pn][b], boxes))
        cand = oracle.merge_boxes_decoded(dec)
        det, idx = oracle.nms_classwise(cand, wl.num_classes, return_indices=True)
        out['cand_scores_%d' % b] = cand.scores
        out['cand_classes_%d' % b] = cand.class_ids
        out['cand_boxes_%d' % b] = cand.boxes
        out['det_index_%d' % b] = idx
        out['det_boxes_%d' % b] = det.boxes
    np.savez_compressed(os.path.join(HERE, 'oracle_shapes_case.npz'), **out)


def reference_shapes_images():
    """reference_shapes_images.npz: six 200x200 images drawn by the REAL data_loaders/shapes.py (random_image +
    draw_shape with OpenCV), with their specifications -- what rn_shapes_render has to reproduce."""
    sys.path.insert(0, REF)
    from data_loaders import shapes as ref_shapes
    out = {}
    for seed in range(6):
        random.seed(100 + seed)
        size = (200, 200)
        bg, shp = ref_shapes.random_image(size)
        image = np.ones([*size, 3], np.uint8) * np.array(bg).reshape(1, 1, 3).astype(np.uint8)
        for shape, color, dims in shp:
            image = ref_shapes.draw_shape(image, shape, dims, color)
        out['image_%d' % seed] = image
        out['bg_%d' % seed] = np.array(bg)
        out['n_%d' % seed] = np.int32(len(shp))
        out['kinds_%d' % seed] = np.array([['square', 'triangle', 'circle'].index(s[0]) for s in shp], np.int32)
        out['colors_%d' % seed] = np.array([s[1] for s in shp], np.int32).reshape(-1, 3)
        out['dims_%d' % seed] = np.array([s[2] for s in shp], np.int32).reshape(-1, 3)
    np.savez_compressed(os.path.join(HERE, 'reference_shapes_images.npz'), **out)


if __name__ == '__main__':
    reference_levels()
    reference_shapes()
    reference_shapes_images()
    reference_tests()
    oracle_case()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
