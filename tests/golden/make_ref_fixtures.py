"""Runs the REFERENCE'S OWN SOURCE (unmodified /root/reference/{dataset,utils,losses,augmentation}.py, executed on
the NumPy TensorFlow stand-in of oracle/tf_stub, see oracle/ref_exec.py) on seeded inputs and commits what it
returns.  Build container only (needs /root/reference); the fixtures travel to the GPU box.

    python tests/golden/make_ref_fixtures.py

Writes tests/golden/ref_exec_<case>.npz for case in CASES.  Per case (B images, levels P3..P7):
  inputs    image_size, boxes_<i>, classes_<i>, logits_<pn> [B,H,W,A,C], regs_<pn> [B,H,W,A,4]
  labels    dataset.build_labels per image (dataset.py:126-142): cls_<pn> (uint8 one-hot), reg_<pn>, mask_<pn>
  masking   utils.process_labels_and_logits (utils.py:240-284): lab_prob, lab_reg, lab_regpp, log_unscaled,
            log_prob, log_reg, log_regpp  -- the `detection_trainable` members, merged P3..P7
  losses    losses.loss (losses.py:155-175) -> loss_class, loss_regr and, through tf.gradients over the whole
            chain (masking included), grad_logits_<pn>, grad_regs_<pn>;
            every other loss function of losses.py on the same masked tensors, reduced the way the
            commented-out alternatives in classification_loss reduce them (losses.py:119-138), value and
            gradient w.r.t. the masked logits: focal, bce, balanced_bce, dice, jaccard, fixed_iou, focal_softmax
  decode    per image, train.py:68-75's chain utils.boxes_decode per level -> merge_boxes_decoded ->
            nms_classwise on (log_prob level maps, regression_postprocessed level maps): det_boxes_<i>,
            det_scores_<i>, det_classes_<i>, and the pre-NMS candidates cand_*_<i>
  flip      augmentation.flip of image 0's sample: flip_cls_<pn>, flip_reg_<pn>, flip_mask_<pn>
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import ref_exec  # noqa: E402

# name: (image (h, w), classes, images, GT per image, seed)
CASES = {
    'shapes': ((300, 300), 3, 1, None, 11),     # normalised GT of the real Shapes generator (reference_shapes.npz)
    'voc_small': ((160, 224), 8, 2, 8, 12),     # C % 8 == 0: the streaming CUDA kernels take it (float32 and bf16)
    'wide': ((96, 352), 4, 2, 5, 13),           # non-square, ceil() grid sizes, a single-row P7
}


def ground_truth(case, rng):
    (h, w), C, B, n_gt, _ = CASES[case]
    out = []
    if n_gt is None:
        shp = np.load(os.path.join(HERE, 'reference_shapes.npz'))
        for i in range(B):
            boxes = (shp['boxes_%d' % i] / 500.0).astype(np.float32)   # generated for a 500 x 500 canvas
            out.append((boxes, shp['classes_%d' % i].astype(np.int32)))
        return out
    for _ in range(B):
        n = int(rng.integers(max(1, n_gt // 2), n_gt + 1))
        ctr = rng.uniform(0.1, 0.9, (n, 2))
        size_px = 2.0 ** rng.uniform(4.0, np.log2(0.7 * min(h, w)), n)
        aspect = 2.0 ** rng.uniform(-1.2, 1.2, n)
        hh = size_px * np.sqrt(aspect) / h
        ww = size_px / np.sqrt(aspect) / w
        boxes = np.stack([ctr[:, 0] - hh / 2, ctr[:, 1] - ww / 2, ctr[:, 0] + hh / 2, ctr[:, 1] + ww / 2], -1)
        boxes = np.clip(boxes, 0.0, 1.0).astype(np.float32)
        out.append((boxes, rng.integers(0, C, n).astype(np.int32)))
    return out


def run_case(ref, case):
    tf = ref.tf
    (h, w), C, B, _, seed = CASES[case]
    rng = np.random.default_rng(seed)
    levels = ref.levels.build_levels()
    names = list(levels.keys())
    image_size = tf.convert_to_tensor(np.array([h, w], np.int32))
    gt = ground_truth(case, rng)
    out = {'image_size': np.array([h, w], np.int32), 'num_classes': np.int32(C), 'batch': np.int32(B),
           'levels': np.array(names)}

    # ---- dataset.build_labels, one image at a time (dataset.py:167)
    per_image = []
    for i, (boxes, classes) in enumerate(gt):
        out['boxes_%d' % i], out['classes_%d' % i] = boxes, classes
        cls, reg, mask = ref.dataset.build_labels(image_size, tf.convert_to_tensor(classes), tf.convert_to_tensor(boxes),
                                                  levels=levels, num_classes=C)
        per_image.append((cls, reg, mask))
    labels_cls = {pn: tf.stack([p[0][pn] for p in per_image], 0) for pn in names}
    labels_reg = {pn: tf.stack([p[1][pn] for p in per_image], 0) for pn in names}
    labels_mask = {pn: tf.stack([p[2][pn] for p in per_image], 0) for pn in names}
    for pn in names:
        c = labels_cls[pn].numpy()
        assert set(np.unique(c)) <= {0.0, 1.0}
        out['cls_' + pn] = c.astype(np.uint8)
        out['reg_' + pn] = labels_reg[pn].numpy()
        out['mask_' + pn] = labels_mask[pn].numpy()

    # ---- head outputs: background logits around the prior, positives pushed up; noisy regressions
    logits, regs = {}, {}
    for pn in names:
        c = out['cls_' + pn].astype(np.float32)
        x = rng.normal(-4.595, 1.5, c.shape) + c * rng.normal(6.0, 1.0, c.shape)
        r = out['reg_' + pn] + np.where(c.max(-1, keepdims=True) > 0, 0.1, 0.3) * rng.normal(0, 1, out['reg_' + pn].shape)
        r = np.where(np.isfinite(r), r, 0.0)
        logits[pn] = tf.convert_to_tensor(x.astype(np.float32))
        regs[pn] = tf.convert_to_tensor(r.astype(np.float32))
        out['logits_' + pn], out['regs_' + pn] = logits[pn].numpy(), regs[pn].numpy()

    # ---- utils.process_labels_and_logits (train.py:218)
    features = {'image': tf.zeros((B, h, w, 3)), 'detection': {'classifications': labels_cls, 'regressions': labels_reg},
                'trainable_masks': labels_mask}
    head = {'detection': {'classifications': logits, 'regressions': regs}}
    lab, log = ref.utils.process_labels_and_logits(labels=features, logits=head, levels=levels)
    lt, gt_ = lab['detection_trainable'], log['detection_trainable']
    out['lab_prob'] = lt.classification.prob.numpy()
    out['lab_reg'] = lt.regression.numpy()
    out['lab_regpp'] = lt.regression_postprocessed.numpy()
    out['log_unscaled'] = gt_.classification.unscaled.numpy()
    out['log_prob'] = gt_.classification.prob.numpy()
    out['log_reg'] = gt_.regression.numpy()
    out['log_regpp'] = gt_.regression_postprocessed.numpy()

    # ---- losses.loss (train.py:220) + gradients through the whole chain
    class_loss, regr_loss = ref.losses.loss(labels=lt, logits=gt_)
    out['loss_class'], out['loss_regr'] = class_loss.numpy(), regr_loss.numpy()
    total = class_loss + regr_loss
    grads = tf.gradients(total, [logits[pn] for pn in names] + [regs[pn] for pn in names])
    for k, pn in enumerate(names):
        out['grad_logits_' + pn] = grads[k].numpy()
        out['grad_regs_' + pn] = grads[len(names) + k].numpy()

    # ---- every loss function of losses.py on the masked tensors
    z = tf.convert_to_tensor(out['lab_prob'])
    fg = ref.utils.classmap_decode(z).fg_mask
    out['fg_mask'] = fg.numpy()
    L = ref.losses

    def record(name, fn):
        x = tf.convert_to_tensor(out['log_unscaled'])   # a fresh leaf per loss
        val = fn(x)
        out['val_' + name] = val.numpy()
        out['dx_' + name] = tf.gradients(val, x)[0].numpy()

    num_fg = tf.reduce_sum(tf.to_float(fg))
    # losses.py:119-122 (commented-out focal branch of classification_loss)
    record('focal', lambda x: tf.reduce_sum(L.focal_sigmoid_cross_entropy_with_logits(labels=z, logits=x)) /
           tf.maximum(num_fg, 1.0))
    record('bce', lambda x: tf.reduce_mean(tf.nn.sigmoid_cross_entropy_with_logits(labels=z, logits=x)))   # :124-125
    record('balanced_bce', lambda x: tf.reduce_mean(
        L.balanced_sigmoid_cross_entropy_with_logits(labels=z, logits=x, axis=0)))                         # :127-128
    record('dice', lambda x: tf.reduce_mean(L.dice_loss(labels=z, logits=x, smooth=0., axis=0)))            # :130-131
    record('jaccard', lambda x: tf.reduce_mean(L.jaccard_loss(labels=z, logits=x, axis=0)))                 # :133-134
    record('fixed_iou', lambda x: tf.reduce_mean(L.fixed_iou_loss(z, x, axis=0, smooth=1e-7)))              # :136-137
    record('focal_softmax', lambda x: tf.reduce_sum(L.focal_softmax_cross_entropy_with_logits(labels=z, logits=x)))
    record('bce_dice', lambda x: L.classification_loss(labels=z, logits=x, fg_mask=fg))                     # as committed
    xr = tf.convert_to_tensor(out['log_reg'])
    rl = L.regression_loss(labels=tf.convert_to_tensor(out['lab_reg']), logits=xr, fg_mask=fg)
    out['val_huber'] = rl.numpy()
    out['dx_huber'] = tf.gradients(rl, xr)[0].numpy()
    # elementwise focal values (the function's own return shape)
    out['elem_focal'] = L.focal_sigmoid_cross_entropy_with_logits(
        labels=z, logits=tf.convert_to_tensor(out['log_unscaled'])).numpy()

    # ---- inference chain per image (train.py:68-75): boxes_decode per level, merge, nms_classwise
    det = log['detection']
    for i in range(B):
        decoded = [ref.utils.boxes_decode(det.classification.prob[pn][i], det.regression_postprocessed[pn][i])
                   for pn in names]
        merged = ref.utils.merge_boxes_decoded(decoded)
        out['cand_boxes_%d' % i] = merged.boxes.numpy()
        out['cand_scores_%d' % i] = merged.scores.numpy()
        out['cand_classes_%d' % i] = merged.class_ids.numpy()
        kept = ref.utils.nms_classwise(merged, num_classes=C)
        out['det_boxes_%d' % i] = kept.boxes.numpy()
        out['det_scores_%d' % i] = kept.scores.numpy()
        out['det_classes_%d' % i] = kept.class_ids.numpy()

    # ---- augmentation.flip of image 0's sample (dataset.py:182-204)
    sample = {'image': tf.zeros((h, w, 3)),
              'detection': {'classifications': {pn: per_image[0][0][pn] for pn in names},
                            'regressions': {pn: per_image[0][1][pn] for pn in names}},
              'trainable_masks': {pn: per_image[0][2][pn] for pn in names}}
    flipped = ref.augmentation.flip(sample)
    for pn in names:
        out['flip_cls_' + pn] = flipped['detection']['classifications'][pn].numpy().astype(np.uint8)
        out['flip_reg_' + pn] = flipped['detection']['regressions'][pn].numpy()
        out['flip_mask_' + pn] = flipped['trainable_masks'][pn].numpy()
    return out


def main():
    ref = ref_exec.load()
    for case in CASES:
        out = run_case(ref, case)
        path = os.path.join(HERE, 'ref_exec_%s.npz' % case)
        np.savez_compressed(path, **out)
        print('%s: %d arrays, %.1f KiB; class_loss %.6f regr_loss %.6f, M=%d, fg=%d, detections %s' % (
            path, len(out), os.path.getsize(path) / 1024, out['loss_class'], out['loss_regr'],
            out['lab_prob'].shape[0], int(out['fg_mask'].sum()),
            [len(out['det_scores_%d' % i]) for i in range(int(out['batch']))]))


if __name__ == '__main__':
    main()
