"""Shared helpers for the parity tests: run the oracle on the same seeded inputs and bring its
dense per-level outputs into the merged level-major layout the CUDA path uses."""
from __future__ import annotations

import numpy as np

import rn_oracle as oracle

from retinanet_tensorflow_b200 import synthetic

LEVEL_NAMES = ['P3', 'P4', 'P5', 'P6', 'P7']


def image_size_of(wl):
    return np.array([wl.image_h, wl.image_w], np.int32)


def oracle_labels_batch(wl, batch, first_image=0, levels=None, anchor_size_mode='tf1_int_trunc'):
    """Per level stacked oracle labels for `batch` images:
    cls {pn: [B,H,W,A,C]}, reg {pn: [B,H,W,A,4]}, mask {pn: [B,H,W,A]}, plus aux
    idx {pn: [B,H,W,A]} int64 and iou {pn: [B,H,W,A]} f32."""
    levels = oracle.build_levels() if levels is None else levels
    out = {k: {pn: [] for pn in levels} for k in ('cls', 'reg', 'mask', 'idx', 'iou')}
    for i in range(batch):
        boxes, classes = synthetic.gt_for_image(wl, first_image + i)
        for pn in levels:
            c, r, m, idx, iou = oracle.level_labels(image_size_of(wl), classes, boxes, levels[pn], 2 ** int(pn[-1]),
                                                    wl.num_classes, anchor_size_mode=anchor_size_mode,
                                                    return_aux=True)
            for k, v in zip(('cls', 'reg', 'mask', 'idx', 'iou'), (c, r, m, idx, iou)):
                out[k][pn].append(v)
    return {k: {pn: np.stack(v, 0) for pn, v in d.items()} for k, d in out.items()}


def merge_levels(per_level, trailing=0):
    """{pn: [B,H,W,A,(K)]} -> merged level-major [R,(K)]."""
    parts = []
    for pn in per_level:
        v = per_level[pn]
        parts.append(v.reshape(-1, v.shape[-1]) if trailing else v.reshape(-1))
    return np.concatenate(parts, 0)


def compact_labels(cls_merged, mask_merged):
    """Dense one-hot [R,C] + trainable mask [R] -> compact int32 labels (class | -1 | -2)."""
    lab = np.where(cls_merged.max(-1) > 0, cls_merged.argmax(-1), -1).astype(np.int32)
    lab[~mask_merged] = -2
    return lab


def split_levels(merged, ref_per_level):
    """Inverse of merge_levels, shaped like the reference dict."""
    out, off = {}, 0
    for pn, v in ref_per_level.items():
        n = int(np.prod(v.shape[:4]))
        tail = merged.shape[1:]
        out[pn] = merged[off:off + n].reshape(*v.shape[:4], *tail)
        off += n
    return out


def assert_close(actual, expected, rtol=1e-5, atol=0.0, what=''):
    actual = np.asarray(actual)
    expected = np.asarray(expected)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    err = np.abs(actual.astype(np.float64) - expected.astype(np.float64))
    tol = atol + rtol * np.abs(expected.astype(np.float64))
    bad = ~(err <= tol)
    if bad.any():
        i = np.unravel_index(np.argmax(np.where(bad, err - tol, -np.inf)), err.shape)
        raise AssertionError('%s: %d / %d elements out of tolerance (rtol=%g atol=%g); worst at %s: got %r want %r'
                             % (what, int(bad.sum()), bad.size, rtol, atol, i, actual[i], expected[i]))
