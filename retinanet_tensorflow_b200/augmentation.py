"""Horizontal flip of label maps -- host mirror of the reference's augmentation.py (5-22).

`flip(input)` keeps the reference's dict-in / dict-out signature (reverse the W axis of the
image and of every per-level label map, negate dx).  `flip_compact` is the form the GPU path
uses: it flips the compact K1 outputs (label, deltas) of a whole batch in one launch, which is
how the reference's `[orig, flipped]` sample (dataset.py:182-204) is produced on the device.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._tensor import as_cuda
from . import utils


def flip_compact(pyr, batch, label=None, deltas=None):
    """label int32 [R] and/or deltas float32 [R, 4] in the merged level-major layout of `pyr`
    -> flipped copies (augmentation.py:7-13)."""
    def maps(t, nbytes):
        m = _lib.Maps()
        if t is not None:
            _lib.call('rn_maps_from_merged', C.byref(pyr), batch, _lib.ptr(t), nbytes, C.byref(m))
        return m

    out_label = torch.empty_like(label) if label is not None else None
    out_deltas = torch.empty_like(deltas) if deltas is not None else None
    _lib.call('rn_flip_labels', C.byref(pyr), batch, C.byref(maps(label, 4)), C.byref(maps(deltas, 16)),
              C.byref(maps(out_label, 4)), C.byref(maps(out_deltas, 16)), _lib.stream_ptr())
    return out_label, out_deltas


def _flip_w(x):
    return torch.flip(as_cuda(x, None), dims=[1])  # layout move only


def flip(input):
    """augmentation.flip (augmentation.py:5-22) on the reference's dense dict structure:
    per-level maps are [H, W, A, .] (one image), the W axis is axis 1."""
    from ._pyramid import pyramid_from_grids
    import numpy as np
    regressions = {}
    for pn, r in input['detection']['regressions'].items():
        r = as_cuda(r)
        h, w, a, _ = r.shape
        pyr = pyramid_from_grids([(h, w)], np.zeros((1, a, 2), np.float32))
        _, flipped = flip_compact(pyr, 1, None, r.reshape(-1, 4))
        regressions[pn] = flipped.reshape(h, w, a, 4)
    return {
        'image': _flip_w(input['image']),
        'detection': {
            'classifications': utils.dict_map(_flip_w, input['detection']['classifications']),
            'regressions': regressions,
        },
        'trainable_masks': utils.dict_map(_flip_w, input['trainable_masks']),
    }
