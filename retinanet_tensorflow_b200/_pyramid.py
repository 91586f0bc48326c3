"""Host-side pyramid descriptors (rn_pyramid_t) built from the reference's `levels` objects."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

ANCHOR_SIZE_MODES = {'tf1_int_trunc': _lib.RN_SIZE_TF1_INT_TRUNC, 'float': _lib.RN_SIZE_FLOAT}


def level_stride(name):
    """`factor=2**int(pn[-1])` (dataset.py:133)."""
    return 2 ** int(name[-1])


_PYRAMID_CACHE = []   # [(level objects, key, pyramid)]: the per-image drop-in calls ask for the same pyramid again and again


def pyramid_for(image_size, levels, anchor_size_mode='tf1_int_trunc', names=None, strides=None):
    """rn_pyramid_t for an image size and a `Levels`-like mapping (dict order = level order).  The last few
    results are kept per set of level objects (treated as immutable, like the reference's `build_levels()`
    result); callers must not modify the returned struct."""
    names = list(levels) if names is None else list(names)
    if strides is None:
        strides = [level_stride(n) for n in names]
    key = (int(image_size[0]), int(image_size[1]), anchor_size_mode, tuple(names), tuple(int(v) for v in strides))
    objs = [levels[n] for n in names]
    for held, k, cached in _PYRAMID_CACHE:
        if k == key and all(a is b for a, b in zip(held, objs)):
            return cached
    pyr = _pyramid_build(image_size, levels, anchor_size_mode, names, strides)
    _PYRAMID_CACHE.append((objs, key, pyr))   # (list append / slice are atomic under the GIL)
    if len(_PYRAMID_CACHE) > 16:
        del _PYRAMID_CACHE[0]
    return pyr


def _pyramid_build(image_size, levels, anchor_size_mode, names, strides):
    sizes = np.ascontiguousarray(np.stack([np.asarray(levels[n].anchor_sizes, np.float64) for n in names], 0))
    if sizes.ndim != 3 or sizes.shape[2] != 2:
        raise ValueError('anchor_sizes must be [A, 2] per level')
    st = np.ascontiguousarray(np.asarray(strides, np.int32))
    pyr = _lib.Pyramid()
    h, w = int(image_size[0]), int(image_size[1])
    _lib.call('rn_pyramid_build', h, w, len(names), st.ctypes.data_as(C.c_void_p), sizes.ctypes.data_as(C.c_void_p),
              sizes.shape[1], ANCHOR_SIZE_MODES[anchor_size_mode], C.byref(pyr))
    return pyr


def pyramid_from_grids(grids, anchor_sizes_norm):
    """rn_pyramid_t from explicit [(H, W), ...] grids and normalised float32 anchor sizes
    [L, A, 2] -- what utils.anchor_boxmap / regression_postprocess are handed (utils.py:47,108)."""
    g = np.ascontiguousarray(np.asarray(grids, np.int32).reshape(-1, 2))
    a = np.ascontiguousarray(np.asarray(anchor_sizes_norm, np.float32))
    if a.ndim == 2:
        a = a[None]
    if a.shape[0] != g.shape[0] or a.shape[2] != 2:
        raise ValueError('anchor sizes must be [L, A, 2]')
    pyr = _lib.Pyramid()
    _lib.call('rn_pyramid_from_grids', g.shape[0], g.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
              a.shape[1], C.byref(pyr))
    return pyr


def level_shapes(pyr):
    return [(pyr.level[l].grid_h, pyr.level[l].grid_w, pyr.level[l].num_anchors) for l in range(pyr.num_levels)]


def level_rows(pyr, batch):
    return [batch * h * w * a for h, w, a in level_shapes(pyr)]


def split_merged(pyr, batch, merged, trailing=()):
    """Per-level views [B, H, W, A, *trailing] of a merged level-major buffer [B*A_total, *trailing]."""
    out, off = [], 0
    for h, w, a in level_shapes(pyr):
        n = batch * h * w * a
        out.append(merged[off:off + n].view(batch, h, w, a, *trailing))
        off += n
    return out
