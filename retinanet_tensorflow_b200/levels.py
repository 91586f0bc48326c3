"""Anchor-pyramid specification -- host mirror of the reference's levels.py (same names and
values: levels.py:5-59).  Pure host metadata (float64 NumPy like the reference); the device
side consumes it through `rn_pyramid_build` (see _pyramid.py)."""
from __future__ import annotations

import itertools

import numpy as np

_BASE_SIZES = (('P3', 32), ('P4', 64), ('P5', 128), ('P6', 256), ('P7', 512))  # levels.py:11-15


def compute_box_size(base_size, aspect_ratio, scale_ratio):
    """(h, w) in pixels of one anchor: area base_size^2 / prod(aspect) stretched by the aspect
    pair, times the scale (levels.py:49-52)."""
    ar = np.array(aspect_ratio)
    unit = np.sqrt(base_size ** 2 / ar.prod())
    return unit * ar * scale_ratio


class Level(object):
    """One pyramid level: `anchor_sizes` is [A, 2] float64 (h, w) px, aspect-major /
    scale-minor (levels.py:32-44)."""

    def __init__(self, anchor_size, anchor_aspect_ratios, anchor_scale_ratios):
        self._anchor_size = anchor_size
        self._anchor_aspect_ratios = list(anchor_aspect_ratios)
        self._anchor_scale_ratios = list(anchor_scale_ratios)

    @property
    def anchor_sizes(self):
        pairs = itertools.product(self._anchor_aspect_ratios, self._anchor_scale_ratios)
        return np.stack([compute_box_size(self._anchor_size, a, s) for a, s in pairs], 0)


class Levels(object):
    """P3..P7 with base sizes 32..512 (levels.py:5-29)."""

    def __init__(self, anchor_aspect_ratios, anchor_scale_ratios):
        self._anchor_aspect_ratios = anchor_aspect_ratios
        self._anchor_scale_ratios = anchor_scale_ratios
        self._levels = {name: Level(size, anchor_aspect_ratios, anchor_scale_ratios) for name, size in _BASE_SIZES}

    @property
    def num_anchors(self):
        return len(self._anchor_aspect_ratios) * len(self._anchor_scale_ratios)

    def keys(self):
        return self._levels.keys()

    def __getitem__(self, item):
        return self._levels[item]

    def __iter__(self):
        return iter(self.keys())


def build_levels():
    """3 aspect ratios x 3 octave scales = 9 anchors per cell (levels.py:55-59)."""
    return Levels([(1, 2), (1, 1), (2, 1)], [2 ** 0, 2 ** (1 / 3), 2 ** (2 / 3)])
