"""Tensor plumbing shared by the reference-facing modules: anything array-like or
DLPack-capable in, contiguous CUDA torch tensors out.  torch is used for device memory and
streams only; all arithmetic of the path happens in librndense.so."""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import torch

from . import _lib
from ._pyramid import pyramid_from_grids


def as_cuda(x, dtype=torch.float32):
    """torch CUDA tensor (contiguous) from a torch tensor, a DLPack exporter (e.g. an eager
    tf.Tensor), a NumPy array or nested lists."""
    dev = _lib.require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    elif hasattr(x, '__dlpack__') and not isinstance(x, np.ndarray):
        t = torch.from_dlpack(x)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
    if t.device.type != 'cuda':
        t = t.to(dev)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def host_ints(x):
    """Small integer vectors (image_size, grid_size) are host metadata."""
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    elif hasattr(x, '__dlpack__') and not isinstance(x, np.ndarray):
        x = torch.from_dlpack(x).cpu().numpy()
    return [int(v) for v in np.asarray(x).reshape(-1)]


@functools.lru_cache(maxsize=64)
def _flat_pyramid_plain(rows):
    return pyramid_from_grids([(rows, 1)], np.zeros((1, 1, 2), np.float32))


def flat_pyramid(rows, anchor_hw=None):
    """A one-level pyramid whose `rows` anchors are just rows (geometry unused).  Callers do not modify it."""
    if anchor_hw is None:
        return _flat_pyramid_plain(max(int(rows), 1))
    return pyramid_from_grids([(max(int(rows), 1), 1)], np.asarray(anchor_hw, np.float32).reshape(1, -1, 2))


def new_status():
    return torch.zeros((1,), dtype=torch.int32, device=_lib.require_cuda())


def boxmap_op(pyr, batch, op, x):
    """rn_boxmap_op on one-level input `x` [.., 4] (already shaped for `pyr`)."""
    out = torch.empty_like(x)
    _lib.call('rn_boxmap_op', C.byref(pyr), batch, op, C.byref(_lib.maps_of([x])), C.byref(_lib.maps_of([out])),
              _lib.stream_ptr())
    return out
