"""Multi-GPU plumbing of the path: image sharding and the one exchange step.

The path shards by image (assignment, decode and NMS are per image).  Only the loss couples
images: its normalisers (M, num_fg, per-class sums) and its sums span the batch
(utils.py:270-278, losses.py:130,139,150).  Every rank therefore keeps the float64 statistics
vector of include/rndense.h and all-reduces (sum) slices of it: the label counts before the
gradient pass, the dice sums between its two phases, the loss sums after it.  ~1 KB per step,
latency-bound; NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch

from . import _lib

H = _lib.RN_STATS_HEADER


def shard_images(global_batch, rank, world):
    """Contiguous image range [first, first + count) of `rank`."""
    base, rem = divmod(int(global_batch), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def stats_length(num_classes):
    return H + 3 * int(num_classes)


def slices(num_classes):
    """The three exchanged slices of the stats vector."""
    c = int(num_classes)
    return {'counts': slice(0, H + c), 'dice_sums': slice(H + c, H + 3 * c),
            'loss_sums': slice(_lib.RN_STAT_CLS_SUM, _lib.RN_STAT_HUBER_SUM + 1)}


def all_reduce_sum(t, group=None):
    """Sum `t` over the ranks in place; no-op without an initialised multi-rank group."""
    dist = torch.distributed
    if group is False or not (dist.is_available() and dist.is_initialized()):
        return t
    if dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def finalize(stats, num_classes, mode):
    """Host mirror of rn_loss_finalize (float64 torch ops on any device):
    -> (class_loss, regr_loss, num_fg, M) from an all-reduced stats vector."""
    c = int(num_classes)
    s = stats.double()
    m, nfg = s[_lib.RN_STAT_TRAINABLE], s[_lib.RN_STAT_NUM_FG]
    base = mode & 15
    cls = s.new_zeros(())
    if base == _lib.RN_CLS_FOCAL:
        cls = cls + s[_lib.RN_STAT_CLS_SUM] / torch.clamp(nfg, min=1.0)
    elif base in (_lib.RN_CLS_BCE, _lib.RN_CLS_BALANCED_BCE):
        cls = cls + (s[_lib.RN_STAT_CLS_SUM] / (m * c) if float(m) > 0 else 0.0)
    if mode & _lib.RN_CLS_DICE_FLAG:
        npos, inter, psum = s[H:H + c], s[H + c:H + 2 * c], s[H + 2 * c:H + 3 * c]
        cls = cls + (1.0 - 2.0 * inter / (npos + psum)).mean()
    reg = s[_lib.RN_STAT_HUBER_SUM] / (4.0 * nfg) if float(nfg) > 0 else s.new_zeros(())
    return cls, reg, nfg, m
