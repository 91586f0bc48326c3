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


PHASES = {'counts': 0, 'dice_sums': 1, 'loss_sums': 2}


def phase_slice(phase, num_classes):
    """(offset, length) of the stats-vector slice a phase exchanges."""
    sl = slices(num_classes)[phase]
    return sl.start, sl.stop - sl.start


class PeerExchange(object):
    """The stats all-reduce as this library's own kernel over NVLink peer memory
    (`rn_stats_exchange`): every rank stores its slice into every peer's symmetric-memory mailbox,
    publishes an epoch flag and sums the `world` messages in rank order.  No NCCL call on the step's
    critical path; results are bit-identical on all ranks.  Needs torch symmetric memory
    (CUDA VMM peer mappings inside one NVLink domain)."""

    SLOTS = 6  # 3 phases x epoch parity

    def __init__(self, num_classes, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        tdist = torch.distributed
        self.group = group if group is not None else tdist.group.WORLD
        self.world = tdist.get_world_size(self.group)
        self.rank = tdist.get_rank(self.group)
        self.num_classes = int(num_classes)
        self.len_max = stats_length(num_classes)
        nbytes = int(_lib.lib.rn_stats_mailbox_bytes(self.world, self.len_max, self.SLOTS))
        self.buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.int64, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.buf.data_ptr():
            raise RuntimeError('symmetric-memory rendezvous returned unexpected peer pointers')
        self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.epochs = torch.zeros((len(PHASES),), dtype=torch.int64, device=device)   # advanced by the kernel
        torch.cuda.synchronize(device)
        tdist.barrier(self.group)     # every mailbox is zeroed before anybody posts
        self._C = C

    def all_reduce(self, stats, phase, status=None):
        offset, length = phase_slice(phase, self.num_classes)
        _lib.call('rn_stats_exchange_auto', _lib.ptr(stats), offset, length, self.len_max, self.world, self.rank,
                  _lib.ptr(self.peer_ptrs), PHASES[phase], len(PHASES), _lib.ptr(self.epochs), _lib.ptr(status),
                  _lib.stream_ptr())
        return stats
