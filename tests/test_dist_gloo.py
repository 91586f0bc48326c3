"""N>1 path on CPU: two gloo ranks shard a batch by image, exchange the statistics vector with
the product's own exchange helper, and must reproduce the single-process loss over the whole
batch (global normalisers).  The per-shard numbers come from the oracle
(this is a test), the exchange / finalisation code is the product's (dist.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_stats(oracle, helpers, synthetic, wl, first, count, mode_code, seed):
    """Stats vector of one shard, laid out as include/rndense.h says."""
    from retinanet_tensorflow_b200 import _lib
    ref = helpers.oracle_labels_batch(wl, count, first_image=first)
    cls = helpers.merge_levels(ref['cls'], 1)
    mask = helpers.merge_levels(ref['mask'])
    reg = helpers.merge_levels(ref['reg'], 1)
    rng = np.random.default_rng(seed + first)
    x = rng.normal(-3, 1.5, cls.shape).astype(np.float32)
    r = (reg + rng.normal(0, 0.3, reg.shape)).astype(np.float32)
    z, xm = cls[mask], x[mask]
    fg = z.max(-1) > 0.5
    c = wl.num_classes
    s = np.zeros(_lib.RN_STATS_HEADER + 3 * c)
    s[0], s[1] = mask.sum(), fg.sum()
    s[2] = oracle.focal_sigmoid_cross_entropy_with_logits(z, xm).astype(np.float64).sum()
    s[3] = (oracle.huber(reg[mask], r[mask]) * fg[:, None]).astype(np.float64).sum()
    p = oracle.sigmoid(xm).astype(np.float64)
    s[8:8 + c] = z.sum(0)
    s[8 + c:8 + 2 * c] = (z * p).sum(0)
    s[8 + 2 * c:8 + 3 * c] = p.sum(0)
    return s, (z, xm, reg[mask], r[mask], fg)


def _worker(rank, world, port, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
    import torch
    import torch.distributed as tdist
    import helpers
    import rn_oracle as oracle
    from retinanet_tensorflow_b200 import _lib, dist, synthetic
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    tdist.init_process_group('gloo', rank=rank, world_size=world)
    wl = synthetic.WORKLOADS['shapes']
    global_batch = 3  # uneven split on purpose
    first, count = dist.shard_images(global_batch, rank, world)
    mode = _lib.RN_CLS_FOCAL | _lib.RN_CLS_DICE_FLAG
    s, _ = _shard_stats(oracle, helpers, synthetic, wl, first, count, mode, seed=77)
    local = torch.from_numpy(s)
    sl = dist.slices(wl.num_classes)
    # the three exchanges of a step, in the order DenseHeadPath issues them: label counts (whole
    # vector, sums still zero) -> dice sums -> loss sums
    stats = torch.zeros_like(local)
    stats[sl['counts']] = local[sl['counts']]
    stats[sl['loss_sums']] = 0
    dist.all_reduce_sum(stats)
    stats[sl['dice_sums']] = local[sl['dice_sums']]
    dist.all_reduce_sum(stats[sl['dice_sums']])
    stats[sl['loss_sums']] = local[sl['loss_sums']]
    dist.all_reduce_sum(stats[sl['loss_sums']])
    cls_loss, reg_loss, nfg, m = dist.finalize(stats, wl.num_classes, mode)
    if rank == 0:
        # single-process reference over the whole batch
        parts = [_shard_stats(oracle, helpers, synthetic, wl, *dist.shard_images(global_batch, r, world), mode, 77)[1]
                 for r in range(world)]
        z, x, t, r_, fg = (np.concatenate([p[i] for p in parts], 0) for i in range(5))
        want_cls = oracle.classification_loss(z, x, fg, mode='focal_dice')
        want_reg = oracle.regression_loss(t, r_, fg)
        np.save(out, np.array([float(cls_loss), float(reg_loss), float(nfg), float(m), want_cls, want_reg, fg.sum(),
                               z.shape[0]]))
    tdist.barrier()
    tdist.destroy_process_group()


def test_two_rank_stats_exchange_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / 'res.npy')
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    np.testing.assert_allclose(got[0], got[4], rtol=1e-5)
    np.testing.assert_allclose(got[1], got[5], rtol=1e-5)
    assert got[2] == got[6] and got[3] == got[7]


def test_shard_images_covers_the_batch():
    from retinanet_tensorflow_b200 import dist
    for batch in (1, 3, 64, 65):
        for world in (1, 2, 4, 8):
            ranges = [dist.shard_images(batch, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and sum(c for _, c in ranges) == batch
            for (f0, c0), (f1, _) in zip(ranges, ranges[1:]):
                assert f0 + c0 == f1
