"""Randomised fuzz of the one-launch detection tail (GPU): rn_detect_tail against rn_select_finalize +
rn_nms_classwise on the candidates the fused loss pass leaves in the workspace, for random image sizes,
class counts, thresholds, capacities, per-level top-k, per-class / total output limits and logit distributions (sparse and
dense firing, several classes per anchor, exact score ties).

    python tests/fuzz_tail.py --cases 200 --seed 0        # on a GPU box

The candidates are appended by atomics in a different order on every run, so two runs of the fused pass
also check that both tails are independent of that order.  Cases whose candidate lists overflow (an
arbitrary subset survives) only compare counts and flags."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

CLASS_COUNTS = [4, 8, 12, 20, 24, 80, 96]


def _snapshot(path, batch):
    import torch
    out = {'cand_count': path.cand_count.clone(), 'det_count': path.det_count.clone(), 'status': int(path.status.item())}
    for k in ('cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes'):
        out[k] = [getattr(path, k)[b, :int(path.cand_count[b])].clone() for b in range(batch)]
    for k in ('det_index', 'det_boxes', 'det_scores', 'det_classes'):
        out[k] = [getattr(path, k)[b, :int(path.det_count[b])].clone() for b in range(batch)]
    return out


def run_case(seed):
    import torch
    from retinanet_tensorflow_b200 import _lib
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(40, 360)), int(rng.integers(40, 360))
    c = int(rng.choice(CLASS_COUNTS))
    batch = int(rng.integers(1, 5))
    thr = float(rng.choice([0.3, 0.5, 0.5, 0.9]))
    cap = int(rng.choice([200, 1000, 4096, 8192, 8192]))
    mu = float(rng.choice([-4.595, -4.595, -3.5]))
    sigma = float(rng.choice([0.5, 1.0, 1.5, 2.0]))
    per_class = int(rng.choice([1000, 1000, 1, 7]))
    ocap = int(rng.choice([cap, cap, 50]))
    nms_iou = float(rng.choice([0.5, 0.5, 0.0, 0.3, 1.0]))
    spread = float(rng.choice([0.05, 0.2, 0.6]))
    topk = int(rng.choice([0, 0, 5, 60, 1000]))
    desc = 'seed=%d %dx%d C=%d B=%d thr=%.2f cap=%d mu=%.1f sigma=%.1f per_class=%d ocap=%d iou=%.1f spread=%.2f topk=%d' % (
        seed, h, w, c, batch, thr, cap, mu, sigma, per_class, ocap, nms_iou, spread, topk)
    path = DenseHeadPath((h, w), c, batch, 4, score_thr=thr, capacity=cap, out_capacity=ocap, nms_iou=nms_iou,
                         max_output_per_class=per_class, topk_per_level=topk, process_group=False, one_launch_tail=False)
    gt_boxes = np.zeros((batch, 4, 4), np.float32)
    gt_boxes[:, 0] = [0.2, 0.2, 0.7, 0.8]
    path.assign(torch.from_numpy(gt_boxes).cuda(), torch.zeros((batch, 4), dtype=torch.int32, device='cuda'),
                torch.ones((batch,), dtype=torch.int32, device='cuda'))
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.empty((path.rows_total, c), device='cuda').normal_(mu, sigma, generator=g)
    rows = torch.randint(0, path.rows_total, (max(1, path.rows_total // 40),), device='cuda', generator=g)
    x[rows, 0] = 3.0           # exact ties in probability space, within a row and across 16-byte units
    x[rows, c - 1] = 3.0
    x[rows[::3], c // 2] = 2.0
    regs = torch.empty((path.rows_total, 4), device='cuda').normal_(0, spread, generator=g)
    assert path.fused_select_supported(), desc
    path.loss_and_select(x, regs, finalize=False)
    path.select_finalize(regs)
    path.nms()
    torch.cuda.synchronize()
    want = _snapshot(path, batch)
    for k in ('cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes', 'det_index', 'det_boxes', 'det_scores',
              'det_classes', 'cand_count', 'det_count', 'status'):
        getattr(path, k).zero_()
    path.one_launch_tail = True
    assert path.tail_fused_supported(), desc
    path.loss_and_select(x, regs, finalize=False)
    path.detect_tail(regs)
    torch.cuda.synchronize()
    got = _snapshot(path, batch)
    ov = _lib.RN_FLAG_CAND_OVERFLOW
    assert got['status'] & ov == want['status'] & ov, (desc, got['status'], want['status'])
    if want['status'] & ov:   # an arbitrary subset of the candidates survived: nothing further is comparable
        assert int(got['cand_count'].max()) <= cap, desc
        return desc + ' (overflow)'
    assert got['status'] == want['status'], (desc, got['status'], want['status'])
    assert torch.equal(got['cand_count'], want['cand_count']), (desc, got['cand_count'], want['cand_count'])
    assert torch.equal(got['det_count'], want['det_count']), (desc, got['det_count'], want['det_count'])
    for k in ('cand_anchor', 'cand_scores', 'cand_classes', 'cand_boxes', 'det_index', 'det_boxes', 'det_scores',
              'det_classes'):
        for b in range(batch):
            assert torch.equal(got[k][b], want[k][b]), (desc, k, b)
    return desc


def run(cases, seed, verbose=False):
    compared = 0
    for i in range(cases):
        try:
            compared += not run_case(seed * 100003 + i).endswith('(overflow)')
        except Exception:
            print('FAILED tail case seed=%d' % (seed * 100003 + i))
            raise
        if verbose and (i + 1) % 50 == 0:
            print('%d cases ok (%d compared in full)' % (i + 1, compared), flush=True)
    return cases


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', type=int, default=200)
    ap.add_argument('--seed', type=int, default=0)
    a = ap.parse_args()
    print('FUZZ_TAIL OK: %d cases' % run(a.cases, a.seed, verbose=True))
