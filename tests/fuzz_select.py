"""Randomised fuzz of candidate selection (GPU): the standalone select kernel and the selection fused
into the loss kernel, on random image sizes, class counts, thresholds, capacities, logit distributions
(sparse and dense firing) and dtypes, against utils.boxes_decode's rule restated in NumPy:

    score = max_c sigmoid(x), class = first argmax, keep score > thr, ascending anchor order per image.

    python tests/fuzz_select.py --cases 200 --seed 0        # on a GPU box

Without overflow the candidate lists must be identical (anchor, class, score bit-exact).  With more
candidates than `capacity` the status word must carry RN_FLAG_CAND_OVERFLOW, the count must equal the
capacity and every reported candidate must be a true one."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

CLASS_COUNTS = [1, 3, 4, 8, 12, 20, 24, 80]


def run_case(seed):
    import ctypes as C
    import torch
    from retinanet_tensorflow_b200 import _lib
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(40, 300)), int(rng.integers(40, 300))
    c = int(rng.choice(CLASS_COUNTS))
    batch = int(rng.integers(1, 5))
    thr = float(rng.choice([0.05, 0.3, 0.5, 0.9]))
    cap = int(rng.choice([64, 1000, 16384]))
    mu = float(rng.choice([-4.595, -2.0, 0.0]))
    sigma = float(rng.choice([0.5, 1.0, 2.0, 5.0]))
    bf16 = bool(c % 8 == 0 and rng.uniform() < 0.4)
    fused = bool(rng.uniform() < 0.5)
    mode = str(rng.choice(['focal', 'bce', 'bce_dice']))
    desc = 'seed=%d %dx%d C=%d B=%d thr=%.2f cap=%d mu=%.1f sigma=%.1f bf16=%d fused=%d mode=%s' % (
        seed, h, w, c, batch, thr, cap, mu, sigma, bf16, fused, mode)
    path = DenseHeadPath((h, w), c, batch, 4, cls_mode=mode, score_thr=thr, capacity=cap, process_group=False)
    gt_boxes = np.zeros((batch, 4, 4), np.float32)
    gt_boxes[:, 0] = [0.2, 0.2, 0.7, 0.8]
    gt_classes = np.zeros((batch, 4), np.int32)
    gt_counts = np.ones((batch,), np.int32)
    path.assign(torch.from_numpy(gt_boxes).cuda(), torch.from_numpy(gt_classes).cuda(), torch.from_numpy(gt_counts).cuda())
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.empty((path.rows_total, c), device='cuda').normal_(mu, sigma, generator=g)
    # exact ties in probability space, within a row and across float4 units
    rows = torch.randint(0, path.rows_total, (max(1, path.rows_total // 50),), device='cuda', generator=g)
    x[rows, 0] = 3.0
    x[rows, c - 1] = 3.0
    if bf16:
        x = x.bfloat16()
    regs = torch.empty((path.rows_total, 4), device='cuda').normal_(0, 0.2, generator=g)
    fused = fused and path.fused_select_supported()
    if fused:
        path.loss_and_select(x, regs)
    else:
        path.select(x, regs)
    torch.cuda.synchronize()
    status = int(path.status.item())
    path.status.zero_()

    xf = x.float()
    p = torch.empty_like(xf)
    _lib.call('rn_sigmoid', xf.numel(), _lib.ptr(xf), _lib.ptr(p), _lib.stream_ptr())
    p_lv = path.per_level(p, (c,))
    for b in range(batch):
        pi = np.concatenate([p_lv[pn][b].reshape(-1, c).cpu().numpy() for pn in path.names])
        score = pi.max(-1)
        cls = pi.argmax(-1)           # first maximum
        want = np.nonzero(score > np.float32(thr))[0]
        n = int(path.cand_count[b])
        a = path.cand_anchor[b, :n].cpu().numpy()
        k = path.cand_classes[b, :n].cpu().numpy()
        s = path.cand_scores[b, :n].cpu().numpy()
        overflow = bool(status & _lib.RN_FLAG_CAND_OVERFLOW)
        if want.shape[0] > cap:
            assert overflow, desc
        if overflow and (n != want.shape[0] or fused):
            assert n <= cap and (n == cap or fused), (desc, n, cap)
            assert np.all(np.diff(a) > 0), desc
            assert np.isin(a, want).all(), desc
        else:
            assert n == want.shape[0], (desc, b, n, want.shape[0])
            np.testing.assert_array_equal(a, want, err_msg=desc)
        if fused and overflow:
            # the fused pass records one entry per firing float4 unit; when some were dropped an anchor's best unit
            # may be among them, so only "a true candidate with at most its true score" can be claimed
            assert np.all(s <= score[a]) and np.all(s > np.float32(thr)), desc
        else:
            np.testing.assert_array_equal(k, cls[a], err_msg=desc)
            np.testing.assert_array_equal(s, score[a], err_msg=desc)
    return desc


def run(cases, seed, verbose=False):
    for i in range(cases):
        try:
            run_case(seed * 100003 + i)
        except Exception:
            print('FAILED select case seed=%d' % (seed * 100003 + i))
            raise
        if verbose and (i + 1) % 50 == 0:
            print('%d cases ok' % (i + 1), flush=True)
    return cases


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', type=int, default=200)
    ap.add_argument('--seed', type=int, default=0)
    a = ap.parse_args()
    print('FUZZ_SELECT OK: %d cases' % run(a.cases, a.seed, verbose=True))
