"""Randomised parity fuzz (GPU vs oracle), bit-exact checks only.  Script + importable helpers:

    python tests/fuzz_parity.py --cases 300 --seed 0        # on a GPU box

Assignment cases use adversarial ground truth: duplicates (first must win), boxes equal to an
anchor (IoU exactly 1), boxes scaled so that the IoU with an anchor sits on the 0.4 / 0.5
thresholds, slivers, near-image-sized boxes, odd image sizes and anchor sets.  NMS cases use heavy
score ties, duplicate boxes and zero-area boxes."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np


def make_assign_case(seed):
    rng = np.random.default_rng(seed)
    import rn_oracle as oracle
    h = int(rng.integers(64, 420))
    w = int(rng.integers(64, 420))
    c = int(rng.integers(1, 12))
    n = int(rng.integers(1, 40))
    levels = oracle.build_levels()
    mode = 'tf1_int_trunc' if rng.uniform() < 0.7 else 'float'
    boxes = []
    for _ in range(n):
        kind = rng.integers(0, 6)
        if kind == 0 and boxes:                       # duplicate of an earlier box
            boxes.append(boxes[int(rng.integers(0, len(boxes)))].copy())
            continue
        if kind in (1, 2):                             # derived from a real anchor
            pn = list(levels)[int(rng.integers(0, 5))]
            stride = 2 ** int(pn[-1])
            gh, gw = int(np.ceil(h / stride)), int(np.ceil(w / stride))
            sizes = oracle.anchor_size_normalised(levels[pn].anchor_sizes, np.array([h, w], np.int32), mode)
            a = int(rng.integers(0, 9))
            cy = (int(rng.integers(0, gh)) + 0.5) / gh
            cx = (int(rng.integers(0, gw)) + 0.5) / gw
            ah, aw = float(sizes[a, 0]), float(sizes[a, 1])
            if kind == 2:                              # concentric box with IoU ~ 0.4 / 0.5 / 0.45
                t = float(rng.choice([0.4, 0.5, 0.45, 0.5000001, 0.3999999]))
                s = np.sqrt(t) if rng.uniform() < 0.5 else 1 / np.sqrt(t)
                ah, aw = ah * s, aw * s
            boxes.append(np.array([cy - ah / 2, cx - aw / 2, cy + ah / 2, cx + aw / 2]))
            continue
        ctr = rng.uniform(0.0, 1.0, 2)
        if kind == 3:                                  # sliver
            sz = np.array([rng.uniform(0.002, 0.01), rng.uniform(0.2, 0.9)])[rng.permutation(2)]
        elif kind == 4:                                # near image-sized
            sz = rng.uniform(0.7, 1.0, 2)
        else:
            sz = np.exp(rng.uniform(np.log(0.02), np.log(0.6), 2))
        boxes.append(np.array([ctr[0] - sz[0] / 2, ctr[1] - sz[1] / 2, ctr[0] + sz[0] / 2, ctr[1] + sz[1] / 2]))
    boxes = np.stack(boxes).astype(np.float32)
    boxes[:, 2:] = np.maximum(boxes[:, 2:], boxes[:, :2] + np.float32(1e-4))   # keep tl < br after rounding
    classes = rng.integers(-1, c + 1, boxes.shape[0]).astype(np.int32)         # includes out-of-range ids
    return {'seed': seed, 'h': h, 'w': w, 'c': c, 'mode': mode, 'boxes': boxes, 'classes': classes}


def oracle_assign(case):
    import helpers
    import rn_oracle as oracle
    levels = oracle.build_levels()
    size = np.array([case['h'], case['w']], np.int32)
    out = {k: {} for k in ('cls', 'mask', 'idx', 'iou', 'reg')}
    for pn in levels:
        c, r, m, idx, iou = oracle.level_labels(size, case['classes'], case['boxes'], levels[pn], 2 ** int(pn[-1]),
                                                case['c'], anchor_size_mode=case['mode'], return_aux=True)
        out['cls'][pn], out['mask'][pn], out['idx'][pn], out['iou'][pn] = c[None], m[None], idx[None], iou[None]
        out['reg'][pn] = r[None]
    lab = helpers.compact_labels(helpers.merge_levels(out['cls'], 1), helpers.merge_levels(out['mask']))
    return (case['seed'], lab, helpers.merge_levels(out['idx']).astype(np.int32), helpers.merge_levels(out['iou']),
            helpers.merge_levels(out['reg'], 1))


def gpu_assign(case):
    import torch
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    n = case['boxes'].shape[0]
    # the CTA decomposition of the assign kernel varies with the case (results must not depend on it)
    patch = [(0, 0), (1, 1), (2, 2), (1, 3), (3, 2), (4, 4)][case['seed'] % 6]
    path = DenseHeadPath((case['h'], case['w']), case['c'], 1, n, anchor_size_mode=case['mode'], keep_iou=True,
                         process_group=False, assign_patch=patch)
    path.assign(torch.from_numpy(case['boxes'][None]).cuda(), torch.from_numpy(case['classes'][None]).cuda(),
                torch.tensor([n], dtype=torch.int32).cuda())
    return path.label.cpu().numpy(), path.matched.cpu().numpy(), path.iou_max.cpu().numpy(), path.deltas.cpu().numpy()


def make_nms_case(seed):
    rng = np.random.default_rng(10_000 + seed)
    n = int(rng.integers(1, 1500))
    c = int(rng.integers(1, 6))
    ctr = rng.uniform(0.2, 0.8, (n, 2))
    sz = rng.uniform(0.02, 0.3, (n, 2))
    boxes = np.concatenate([ctr - sz / 2, ctr + sz / 2], -1).astype(np.float32)
    dup = rng.uniform(size=n) < 0.2
    boxes[dup] = boxes[rng.integers(0, n, dup.sum())]
    flat = rng.uniform(size=n) < 0.03
    boxes[flat, 2] = boxes[flat, 0]                      # zero-area boxes (IoU 0 with everything)
    levels_of_score = int(rng.integers(2, 40))
    scores = (rng.integers(0, levels_of_score, n) / levels_of_score * 0.5 + 0.5).astype(np.float32)   # heavy ties
    classes = rng.integers(0, c, n).astype(np.int64)
    return {'seed': seed, 'boxes': boxes, 'scores': scores, 'classes': classes, 'c': c,
            'max_out': int(rng.choice([1000, 1000, 5, 37]))}


def oracle_nms(case):
    import rn_oracle as oracle
    cand = oracle.BoxesDecoded(case['boxes'], case['scores'], case['classes'])
    _, idx = oracle.nms_classwise(cand, case['c'], max_output_size=case['max_out'], return_indices=True)
    return case['seed'], idx


def gpu_nms(case):
    from test_gpu_decode import _nms_raw
    got, _ = _nms_raw(case['boxes'], case['scores'], case['classes'], case['c'], max_out=case['max_out'])
    return got


def run(cases, seed, workers):
    import multiprocessing as mp
    a_cases = [make_assign_case(seed * 100003 + i) for i in range(cases)]
    n_cases = [make_nms_case(seed * 100003 + i) for i in range(cases)]
    with mp.get_context('spawn').Pool(workers) as pool:
        a_ref = pool.map_async(oracle_assign, a_cases, chunksize=1)
        n_ref = pool.map_async(oracle_nms, n_cases, chunksize=1)
        a_gpu = [gpu_assign(c) for c in a_cases]
        n_gpu = [gpu_nms(c) for c in n_cases]
        a_ref, n_ref = a_ref.get(), n_ref.get()
    bad = 0
    stats = {'pos': 0, 'ign': 0, 'anchors': 0, 'kept': 0}
    for case, (lab, idx, iou, reg), (_, o_lab, o_idx, o_iou, o_reg) in zip(a_cases, a_gpu, a_ref):
        ok = np.array_equal(lab, o_lab) and np.array_equal(idx, o_idx) and np.array_equal(iou, o_iou)
        # encoded deltas (tolerance claim: 1e-5 relative + 1e-7): same non-finite pattern, close elsewhere
        fin = np.isfinite(o_reg)
        if not (np.array_equal(fin, np.isfinite(reg)) and
                np.all(np.abs(reg[fin] - o_reg[fin]) <= 1e-5 * np.abs(o_reg[fin]) + 1e-7)):
            bad += 1
            print('DELTAS MISMATCH seed', case['seed'], 'size', case['h'], case['w'],
                  float(np.abs(reg[fin] - o_reg[fin]).max()) if fin.any() else None)
        stats['pos'] += int((o_lab >= 0).sum())
        stats['ign'] += int((o_lab == -2).sum())
        stats['anchors'] += o_lab.shape[0]
        if not ok:
            bad += 1
            d = np.nonzero((lab != o_lab) | (idx != o_idx) | (iou != o_iou))[0]
            print('ASSIGN MISMATCH seed', case['seed'], 'size', case['h'], case['w'], 'first diffs', d[:5],
                  lab[d[:5]], o_lab[d[:5]], idx[d[:5]], o_idx[d[:5]], iou[d[:5]], o_iou[d[:5]])
    for case, got, (_, want) in zip(n_cases, n_gpu, n_ref):
        stats['kept'] += int(want.shape[0])
        if not np.array_equal(got, want):
            bad += 1
            print('NMS MISMATCH seed', case['seed'], got[:10], want[:10])
    print('FUZZ cases=%d assign: %d anchors, %d positives, %d ignored; nms: %d kept; mismatches=%d' % (
        cases, stats['anchors'], stats['pos'], stats['ign'], stats['kept'], bad))
    return bad


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', type=int, default=100)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--workers', type=int, default=os.cpu_count() or 4)
    a = ap.parse_args()
    sys.exit(1 if run(a.cases, a.seed, a.workers) else 0)
