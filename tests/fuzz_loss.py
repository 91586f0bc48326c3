"""Randomised parity fuzz of the loss kernels (GPU, through the C ABI) against the oracle.

    python tests/fuzz_loss.py --cases 400 --seed 0        # on a GPU box

Every case draws: number of classes (vector and scalar layouts, bf16-able or not), 1-5 row
segments of arbitrary length (NOT pyramid shaped: odd tails, fewer rows than threads), a label mix
(no positives / few / many, no ignored / some / all), a logit distribution (around the prior, centred,
wide, saturated), a loss mode, float32 or bfloat16 logits, and the fast or the generic kernel.
Checked: classification loss, num_fg, M and the full gradient against rn_oracle within the
tolerances of tests/test_gpu_loss.py; Huber loss and gradient on random regression rows."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

MODES = ['focal', 'bce', 'balanced_bce', 'bce_dice', 'balanced_bce_dice', 'focal_dice']
CLASS_COUNTS = [1, 2, 3, 4, 5, 8, 12, 16, 20, 24, 40, 80, 96]


def make_case(seed):
    rng = np.random.default_rng(seed)
    c = int(rng.choice(CLASS_COUNTS))
    nseg = int(rng.integers(1, 6))
    big = rng.uniform() < 0.15
    seg_rows = [int(rng.integers(1, 40000 if big else 1500)) for _ in range(nseg)]
    rows = sum(seg_rows)
    p_pos = float(rng.choice([0.0, 0.01, 0.3]))
    p_ign = float(rng.choice([0.0, 0.05, 0.5, 1.0], p=[0.4, 0.4, 0.15, 0.05]))
    u = rng.uniform(size=rows)
    lab = np.full(rows, -1, np.int32)
    lab[u < p_pos] = rng.integers(0, c, size=int((u < p_pos).sum()))
    lab[u > 1.0 - p_ign] = -2
    mu = float(rng.choice([-4.595, -1.0, 0.0, 2.0], p=[0.5, 0.2, 0.2, 0.1]))
    sigma = float(rng.choice([0.3, 1.0, 2.0, 6.0, 20.0], p=[0.2, 0.35, 0.25, 0.15, 0.05]))
    x = rng.normal(mu, sigma, (rows, c)).astype(np.float32)
    pos = np.nonzero(lab >= 0)[0]
    x[pos, lab[pos]] += rng.normal(6.0, 1.0, pos.shape[0]).astype(np.float32)
    mode = str(rng.choice(MODES))
    bf16 = bool(c % 8 == 0 and rng.uniform() < 0.35)
    generic = bool(not bf16 and rng.uniform() < 0.2)
    head = bool(rng.uniform() < 0.5)   # the one-launch entry point (rn_head_loss) instead of the separate calls
    pred = rng.normal(0, 1.5, (rows, 4)).astype(np.float32)
    target = rng.normal(0, 1.0, (rows, 4)).astype(np.float32)
    return dict(c=c, seg_rows=seg_rows, lab=lab, x=x, mode=mode, bf16=bf16, generic=generic, head=head, pred=pred,
                target=target, desc='C=%d segs=%s mode=%s bf16=%d generic=%d head=%d mu=%.1f sigma=%.1f pos=%.2f ign=%.2f'
                % (c, seg_rows, mode, bf16, generic, head, mu, sigma, p_pos, p_ign))


def run_case(case):
    import torch
    import rn_oracle as oracle
    import helpers
    from retinanet_tensorflow_b200 import _lib
    from retinanet_tensorflow_b200.pipeline import CLS_MODES
    c, seg_rows, lab, mode = case['c'], case['seg_rows'], case['lab'], case['mode']
    mode_bits = CLS_MODES[mode]
    run_bits = mode_bits | (_lib.RN_CLS_GENERIC_FLAG if case['generic'] else 0)   # per-call kernel choice
    dev = torch.device('cuda')
    x = torch.from_numpy(case['x']).to(dev)
    if case['bf16']:
        x = x.bfloat16()
    x_ref = x.float().cpu().numpy()
    lab_t = torch.from_numpy(lab).to(dev)
    pred, target = torch.from_numpy(case['pred']).to(dev), torch.from_numpy(case['target']).to(dev)
    dlogits, dreg = torch.full_like(x, 7.0), torch.full_like(pred, 7.0)

    def split(t):
        out, o = [], 0
        for r in seg_rows:
            out.append(t[o:o + r])
            o += r
        return out

    rows = _lib.rows_of(seg_rows)
    m_x, m_lab, m_dx = _lib.maps_of(split(x)), _lib.maps_of(split(lab_t)), _lib.maps_of(split(dlogits))
    m_pred, m_tgt, m_dreg = _lib.maps_of(split(pred)), _lib.maps_of(split(target)), _lib.maps_of(split(dreg))
    stats = torch.zeros((_lib.RN_STATS_HEADER + 3 * c,), dtype=torch.float64, device=dev)
    out = torch.zeros((4,), dtype=torch.float32, device=dev)
    ws_bytes = _lib.lib.rn_loss_workspace_bytes(c)
    ws = torch.zeros((ws_bytes,), dtype=torch.uint8, device=dev)
    hws = torch.zeros((_lib.lib.rn_loss_workspace_bytes(1),), dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr()
    dt = _lib.RN_DTYPE_BF16 if case['bf16'] else _lib.RN_DTYPE_F32
    _lib.call('rn_label_stats', C.byref(rows), c, C.byref(m_lab), _lib.ptr(stats), st)
    if mode_bits & _lib.RN_CLS_DICE_FLAG:
        _lib.call('rn_cls_sums', C.byref(rows), c, C.byref(m_x), C.byref(m_lab), 0, dt, _lib.ptr(stats),
                  _lib.ptr(ws), ws_bytes, st)
    if case['head']:
        _lib.call('rn_head_loss', C.byref(rows), c, C.byref(m_x), C.byref(m_lab), run_bits, 2.0, 0.25, 1e-7,
                  C.byref(m_dx), dt, C.byref(m_pred), C.byref(m_tgt), C.byref(m_dreg), _lib.ptr(stats), None,
                  _lib.ptr(ws), ws_bytes, _lib.ptr(out), 0.0, 0, None, None, None, 0, st)
    else:
        _lib.call('rn_cls_loss_fwd_bwd', C.byref(rows), c, C.byref(m_x), C.byref(m_lab), 0, run_bits, 2.0, 0.25,
                  1e-7, C.byref(m_dx), dt, _lib.ptr(stats), _lib.ptr(ws), ws_bytes, st)
        _lib.call('rn_huber_fwd_bwd', C.byref(rows), C.byref(m_pred), C.byref(m_tgt), C.byref(m_lab),
                  C.byref(_lib.Maps()), C.byref(m_dreg), _lib.ptr(stats), _lib.ptr(hws), hws.numel(), st)
        _lib.call('rn_loss_finalize', c, mode_bits, _lib.ptr(stats), _lib.ptr(out), st)
    torch.cuda.synchronize()
    got = out.cpu().numpy()

    mask = lab != -2
    cls = np.zeros((lab.shape[0], c), np.float32)
    pos = np.nonzero(lab >= 0)[0]
    cls[pos, lab[pos]] = 1.0
    fg = (lab >= 0)[mask]
    assert got[2] == fg.sum() and got[3] == mask.sum(), ('counts', got, fg.sum(), mask.sum())
    g = dlogits.float().cpu().numpy()
    gr = dreg.cpu().numpy()
    if mask.sum() == 0:   # every anchor ignored: the reference's means are 0/0; gradients must be zero
        assert not g.any() and not gr.any()
        return
    assert np.isfinite(g).all() and np.isfinite(got).all()
    with np.errstate(all='ignore'):
        want = oracle.classification_loss(cls[mask], x_ref[mask], fg, mode=mode)
        g_ref = np.zeros_like(x_ref)
        g_ref[mask] = oracle.classification_loss_grad(cls[mask], x_ref[mask], fg, mode=mode)
        want_r = oracle.regression_loss(case['target'][mask], case['pred'][mask], fg)
        gr_ref = np.zeros_like(case['pred'])
        gr_ref[mask] = oracle.regression_loss_grad(case['target'][mask], case['pred'][mask], fg)
    if np.isfinite(want):   # dice with an empty class is 0/0 in the reference too
        np.testing.assert_allclose(got[0], want, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got[1], want_r, rtol=1e-5, atol=1e-7)
    ok = np.isfinite(g_ref)
    scale = float(np.abs(g_ref[ok]).max()) if ok.any() else 0.0
    # focal, saturated logits (x > 9): the reference's own 1 - p is quantised (see test_gpu_loss.py)
    sat = (x_ref > 9.0) if 'focal' in mode else np.zeros_like(ok)
    if 'dice' in mode:
        # the dice gradient carries p (1 - p) (TF's SigmoidGrad y (1 - y)): for p -> 1 the float32 `1 - p` has few
        # significant bits (x = 5.3: a 1-ulp change of p moves it by 1.2e-5 relative) on ANY implementation
        sat = sat | (x_ref > 4.0)
    rtol = 2.0 ** -7 if case['bf16'] else 1e-5
    helpers.assert_close(np.where(sat | ~ok, 0, g), np.where(sat | ~ok, 0, g_ref), rtol=rtol, atol=1e-6 * scale + 1e-12,
                         what='dlogits')
    helpers.assert_close(gr, gr_ref, rtol=1e-5, atol=1e-6 * float(np.abs(gr_ref).max()) + 1e-12, what='dreg')


def run(cases, seed, verbose=False):
    done = 0
    for i in range(cases):
        case = make_case(seed * 100003 + i)
        try:
            run_case(case)
        except Exception:
            print('FAILED case %d: %s' % (i, case['desc']))
            raise
        done += 1
        if verbose and done % 50 == 0:
            print('%d cases ok' % done, flush=True)
    return done


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', type=int, default=400)
    ap.add_argument('--seed', type=int, default=0)
    a = ap.parse_args()
    n = run(a.cases, a.seed, verbose=True)
    print('FUZZ_LOSS OK: %d cases' % n)
