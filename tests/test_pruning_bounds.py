"""The exact-arithmetic shortcuts of the assignment and NMS kernels, restated in NumPy float32 and checked against the
reference's own decision (IEEE division) on adversarial random boxes -- CPU only.

The kernels avoid most IEEE divisions with conservative tests (csrc/rn_assign.cu, csrc/rn_decode.cu):

  assign, a pair (anchor, GT) may only be skipped when it cannot beat the running best:
      fl(inter / uni) > best   ==>   fl(inter * c) > fl(best * uni)            (c = 1 + 2e-6)
      fl(best * max(aa, ga)) > fl(min(aa, ga) * c)   ==>   fl(inter / uni) <= best
  NMS, `fl(inter / uni) > thr` is decided from the products outside a 2e-6 band:
      fl(inter) > fl(fl(thr * uni) * c)   ==>   fl(inter / uni) > thr
      fl(inter * c) < fl(thr * uni)       ==>   not (fl(inter / uni) > thr)

Every operation below is a separate float32 rounding, like the kernels' __fmul_rn / __fsub_rn / __fdiv_rn."""
import numpy as np

F = np.float32
C = F(1.000002)


def _boxes(rng, n, scale):
    ctr = rng.uniform(0.0, 1.0, (n, 2))
    size = np.exp(rng.uniform(np.log(1e-3), np.log(0.9), (n, 2)))
    b = np.concatenate([ctr - size / 2, ctr + size / 2], -1) * scale
    return b.astype(F)


def _near_duplicates(rng, a):
    """Boxes a few ulps around `a`: the cases in which rounding decides."""
    jitter = rng.integers(-3, 4, a.shape).astype(np.int32)
    return (a.view(np.int32) + jitter).view(F)


def _pair_terms(a, g):
    yt, xl = np.maximum(a[:, 0], g[:, 0]), np.maximum(a[:, 1], g[:, 1])
    yb, xr = np.minimum(a[:, 2], g[:, 2]), np.minimum(a[:, 3], g[:, 3])
    ok = ~((yb < yt) | (xr < xl))
    inter = ((yb - yt).astype(F) * (xr - xl).astype(F)).astype(F)
    aa = ((a[:, 2] - a[:, 0]).astype(F) * (a[:, 3] - a[:, 1]).astype(F)).astype(F)
    ga = ((g[:, 2] - g[:, 0]).astype(F) * (g[:, 3] - g[:, 1]).astype(F)).astype(F)
    uni = ((aa + ga).astype(F) - inter).astype(F)
    return ok, inter, uni, aa, ga


def _cases(seed, n=200_000):
    rng = np.random.default_rng(seed)
    for scale in (1.0, 800.0):           # normalised and pixel coordinates
        a = _boxes(rng, n, scale)
        g = _boxes(rng, n, scale)
        near = rng.uniform(size=n) < 0.5
        g[near] = _near_duplicates(rng, a[near])     # IoU ~ 1 and exact ties
        shift = rng.uniform(size=n) < 0.25
        g[shift, 1] += (a[shift, 3] - a[shift, 1]) * F(0.5)   # IoU ~ 1/3 .. 1/2: around the thresholds
        g[shift, 3] += (a[shift, 3] - a[shift, 1]) * F(0.5)
        yield rng, a, g


def test_assign_product_test_never_skips_a_winner():
    for rng, a, g in _cases(1):
        ok, inter, uni, aa, ga = _pair_terms(a, g)
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = (inter / uni).astype(F)
        # running best: an earlier IoU of the same magnitude (ties and near ties are the dangerous cases)
        pick = rng.integers(0, 6, iou.shape)
        best = np.select([pick == 0, pick == 1, pick == 2, pick == 3, pick == 4],
                         [iou, np.nextafter(iou, F(0)), np.nextafter(iou, F(2)), (iou * F(1.00001)).astype(F),
                          (iou * F(2)).astype(F)], F(0)).astype(F)
        wins = ok & (iou > best)
        passes = ((inter * C).astype(F) > (best * uni).astype(F)) | (inter < F(1e-30))
        assert not np.any(wins & ~passes)
        assert wins.sum() > 1000 and (ok & ~passes).sum() > 1000      # both outcomes occur


def test_assign_area_ratio_bound_only_skips_losers():
    for rng, a, g in _cases(2):
        ok, inter, uni, aa, ga = _pair_terms(a, g)
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = np.where(ok, (inter / uni).astype(F), F(0))
        mn, mx = np.minimum(aa, ga), np.maximum(aa, ga)
        ratio = (mn / mx).astype(F)
        # running bests around the bound itself
        best = (ratio * rng.choice([F(0.5), F(0.999999), F(1.0), F(1.000001), F(1.000003), F(1.5)], ratio.shape)).astype(F)
        skipped = (best * mx).astype(F) > (mn * C).astype(F)
        assert not np.any(skipped & (iou > best))
        assert skipped.sum() > 1000 and (~skipped).sum() > 1000


def test_nms_band_test_equals_the_division():
    for rng, a, g in _cases(3):
        ok, inter, uni, aa, ga = _pair_terms(a, g)
        ih = (np.minimum(a[:, 2], g[:, 2]) - np.maximum(a[:, 0], g[:, 0])).astype(F)
        iw = (np.minimum(a[:, 3], g[:, 3]) - np.maximum(a[:, 1], g[:, 1])).astype(F)
        pos = (ih > 0) & (iw > 0) & (aa > 0) & (ga > 0)
        for thr in (F(0.0), F(0.3), F(0.5), F(1.0)):
            with np.errstate(divide='ignore', invalid='ignore'):
                want = pos & ((inter / uni).astype(F) > thr)          # TF's CPU kernel: IoU > thr
            rhs = (thr * uni).astype(F)
            sure_yes = (inter >= F(1e-30)) & (inter > (rhs * C).astype(F))
            sure_no = (inter >= F(1e-30)) & ~sure_yes & ((inter * C).astype(F) < rhs)
            got = np.where(sure_yes, True, np.where(sure_no, False, want)) & pos
            assert np.array_equal(got, want), thr
            assert not np.any(sure_yes & pos & ~want) and not np.any(sure_no & want)
            if 0 < thr < 1:
                assert (sure_yes & pos).sum() > 100 and (sure_no & pos).sum() > 100
