"""The pieces behind the reference's dense call shapes (GPU, through the C ABI):
  * rn_mask_scan / rn_mask_gather -- tf.boolean_mask per level + concat (utils.py:258-284) -- against torch's
    boolean indexing, bit-exact, for every row width the modules use;
  * the streaming loss kernel on DENSE float labels [rows, C] (losses.py's own signatures) against the generic
    kernel and the oracle."""
import ctypes as C

import numpy as np
import pytest

import helpers
import rn_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('width,dtype', [(1, 'uint8'), (3, 'float32'), (4, 'float32'), (20, 'float32'), (80, 'float32'),
                                         (4, 'int64')])
@pytest.mark.parametrize('keep', [0.0, 0.03, 0.5, 0.99, 1.0])
def test_mask_gather_equals_boolean_mask(width, dtype, keep):
    import torch
    from retinanet_tensorflow_b200 import _lib
    g = torch.Generator(device='cuda').manual_seed(int(1000 * keep) + width)
    seg_rows = [1, 1023, 1024, 1025, 5000, 37]
    tdt = getattr(torch, dtype)
    srcs, masks = [], []
    for r in seg_rows:
        if tdt.is_floating_point:
            srcs.append(torch.randn((r, width), device='cuda', generator=g).to(tdt))
        else:
            srcs.append(torch.randint(0, 200, (r, width), device='cuda', generator=g).to(tdt))
        masks.append(torch.rand((r,), device='cuda', generator=g) < keep)
    rows = _lib.rows_of(seg_rows)
    nbytes = int(_lib.lib.rn_mask_workspace_bytes(C.byref(rows)))
    ws = torch.empty((max(nbytes, 4),), dtype=torch.uint8, device='cuda')
    count = torch.zeros((1,), dtype=torch.int64, device='cuda')
    mm = _lib.maps_of([m.view(torch.uint8) for m in masks])
    _lib.call('rn_mask_scan', C.byref(rows), C.byref(mm), _lib.ptr(ws), nbytes, _lib.ptr(count), _lib.stream_ptr())
    want = torch.cat([s[m] for s, m in zip(srcs, masks)], 0)
    assert int(count.item()) == want.shape[0]
    out = torch.full((want.shape[0] + 1, width), 7, dtype=tdt, device='cuda')   # one guard row
    if want.shape[0]:
        _lib.call('rn_mask_gather', C.byref(rows), width * srcs[0].element_size(), C.byref(_lib.maps_of(srcs)), C.byref(mm),
                  _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out[:-1], want)
    assert bool((out[-1] == 7).all())


def _dense_case(rows, c, seed, soft=False, sigma=1.5):
    rng = np.random.default_rng(seed)
    z = np.zeros((rows, c), np.float32)
    pos = rng.random(rows) < 0.02
    z[pos, rng.integers(0, c, int(pos.sum()))] = 1.0
    if soft:   # not one-hot: several ones per row and fractional labels (BCE takes any z in [0, 1])
        extra = rng.random((rows, c)) < 0.002
        z[extra] = rng.choice([1.0, 0.25, 0.5], int(extra.sum())).astype(np.float32)
    x = rng.normal(-4.595, sigma, (rows, c)).astype(np.float32)
    x[z == 1.0] += 6.0
    return z, x


def _run_dense(z, x, mode_bits, tile):
    import torch
    from retinanet_tensorflow_b200 import _lib
    rows, c = x.shape
    xd, zd = torch.from_numpy(x).cuda(), torch.from_numpy(z).cuda()
    g = torch.full_like(xd, 7.0)
    stats = torch.zeros((_lib.RN_STATS_HEADER + 3 * c,), dtype=torch.float64, device='cuda')
    nb = _lib.lib.rn_loss_workspace_bytes(c)
    ws = torch.zeros((nb,), dtype=torch.uint8, device='cuda')
    out = torch.zeros((4,), dtype=torch.float32, device='cuda')
    seg = _lib.rows_of([rows])
    st = _lib.stream_ptr()
    run_bits = mode_bits | (0 if tile else _lib.RN_CLS_GENERIC_FLAG)   # per-call kernel choice
    _lib.call('rn_label_stats_dense', rows, c, _lib.ptr(zd), _lib.ptr(stats), st)
    if mode_bits & _lib.RN_CLS_DICE_FLAG:
        _lib.call('rn_cls_sums', C.byref(seg), c, C.byref(_lib.maps_of([xd])), C.byref(_lib.maps_of([zd])), 1,
                  _lib.RN_DTYPE_F32, _lib.ptr(stats), _lib.ptr(ws), nb, st)
    _lib.call('rn_cls_loss_fwd_bwd', C.byref(seg), c, C.byref(_lib.maps_of([xd])), C.byref(_lib.maps_of([zd])), 1,
              run_bits, 2.0, 0.25, 1e-7, C.byref(_lib.maps_of([g])), _lib.RN_DTYPE_F32, _lib.ptr(stats), _lib.ptr(ws),
              nb, st)
    _lib.call('rn_loss_finalize', c, mode_bits, _lib.ptr(stats), _lib.ptr(out), st)
    torch.cuda.synchronize()
    return out.cpu().numpy(), g.cpu().numpy()


@pytest.mark.parametrize('mode', ['focal', 'bce', 'balanced_bce', 'bce_dice', 'focal_dice'])
@pytest.mark.parametrize('rows,c', [(5000, 80), (777, 4), (20000, 20)])
def test_dense_label_streaming_kernel_matches_oracle_and_generic(rows, c, mode):
    from retinanet_tensorflow_b200.pipeline import CLS_MODES
    z, x = _dense_case(rows, c, seed=rows + c)
    bits = CLS_MODES[mode]
    out_t, g_t = _run_dense(z, x, bits, tile=True)
    out_g, g_g = _run_dense(z, x, bits, tile=False)
    fg = oracle.classmap_decode(z).fg_mask
    want = oracle.classification_loss(z, x, fg, mode=mode)
    g_ref = oracle.classification_loss_grad(z, x, fg, mode=mode)
    np.testing.assert_allclose(out_t[0], want, rtol=1e-5)
    np.testing.assert_allclose(out_g[0], want, rtol=1e-5)
    tol = dict(rtol=1e-5, atol=1e-6 * float(np.abs(g_ref).max()))
    helpers.assert_close(g_t, g_ref, what='streaming kernel, dense labels', **tol)
    helpers.assert_close(g_g, g_ref, what='generic kernel, dense labels', **tol)


def test_dense_soft_labels_bce():
    """BCE accepts labels that are not one-hot (several ones per row, fractional values): those units take the
    careful path of the streaming kernel."""
    from retinanet_tensorflow_b200.pipeline import CLS_MODES
    z, x = _dense_case(6000, 16, seed=5, soft=True)
    out_t, g_t = _run_dense(z, x, CLS_MODES['bce'], tile=True)
    fg = oracle.classmap_decode(z).fg_mask
    np.testing.assert_allclose(out_t[0], oracle.classification_loss(z, x, fg, mode='bce'), rtol=1e-5)
    g_ref = oracle.classification_loss_grad(z, x, fg, mode='bce')
    helpers.assert_close(g_t, g_ref, rtol=1e-5, atol=1e-6 * float(np.abs(g_ref).max()), what='soft labels')
