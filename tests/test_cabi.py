"""CPU-side checks of the drop-in boundary: librndense.so loads without a GPU, exports every
symbol include/rndense.h declares, its HOST functions (pyramid description) agree bit-for-bit
with the oracle's float32 constants, argument validation returns the documented status codes,
and the product path refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rn_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'rndense.h')).read()
    return sorted(set(re.findall(r'RN_API\s+[\w\s\*]+?\b(rn_\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from retinanet_tensorflow_b200 import _lib
    names = _header_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(_lib.lib, name), 'librndense.so does not export %s' % name
    assert sorted(_lib.PROTOTYPES) == names, 'ctypes prototypes and header differ'
    assert _lib.lib.rn_version() == 100
    assert _lib.lib.rn_status_string(-3) == b'capacity exceeded'


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, 'include', 'rndense.h')).read()
    for cite in ('dataset.py:43-142', 'utils.py:62-97', 'utils.py:108-117', 'utils.py:183-195', 'utils.py:198-220',
                 'losses.py:144-152', 'losses.py:6-15', 'augmentation.py:5-22', 'levels.py:38-44'):
        assert cite in text, cite


@pytest.mark.parametrize('size,mode', [((500, 500), 'tf1_int_trunc'), ((800, 1333), 'tf1_int_trunc'),
                                        ((1024, 1024), 'float'), ((500, 667), 'float'), ((33, 17), 'tf1_int_trunc')])
def test_pyramid_constants_match_oracle_bit_for_bit(size, mode):
    from retinanet_tensorflow_b200._pyramid import pyramid_for
    from retinanet_tensorflow_b200.levels import build_levels
    levels = build_levels()
    pyr = pyramid_for(size, levels, mode)
    assert pyr.num_levels == 5
    off = 0
    for l, pn in enumerate(levels):
        lv = pyr.level[l]
        grid = oracle.grid_size_for(np.array(size, np.int32), 2 ** int(pn[-1]))
        assert (lv.grid_h, lv.grid_w) == (int(grid[0]), int(grid[1]))
        sizes = oracle.anchor_size_normalised(levels[pn].anchor_sizes, np.array(size, np.int32), mode)
        np.testing.assert_array_equal(np.array(lv.anchor_h[:9], np.float32), sizes[:, 0])
        np.testing.assert_array_equal(np.array(lv.anchor_w[:9], np.float32), sizes[:, 1])
        y, x = oracle._grid_axes(grid)
        got_y = (np.float32(lv.y0) + (np.float32(lv.y_step) * np.arange(lv.grid_h, dtype=np.float32))).astype(np.float32)
        got_x = (np.float32(lv.x0) + (np.float32(lv.x_step) * np.arange(lv.grid_w, dtype=np.float32))).astype(np.float32)
        np.testing.assert_array_equal(got_y, y)
        np.testing.assert_array_equal(got_x, x)
        assert lv.anchor_offset == off
        off += lv.grid_h * lv.grid_w * 9
    assert pyr.anchors_per_image == off
    if size == (800, 1333):
        assert off == 200700
    if size == (500, 500):
        assert off == 47961


def test_levels_mirror_matches_real_reference():
    from retinanet_tensorflow_b200.levels import Level, build_levels, compute_box_size
    ref = np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_levels.npz'))
    lv = build_levels()
    assert list(lv.keys()) == list(ref['names']) and lv.num_anchors == 9 and list(iter(lv)) == list(lv.keys())
    for pn in lv:
        np.testing.assert_array_equal(lv[pn].anchor_sizes, ref['sizes_' + pn])
    np.testing.assert_array_equal(compute_box_size(32, (1, 2), 1), ref['box_32_1x2_1'])
    np.testing.assert_array_equal(Level(32, [(1, 4)], [2 ** 0, 2 ** 1]).anchor_sizes, [[16, 64], [32, 128]])


def test_argument_validation_status_codes():
    from retinanet_tensorflow_b200 import _lib
    pyr = _lib.Pyramid()
    strides = np.array([8], np.int32)
    sizes = np.ones((1, 1, 2), np.float64)
    rc = _lib.lib.rn_pyramid_build(0, 10, 1, strides.ctypes.data_as(C.c_void_p), sizes.ctypes.data_as(C.c_void_p), 1,
                                   0, C.byref(pyr))
    assert rc == -2 and b'image size' in _lib.lib.rn_last_error()
    rc = _lib.lib.rn_pyramid_build(10, 10, 99, strides.ctypes.data_as(C.c_void_p),
                                   sizes.ctypes.data_as(C.c_void_p), 1, 0, C.byref(pyr))
    assert rc == -2
    rc = _lib.lib.rn_pyramid_build(10, 10, 1, strides.ctypes.data_as(C.c_void_p), sizes.ctypes.data_as(C.c_void_p),
                                   1, 7, C.byref(pyr))
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.call('rn_anchor_generate', None, None, None, None)
    # NULL inputs are rejected before any launch (no GPU needed to get the error)
    rows = _lib.rows_of([4])
    assert _lib.lib.rn_cls_loss_fwd_bwd(C.byref(rows), 3, None, None, 0, 1, 2.0, 0.25, 1e-7, None, 0, None, None, 0,
                                        None) == -1
    assert _lib.lib.rn_nms_classwise(1, 1, 8, None, None, None, None, 0.5, 10, 8, None, None, None, None, None, None,
                                     None, 0, None) == -1
    assert _lib.lib.rn_loss_workspace_bytes(80) >= 2048 * 160 * 8
    assert _lib.lib.rn_launch_count() >= 0


def test_rows_and_maps_helpers():
    from retinanet_tensorflow_b200 import _lib
    from retinanet_tensorflow_b200._pyramid import pyramid_for
    from retinanet_tensorflow_b200.levels import build_levels
    pyr = pyramid_for((500, 500), build_levels())
    rows = _lib.Rows()
    _lib.call('rn_rows_from_pyramid', C.byref(pyr), 3, C.byref(rows))
    assert rows.num_segments == 5 and rows.images == 3
    assert [rows.rows[i] for i in range(5)] == [3 * 9 * n * n for n in (63, 32, 16, 8, 4)]
    maps = _lib.Maps()
    _lib.call('rn_maps_from_merged', C.byref(pyr), 3, C.c_void_p(1 << 20), 16, C.byref(maps))
    assert maps.p[0] == 1 << 20 and maps.p[1] == (1 << 20) + 3 * 9 * 63 * 63 * 16


def test_new_entry_points_validate_their_arguments():
    """rn_assign_encode_patch / rn_detect_tail reject bad arguments before touching the device (status codes, no
    launch), and the capability query follows the documented limits."""
    from retinanet_tensorflow_b200 import _lib
    from retinanet_tensorflow_b200._pyramid import pyramid_for
    from retinanet_tensorflow_b200.levels import build_levels
    pyr = pyramid_for((500, 500), build_levels())
    lib = _lib.lib
    assert lib.rn_detect_tail_supported(80, 8192) == 1 and lib.rn_detect_tail_supported(80, 8193) == 0
    assert lib.rn_detect_tail_supported(0, 100) == 0 and lib.rn_detect_tail_supported(3, 1) == 1
    m = _lib.Maps()
    # NULL maps / arrays
    assert lib.rn_detect_tail(C.byref(pyr), 1, 3, C.byref(m), 1, 64, 1, 0, None, None, None, None, None, 0.5, 10, 64, None,
                              None, None, None, None, None, None, 0, None) == -1
    assert b'rn_detect_tail' in lib.rn_last_error()
    # patch shape: both zero or both in 1..64
    for sy, sx in ((0, 2), (2, 0), (-1, 1), (65, 1)):
        assert lib.rn_assign_encode_patch(C.byref(pyr), 1, None, None, None, 4, 3, 0.4, 0.5, C.byref(m), C.byref(m),
                                          C.byref(m), C.byref(m), None, None, sy, sx, None) == -1
        assert b'patch_sy' in lib.rn_last_error()


def test_pyramid_cache_keys_on_levels_size_and_mode():
    """The drop-in modules ask for the same rn_pyramid_t once per image: the host keeps the last few, keyed on the
    level objects (identity), the image size, the anchor-size mode, names and strides."""
    from retinanet_tensorflow_b200._pyramid import pyramid_for
    from retinanet_tensorflow_b200.levels import build_levels
    la, lb = build_levels(), build_levels()
    p1 = pyramid_for((500, 667), la)
    assert pyramid_for((500, 667), la) is p1
    assert pyramid_for([500, 667], la, names=list(la)) is p1
    assert pyramid_for((500, 667), lb) is not p1                      # other level objects
    assert pyramid_for((500, 668), la) is not p1                      # other image size
    pf = pyramid_for((500, 667), la, 'float')
    assert pf is not p1 and pf.level[0].anchor_h[0] != p1.level[0].anchor_h[0]
    one = pyramid_for((500, 667), {'L': la['P4']}, 'tf1_int_trunc', ['L'], [16])
    assert one.num_levels == 1 and one.level[0].grid_h == 32 and one is not p1
    assert pyramid_for((500, 667), {'L': la['P4']}, 'tf1_int_trunc', ['L'], [16]) is one   # same level object, new dict
    assert pyramid_for((500, 667), {'L': la['P4']}, 'tf1_int_trunc', ['L'], [8]) is not one


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('this check is for the GPU-less container')
    from retinanet_tensorflow_b200 import dataset, losses, utils
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        DenseHeadPath((500, 500), 3, 1, 4)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        utils.iou(np.zeros((1, 4), np.float32), np.zeros((1, 4), np.float32))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        dataset.build_labels([500, 500], [0], np.array([[0.1, 0.1, 0.5, 0.5]], np.float32), None, 3)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        losses.regression_loss([[3.]], [[1.]], [True])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'retinanet_tensorflow_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'rn_oracle' not in text and 'import oracle' not in text, f


def test_pyramid_constants_property():
    """rn_pyramid_build (host C) against the oracle for random image sizes and both anchor-size modes:
    grids, anchor sizes and the float32 linspace constants bit for bit (hypothesis, 60 examples)."""
    hyp = pytest.importorskip('hypothesis')
    from hypothesis import strategies as st
    from retinanet_tensorflow_b200._pyramid import pyramid_for
    from retinanet_tensorflow_b200.levels import build_levels
    levels = build_levels()

    @hyp.settings(max_examples=60, deadline=None, database=None)
    @hyp.given(st.integers(9, 2100), st.integers(9, 2100), st.sampled_from(['tf1_int_trunc', 'float']))
    def check(h, w, mode):
        pyr = pyramid_for((h, w), levels, mode)
        off = 0
        for l, pn in enumerate(levels):
            lv = pyr.level[l]
            grid = oracle.grid_size_for(np.array((h, w), np.int32), 2 ** int(pn[-1]))
            assert (lv.grid_h, lv.grid_w) == (int(grid[0]), int(grid[1]))
            sizes = oracle.anchor_size_normalised(levels[pn].anchor_sizes, np.array((h, w), np.int32), mode)
            np.testing.assert_array_equal(np.array(lv.anchor_h[:9], np.float32), sizes[:, 0])
            np.testing.assert_array_equal(np.array(lv.anchor_w[:9], np.float32), sizes[:, 1])
            y, x = oracle._grid_axes(grid)
            gy = (np.float32(lv.y0) + (np.float32(lv.y_step) * np.arange(lv.grid_h, dtype=np.float32))).astype(np.float32)
            gx = (np.float32(lv.x0) + (np.float32(lv.x_step) * np.arange(lv.grid_w, dtype=np.float32))).astype(np.float32)
            np.testing.assert_array_equal(gy, y)
            np.testing.assert_array_equal(gx, x)
            assert lv.anchor_offset == off
            off += lv.grid_h * lv.grid_w * 9
        assert pyr.anchors_per_image == off

    check()
