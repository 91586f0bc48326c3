"""the Shapes generator's host side reproduces the reference's `random` stream (CPU); the device
rasteriser and the box drawing are checked against OpenCV output of the REAL reference (GPU)."""
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NAMES = ['square', 'triangle', 'circle']


def golden():
    return np.load(os.path.join(HERE, 'golden', 'reference_shapes_images.npz'))


def test_host_specs_follow_the_reference_random_stream():
    """Same seed -> same background, shapes, colours and dimensions as data_loaders/shapes.py:155-176 (fixture
    written by the real module)."""
    from retinanet_tensorflow_b200 import synthetic_shapes as ss
    fx = golden()
    for seed in range(6):
        random.seed(100 + seed)
        bg, shapes = ss.random_image((200, 200))
        assert np.array_equal(bg, fx['bg_%d' % seed])
        assert len(shapes) == int(fx['n_%d' % seed])
        assert [NAMES.index(s[0]) for s in shapes] == fx['kinds_%d' % seed].tolist()
        assert np.array_equal(np.array([s[1] for s in shapes]).reshape(-1, 3), fx['colors_%d' % seed])
        assert np.array_equal(np.array([s[2] for s in shapes]).reshape(-1, 3), fx['dims_%d' % seed])


def test_reference_shapes_gt_fixture_is_reproduced():
    """tests/golden/reference_shapes.npz (GT boxes / classes of the real generator at 500x500, round 1) through the
    new host generator."""
    from retinanet_tensorflow_b200 import synthetic_shapes as ss
    fx = np.load(os.path.join(HERE, 'golden', 'reference_shapes.npz'))
    for seed in range(8):
        random.seed(seed)
        _, shapes = ss.random_image((500, 500))
        boxes = np.array([[y - s, x - s, y + s, x + s] for _, _, (x, y, s) in shapes], np.float64)
        assert np.array_equal(boxes, fx['boxes_%d' % seed])
        assert [NAMES.index(s[0]) for s in shapes] == fx['classes_%d' % seed].tolist()


@pytest.mark.gpu
def test_render_matches_opencv_images_of_the_reference():
    """Squares and circles pixel for pixel; triangles up to isolated edge pixels of cv2.fillPoly (< 0.6 % of the
    image, and only where the reference image has a triangle edge)."""
    import torch
    from retinanet_tensorflow_b200 import synthetic_shapes as ss
    fx = golden()
    gen = ss.Shapes(image_size=(200, 200))
    for seed in range(6):
        n = int(fx['n_%d' % seed])
        specs = np.zeros((1, ss.MAX_SHAPES, 8), np.int32)
        for j in range(n):
            x, y, s = fx['dims_%d' % seed][j]
            specs[0, j] = (fx['kinds_%d' % seed][j], x, y, s, *fx['colors_%d' % seed][j], 0)
        img = gen.render(specs, np.array([n], np.int32), fx['bg_%d' % seed].astype(np.uint8)[None])[0].cpu().numpy()
        ref = fx['image_%d' % seed]
        diff = (img != ref).any(-1)
        if 1 not in fx['kinds_%d' % seed].tolist():
            assert not diff.any(), 'seed %d: squares / circles must match OpenCV exactly' % seed
        else:
            assert diff.mean() < 0.006, (seed, diff.mean())
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_batch_feeds_the_dense_head_path():
    """Shapes.batch -> DenseHeadPath.assign: device GT tensors in the padded layout, images rendered on the device."""
    import torch
    from retinanet_tensorflow_b200 import synthetic_shapes as ss
    from retinanet_tensorflow_b200.pipeline import DenseHeadPath
    random.seed(5)
    gen = ss.Shapes(image_size=(500, 500))
    data = gen.batch(4)
    assert tuple(data['image'].shape) == (4, 500, 500, 3) and data['image'].dtype == torch.uint8
    path = DenseHeadPath((500, 500), gen.num_classes, 4, ss.MAX_SHAPES, process_group=False)
    path.assign(data['gt_boxes'], data['gt_classes'], data['gt_counts'])
    path.check_status()
    assert int((path.label >= 0).sum()) > 0


@pytest.mark.gpu
def test_draw_boxes_matches_opencv_rectangles():
    """utils.draw_bounding_boxes' rectangle (utils.py:125-129) on device detection lists vs cv2.rectangle."""
    import cv2
    import torch
    from retinanet_tensorflow_b200 import synthetic_shapes as ss
    rng = np.random.default_rng(0)
    B, K, H, W, C = 3, 7, 120, 160, 5
    boxes = np.zeros((B, K, 4), np.float32)
    ctr = rng.uniform(0.2, 0.8, (B, K, 2))
    sz = rng.uniform(0.05, 0.5, (B, K, 2))
    boxes[..., :2] = ctr - sz / 2
    boxes[..., 2:] = ctr + sz / 2
    classes = rng.integers(0, C, (B, K)).astype(np.int32)
    counts = np.array([7, 3, 0], np.int32)
    image = np.full((B, H, W, 3), 30, np.uint8)
    pal = np.clip(np.rint(ss.class_palette(C)), 0, 255).astype(np.uint8)
    want = image.copy()
    for b in range(B):
        px = (boxes[b] * np.array([H, W, H, W])).round().astype(np.int32)
        for k in range(counts[b]):
            want[b] = cv2.rectangle(np.ascontiguousarray(want[b]), (int(px[k, 1]), int(px[k, 0])), (int(px[k, 3]), int(px[k, 2])),
                                    tuple(int(v) for v in pal[classes[b, k]]), 1)
    got = ss.draw_bounding_boxes(torch.from_numpy(image).cuda(), torch.from_numpy(boxes).cuda(),
                                 torch.from_numpy(classes).cuda(), torch.from_numpy(counts).cuda(), num_classes=C).cpu().numpy()
    # overlapping outlines: the later box wins in OpenCV's sequential drawing; the device draws boxes concurrently,
    # so compare the SET of outlined pixels and the colour wherever only one box passes
    assert np.array_equal((got != 30).any(-1), (want != 30).any(-1))
    assert ((got != want).any(-1)).mean() < 0.01
