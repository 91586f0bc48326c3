"""Index arithmetic of the assignment kernel's work decomposition (csrc/rn_assign.cu), emulated in Python -- CPU only.

A CTA owns sy x sx sub-patches of 8 x 8 cells; thread `tid` of a sub-patch handles cell `tid // tpc` (tpc threads per
cell, 3 anchors each); a warp rejects GT boxes against the bounding box of its cells [r_lo..r_hi] x [x_lo..x_hi] and
skips a sub-patch when `sy0 + r_lo >= grid_h or sx0 + x_lo >= grid_w`.  Checked here: every cell of the grid is
handled exactly once, the warp bounds contain every cell of the warp, and the skip rule never drops a valid cell."""
import itertools

import pytest

TILE = 8


def _patches(grid_h, grid_w, sy, sx):
    tiles_x = -(-grid_w // (TILE * sx))
    tiles_y = -(-grid_h // (TILE * sy))
    for t in range(tiles_x * tiles_y):
        py = t // tiles_x
        yield py * TILE * sy, (t - py * tiles_x) * TILE * sx


@pytest.mark.parametrize('grid', [(1, 1), (7, 11), (13, 21), (25, 42), (63, 84), (100, 167)])
@pytest.mark.parametrize('patch', [(1, 1), (2, 2), (2, 1), (1, 3), (3, 2), (5, 7)])
@pytest.mark.parametrize('tpc', [1, 2, 3, 6])
def test_every_cell_once_and_warp_bounds_hold(grid, patch, tpc):
    grid_h, grid_w = grid
    sy, sx = patch
    threads = TILE * TILE * tpc
    seen = {}
    for ty0, tx0 in _patches(grid_h, grid_w, sy, sx):
        for warp in range(threads // 32):
            c_lo = (warp * 32) // tpc
            c_hi = min((warp * 32 + 31) // tpc, TILE * TILE - 1)
            r_lo, r_hi = c_lo // TILE, c_hi // TILE
            x_lo = c_lo % TILE if r_lo == r_hi else 0
            x_hi = c_hi % TILE if r_lo == r_hi else TILE - 1
            sy0 = ty0
            while sy0 < ty0 + sy * TILE:
                sx0 = tx0
                while sx0 < tx0 + sx * TILE:
                    skipped = sy0 + r_lo >= grid_h or sx0 + x_lo >= grid_w
                    for lane in range(32):
                        tid = warp * 32 + lane
                        cell = tid // tpc
                        cy, cx = sy0 + cell // TILE, sx0 + cell % TILE
                        ok = cy < grid_h and cx < grid_w
                        if ok:
                            assert not skipped, (grid, patch, tpc, warp, cy, cx)
                            assert r_lo <= cell // TILE <= r_hi and x_lo <= cell % TILE <= x_hi
                            seen[(cy, cx, tid % tpc)] = seen.get((cy, cx, tid % tpc), 0) + 1
                    sx0 += TILE
                sy0 += TILE
    want = {(y, x, s): 1 for y, x, s in itertools.product(range(grid_h), range(grid_w), range(tpc))}
    assert seen == want
