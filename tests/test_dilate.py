"""Z_optimization._dilate_rect: the rectangle dilation of the non-local editing mask (the reference calls cv2.dilate(image_mask, np.ones([16, 16])),
Z_optimization.py:358; cv2 is not a dependency here).  OpenCV's definition: dst(x, y) = max over the element of src(x + x' - anchor, y + y' - anchor),
anchor = k // 2, pixels outside the image ignored."""
import numpy as np

from Z_optimization import _dilate_rect


def test_single_pixel_becomes_the_anchored_rectangle():
    m = np.zeros((40, 50), dtype=np.float32)
    m[20, 25] = 1
    d = _dilate_rect(m, 16)
    ys, xs = np.nonzero(d)
    assert (ys.min(), ys.max(), xs.min(), xs.max()) == (20 - 7, 20 + 8, 25 - 7, 25 + 8)     # source offsets -8..+7 seen from the destination
    assert d.dtype == m.dtype and set(np.unique(d)) == {0.0, 1.0}


def test_borders_and_grey_values():
    m = np.zeros((12, 12), dtype=np.float32)
    m[0, 0], m[11, 11], m[5, 5] = 0.25, 0.5, 1.0
    d = _dilate_rect(m, 3)
    ref = np.zeros_like(m)
    for y in range(12):
        for x in range(12):
            ref[y, x] = max(m[yy, xx] for yy in range(max(0, y - 1), min(12, y + 2)) for xx in range(max(0, x - 1), min(12, x + 2)))
    assert np.array_equal(d, ref)
