"""esr_b200.colors (the HSV conversion of the scribble tool; the reference takes it from skimage.color) against the standard library's
hexcone model, on [0,1] and on [0,255] inputs (the reference converts 255-scaled images, Z_optimization.py:419)."""
import colorsys

import numpy as np

from esr_b200.colors import hsv2rgb, rgb2hsv


def test_hsv_roundtrip_and_colorsys():
    rng = np.random.RandomState(0)
    x = rng.rand(31, 17, 3)
    x[0, 0] = 0.4                      # grey: hue and saturation 0
    x[0, 1] = (0.9, 0.9, 0.1)          # two channels tie for the maximum
    h = rgb2hsv(x)
    ref = np.array([colorsys.rgb_to_hsv(*p) for p in x.reshape(-1, 3)]).reshape(x.shape)
    assert np.abs(h - ref).max() < 1e-12
    assert np.abs(hsv2rgb(h) - x).max() < 1e-12
    h255 = rgb2hsv(255 * x)
    assert np.abs(h255[..., :2] - h[..., :2]).max() < 1e-12 and np.abs(h255[..., 2] - 255 * h[..., 2]).max() < 1e-9
    assert np.abs(hsv2rgb(h255) - 255 * x).max() < 1e-9
    assert tuple(rgb2hsv(np.zeros((2, 2, 3)))[0, 0]) == (0.0, 0.0, 0.0)
