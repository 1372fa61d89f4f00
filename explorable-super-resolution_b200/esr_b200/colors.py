"""RGB <-> HSV on numpy arrays, for the scribble tool of Z_optimizer (the reference takes them from skimage.color, Z_optimization.py:4,419-421:
brightening / darkening scribbles scale the V channel of the current output).  The standard hexcone model; arrays [..., 3], any float range
(V = max, S = (max - min) / max, H in [0, 1))."""
import numpy as np


def rgb2hsv(rgb):
    arr = np.asarray(rgb, dtype=np.float64)
    v = arr.max(-1)
    delta = v - arr.min(-1)
    with np.errstate(invalid='ignore', divide='ignore'):
        s = np.where(delta == 0, 0.0, delta / v)
        r, g, b = arr[..., 0], arr[..., 1], arr[..., 2]
        h = np.zeros_like(v)
        # the same precedence as a sequence of masked assignments red, green, blue: the LAST matching channel wins
        h = np.where(r == v, (g - b) / delta, h)
        h = np.where(g == v, 2.0 + (b - r) / delta, h)
        h = np.where(b == v, 4.0 + (r - g) / delta, h)
        h = (h / 6.0) % 1.0
    h = np.where(delta == 0, 0.0, h)
    out = np.stack([h, s, v], -1)
    out[np.isnan(out)] = 0
    return out


def hsv2rgb(hsv):
    arr = np.asarray(hsv, dtype=np.float64)
    h, s, v = arr[..., 0], arr[..., 1], arr[..., 2]
    hi = np.floor(h * 6)
    f = h * 6 - hi
    p = v * (1 - s)
    q = v * (1 - f * s)
    t = v * (1 - (1 - f) * s)
    sector = hi.astype(np.int64) % 6
    table = [(v, t, p), (q, v, p), (p, v, t), (p, q, v), (t, p, v), (v, p, q)]
    out = np.zeros(arr.shape, dtype=np.float64)
    for k, (r_, g_, b_) in enumerate(table):
        m = sector == k
        out[..., 0] = np.where(m, r_, out[..., 0])
        out[..., 1] = np.where(m, g_, out[..., 1])
        out[..., 2] = np.where(m, b_, out[..., 2])
    return out
