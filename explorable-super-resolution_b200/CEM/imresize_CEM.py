"""Integer-factor image resizing with the CEM's anti-aliasing kernels (numpy, init-time only).

Mirror of the reference's CEM/imresize_CEM.py surface (`imresize`, `calc_strides`, `Cubic_Kernel`):
same arguments, same conventions, same numbers — the taps produced here are what the CUDA CEM kernels
(csrc/aux_kernels.cuh) consume.  Conventions restated from the reference:
  * imresize_CEM.py:89-102  calc_strides: for scale s the sample of every s-cell sits at offset
    pre = s - floor(s/2) - 1, followed by post = floor(s/2) zeros (align_center derives it from the image);
  * imresize_CEM.py:104-110 Cubic_Kernel: 2-D impulse response of OpenCV's bicubic (Keys, a = -0.75)
    upscaling by s, trimmed to its support;
  * imresize_CEM.py:20-45   for even s the kernel is zero-padded by one row/col so that it stays centred;
    down-scaling uses the same kernel rotated by 180 degrees and divided by s^2.
The cubic taps are computed without OpenCV by following its float32 coefficient arithmetic (bit-identical
to the reference's cv2.resize result for s = 2, 3, 4, 8; pinned by tests/test_cem_design.py)."""
import numpy as np
from scipy.signal import convolve2d

def calc_strides(array, factor, align_center=False):
    s = int(np.maximum(factor, 1 / factor))
    if align_center:
        half = np.ceil(np.array(array.shape[:2]) / 2 * (factor if factor > 1 else 1))
        pre = np.mod(half, s)
        pre[pre == 0] = s
        pre = (pre - 1).astype(np.int32)
        post = (s - pre - 1).astype(np.int32)
    else:
        post = (np.floor(s / 2) * np.ones([2])).astype(np.int32)
        pre = (s - post - 1).astype(np.int32)
    return pre, post


def _cv2_cubic_coeffs(x):
    """The four Keys (a = -0.75) taps for fractional offset x, evaluated in float32 exactly as OpenCV's
    interpolateCubic does (its coefficient type is float even for float64 images)."""
    f = np.float32
    a, one = f(-0.75), f(1)
    c0 = ((a * (x + one) - f(5) * a) * (x + one) + f(8) * a) * (x + one) - f(4) * a
    c1 = ((a + f(2)) * x - (a + f(3))) * x * x + one
    c2 = ((a + f(2)) * (one - x) - (a + f(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return (c0, c1, c2, c3)


def Cubic_Kernel(sf):
    """Impulse response of bicubic x`sf` up-scaling, trimmed to its support: what
    cv2.resize(delta_11x11, INTER_CUBIC) returns in the reference (imresize_CEM.py:104-110), computed
    here without OpenCV by following its arithmetic (float32 source coordinate and taps)."""
    sf = int(sf)
    n, c = 11, 5  # delta image size / position used by the reference
    row = np.zeros(sf * n, dtype=np.float64)
    for dx in range(sf * n):
        fx = np.float32((dx + 0.5) * (1.0 / sf) - 0.5)
        sx = int(np.floor(fx))
        k = c - sx + 1  # which of the 4 taps lands on the delta
        if 0 <= k <= 3:
            row[dx] = float(_cv2_cubic_coeffs(np.float32(fx - np.float32(sx)))[k])
    nz = np.nonzero(row)[0]
    taps = row[nz[0]:nz[-1] + 1]
    return np.outer(taps, taps)


def _default_kernel(sf, blur_sigma=None):
    k = Cubic_Kernel(sf)
    if blur_sigma is not None:
        from scipy.signal.windows import gaussian
        from scipy.stats import norm
        size = int(1 + 2 * np.ceil(-1 * norm.ppf(0.005, scale=blur_sigma)))
        g = gaussian(size, blur_sigma).reshape([1, size]) * gaussian(size, blur_sigma).reshape([size, 1])
        k = convolve2d(k, g / np.sum(g))
    return k


def imresize(im, scale_factor=None, output_shape=None, kernel=None, align_center=False, return_upscale_kernel=False,
             use_zero_padding=False, antialiasing=True, kernel_shift_flag=False):
    """Same contract as the reference's imresize (imresize_CEM.py:8-87): integer up/down factors only,
    kernels cached per factor on the function object, edge (replicate) padding unless use_zero_padding."""
    if isinstance(kernel, np.ndarray):
        raise NotImplementedError('externally estimated (non-default) kernels are not wired yet (SURVEY 8f-4)')
    assert kernel is None or any(w in kernel for w in ['cubic', 'blurry_cubic', 'reset_2_default'])
    cache = imresize.__dict__.setdefault('kernels', {})
    if scale_factor is None:
        scale_factor = [output_shape[0] / im.shape[0]]
    elif not isinstance(scale_factor, list):
        scale_factor = [scale_factor]
    f = scale_factor[0]
    assert np.round(f) == f or np.round(1 / f) == 1 / f, 'Only supporting integer downsampling or upsampling rates'
    assert len(scale_factor) == 1 or scale_factor[0] == scale_factor[1]
    s = int(np.maximum(f, 1 / f))
    pre, post = calc_strides(im, f, align_center)
    pad_after = np.maximum(0, pre - post)
    pad_before = np.maximum(0, post - pre)
    if str(s) not in cache or kernel == 'reset_2_default':
        sigma = float(kernel[len('blurry_cubic_'):]) if (kernel is not None and 'blurry_cubic' in kernel) else None
        cache[str(s)] = _default_kernel(s, sigma)
    aa = np.pad(cache[str(s)], ((pad_before[0], pad_after[0]), (pad_before[1], pad_after[1])), mode='constant')
    if f < 1:
        aa = np.rot90(aa * f ** 2, 2)
    if return_upscale_kernel:
        return aa
    assert output_shape is None or np.all(f * np.array(im.shape[:2]) == output_shape[:2])
    half = np.floor(np.array(aa.shape) / 2).astype(np.int32)
    target = f * np.array(im.shape[:2])
    assert np.all(target == np.round(target)), 'Seems like an attempt to downscale with a factor inducing a non-integer image size'
    target = target.astype(np.int32)
    if im.ndim < 3:
        im = np.expand_dims(im, -1)

    def filt(x):
        if use_zero_padding:
            return convolve2d(x, aa, 'same')
        return convolve2d(np.pad(x, ((half[0], half[0]), (half[1], half[1])), mode='edge'), aa, 'valid')

    chans = []
    for c in range(im.shape[2]):
        if f > 1:
            stuffed = np.zeros(target, dtype=np.float64)
            stuffed[pre[0]::s, pre[1]::s] = im[:, :, c]
            chans.append(filt(stuffed))
        else:
            chans.append(filt(im[:, :, c])[pre[0]::s, pre[1]::s])
    return np.squeeze(np.stack(chans, -1))
