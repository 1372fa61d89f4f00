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


def Return_Filter_Energy_Distribution(filter):
    """fraction of the filter's L2 norm left after peeling 0, 1, 2, ... frames off its border (imresize_CEM.py:177-179)"""
    n = filter.shape[0]
    norms = [np.sqrt(np.sum(filter ** 2))] + [np.sqrt(np.sum(filter[f:-f, f:-f] ** 2)) for f in range(1, int(np.ceil(n / 2)))]
    return norms / norms[0]


def Center_Mass(kernel, ds_factor):
    """Zero-pads a square kernel so that its centre of mass sits at the centre, keeps it square, then crops equal frames holding
    less than 1 % of its norm such that (size - 1 [+1 for even factors]) is a multiple of ds_factor; unit sum (imresize_CEM.py:135-175)."""
    assert kernel.shape[0] == kernel.shape[1], 'Currently supporting only square kernels'
    n = kernel.shape[0]
    gx, gy = np.meshgrid(np.arange(n), np.arange(n))
    cx, cy = float(convolve2d(gx, kernel, mode='valid')[0, 0]) + 1, float(convolve2d(gy, kernel, mode='valid')[0, 0]) + 1
    x_pad, y_pad = 2 * (n / 2 - cx), 2 * (n / 2 - cy)
    diff = np.round(np.abs(y_pad)) - np.round(np.abs(x_pad))
    pads = {'x': [np.maximum(0, -x_pad), np.maximum(0, x_pad)], 'y': [np.maximum(0, -y_pad), np.maximum(0, y_pad)]}
    rnd = lambda v: int(np.round(v))

    def widen(pre, post, extra):      # split the padding the other axis needs, minding which side the rounding favoured
        lean_right = np.round(post) - post - (np.round(pre) - pre)
        pre, post = rnd(pre), rnd(post)
        big, small = int(np.ceil(extra / 2)), int(np.floor(extra / 2))
        return (pre + small, post + big) if lean_right > 0 else (pre + big, post + small)
    if diff > 0:
        pads['y'] = [rnd(pads['y'][0]), rnd(pads['y'][1])]
        pads['x'] = list(widen(pads['x'][0], pads['x'][1], diff))
    elif diff < 0:
        pads['x'] = [rnd(pads['x'][0]), rnd(pads['x'][1])]
        pads['y'] = list(widen(pads['y'][0], pads['y'][1], -diff))
    kernel = np.pad(kernel, ((rnd(pads['y'][0]), rnd(pads['y'][1])), (rnd(pads['x'][0]), rnd(pads['x'][1]))), mode='constant')
    assert kernel.shape[0] == kernel.shape[1], 'I caused the kernel to stop being a square...'
    margins = np.argwhere(Return_Filter_Energy_Distribution(kernel) < 0.99)[0][0] * np.ones([2]).astype(np.int32)
    side = 0
    while np.mod(kernel.shape[0] - np.sum(margins) - 1 + np.mod(ds_factor + 1, 2), ds_factor) != 0:
        margins[side] -= 1
        side = (side + 1) % 2
    kernel = kernel[margins[0]:-margins[1], margins[0]:-margins[1]]
    return kernel / np.sum(kernel)


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
    given = isinstance(kernel, np.ndarray)
    assert kernel is None or given or any(w in kernel for w in ['cubic', 'blurry_cubic', 'reset_2_default'])
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
    if given:
        # an externally estimated DOWN-scaling kernel (KernelGAN, GUI.py:1594-1603; imresize_CEM.py:23-33): stored as the up-scaling
        # kernel of this factor - rotated, re-centred on its centre of mass, cropped to 99 % of its energy - until 'reset_2_default'
        if str(s) in cache:
            print('Overriding previous kernel with given kernel...')
        assert np.abs(1 - np.sum(kernel)) < np.finfo(np.float32).eps, 'Supplied non-default kernel does not sum to 1'
        k = Center_Mass(np.rot90(kernel, 2), ds_factor=s) * s ** 2
        assert k.shape[0] == k.shape[1], 'Only square kernels supported for now'
        assert np.all(np.mod(k.shape + pad_after + pad_before - 1, s) == 0), 'Convolution-invalidated size should be an integer multiplication of sf_4_kernel'
        cache[str(s)] = k
    elif str(s) not in cache or kernel == 'reset_2_default':
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
