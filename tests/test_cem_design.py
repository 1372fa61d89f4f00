"""Host-side CEM filter design (explorable-super-resolution_b200/CEM) against the reference's own numbers."""
import numpy as np
import pytest

from util import golden


@pytest.mark.parametrize('s', [2, 3, 4])
def test_filters_bit_identical_to_reference(s):
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    g = golden('cem_x%d' % s)
    cem = CEMnet(Get_CEM_Conf(s))
    assert cem.ds_kernel.shape == g['ds_kernel'].shape and np.array_equal(cem.ds_kernel, g['ds_kernel'])
    assert cem.inv_hTh.shape == g['inv_hTh'].shape and np.abs(cem.inv_hTh - g['inv_hTh']).max() < 1e-15
    assert [cem.invalidity_margins_LR, cem.invalidity_margins_HR, cem.ds_kernel_invalidity_half_size_LR,
            cem.inv_hTh_invalidity_half_size] == list(g['margins'])


def test_strides_convention():
    from CEM.imresize_CEM import calc_strides
    assert [list(map(int, calc_strides(None, s)[0])) for s in (2, 3, 4, 8)] == [[0, 0], [1, 1], [1, 1], [3, 3]]
    assert [list(map(int, calc_strides(None, s)[1])) for s in (2, 3, 4, 8)] == [[1, 1], [1, 1], [2, 2], [4, 4]]


def test_default_filters_are_rank_one():
    from CEM.CEMnet import CEMnet, Get_CEM_Conf, _separable_terms
    cem = CEMnet(Get_CEM_Conf(4))
    for k in (cem.ds_kernel, cem.inv_hTh):
        kv, kh = _separable_terms(k)
        assert kv.shape[0] == 1
        assert np.abs(kv.astype(np.float64).T @ kh.astype(np.float64) - k).max() < 1e-6 * np.abs(k).max()
    # a generic (estimated) kernel needs more terms and is still represented exactly
    rng = np.random.RandomState(0)
    k = rng.rand(9, 9)
    kv, kh = _separable_terms(k)
    assert kv.shape[0] > 1 and np.abs(kv.astype(np.float64).T @ kh.astype(np.float64) - k).max() < 1e-5


def test_numpy_projection_utilities_run():
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    cem = CEMnet(Get_CEM_Conf(4))
    rng = np.random.RandomState(1)
    lr, hr = rng.rand(12, 16, 3), rng.rand(48, 64, 3)
    out = cem.Enforce_DT_on_Image_Pair(lr, hr)
    assert out.shape == hr.shape
    from CEM.imresize_CEM import imresize
    assert np.abs(imresize(out, [1 / 4]) - lr)[3:-3, 3:-3].max() < 1e-4


@pytest.mark.parametrize('tag,s', [('x4', 4), ('x2', 2)])
def test_estimated_kernel_design_matches_reference(tag, s):
    """CEMnet(upscale_kernel=ndarray) - the externally estimated (non-separable, off-centre) down-scaling kernel of the GUI's
    KernelGAN route (GUI.py:1594-1603): centre-of-mass re-centring, energy cropping, hTh inversion with the 0.1 magnitude bound."""
    from CEM.CEMnet import CEMnet, Get_CEM_Conf, _separable_terms
    from CEM.imresize_CEM import imresize
    g = golden('cem_estimated_kernel')
    conf = Get_CEM_Conf(s)
    conf.lower_magnitude_bound = 0.1
    try:
        cem = CEMnet(conf, upscale_kernel=g[tag + ':kernel'])
        assert cem.ds_kernel.shape == g[tag + ':ds_kernel'].shape and np.abs(cem.ds_kernel - g[tag + ':ds_kernel']).max() < 1e-15
        assert cem.inv_hTh.shape == g[tag + ':inv_hTh'].shape and np.abs(cem.inv_hTh - g[tag + ':inv_hTh']).max() < 1e-12
        assert [cem.invalidity_margins_LR, cem.invalidity_margins_HR] == list(g[tag + ':margins'])
        # the CUDA filters take separable terms: this kernel needs several and is still represented to fp32 accuracy
        kv, kh = _separable_terms(cem.ds_kernel)
        assert kv.shape[0] > 1
        assert np.abs(kv.astype(np.float64).T @ kh.astype(np.float64) - cem.ds_kernel).max() < 1e-6 * np.abs(cem.ds_kernel).max()
        # the given kernel stays in force for later calls with this factor until it is reset (imresize_CEM.py:24-33)
        assert np.array_equal(CEMnet(conf).ds_kernel, cem.ds_kernel)
    finally:
        imresize(None, [s, s], return_upscale_kernel=True, kernel='reset_2_default')
    assert np.array_equal(CEMnet(Get_CEM_Conf(s)).ds_kernel, golden('cem_x%d' % s)['ds_kernel'])


def test_imresize_matches_reference_on_images():
    """CEM.imresize_CEM.imresize as an image resizer (data pipeline / GUI use) against the unmodified reference's outputs"""
    from CEM.imresize_CEM import imresize
    g = golden('imresize_cases')
    cases = [('down4', dict(scale_factor=1 / 4)), ('up4', dict(scale_factor=4)), ('down2', dict(scale_factor=[0.5])), ('up3', dict(scale_factor=3)),
             ('down3', dict(scale_factor=1 / 3)), ('down4_zero', dict(scale_factor=1 / 4, use_zero_padding=True)),
             ('up2_center', dict(scale_factor=2, align_center=True)), ('down2_center', dict(scale_factor=0.5, align_center=True)),
             ('up4_shape', dict(output_shape=[24, 28]))]
    for tag, kw in cases:
        out = imresize(g[tag + ':in'], **kw)
        assert out.shape == g[tag + ':out'].shape, tag
        assert np.abs(out - g[tag + ':out']).max() < 1e-12, (tag, np.abs(out - g[tag + ':out']).max())
