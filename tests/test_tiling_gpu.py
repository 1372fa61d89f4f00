"""Tiled inference (esr_b200.tiling): with an overlap that covers the receptive field of a shallow generator the stitched
output equals the one-pass eval-mode forward (CEM replicate padding at the image border included), with and without a latent map."""
import pytest
import torch

from esr_b200.tiling import _tile_ranges


def test_tile_ranges_cover_the_axis_once():
    for n, tile, ov in [(100, 32, 8), (64, 64, 16), (65, 64, 16), (7, 3, 5)]:
        r = _tile_ranges(n, tile, ov)
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))                       # cores tile the axis without gaps or overlaps
        assert all(hi - lo <= tile and ea <= lo and eb >= hi and ea >= 0 and eb <= n for lo, hi, ea, eb in r)
        assert all(ea == max(0, lo - ov) and eb == min(n, hi + ov) for lo, hi, ea, eb in r)


@pytest.mark.gpu
@pytest.mark.parametrize('z', [0, 3])
def test_tiled_forward_matches_one_pass(z):
    import models.modules.architecture as arch
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    from esr_b200 import ops, precision
    from esr_b200.tiling import tiled_forward
    ops.device_check()
    torch.manual_seed(5)
    S = 4
    net = arch.RRDBNet(3, 3, 32, 1, upscale=S, latent_input='all_layers,HR_downscaled' if z else None, num_latent_channels=z)
    wrapped = CEMnet(Get_CEM_Conf(S)).WrapArchitecture_PyTorch(net, None).to('cuda')
    for p in wrapped.parameters():
        p.requires_grad_(False)
    wrapped.eval()
    h, w = 88, 120
    lr = torch.rand(2, 3, h, w, device='cuda')
    z_hr = (2 * torch.rand(2, z, S * h, S * w, device='cuda') - 1) if z else None
    with precision.use('parity'), torch.no_grad():
        if z:
            x = torch.cat([z_hr.contiguous().view(2, z * S * S, h, w), lr], 1)
        else:
            x = lr
        ref = wrapped(x)
        # receptive field of this generator: 1 (fea) + 15 (one RRDB) + 1 (LR_conv) + up / HR convs (< 2) + CEM filters (<= 24) LR pixels
        out = tiled_forward(wrapped, lr, z_hr, tile=40, overlap=44)
    assert out.shape == ref.shape == (2, 3, S * h, S * w)
    err = (out - ref).abs().max().item()
    print('tiled vs one pass (z=%d): max abs diff %.2e' % (z, err))
    assert err < 1e-4
    with precision.use('parity'), torch.no_grad():
        coarse = tiled_forward(wrapped, lr, z_hr, tile=40, overlap=4)            # too little overlap: seams must show (the test has teeth)
    assert (coarse - ref).abs().max().item() > 10 * max(err, 1e-6)
