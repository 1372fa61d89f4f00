"""Weight / bias gradient kernel (csrc/conv3x3_wgrad.cuh) against torch's fp64 autograd on fp16-exact operands
(only the accumulation order differs: 1e-5 of the gradient's range)."""
import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _watchdog():
    yield
    from esr_b200 import lib
    wd = lib.watchdog()
    assert wd[0] == 0, 'pipeline watchdog fired: %r' % (wd,)


@pytest.mark.parametrize('n,cin,cout,h,w,lead,dtype', [
    (1, 64, 32, 20, 64, 0, torch.float16),
    (1, 64, 32, 1, 64, 0, torch.float16),      # single row
    (2, 96, 32, 33, 52, 0, torch.float16),     # config-3 width: partial K step
    (1, 192, 64, 40, 128, 0, torch.float16),   # two strips, two n-blocks, 5 M chunks
    (1, 64, 64, 30, 70, 0, torch.float16),     # N = 192, second strip 6 pixels wide
    (3, 160, 32, 70, 200, 0, torch.bfloat16),  # ranges cross images and strips
    (1, 64, 3, 25, 64, 0, torch.float16),      # image conv: padded gradient plane
    (1, 3, 64, 25, 64, 0, torch.float16),
    (1, 67, 32, 20, 84, 3, torch.float16),     # latent channels in their own plane group
    (1, 64, 256, 16, 64, 0, torch.float16),    # 8 n-blocks
    (1, 256, 128, 20, 64, 0, torch.float16),   # wide conv: two input-channel slices
    (1, 512, 512, 8, 8, 0, torch.float16),     # VGG-size conv: three slices, 16 n-blocks
])
def test_wgrad_matches_autograd(n, cin, cout, h, w, lead, dtype):
    from esr_b200 import ops
    ops.device_check()
    g = torch.Generator().manual_seed(cin * 7 + cout + h + w)
    x = torch.randn(n, cin, h, w, generator=g).to(dtype).double().to(DEV)
    gy = torch.randn(n, cout, h, w, generator=g).to(dtype).double().to(DEV)
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, device=DEV, requires_grad=True)
    b = torch.zeros(cout, dtype=torch.float64, device=DEV, requires_grad=True)
    F.conv2d(x, wt, b, padding=1).backward(gy)
    if lead:
        x16 = torch.zeros((n, 1 + ops.planes_for(cin - lead), h, w, 8), dtype=dtype, device=DEV)
        ops.pack_nchw(x[:, :lead].float(), dst16=x16, plane_off=0)
        ops.pack_nchw(x[:, lead:].float(), dst16=x16, plane_off=1)
    else:
        x16, _ = ops.pack_nchw(x.float(), dtype=dtype)
    gy16, _ = ops.pack_nchw(gy.float(), dtype=dtype)
    dw, db = ops.conv3x3_wgrad(x16, gy16, cout, cin, lead=lead)
    assert rel_err(dw, wt.grad)[0] < 1e-5, rel_err(dw, wt.grad)
    assert rel_err(db, b.grad)[0] < 1e-5
    # accumulate + scale
    dw2, db2 = ops.conv3x3_wgrad(x16, gy16, cout, cin, lead=lead, dw=dw.clone(), db=db.clone(), scale=0.5, accumulate=True)
    assert rel_err(dw2, 1.5 * wt.grad)[0] < 1e-5
    assert rel_err(db2, 1.5 * b.grad)[0] < 1e-5
