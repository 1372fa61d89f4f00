"""Parity of the row-streaming conv kernel (csrc/conv3x3_rows.cuh) against an fp64 torch convolution on fp16-exact
inputs (only the accumulation order differs: 1e-5), across the shapes that exercise its corner cases: strip borders,
image top/bottom, ranges that cross strips and images, accumulator-ring wrap, odd plane counts, several n-blocks."""
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


def _ops():
    from esr_b200 import ops
    ops.device_check()
    return ops


@pytest.mark.parametrize('n,cin,cout,h,w,dtype', [
    (1, 64, 32, 40, 128, torch.float16),      # one strip, exact width
    (1, 64, 32, 1, 128, torch.float16),       # single row image
    (1, 64, 32, 2, 128, torch.float16),
    (1, 64, 32, 3, 70, torch.float16),        # narrow strip, fewer rows than CTAs
    (2, 64, 32, 300, 256, torch.float16),     # ranges cross strips and images, ring wraps many times
    (1, 96, 32, 200, 130, torch.float16),     # second strip is 2 pixels wide
    (1, 160, 32, 150, 127, torch.float16),
    (1, 192, 64, 150, 256, torch.float16),    # two n-blocks of 32 (weights for N=192 do not fit)
    (1, 64, 64, 170, 128, torch.float16),     # N = 192 resident
    (3, 64, 64, 64, 200, torch.bfloat16),
    (1, 64, 3, 90, 256, torch.float16),       # n-block 16
    (1, 3, 64, 90, 256, torch.float16),       # one (padded) plane pair
    (1, 72, 32, 90, 128, torch.float16),      # 9 planes: odd tail chunk
    (1, 64, 256, 40, 128, torch.float16),     # 8 n-blocks
    (16, 64, 32, 256, 256, torch.float16),    # config-2 shape
])
def test_rows_conv_matches_fp64_reference(n, cin, cout, h, w, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(n * 1000 + cin + cout + h + w)
    x = torch.randn(n, cin, h, w, generator=g).to(dtype).float().to(DEV)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).to(dtype).float().to(DEV)
    b = (torch.randn(cout, generator=g) * 0.1).to(DEV)
    ref = F.leaky_relu(F.conv2d(x.double(), wt.double(), b.double(), padding=1), 0.2).float()
    x16, _ = ops.pack_nchw(x, dtype=dtype)
    pc = ops.PackedConv(wt, b, dtype=dtype)
    assert pc.wrows is not None
    out32 = torch.full((n, ops.planes_for(cout), h, w, 8), float('nan'), dtype=torch.float32, device=DEV)
    ops.conv3x3(x16, pc, lrelu=True, out32=out32, rows='force')
    got = ops.unpack_planes(out32, cout)
    assert not torch.isnan(got).any()
    emax, el2 = rel_err(got, ref)
    assert emax < 1e-5 and el2 < 1e-5, (emax, el2)


def test_rows_conv_equals_tile_kernel_bitwise_stores():
    """same epilogue code, same fp32 accumulators up to summation order: fused residual / dual store paths agree"""
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    n, cin, cout, h, w = 2, 192, 64, 77, 256
    x16, _ = ops.pack_nchw(torch.randn(n, cin, h, w, generator=g).to(DEV))
    pc = ops.PackedConv((torch.randn(cout, cin, 3, 3, generator=g) / 40).to(DEV), (torch.randn(cout, generator=g) * 0.1).to(DEV))
    _, r2 = ops.pack_nchw(torch.randn(n, cout, h, w, generator=g).to(DEV), want16=False, want32=True)
    outs = []
    for mode in ('force', False):
        o16 = torch.zeros((n, 24, h, w, 8), dtype=torch.float16, device=DEV)
        o32 = torch.zeros((n, 8, h, w, 8), dtype=torch.float32, device=DEV)
        ops.conv3x3(x16, pc, alpha=0.04, res1=x16, res1_off=3, beta1=0.2, res2=r2, beta2=1.0, out16=o16, out16_off=16, out32=o32, rows=mode)
        outs.append((o16, o32))
    assert rel_err(outs[0][1], outs[1][1])[0] < 1e-5
    assert float(outs[0][0][:, :16].abs().max()) == 0.0
    assert rel_err(outs[0][0].float(), outs[1][0].float())[0] < 2e-3


def test_rows_dgrad_operand():
    """transpose_flip image through the row kernel = gradient of the conv w.r.t. its input"""
    ops = _ops()
    g = torch.Generator().manual_seed(12)
    n, cin, cout, h, w = 1, 64, 32, 60, 128
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / 24).half().float().to(DEV)
    gy = torch.randn(n, cout, h, w, generator=g).half().float().to(DEV)
    ref = F.conv_transpose2d(gy.double(), wt.double(), padding=1).float()
    g16, _ = ops.pack_nchw(gy)
    pt = ops.PackedConv(wt, None, transpose_flip=True)
    out32 = torch.zeros((n, cin // 8, h, w, 8), dtype=torch.float32, device=DEV)
    ops.conv3x3(g16, pt, out32=out32, rows='force')
    assert rel_err(ops.unpack_planes(out32, cin), ref)[0] < 1e-5


@pytest.mark.parametrize('cin,cout,up2,lrelu', [(64, 32, False, True), (96, 32, False, False), (64, 64, False, True), (64, 64, True, True),
                                                 (192, 64, False, True)])
def test_rows_fast_epilogue_out16(cin, cout, up2, lrelu):
    """specialised epilogue 1 (bias + LeakyReLU -> 16-bit planes, plain / nearest-x2 store) == generic tile kernel, bit for bit
    up to the fp32 summation order (compared after rounding to fp16: at most 1 ulp apart)"""
    ops = _ops()
    g = torch.Generator().manual_seed(21 + cin + cout)
    n, h, w = 2, 45, 256
    x16, _ = ops.pack_nchw(torch.randn(n, cin, h, w, generator=g).to(DEV))
    pc = ops.PackedConv((torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).to(DEV), (torch.randn(cout, generator=g) * 0.1).to(DEV))
    f = 2 if up2 else 1
    outs = []
    for mode in ('force', False):
        o16 = torch.zeros((n, cout // 8 + 2, f * h, f * w, 8), dtype=torch.float16, device=DEV)
        ops.conv3x3(x16, pc, lrelu=lrelu, out16=o16, out16_off=1, up2=up2, rows=mode)
        outs.append(o16)
    a, b = outs[0].float(), outs[1].float()
    assert float(a[:, 0].abs().max()) == 0.0 and float(a[:, -1].abs().max()) == 0.0
    assert float(b.abs().max()) > 0.1
    assert rel_err(a, b)[0] < 2e-3 and rel_err(a, b)[1] < 2e-4


def test_rows_fast_epilogue_residuals():
    """specialised epilogue 2 (alpha*acc + beta1*res16 [+ beta2*res32] -> out16 [+ out32]) against the fp64 formula"""
    ops = _ops()
    g = torch.Generator().manual_seed(31)
    n, cin, cout, h, w = 1, 192, 64, 50, 256
    x = torch.randn(n, cin, h, w, generator=g).half().float().to(DEV)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / 40).half().float().to(DEV)
    b = (torch.randn(cout, generator=g) * 0.1).to(DEV)
    r2 = torch.randn(n, cout, h, w, generator=g).to(DEV)
    acc = F.conv2d(x.double(), wt.double(), b.double(), padding=1)
    x16, _ = ops.pack_nchw(x)
    _, r2p = ops.pack_nchw(r2, want16=False, want32=True)
    pc = ops.PackedConv(wt, b)
    # dense-block conv5: residual = the first 64 channels of the conv's own input
    o16 = torch.zeros((n, 8, h, w, 8), dtype=torch.float16, device=DEV)
    ops.conv3x3(x16, pc, alpha=0.2, res1=x16, res1_off=0, beta1=1.0, out16=o16, rows='force')
    ref = (0.2 * acc + x[:, :64].double()).float()
    assert rel_err(ops.unpack_planes(o16, cout), ref)[0] < 1e-3
    # third block of an RRDB: + fp32 trunk, dual store
    o32 = torch.zeros((n, 8, h, w, 8), dtype=torch.float32, device=DEV)
    ops.conv3x3(x16, pc, alpha=0.04, res1=x16, res1_off=0, beta1=0.2, res2=r2p, beta2=1.0, out16=o16, out32=o32, rows='force')
    ref = (0.04 * acc + 0.2 * x[:, :64].double() + r2.double()).float()
    got = ops.unpack_planes(o32, cout)
    assert rel_err(got, ref)[0] < 1e-5
    assert torch.equal(ops.unpack_planes(o16, cout), got.half().float())


def test_rows_fast_epilogue_shortcut_add_and_image_store():
    """LR_conv (+ fp32 ShortcutBlock residual, nearest-x2 store) and HR_conv1 (NCHW fp32 image) epilogues == tile kernel"""
    ops = _ops()
    g = torch.Generator().manual_seed(41)
    n, h, w = 2, 37, 256
    x16, _ = ops.pack_nchw(torch.randn(n, 64, h, w, generator=g).to(DEV))
    _, f32 = ops.pack_nchw(torch.randn(n, 64, h, w, generator=g).to(DEV), want16=False, want32=True)
    pc = ops.PackedConv((torch.randn(64, 64, 3, 3, generator=g) / 24).to(DEV), (torch.randn(64, generator=g) * 0.1).to(DEV))
    outs = []
    for mode in ('force', False):
        o16 = torch.zeros((n, 8, 2 * h, 2 * w, 8), dtype=torch.float16, device=DEV)
        ops.conv3x3(x16, pc, res1=f32, beta1=1.0, out16=o16, up2=True, rows=mode)
        outs.append(o16.float())
    assert rel_err(outs[0], outs[1])[0] < 2e-3 and rel_err(outs[0], outs[1])[1] < 2e-4
    p3 = ops.PackedConv((torch.randn(3, 64, 3, 3, generator=g) / 24).to(DEV), (torch.randn(3, generator=g) * 0.1).to(DEV))
    imgs = []
    for mode in ('force', False):
        o = torch.full((n, 3, h, w), float('nan'), device=DEV)
        ops.conv3x3(x16, p3, out_nchw=o, rows=mode)
        imgs.append(o)
    assert not torch.isnan(imgs[0]).any()
    assert rel_err(imgs[0], imgs[1])[0] < 1e-5


def test_rows_mask_only_epilogue_equals_generic():
    """gradient-slice launches of the dense-block backward (acc * LeakyReLU'(saved activation) -> 16-bit planes): the
    specialised epilogue against the generic one of the tile kernel"""
    ops = _ops()
    g = torch.Generator().manual_seed(51)
    n, cin, cout, h, w = 2, 96, 32, 41, 256
    x16, _ = ops.pack_nchw(torch.randn(n, cin, h, w, generator=g).to(DEV))
    act16, _ = ops.pack_nchw(torch.randn(n, cout, h, w, generator=g).to(DEV))
    pc = ops.PackedConv((torch.randn(cout, cin, 3, 3, generator=g) / 30).to(DEV), None)
    outs = []
    for mode in ('force', False):
        o16 = torch.zeros((n, 6, h, w, 8), dtype=torch.float16, device=DEV)
        ops.conv3x3(x16, pc, mask16=act16, mask_slope=0.2, out16=o16, out16_off=1, rows=mode)
        outs.append(o16.float())
    assert float(outs[0][:, 0].abs().max()) == 0.0 and float(outs[0][:, 5].abs().max()) == 0.0
    assert rel_err(outs[0], outs[1])[0] < 2e-3 and rel_err(outs[0], outs[1])[1] < 2e-4
