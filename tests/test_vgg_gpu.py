"""VGG19 feature extractor (perceptual loss network) against the same layer stack evaluated by PyTorch in fp32
(`torchvision.models.vgg19().features[:35]` is exactly this nn.Sequential; built by hand so the test does not need
torchvision).  Weights: torchvision's own initialisation (no pretrained weights offline), rounded to fp16.

Tolerances: fp16 engine (the default; its backward is loss-scaled) - features rel-L2 <= 3e-3 (16 fused conv stages on
10-bit-mantissa operands); bf16 engine - rel-L2 <= 3e-2.  Input gradients pass through 15 ReLUs and 4 poolings whose
decisions can flip under rounding (random weights put many pre-activations near 0), so they are held to
cosine >= 0.99 and rel-L2 <= 0.15 (fp16; measured 0.995 / 0.10) and cosine >= 0.9 (bf16; measured 0.94), also for a
gradient as small as a mean-reduced loss produces (1e-7).  The pooling kernels themselves are bit-exact."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    from esr_b200 import lib
    wd = lib.watchdog()
    assert wd[0] == 0, 'pipeline watchdog fired: %r' % (wd,)


def _netF(dtype):
    import models.modules.architecture as arch
    torch.manual_seed(0)
    net = arch.VGGFeatureExtractor(feature_layer=34, use_bn=False, use_input_norm=True, arch_config='untrained').to(DEV)
    with torch.no_grad():
        for p in net.features.parameters():
            p.copy_(p.half().float())
        for m in net.features:
            if isinstance(m, nn.Conv2d):
                m.bias.normal_(0, 0.05)
                m.bias.copy_(m.bias.half().float())
    net.compute_dtype = dtype
    return net


def _reference(net, x):
    """the nn.Sequential itself, fp32"""
    return net.features((x - net.mean) / net.std)


def test_state_dict_keys_match_torchvision_layout():
    net = _netF(torch.float16)
    keys = [k for k in net.state_dict().keys() if k.startswith('features')]
    convs = [0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28, 30, 32, 34]
    assert keys == [('features.%d.%s' % (i, s)) for i in convs for s in ('weight', 'bias')]
    assert all(not p.requires_grad for p in net.features.parameters())
    assert net.features[34].weight.shape == (512, 512, 3, 3) and isinstance(net.features[4], nn.MaxPool2d)


@pytest.mark.parametrize('dtype,tol', [(torch.float16, 3e-3), (torch.bfloat16, 3e-2)])
def test_features_match_fp32_reference(dtype, tol):
    from esr_b200 import ops
    ops.device_check()
    net = _netF(dtype)
    x = torch.rand(2, 3, 64, 96, device=DEV)
    with torch.no_grad():
        got = net(x)
        ref = _reference(net, x)
    assert got.shape == ref.shape == (2, 512, 4, 6)
    emax, el2 = rel_err(got, ref)
    assert el2 < tol and emax < 4 * tol, (emax, el2)


@pytest.mark.parametrize('dtype,cos_min,l2_max,gscale', [(torch.float16, 0.99, 0.15, 1.0), (torch.float16, 0.99, 0.15, 1e-7),
                                                          (torch.bfloat16, 0.9, 0.5, 1.0)])
def test_input_gradient_matches_autograd(dtype, cos_min, l2_max, gscale):
    from esr_b200 import ops
    ops.device_check()
    net = _netF(dtype)
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(2, 3, 64, 64, generator=g).to(DEV)
    wt = torch.randn(2, 512, 4, 4, generator=g).to(DEV) * gscale
    x = x0.clone().requires_grad_(True)
    (net(x) * wt).sum().backward()
    xr = x0.clone().requires_grad_(True)
    (_reference(net, xr) * wt).sum().backward()
    cos = F.cosine_similarity(x.grad.flatten().double(), xr.grad.flatten().double(), dim=0).item()
    emax, el2 = rel_err(x.grad, xr.grad)
    print('vgg input gradient %s: cosine %.4f rel-L2 %.3e' % (dtype, cos, el2))
    assert cos > cos_min and el2 < l2_max, (cos, el2)


def test_pooling_kernels_bit_exact():
    from esr_b200 import ops
    ops.device_check()
    g = torch.Generator().manual_seed(4)
    a = F.relu(torch.randn(2, 24, 12, 20, generator=g)).half().float().to(DEV).requires_grad_(True)
    p16, _ = ops.pack_nchw(a.detach())
    pooled = ops.maxpool2x2(p16)
    ref = F.max_pool2d(a, 2, 2)
    assert torch.equal(ops.unpack_planes(pooled, 24), ref.detach())
    go = torch.randn(2, 24, 6, 10, generator=g).half().float().to(DEV)
    ref.backward(go)
    g16, _ = ops.pack_nchw(go)
    gin = ops.unpack_planes(ops.maxpool2x2_bwd(g16, p16), 24)
    # fused with the ReLU derivative: positions whose activation is 0 get no gradient (torch would give the first zero of an
    # all-zero window the gradient, and the ReLU behind it kills it)
    assert torch.equal(gin, a.grad * (a.detach() > 0))


def test_define_F_and_detached_real_features():
    from models import networks

    class ND(dict):
        def __missing__(self, k):
            return None
    netF = networks.define_F(ND(gpu_ids=[0]), use_bn=False, arch_config='untrained')
    assert type(netF).__name__ == 'SingleDeviceDataParallel' and not netF.training
    x = torch.rand(1, 3, 32, 32, device=DEV)
    f = netF(x)
    assert f.shape == (1, 512, 2, 2) and not f.requires_grad
