"""debug: capture every wgrad call of a training backward and check the kernel against torch on the SAME operands"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, 'explorable-super-resolution_b200'), REPO, os.path.join(REPO, 'tests')):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
from util import golden, mirror_rrdb, rel_err
from esr_b200 import ops
from CEM.CEMnet import CEMnet, Get_CEM_Conf

g = golden('wgrad_kinkfree_plain_train')
net = mirror_rrdb(g)
wrapped = CEMnet(Get_CEM_Conf(4)).WrapArchitecture_PyTorch(net, None).to('cuda')
wrapped.train()
calls = []
real = ops.conv3x3_wgrad
def spy(x16, gy16, cout, cin, **kw):
    dw, db = real(x16, gy16, cout, cin, **kw)
    lead, gy_off, scale = kw.get('lead', 0), kw.get('gy_off', 0), kw.get('scale', 1.0)
    cp = (cin + 7) // 8
    x = ops.unpack_planes(x16[:, :cp].contiguous(), cin).double()
    gy = ops.unpack_planes(gy16[:, gy_off:gy_off + (cout + 7) // 8].contiguous(), cout).double()
    with torch.enable_grad():
        wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, device='cuda', requires_grad=True)
        b = torch.zeros(cout, dtype=torch.float64, device='cuda', requires_grad=True)
        F.conv2d(x, wt, b, padding=1).backward(gy * scale)
    calls.append((cout, cin, gy_off, tuple(x16.shape), tuple(gy16.shape), rel_err(dw, wt.grad)[0], rel_err(db, b.grad)[0],
                  float(gy.abs().max()), float(gy.abs().mean())))
    return dw, db
ops.conv3x3_wgrad = spy
import esr_b200.engine as E
E.ops.conv3x3_wgrad = spy
x = torch.from_numpy(g['x']).cuda()
out = wrapped(x)
(out * torch.from_numpy(g['wt']).cuda()).sum().backward()
for c in calls:
    print('cout %3d cin %3d gy_off %2d x%s gy%s  kernel-vs-torch dw %.2e db %.2e   |gy| max %.3e mean %.3e' % c)
names = [n for n, _ in net.named_parameters()]
for name, p in net.named_parameters():
    ref = torch.from_numpy(g['g:' + name])
    e = rel_err(p.grad.cpu(), ref)
    print('%-45s max %.3e l2 %.3e   |ref| max %.3e' % (name, e[0], e[1], float(ref.abs().max())))
