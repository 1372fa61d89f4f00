import os, sys
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(R, "explorable-super-resolution_b200")); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, numpy as np, math
import torch.nn.functional as F
from util import golden, golden_state_dict, mirror_rrdb, rel_err
from oracle import esr_oracle as O
from esr_b200 import ops
g, gw = golden('grad_rrdb_latent'), golden('rrdb_latent_x4')
sd = golden_state_dict(gw)
x = torch.from_numpy(g['x']); wt = torch.from_numpy(g['wt'])
# oracle with separate z_hr / z_lr leaves
lat, img = torch.split(x, [48, 3], dim=1)
z_hr = lat.reshape(1, 3, 64, 48).clone().requires_grad_(True)
z_lr_leaf = F.interpolate(z_hr.detach(), scale_factor=0.25, mode='bilinear', align_corners=False, recompute_scale_factor=False).requires_grad_(True)
def fwd(z_hr, z_lr, img):
    p='model.'; nf=32; nb=1; z=3
    xx = torch.cat([z_lr, img], 1)
    fea = O._conv(xx, sd, p+'0', False)
    out = torch.cat([z_lr, fea], 1)
    out = O.rrdb_forward(out, sd, p+'1.sub.0', nf, z)
    out = torch.cat([z_lr, out], 1)
    out = O._conv(out, sd, p+'1.sub.1', False)
    out = fea + out
    for idx in (2, 3):
        out = F.interpolate(out, scale_factor=2, mode='nearest'); out = O._conv(out, sd, p+'%d.1'%idx, True)
    out = torch.cat([z_hr, out], 1); out = O._conv(out, sd, p+'4', True)
    out = torch.cat([z_hr, out], 1); return O._conv(out, sd, p+'6', False)
img_l = img.clone().requires_grad_(True)
y = fwd(z_hr, z_lr_leaf, img_l); (y*wt).sum().backward()
print('oracle: |g z_hr| max', z_hr.grad.abs().max().item(), '|g z_lr| max', z_lr_leaf.grad.abs().max().item(), '|g img|', img_l.grad.abs().max().item())
net = mirror_rrdb(gw).cuda()
for p_ in net.parameters(): p_.requires_grad_(False)
eng = net.engine()
out, sv = eng.forward(x.cuda(), save=True)
print('fwd err', rel_err(out.cpu(), y.detach()))
# monkeypatch latent_grad to capture planes
cap = {}
orig = ops.latent_grad
def capt(gz_hr, gz_lr, *a):
    cap['hr'] = ops.unpack_planes(gz_hr, 3).cpu(); cap['lr'] = ops.unpack_planes(gz_lr, 3).cpu(); return orig(gz_hr, gz_lr, *a)
ops.latent_grad = capt
import esr_b200.engine as E; E.ops.latent_grad = capt
gx = eng.backward_input(wt.cuda(), sv).cpu()
print('g z_hr err', rel_err(cap['hr'], z_hr.grad))
print('g z_lr err', rel_err(cap['lr'], z_lr_leaf.grad))
print('g img err', rel_err(gx[:, 48:], img_l.grad))
d = (cap['lr'] - z_lr_leaf.grad).abs(); i = d.argmax(); print('lr worst idx', np.unravel_index(i.item(), d.shape), d.max().item())
d = (cap['hr'] - z_hr.grad).abs(); i = d.argmax(); print('hr worst idx', np.unravel_index(i.item(), d.shape), d.max().item())
ref = torch.from_numpy(g['gx'])
print('total', rel_err(gx, ref), 'z part', rel_err(gx[:, :48], ref[:, :48]))
