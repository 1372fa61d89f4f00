import os, sys
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(R, "explorable-super-resolution_b200")); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, numpy as np
import torch.nn.functional as F
from util import golden, golden_state_dict, mirror_rrdb, rel_err
from oracle import esr_oracle as O
from esr_b200 import ops
import esr_b200.engine as E
g, gw = golden('grad_rrdb_latent'), golden('rrdb_latent_x4')
sd = golden_state_dict(gw)
x = torch.from_numpy(g['x']); wt = torch.from_numpy(g['wt'])
lat, img = torch.split(x, [48, 3], dim=1); img = img.clone().requires_grad_(True)
z_hr = lat.reshape(1, 3, 64, 48)
z_lr = F.interpolate(z_hr, scale_factor=0.25, mode='bilinear', align_corners=False, recompute_scale_factor=False)
p='model.'; nf=32; z=3
keep = {}
def K(name, t): t.retain_grad(); keep[name] = t; return t
xx = torch.cat([z_lr, img], 1)
fea = K('fea', O._conv(xx, sd, p+'0', False))
cur = fea
pre = p+'1.sub.0'
rrdb_in = cur
for j, rn in enumerate(('RDB1','RDB2','RDB3')):
    xin = K('rdb%d_in'%j, cur * 1.0)
    outs = [torch.cat([z_lr, xin], 1)]
    for i in range(5):
        o = O._conv(torch.cat(outs, 1), sd, '%s.%s.convs.%d.0' % (pre, rn, i), i < 4)
        outs.append(K('rdb%d_x%d'%(j,i+1), o))
    cur = outs[-1]*0.2 + xin
rr_out = K('rrdb_out', cur*0.2 + rrdb_in)
t = K('t', O._conv(torch.cat([z_lr, rr_out],1), sd, p+'1.sub.1', False) + fea)
out = t
for idx in (2,3):
    out = F.interpolate(out, scale_factor=2, mode='nearest'); out = K('up%d'%idx, O._conv(out, sd, p+'%d.1'%idx, True))
out = K('hr0', O._conv(torch.cat([z_hr, out],1), sd, p+'4', True))
y = O._conv(torch.cat([z_hr, out],1), sd, p+'6', False)
(y*wt).sum().backward()
net = mirror_rrdb(gw).cuda()
for p_ in net.parameters(): p_.requires_grad_(False)
eng = net.engine()
_, sv = eng.forward(x.cuda(), save=True)
calls = []
real = ops.conv3x3
def rec(x16, pc, **kw):
    real(x16, pc, **kw)
    calls.append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in kw.items() if k in ('out32','out16','tail_first')})
E.ops.conv3x3 = rec
ds = []
real_ds = ops.downsum2x
def rec_ds(*a, **k):
    r = real_ds(*a, **k); ds.append(r); return r
E.ops.downsum2x = rec_ds
gx = eng.backward_input(wt.cuda(), sv)
def m(t): return torch.where(t > 0, torch.ones_like(t), torch.full_like(t, 0.2))
mine_hr0 = ops.unpack_planes(sv.B['hr_b'], 32, plane_off=1).cpu(); orc = keep['hr0'].detach()
mis = (mine_hr0 > 0) != (orc > 0)
print('sign mismatches in saved HR0 activation:', int(mis.sum()), 'of', mis.numel(), 'max |oracle act| among them: %.3e' % (orc[mis].abs().max().item() if mis.any() else 0), 'act range %.3f' % orc.abs().max().item())
gm = ops.unpack_planes(calls[0]['out16'], 32).cpu(); go = keep['hr0'].grad * m(orc)
d = (gm - go).abs(); big = d > 0.02 * go.abs().max()
print('big-error elements:', int(big.sum()), 'of which sign-mismatch:', int((big & mis).sum()))
idx = big.nonzero()[:6]
for t in idx: print('  at', t.tolist(), 'mine %.4f oracle %.4f act_mine %.5f act_orc %.5f' % (gm[tuple(t)], go[tuple(t)], mine_hr0[tuple(t)], orc[tuple(t)]))
print('g pre-act HR0 :', 'max %.2e l2 %.2e' % rel_err(ops.unpack_planes(calls[0]['out16'], 32).cpu(), keep['hr0'].grad * m(keep['hr0'].detach())))
print('g pre-act up3 :', 'max %.2e l2 %.2e' % rel_err(ops.unpack_planes(calls[1]['out16'], 32).cpu(), keep['up3'].grad * m(keep['up3'].detach())))
print('g pre-act up2 :', 'max %.2e l2 %.2e' % rel_err(ops.unpack_planes(ds[0][1], 32).cpu(), keep['up2'].grad * m(keep['up2'].detach())))
print('g t           :', 'max %.2e l2 %.2e' % rel_err(ops.unpack_planes(ds[1][0], 32).cpu(), keep['t'].grad))
def cmp(name, mine, c, off=0):
    print('%-12s' % name, 'max %.2e l2 %.2e' % rel_err(ops.unpack_planes(mine, c, plane_off=off).cpu(), keep[name].grad))
# call order: 0 HR1^T, 1 HR0^T, 2 up1^T, 3 up0^T, 4 LR^T, then RDB3: conv5^T,4,3,2,1 ; RDB2..; RDB1..
cmp('hr0', calls[0]['out16'], 32)          # masked grad wrt hr0 pre-activation?  (grad wrt hr0 output * mask) -- compare loosely
cmp('rrdb_out', calls[4]['out32'], 32)
b = 5
for j in (2,1,0):
    # after conv5^T: G16 planes 20.. hold masked grad of x4 ; gS holds unmasked
    gS5 = calls[b]['out32']
    print('rdb%d' % j, 'gS after conv5^T (x4 slice, unmasked) vs oracle grad x4:', 'max %.2e l2 %.2e' % rel_err(ops.unpack_planes(gS5, 32, plane_off=4+12).cpu(), keep['rdb%d_x4'%j].grad))
    for ii, i in enumerate((3,2,1)):
        gSi = calls[b+1+ii]['out32']
        print('   after conv%d^T: x%d slice' % (i+1, i), 'max %.2e l2 %.2e' % rel_err(ops.unpack_planes(gSi, 32, plane_off=4+4*(i-1)).cpu(), keep['rdb%d_x%d'%(j,i)].grad))
    cmp('rdb%d_in' % j, calls[b+4]['out32'], 32)
    b += 5
