"""Layer-by-layer comparison of the CUDA discriminator engine with the host stand-ins (tests/disc_emul.py) using identical
operand rounding: forward conv outputs / statistics per layer, then every parameter gradient (GPU box)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, 'explorable-super-resolution_b200'), REPO, os.path.join(REPO, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch
import disc_emul as E
from util import golden, rel_err
import models.modules.architecture as arch
from esr_b200 import ops, disc

g = golden('disc_vgg128_nf8')
sd = {k[2:]: (torch.from_numpy(g[k]).float() if g[k].dtype.kind == 'f' else torch.from_numpy(g[k])) for k in g.files if k.startswith('w:')}


class MP:
    def __init__(self):
        self.saved = []

    def setattr(self, o, n, v):
        self.saved.append((o, n, getattr(o, n)))
        setattr(o, n, v)

    def undo(self):
        for o, n, v in reversed(self.saved):
            setattr(o, n, v)


for dtype in (torch.float16, torch.bfloat16):
    x = torch.from_numpy(g['x'].astype(np.float32))
    wt = torch.from_numpy(g['wt'])
    mp = MP()
    E.install(mp)
    emul = arch.Discriminator_VGG_128(3, 8)
    emul.load_state_dict(sd)
    emul.compute_dtype = dtype
    emul.train()
    e_out, e_sv = emul.engine().forward(x, save=True)
    e_gx, e_pl = emul.engine().backward(wt.clone(), e_sv)
    mp.undo()
    net = arch.Discriminator_VGG_128(3, 8)
    net.load_state_dict(sd)
    net.compute_dtype = dtype
    net = net.cuda().train()
    out, sv = net.engine().forward(x.cuda(), save=True)
    gx, pl = net.engine().backward(wt.cuda(), sv)
    print(str(dtype), 'logits', rel_err(out.cpu(), e_out))
    for li, (a, b) in enumerate(zip(sv[0], e_sv[0])):
        cur_mis = (a[0].cpu().float() != b[0].float()).float().mean().item()
        print('  layer %d: input-planes mismatching elements %.2e, y32 %s, mean %.1e invstd %.1e' %
              (li, cur_mis, '%.2e %.2e' % rel_err(a[1].cpu(), b[1]), rel_err(a[2].cpu(), b[2])[0], rel_err(a[3].cpu(), b[3])[0]))
    print('  feat', rel_err(sv[1].cpu(), e_sv[1]), 'h1', rel_err(sv[2].cpu(), e_sv[2]))
    print('  gx', rel_err(gx.cpu(), e_gx))
    for (name, _), a, b in zip(net.named_parameters(), pl, e_pl):
        print('  grad %-24s %.2e %.2e' % ((name,) + rel_err(a.cpu(), b)))
