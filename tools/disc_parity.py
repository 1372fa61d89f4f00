"""Per-tensor parity of the discriminator against the reference's golden fixture (tests/golden/disc_vgg128_nf8.npz) in fp16 and
bf16: prints rel-L2 / cosine of the logits, the image gradient and every parameter gradient (GPU box)."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, 'explorable-super-resolution_b200'), REPO, os.path.join(REPO, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.nn.functional as F
from util import golden, rel_err
import models.modules.architecture as arch

g = golden('disc_vgg128_nf8')
sd = {k[2:]: (torch.from_numpy(g[k]).float() if g[k].dtype.kind == 'f' else torch.from_numpy(g[k])) for k in g.files if k.startswith('w:')}
cos = lambda a, b: F.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()
for dtype in (torch.float16, torch.bfloat16):
    net = arch.Discriminator_VGG_128(3, int(g['cfg'][0]))
    net.load_state_dict(sd)
    net.compute_dtype = dtype
    net = net.cuda().train()
    x = torch.from_numpy(g['x'].astype(np.float32)).cuda().requires_grad_(True)
    out = net(x)
    gs = 64.0 if dtype == torch.float16 else 1.0
    (out * gs * torch.from_numpy(g['wt']).cuda()).sum().backward()
    rows = {'logits': rel_err(out.detach().cpu(), torch.from_numpy(g['out'])), 'gx': rel_err(x.grad.cpu() / gs, torch.from_numpy(g['gx'])) +
            (cos(x.grad.cpu(), torch.from_numpy(g['gx'])),)}
    for name, p in net.named_parameters():
        ref = torch.from_numpy(g['g:' + name])
        rows[name] = rel_err(p.grad.cpu() / gs, ref) + (cos(p.grad.cpu(), ref), float(ref.abs().max()))
    print(str(dtype))
    for k, v in rows.items():
        print('  %-28s' % k, ' '.join('%.3e' % t for t in v))
