import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'explorable-super-resolution_b200'))
from esr_b200 import ops, lib
dev = 'cuda'
torch.manual_seed(0)
for (n, h, w, cin, cout) in [(1, 8, 64, 64, 32), (1, 64, 64, 64, 32), (2, 40, 100, 96, 32), (4, 256, 256, 64, 32), (2, 64, 200, 192, 64), (2, 64, 64, 64, 64), (1, 20, 70, 16, 8), (1, 300, 130, 64, 64)]:
    x = (torch.randn(n, cin // 8, h, w, 8, device=dev) * 0.5).to(torch.bfloat16)
    gy = (torch.randn(n, cout // 8, h, w, 8, device=dev) * 0.5).to(torch.bfloat16)
    dw, db = ops.conv3x3_wgrad(x, gy, cout, cin)
    torch.cuda.synchronize()
    xn = x.float().permute(0, 1, 4, 2, 3).reshape(n, cin, h, w).double()
    gn = gy.float().permute(0, 1, 4, 2, 3).reshape(n, cout, h, w).double()
    ref = torch.nn.grad.conv2d_weight(xn, (cout, cin, 3, 3), gn, padding=1)
    ref_db = gn.sum(dim=(0, 2, 3))
    e = (dw.double() - ref).abs().max().item() / ref.abs().max().item()
    eb = (db.double() - ref_db).abs().max().item() / ref_db.abs().max().item()
    print((n, h, w, cin, cout), 'dW rel err %.2e  db rel err %.2e  nan %d  wd %s' % (e, eb, int(torch.isnan(dw).sum()), lib.watchdog()), flush=True)
    if eb > 1e-3:
        print('  db ', db[:8].tolist()); print('  ref', ref_db[:8].tolist())
