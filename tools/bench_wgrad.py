"""times the weight-gradient launches of one C2 dense block (16 x 256 x 256, bf16) and the HR convs' on the current GPU
  python tools/bench_wgrad.py            (ESR_WGRAD_V1=1 for the three-TMA-copies kernel)"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'explorable-super-resolution_b200'))
from esr_b200 import ops  # noqa: E402

dev = 'cuda'
torch.manual_seed(0)
rows = []
for (n, h, w, cin, cout) in [(16, 256, 256, 64, 32), (16, 256, 256, 96, 32), (16, 256, 256, 128, 32), (16, 256, 256, 160, 32), (16, 256, 256, 192, 64),
                             (16, 256, 256, 64, 64), (4, 1024, 1024, 64, 64), (4, 52, 52, 64, 32), (4, 52, 52, 192, 64)]:
    x = (torch.randn(n, cin // 8, h, w, 8, device=dev) * 0.5).to(torch.bfloat16)
    gy = (torch.randn(n, cout // 8, h, w, 8, device=dev) * 0.5).to(torch.bfloat16)
    dw = torch.zeros(cout, cin, 3, 3, device=dev)
    db = torch.zeros(cout, device=dev)
    for _ in range(3):
        ops.conv3x3_wgrad(x, gy, cout, cin, dw=dw, db=db)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        ops.conv3x3_wgrad(x, gy, cout, cin, dw=dw, db=db)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / iters
    flop = 2.0 * n * h * w * 9 * cin * cout
    # reference on a sub-sample: bias gradient exactly, dW checksum
    ref_db = gy.float().sum(dim=(0, 2, 3)).reshape(-1)
    err_b = ((db - ref_db).abs().max() / ref_db.abs().max().clamp_min(1e-6)).item()
    rows.append((n, h, w, cin, cout, us, flop / us / 1e6, err_b, float(dw.double().abs().sum())))
    print('n%d %dx%d %3d->%2d  %8.1f us (wgrad + reduce + bias)  %7.1f TF/s   db err %.1e   |dW| %.6e' % rows[-1], flush=True)
