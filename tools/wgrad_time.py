"""times the weight-gradient op on the dense-block shapes of C2 (16 x 256 x 256) in isolation"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "explorable-super-resolution_b200"))
import torch
from esr_b200 import ops
dev = torch.device("cuda")
N = int(os.environ.get("N", 16)); H = W = 256
X = (torch.randn(N, 24, H, W, 8, device=dev) * 0.5).bfloat16()
for cin, cout in ((64, 32), (96, 32), (128, 32), (160, 32), (192, 64), (64, 64)):
    G = (torch.randn(N, cout // 8, H, W, 8, device=dev) * 0.5).bfloat16()
    for _ in range(2): ops.conv3x3_wgrad(X, G, cout, cin)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps): ops.conv3x3_wgrad(X, G, cout, cin)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print("wgrad %3d->%2d: %8.1f us  %7.1f TFLOP/s" % (cin, cout, us, 2.0 * N * H * W * cin * cout * 9 / us / 1e6), flush=True)
