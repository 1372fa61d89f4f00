"""forward launches of one C2 dense block (16 x 256 x 256, fp16) on the row kernel, one after the other (cold-ish: a 384 MB dense buffer)
  python tools/bench_conv.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "explorable-super-resolution_b200"))
import torch
from esr_b200 import ops
dev = torch.device("cuda")
n, h, w = 16, 256, 256
D = (torch.randn(n, 24, h, w, 8, device=dev) * 0.5).half()
O = torch.zeros(n, 8, h, w, 8, device=dev, dtype=torch.float16)
for cin, cout in ((64, 32), (96, 32), (128, 32), (160, 32), (192, 64)):
    pc = ops.PackedConv(torch.randn(cout, cin, 3, 3, device=dev) * 0.03, torch.zeros(cout, device=dev))
    if cout == 32:
        f = lambda: ops.conv3x3(D, pc, cin_planes=cin // 8, lrelu=True, out16=D, out16_off=cin // 8)
    else:
        f = lambda: ops.conv3x3(D, pc, cin_planes=cin // 8, alpha=0.2, res1=D, res1_off=0, beta1=1.0, out16=O)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print("%3d->%2d: %7.1f us  %7.1f TFLOP/s" % (cin, cout, us, 2.0 * n * h * w * cin * cout * 9 / us / 1e6), flush=True)
