"""per-launch latency of the conv kernels on small images (the C3 / C4 regime)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "explorable-super-resolution_b200"))
import torch
from esr_b200 import ops
dev = torch.device("cuda")
def run(n, h, w, cin, cout, rows, reps=200):
    D = (torch.randn(n, 24, h, w, 8, device=dev) * 0.5).half()
    pc = ops.PackedConv(torch.randn(cout, cin, 3, 3, device=dev) * 0.03, torch.zeros(cout, device=dev))
    O = torch.zeros(n, 8, h, w, 8, device=dev, dtype=torch.float16)
    f = lambda: ops.conv3x3(D, pc, cin_planes=cin // 8, lrelu=True, out16=O, rows=rows)
    for _ in range(10): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): f()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g, stream=s):
        for _ in range(reps): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print("n=%d %dx%d %3d->%2d rows=%-5s: %6.2f us per launch (graph replay)  %6.1f TFLOP/s" % (n, h, w, cin, cout, rows, us, 2.0 * n * h * w * cin * cout * 9 / us / 1e6), flush=True)
for (n, h, w) in ((8, 84, 84), (4, 52, 52), (1, 128, 128)):
    for cin, cout in ((64, 32), (160, 32), (192, 64)):
        for rows in (False, 'force'):
            run(n, h, w, cin, cout, rows)
