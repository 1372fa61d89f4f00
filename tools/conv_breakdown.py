"""Times every distinct conv launch of the C2 forward in isolation (back-to-back launches, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "explorable-super-resolution_b200"))
import torch
from esr_b200 import ops

dev = torch.device("cuda")
N = int(os.environ.get("N", 16))
def t16(*s): return (torch.randn(*s, device=dev) * 0.5).half()
def t32(*s): return torch.randn(*s, device=dev)
H = W = 256
D = t16(N, 24, H, W, 8); D2 = t16(N, 24, H, W, 8); T = t32(N, 8, H, W, 8); T2 = t32(N, 8, H, W, 8)
def pc(cin, cout): return ops.PackedConv(torch.randn(cout, cin, 3, 3, device=dev) * 0.03, torch.zeros(cout, device=dev))
cases = []
for i in range(4):
    cin = 64 + 32 * i
    p = pc(cin, 32)
    cases.append(("rdb conv%d %d->32 @256" % (i + 1, cin), cin, 32, H, lambda p=p, i=i: ops.conv3x3(D, p, cin_planes=8 + 4 * i, lrelu=True, out16=D, out16_off=8 + 4 * i)))
p5 = pc(192, 64)
cases.append(("rdb conv5 192->64 res16", 192, 64, H, lambda: ops.conv3x3(D, p5, alpha=0.2, res1=D, res1_off=0, out16=D2, out16_off=0)))
cases.append(("rdb3 conv5 192->64 res16+res32+out32", 192, 64, H, lambda: ops.conv3x3(D, p5, alpha=0.04, res1=D, beta1=0.2, res2=T, out16=D2, out32=T2)))
cases.append(("conv5 192->64 fp32 trunk (rdb mode)", 192, 64, H, lambda: ops.conv3x3(D, p5, alpha=0.2, res1=T, out16=D2, out32=T2)))
p64 = pc(64, 64)
U0 = t16(N, 8, 2 * H, 2 * W, 8)
cases.append(("LR_conv 64->64 @256 -> up2 store", 64, 64, H, lambda: ops.conv3x3(D, p64, cin_planes=8, res1=T, out16=U0, up2=True)))
if N <= 16:
    U1 = t16(N, 8, 4 * H, 4 * W, 8); V = t16(N, 8, 4 * H, 4 * W, 8)
    cases.append(("upconv0 64->64 @512 -> up2 store", 64, 64, 2 * H, lambda: ops.conv3x3(U0, p64, lrelu=True, out16=U1, up2=True)))
    cases.append(("upconv1/HR0 64->64 @1024", 64, 64, 4 * H, lambda: ops.conv3x3(U1, p64, lrelu=True, out16=V)))
    p3 = pc(64, 3); out = torch.empty(N, 3, 4 * H, 4 * W, device=dev)
    cases.append(("HR1 64->3 @1024 -> NCHW fp32", 64, 3, 4 * H, lambda: ops.conv3x3(V, p3, out_nchw=out)))
tot = 0
for name, cin, cout, h, fn in cases:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    fl = 2.0 * N * h * h * cin * cout * 9
    print("%-42s %8.1f us  %7.1f TFLOP/s" % (name, us, fl / us / 1e6), flush=True)
