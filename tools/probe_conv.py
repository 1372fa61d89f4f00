"""GPU bring-up probe: one fused conv launch vs torch, several shapes.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "explorable-super-resolution_b200"))
import torch
import torch.nn.functional as F
from esr_b200 import ops, lib

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
ops.device_check()
print("device ok", torch.cuda.get_device_name(0), flush=True)

def run(n, cin, cout, h, w, dtype=torch.float16, lrelu=False, mt=0, p=0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).to(dev)
    b = torch.randn(cout, generator=g).to(dev) * 0.1
    xq = x.to(dtype).float(); wq = wt.to(dtype).float()
    ref = F.conv2d(xq, wq, b, padding=1)
    if lrelu: ref = F.leaky_relu(ref, 0.2)
    x16, _ = ops.pack_nchw(x, dtype=dtype)
    back = ops.unpack_planes(x16, cin)
    e0 = (back - xq).abs().max().item()
    pc = ops.PackedConv(wt, b, dtype=dtype)
    torch.cuda.synchronize(); print("  packed ok", flush=True)
    out32 = torch.zeros((n, ops.planes_for(cout), h, w, 8), dtype=torch.float32, device=dev)
    ops.conv3x3(x16, pc, lrelu=lrelu, out32=out32, tile_mt=mt, tile_p=p)
    torch.cuda.synchronize()
    wd = lib.watchdog()
    if wd[0]:
        print("  WATCHDOG: block %d thread %d bar_off 0x%x parity %d tag %d (1=producer-empty 2=mma-tmem-empty 3=mma-full 4=epi-tmem-full)" % (wd[1], wd[2], wd[3], wd[4], wd[5]), flush=True)
    got = ops.unpack_planes(out32, cout)
    err = (got - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    print(f"n={n} cin={cin} cout={cout} {h}x{w} {dtype} lrelu={lrelu} mt={mt} p={p}: pack_err={e0:.1e} max_err={err:.3e} rel={rel:.3e}", flush=True)
    if rel > 1e-3:
        d = (got - ref).abs()
        idx = (d > 1e-3 * ref.abs().max()).nonzero()
        print("   bad count", idx.shape[0], "of", d.numel(), "first", idx[:8].tolist())
    return rel

cases = [
    (1, 16, 16, 8, 30),
    (1, 16, 16, 16, 30),
    (1, 32, 32, 16, 30),
    (1, 64, 32, 16, 30),
    (2, 64, 64, 40, 70),
    (1, 192, 64, 64, 64),
    (1, 3, 64, 33, 47),
    (1, 64, 3, 33, 47),
]
worst = 0
for c in cases:
    for mt in (1, 4):
        worst = max(worst, run(*c, mt=mt))
worst = max(worst, run(2, 96, 32, 50, 50, dtype=torch.bfloat16, lrelu=True))
worst = max(worst, run(1, 64, 64, 64, 130, mt=2))
print("WORST", worst)
# quick timing of the big trunk convs
for cin, cout in [(64, 32), (192, 64)]:
    n, h, w = 4, 256, 256
    x16 = torch.randn(n, cin // 8, h, w, 8, device=dev).half()
    pc = ops.PackedConv(torch.randn(cout, cin, 3, 3, device=dev) * 0.05, torch.zeros(cout, device=dev))
    o = torch.zeros(n, cout // 8, h, w, 8, device=dev).half()
    for _ in range(3): ops.conv3x3(x16, pc, lrelu=True, out16=o)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10): ops.conv3x3(x16, pc, lrelu=True, out16=o)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    fl = 2 * n * h * w * cin * cout * 9
    print(f"conv {cin}->{cout} {n}x{h}x{w}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
