import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "explorable-super-resolution_b200"))
import torch, torch.nn.functional as F
from esr_b200 import ops
torch.backends.cudnn.allow_tf32 = False
dev = 'cuda'
for cout in (16, 24, 32, 40, 48, 64, 104, 168):
    for (h, w) in ((16, 12), (64, 48), (33, 70)):
        g = torch.Generator().manual_seed(cout + h)
        x = torch.randn(1, 32, h, w, generator=g).half().float().to(dev)
        wt = (torch.randn(cout, 32, 3, 3, generator=g) / 17).half().float().to(dev)
        ref = F.conv2d(x.double(), wt.double(), None, padding=1).float()
        x16, _ = ops.pack_nchw(x)
        pc = ops.PackedConv(wt, torch.zeros(cout, device=dev))
        o = torch.zeros(1, (cout + 7) // 8, h, w, 8, device=dev)
        ops.conv3x3(x16, pc, out32=o)
        got = ops.unpack_planes(o, cout)
        print(cout, (h, w), 'nb_n pad', pc.cout_pad, 'rel err %.2e' % ((got - ref).abs().max() / ref.abs().max()).item())
