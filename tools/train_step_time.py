"""times one generator training step (fwd + CEM + L1 loss + bwd with weight gradients + Adam) at a C2-like shape"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, 'explorable-super-resolution_b200'), REPO):
    sys.path.insert(0, p)
import torch
sys.argv = sys.argv[:1] + sys.argv[1:]
B = int(os.environ.get('B', 8)); LR = int(os.environ.get('LR', 256))
import bench
bench.BATCH = B
model, cem = bench.build_model(torch.device('cuda'))
model.train()
params = [p for p in model.parameters() if p.requires_grad]
print('trainable params', sum(p.numel() for p in params))
opt = torch.optim.Adam(params, lr=1e-4)
x = torch.rand(B, 3, LR, LR, device='cuda'); hr = torch.rand(B, 3, 4 * LR, 4 * LR, device='cuda')
from esr_b200 import lib
def step():
    opt.zero_grad(set_to_none=True)
    out = model(x)
    loss = (out - hr).abs().mean()
    loss.backward()
    opt.step()
    return loss
for _ in range(2): l = step()
torch.cuda.synchronize()
n0 = lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 3
for _ in range(K): l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print('B=%d LR=%d: %.1f ms/step, %.1f HR-MP/s fwd+bwd, loss %.4f, %d launches/step, mem %.1f GB' % (
    B, LR, ms, B * (4 * LR) ** 2 / 1e6 / (ms * 1e-3), l.item(), (lib.launch_count() - n0) // K, torch.cuda.max_memory_allocated() / 2**30))
print('watchdog', lib.watchdog())
# breakdown: forward only (save), backward only
with torch.no_grad():
    torch.cuda.synchronize(); e0.record(); model(x); e1.record(); torch.cuda.synchronize()
print('inference fwd (fp16): %.1f ms' % e0.elapsed_time(e1))
# per-op-class CUDA-event timing of one training step
import esr_b200.engine as E
import esr_b200.ops as O
acc = {}
def wrap(name, fn):
    def f(*a, **k):
        key = name
        if name == 'conv3x3':
            key = 'conv3x3 dgrad' if (k.get('mask16') is not None or k.get('lead_planes') or k.get('res3') is not None or k.get('tail_first') or torch.is_grad_enabled() is False and False) else 'conv3x3'
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = fn(*a, **k); e.record()
        acc.setdefault(key, []).append((s, e))
        return r
    return f
for nm in ('conv3x3', 'conv3x3_wgrad', 'pack_nchw', 'downsum2x', 'planes_add', 'cem_down', 'cem_inv', 'cem_up_add', 'sep_adjoint_2d', 'sum_nchw', 'latent_grad'):
    setattr(O, nm, wrap(nm, getattr(O, nm)))
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True); t3 = torch.cuda.Event(enable_timing=True)
opt.zero_grad(set_to_none=True)
t0.record(); out = model(x); loss = (out - hr).abs().mean(); t1.record(); loss.backward(); t2.record(); opt.step(); t3.record()
torch.cuda.synchronize()
print('forward(save) %.1f ms | backward %.1f ms | adam %.1f ms' % (t0.elapsed_time(t1), t1.elapsed_time(t2), t2.elapsed_time(t3)))
for k, v in acc.items():
    print('  %-16s %5d calls %8.1f ms' % (k, len(v), sum(a.elapsed_time(b) for a, b in v)))
