"""does engine.backward read scratch memory it has not written?  (ESR_POISON fills the scratch with NaN)"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'explorable-super-resolution_b200'))
import torch
import models.modules.architecture as arch

dev = 'cuda'
torch.manual_seed(0)
for (z, pad, h, training) in [(3, 10, 64, False), (3, 0, 64, False), (0, 10, 64, False), (3, 10, 52, False), (3, 0, 48, True), (0, 0, 48, True)]:
    net = arch.RRDBNet(3, 3, 64, 2, upscale=4, latent_input='all_layers,HR_downscaled' if z else None, num_latent_channels=z).to(dev)
    for p in net.parameters():
        p.requires_grad_(training)
    eng = net.engine(training=training)
    cin = 3 + z * 16
    x = torch.rand(8, cin, h, h, device=dev)
    res = []
    for poison in ('0', '1'):
        os.environ['ESR_POISON'] = poison
        out, sv = eng.forward(x, pad=pad, save=True)
        g = torch.randn_like(out)
        torch.manual_seed(1)
        g = torch.randn(out.shape, device=dev)
        gx, grads = eng.backward(g, sv, wgrad=training)
        res.append(gx.clone())
    a, b = res
    print('z=%d pad=%d h=%d training=%s: nan in poisoned gx: latent %d image %d   max diff %.3e' % (
        z, pad, h, training, int(torch.isnan(b[:, :cin - 3]).sum()), int(torch.isnan(b[:, cin - 3:]).sum()),
        float((a - b).nan_to_num(0).abs().max())), flush=True)
