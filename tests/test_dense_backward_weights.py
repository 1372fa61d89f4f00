"""Host logic of the dense-block backward (esr_b200.engine.combine_dense_backward_weights) on CPU: running the combined
convolutions over the concatenated block gradients reproduces torch autograd's gradients of a ResidualDenseBlock_5C
(models/modules/block.py:196-242), including the latent channels concatenated in front of every conv input."""
import pytest
import torch
import torch.nn.functional as F


def _dense_block(ws, bs, x, zlat, scale):
    """block.py:230-235 with latent input: x_i = lrelu(conv_i(cat(z, x, x_1..x_{i-1}))), out = x_5 * scale + x"""
    feats = [x]
    pre = []
    for i in range(5):
        inp = torch.cat(([zlat] if zlat is not None else []) + feats, 1)
        y = F.conv2d(inp, ws[i], bs[i], padding=1)
        pre.append(y)
        if i < 4:
            feats.append(F.leaky_relu(y, 0.2))
    return pre[4] * scale + x, pre


@pytest.mark.parametrize('z', [0, 3])
def test_combined_weights_reproduce_autograd(z):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'explorable-super-resolution_b200'))
    from esr_b200.engine import combine_dense_backward_weights
    torch.manual_seed(0)
    nf, gc, R, scale = 16, 8, 2, 0.2
    ws = [[(torch.randn(gc if i < 4 else nf, z + nf + i * gc, 3, 3, dtype=torch.float64) * 0.1) for i in range(5)] for _ in range(R)]
    bs = [[torch.randn(gc if i < 4 else nf, dtype=torch.float64) * 0.1 for i in range(5)] for _ in range(R)]
    W = [torch.stack([ws[r][i] for r in range(R)]) for i in range(5)]
    comb = combine_dense_backward_weights(W, torch.full((R,), scale, dtype=torch.float64), z, nf, gc)
    for r in range(R):
        x = torch.randn(1, nf, 9, 7, dtype=torch.float64, requires_grad=True)
        zlat = torch.randn(1, z, 9, 7, dtype=torch.float64, requires_grad=True) if z else None
        out, pre = _dense_block(ws[r], bs[r], x, zlat, scale)
        for p in pre:
            p.retain_grad()
        g_out = torch.randn_like(out)
        out.backward(g_out)
        # the chain the engine runs: G = [g_5 | g_4 | ... ], g_5 = dL/d(out) (the block scale sits in the weights)
        G = g_out.clone()
        for i in (4, 3, 2, 1):
            g_i = F.conv2d(G, comb[i][r], padding=1) * torch.where(pre[i - 1] > 0, torch.ones_like(pre[i - 1]), torch.full_like(pre[i - 1], 0.2))   # lrelu'(x_i)
            assert torch.allclose(g_i, pre[i - 1].grad, atol=1e-10), i
            G = torch.cat([G, g_i], 1)
        g0 = F.conv2d(G, comb[0][r], padding=1)
        if z:
            assert g0.shape[1] == 8 + nf and float(g0[:, z:8].abs().max()) == 0.0          # latent rows padded to one plane
            assert torch.allclose(g0[:, :z], zlat.grad, atol=1e-10)
            gx = g0[:, 8:]
        else:
            gx = g0
        assert torch.allclose(gx + g_out, x.grad, atol=1e-10)                              # + the block's own skip connection
