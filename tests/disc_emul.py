"""CPU stand-ins for the esr_b200.ops entry points the discriminator engine calls, operating on the same planar-8 /
space-to-depth layouts with plain torch.  They let the CPU suite exercise DiscEngine's HOST logic (layer bookkeeping, layouts,
4x4-stride-2 weight re-indexing, gradient routing) against autograd; the CUDA kernels themselves are checked by the GPU tests.
Test infrastructure only."""
import torch
import torch.nn.functional as F


def planes_for(c):
    return (c + 7) // 8


def to_planes(x, dtype):
    n, c, h, w = x.shape
    p = planes_for(c)
    buf = torch.zeros((n, p * 8, h, w), dtype=dtype)
    buf[:, :c] = x.to(dtype)
    return buf.view(n, p, 8, h, w).permute(0, 1, 3, 4, 2).contiguous()


def from_planes(pl, c=None):
    n, p, h, w, _ = pl.shape
    x = pl.permute(0, 1, 4, 2, 3).reshape(n, p * 8, h, w)
    return x if c is None else x[:, :c]


def s2d_nchw(x):
    """[N,C,H,W] -> [N,4C,H/2,W/2], channel (py*2+px)*C + c"""
    n, c, h, w = x.shape
    return x.view(n, c, h // 2, 2, w // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, h // 2, w // 2)


def d2s_nchw(x):
    n, c4, h2, w2 = x.shape
    c = c4 // 4
    return x.view(n, 2, 2, c, h2, w2).permute(0, 3, 4, 1, 5, 2).reshape(n, c, 2 * h2, 2 * w2)


class PackedConv:
    def __init__(self, weight, bias, dtype=torch.float32, lead=0, transpose_flip=False, **kw):
        self.t = transpose_flip
        self.dtype = dtype
        self.wpacked = weight          # only its .device is looked at
        self.cout_pad = (weight.shape[1] if transpose_flip else weight.shape[0])
        self.repack(weight, bias)

    def repack(self, weight, bias, queue=None):
        self.w = weight.detach().to(self.dtype).float()     # operand rounding of the tensor-core image, fp32 accumulation
        self.b = None if (bias is None or self.t) else bias.detach().float()


def conv3x3(x16, pc, *, out32=None, out_nchw=None, bias=None, **kw):
    assert not kw, kw
    x = from_planes(x16).float()
    if pc.t:
        y = F.conv_transpose2d(x[:, :pc.w.shape[0]], pc.w, padding=1)
    else:
        b = pc.b if bias is None else bias[:pc.w.shape[0]]
        y = F.conv2d(x[:, :pc.w.shape[1]], pc.w, None, padding=1).float() + b.view(1, -1, 1, 1)
    y = y.float()
    if out32 is not None:
        out32.copy_(to_planes(y, torch.float32))
    if out_nchw is not None:
        out_nchw.copy_(y[:, :out_nchw.shape[1]])


def pack_nchw(src, dtype=torch.float32, **kw):
    return to_planes(src, dtype), None


def bn_stats(y32, c, gamma, beta, eps, momentum, train, running_mean, running_var):
    y = from_planes(y32, c).double()
    if train:
        mean = y.mean((0, 2, 3))
        var = y.var((0, 2, 3), unbiased=False)
        m = y.numel() / c
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * mean.float())
            running_var.mul_(1 - momentum).add_(momentum * (var * m / (m - 1)).float())
    else:
        mean, var = running_mean.double(), running_var.double()
    invstd = 1 / torch.sqrt(var + eps)
    scale = gamma.double() * invstd
    return mean.float(), invstd.float(), scale.float(), (beta.double() - mean * scale).float()


def bn_lrelu_fwd(y32, c, scale, shift, slope, dtype, space_to_depth=False, want16=True, want_nchw=False):
    y = from_planes(y32, c)
    v = F.leaky_relu(y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), slope)
    d16 = None
    if want16:
        if space_to_depth:
            pad = torch.zeros((v.shape[0], planes_for(c) * 8, v.shape[2], v.shape[3]))
            pad[:, :c] = v
            d16 = to_planes(s2d_nchw(pad), dtype)
        else:
            d16 = to_planes(v, dtype)
    return d16, (v.contiguous() if want_nchw else None)


def bn_lrelu_bwd(g, g_layout, y32, c, scale, shift, mean, invstd, slope, dtype, *, has_bn, train=True, gscale=1.0, dgamma=None, dbeta=None,
                 accumulate=False, scratch=None):
    y = from_planes(y32, c)
    if g_layout == 0:
        gg = from_planes(g, c)
    elif g_layout == 1:
        gg = d2s_nchw(from_planes(g))[:, :c]
    else:
        gg = g
    v = lambda t: t.view(1, -1, 1, 1)
    gb = torch.where(y * v(scale) + v(shift) > 0, gg, gg * slope)
    xh = (y - v(mean)) * v(invstd)
    m = y.numel() / c
    s, q = gb.sum((0, 2, 3)), (gb * xh).sum((0, 2, 3))
    if has_bn:
        if dbeta is not None:
            dbeta.copy_(s * gscale)
        if dgamma is not None:
            dgamma.copy_(q * gscale)
    c1 = s / m if (has_bn and train) else torch.zeros_like(s)
    c2 = q / m if (has_bn and train) else torch.zeros_like(q)
    return to_planes(v(scale) * (gb - v(c1) - xh * v(c2)), dtype)


def _layout_to_nchw(g, g_layout, c):
    if g is None:
        return None
    if g_layout == 0:
        return from_planes(g, c)
    if g_layout == 1:
        return d2s_nchw(from_planes(g))[:, :c]
    return g


def bn_tangent_fwd(t32, y32, c, scale, shift, mean, invstd, slope, dtype, *, has_bn, space_to_depth=False, want16=True, want_nchw=False):
    """csrc/disc_kernels.cuh bn_tangent_apply_kernel with torch ops"""
    y, t = from_planes(y32, c), from_planes(t32, c)
    v = lambda a: a.view(1, -1, 1, 1)
    xh = (y - v(mean)) * v(invstd)
    if has_bn:
        c1, c2 = t.mean((0, 2, 3)), (xh * t).mean((0, 2, 3))
    else:
        c1, c2 = torch.zeros(c), torch.zeros(c)
    wv = v(scale) * (t - v(c1) - xh * v(c2))
    wv = torch.where(y * v(scale) + v(shift) > 0, wv, wv * slope)
    d16 = None
    if want16:
        if space_to_depth:
            pad = torch.zeros((wv.shape[0], planes_for(c) * 8, wv.shape[2], wv.shape[3]))
            pad[:, :c] = wv
            d16 = to_planes(s2d_nchw(pad), dtype)
        else:
            d16 = to_planes(wv, dtype)
    return d16, (wv.contiguous() if want_nchw else None), c1, c2


def bn_double_bwd(zb, wb, g_layout, y32, t32, c, scale, shift, mean, invstd, c1, c2, slope, dtype, *, has_bn, gscale=1.0, dgamma=None, dbeta=None,
                  accumulate=False):
    """csrc/disc_kernels.cuh bn_dbl_* kernels with torch ops"""
    y, t = from_planes(y32, c), from_planes(t32, c)
    v = lambda a: a.view(1, -1, 1, 1)
    msk = torch.where(y * v(scale) + v(shift) > 0, 1.0, slope)
    zz, ww = _layout_to_nchw(zb, g_layout, c), _layout_to_nchw(wb, g_layout, c)
    p = msk * zz if zz is not None else torch.zeros_like(y)
    q = msk * ww if ww is not None else torch.zeros_like(y)
    if not has_bn:
        return to_planes(v(scale) * q, dtype), to_planes(v(scale) * p, dtype)
    xh = (y - v(mean)) * v(invstd)
    n = y.numel() / c
    A = t - v(c1) - xh * v(c2)
    S = [a.sum((0, 2, 3)) for a in (p, p * xh, q, q * xh, q * A)]
    if dbeta is not None:
        dbeta.copy_(S[0] * gscale)
    if dgamma is not None:
        dgamma.copy_((S[1] + invstd * S[4]) * gscale)
    qc = q - v(S[2] / n) - xh * v(S[3] / n)
    tb = v(scale) * qc
    yb = v(scale) * (p - v(S[0] / n) - xh * v(S[1] / n)) - v(scale * invstd) * (v(S[4] / n) * xh + v(c2) * qc + v(S[3] / n) * A)
    return to_planes(tb, dtype), to_planes(yb, dtype)


def linear_fwd(x, weight, bias, lrelu=False, slope=0.2):
    y = F.linear(x, weight, bias)
    return F.leaky_relu(y, slope) if lrelu else y


def is_split(dt):
    return False


def logical_planes(t, dt):
    return t.shape[1]


def linear_bwd(g, act, x, weight, slope=0.2, want_gx=True, want_w=True, gscale=1.0, dw=None, db=None, accumulate=False):
    gm = g if act is None else torch.where(act > 0, g, g * slope)
    gw, gb = (gscale * gm.t() @ x if want_w else None), (gscale * gm.sum(0) if want_w else None)
    if want_w and dw is not None:       # written / accumulated in place, like the C-ABI call
        dw.copy_(dw + gw if accumulate else gw)
        db.copy_(db + gb if accumulate else gb)
        gw, gb = dw, db
    return (gm @ weight if want_gx else None), gw, gb


def conv3x3_wgrad(x16, gy16, cout, cin, dw=None, db=None, accumulate=False, scale=1.0, **kw):
    x = from_planes(x16, cin).double()
    gy = from_planes(gy16, cout).double()
    w = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(cout, dtype=torch.float64, requires_grad=True)
    with torch.enable_grad():    # the engine's backward runs under no_grad
        F.conv2d(x, w, b, padding=1).backward(gy)
    gw, gb = scale * w.grad.float(), scale * b.grad.float()
    if dw is not None:           # written / accumulated in place, like the C-ABI call (db=False: no bias gradient wanted)
        dw.copy_(dw + gw if accumulate else gw)
        if db is not None and db is not False:
            db.copy_(db + gb if accumulate else gb)
        return dw, (db if db is not False else None)
    return gw, gb


def install(monkeypatch):
    from esr_b200 import ops
    monkeypatch.setattr(ops, 'run_pack_queue', lambda q: None)
    for name in ('PackedConv', 'conv3x3', 'pack_nchw', 'bn_stats', 'bn_lrelu_fwd', 'bn_lrelu_bwd', 'linear_fwd', 'linear_bwd', 'conv3x3_wgrad',
                 'bn_tangent_fwd', 'bn_double_bwd'):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, 'require_cuda', lambda *a: None)
