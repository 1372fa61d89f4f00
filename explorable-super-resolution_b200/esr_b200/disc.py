"""Discriminator_VGG_128 engine (models/modules/architecture.py:446-508): the critic of the SRRaGAN training step
(models/SRRaGAN_model.py:342-395 D step, :466-477 generator-side GAN term).

Every convolution is a tensor-core launch of the generator's conv kernels: 3x3 convs directly, 4x4 stride-2 convs as 3x3 convs
over the 2x2 space-to-depth image with re-indexed weights (`k4s2_to_3x3`) - the producer's BatchNorm/LeakyReLU kernel stores
straight into the space-to-depth layout, the transposed launch returns the gradient in it and the BatchNorm backward reads it
there, so the re-arrangement never exists as a pass of its own.  BatchNorm2d uses batch statistics in training mode (per
process, like each replica of the reference's nn.DataParallel) and updates the running statistics in place.  The classifier
(Linear 512*4*4 -> 100 -> 1) runs in fp32.  Backward: input gradient (generator's GAN term) and / or parameter gradients
(D step), fp32 master gradients in the reference's parameter shapes.  No PyTorch/cuDNN fallback."""
import torch

from . import ops
from . import optim as flat

SLOPE = 0.2
DIRECT = 'direct'      # this gradient was written straight into the parameter's FlatAdam-registered gradient view


_K4_INDEX = {}


def _k4_index(c, dev):
    """gather tables of the re-indexing for `c` input channels: (src [4c*9] into a flattened [c*16] kernel, valid [4c*9],
    inverse [c*16] into a flattened [4c*9] kernel)"""
    key = (c, str(dev))
    if key not in _K4_INDEX:
        src = torch.zeros((2, 2, c, 3, 3), dtype=torch.long)
        valid = torch.zeros((2, 2, c, 3, 3), dtype=torch.float32)
        inv = torch.zeros((c, 4, 4), dtype=torch.long)
        ch = torch.arange(c)
        for ty in range(3):
            for py in range(2):
                ky = 2 * ty + py - 1
                if not 0 <= ky <= 3:
                    continue
                for tx in range(3):
                    for px in range(2):
                        kx = 2 * tx + px - 1
                        if 0 <= kx <= 3:
                            src[py, px, :, ty, tx] = ch * 16 + ky * 4 + kx
                            valid[py, px, :, ty, tx] = 1.0
                            inv[:, ky, kx] = (((py * 2 + px) * c + ch) * 3 + ty) * 3 + tx
        _K4_INDEX[key] = (src.reshape(-1).to(dev), valid.reshape(-1).to(dev), inv.reshape(-1).to(dev))
    return _K4_INDEX[key]


def k4s2_to_3x3(w):
    """[O, C, 4, 4] stride-2 pad-1 kernel -> [O, 4C, 3, 3] stride-1 pad-1 kernel over the space-to-depth image
    S[(py*2+px)*C + c][Y][X] = x[c][2Y+py][2X+px]:  W'[o, (py,px,c), ty, tx] = W[o, c, 2ty+py-1, 2tx+px-1] (zero outside 0..3)"""
    o, c = w.shape[0], w.shape[1]
    src, valid, _ = _k4_index(c, w.device)
    return (w.reshape(o, c * 16)[:, src] * valid.to(w.dtype)).reshape(o, 4 * c, 3, 3)


def k3x3_to_k4s2(w3):
    """inverse gather of `k4s2_to_3x3` (used on weight gradients): [O, 4C, 3, 3] -> [O, C, 4, 4]"""
    o, c = w3.shape[0], w3.shape[1] // 4
    _, _, inv = _k4_index(c, w3.device)
    return w3.reshape(o, 4 * c * 9)[:, inv].reshape(o, c, 4, 4)


class _Layer:
    __slots__ = ('conv', 'bn', 'k4', 'cin', 'cout')


class DiscEngine:
    def __init__(self, module, dtype=torch.bfloat16):
        self.m = module
        self.dtype = dtype
        self.layers = []
        mods = list(module.features.children())
        i = 0
        while i < len(mods):
            conv = mods[i]
            assert isinstance(conv, torch.nn.Conv2d)
            L = _Layer()
            L.conv, L.bn = conv, None
            L.k4 = tuple(conv.kernel_size) == (4, 4)
            if L.k4:
                assert tuple(conv.stride) == (2, 2) and tuple(conv.padding) == (1, 1)
            else:
                assert tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (1, 1)
            L.cin, L.cout = conv.in_channels, conv.out_channels
            i += 1
            if i < len(mods) and isinstance(mods[i], torch.nn.BatchNorm2d):
                L.bn = mods[i]
                i += 1
            assert isinstance(mods[i], torch.nn.LeakyReLU) and abs(mods[i].negative_slope - SLOPE) < 1e-12
            i += 1
            if L.k4 and L.cin % 8:
                raise NotImplementedError('esr_b200 discriminator: stride-2 convs need a multiple of 8 input channels')
            self.layers.append(L)
        self.fc1, self.fc2 = module.classifier[0], module.classifier[2]
        self._ver = None
        self._pk = self._pkt = None
        self._const = {}

    # ---- parameters ----------------------------------------------------------------------------------------------------------
    def params(self):
        """every parameter in module.parameters() order"""
        return list(self.m.parameters())

    def _packed(self):
        ver = tuple((L.conv.weight._version, L.conv.weight.data_ptr(), L.conv.bias._version) for L in self.layers)
        if ver != self._ver:
            q = []
            ws = [k4s2_to_3x3(L.conv.weight.detach().float()) if L.k4 else L.conv.weight.detach().float() for L in self.layers]
            if self._pk is None or self._pk[0].wpacked.device != ws[0].device:
                self._pk = [ops.PackedConv(w, L.conv.bias, dtype=self.dtype, queue=q) for w, L in zip(ws, self.layers)]
                self._pkt = [ops.PackedConv(w, None, dtype=self.dtype, transpose_flip=True, queue=q) for w in ws]
            else:       # after an optimizer step: same buffers, every conv re-packed by one host call
                for pc, pct, w, L in zip(self._pk, self._pkt, ws, self.layers):
                    pc.repack(w, L.conv.bias, q)
                    pct.repack(w, None, q)
            ops.run_pack_queue(q)
            self._ver = ver
        return self._pk, self._pkt

    def _identity_affine(self, c, dev):
        key = (c, str(dev))
        if key not in self._const:
            one, zero = torch.ones(c, device=dev), torch.zeros(c, device=dev)
            self._const[key] = (zero, one, one, zero)     # mean, invstd, scale, shift
        return self._const[key]

    # ---- forward -------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, save=False):
        x = x.float().contiguous()
        ops.require_cuda(x)
        n, c, h, w = x.shape
        if c != self.layers[0].cin:
            raise ops.L.EsrError('Discriminator_VGG_128 expects %d-channel images' % self.layers[0].cin)
        pk, _ = self._packed()
        dev = x.device
        train = self.m.training
        cur, _ = ops.pack_nchw(x, dtype=self.dtype)
        saved = []
        feat = None
        for li, L in enumerate(self.layers):
            last = li == len(self.layers) - 1
            nn_, pin, hh, ww, _ = cur.shape          # for a stride-2 conv `cur` already is the space-to-depth image
            y32 = torch.empty((n, ops.planes_for(L.cout), hh, ww, 8), dtype=torch.float32, device=dev)
            ops.conv3x3(cur, pk[li], out32=y32)
            if L.bn is not None:
                bn = L.bn
                use_batch = train or bn.running_mean is None
                if train and bn.num_batches_tracked is not None:
                    bn.num_batches_tracked += 1
                mom = bn.momentum if bn.momentum is not None else 0.1
                mean, invstd, scale, shift = ops.bn_stats(y32, L.cout, bn.weight.detach(), bn.bias.detach(), bn.eps, mom, use_batch,
                                                          bn.running_mean, bn.running_var)
            else:
                use_batch = False
                mean, invstd, scale, shift = self._identity_affine(L.cout, dev)
            nxt_k4 = (not last) and self.layers[li + 1].k4
            if nxt_k4 and (hh % 2 or ww % 2):
                raise ops.L.EsrError('Discriminator_VGG_128: image size must be even in front of every stride-2 conv')
            out16, out_nchw = ops.bn_lrelu_fwd(y32, L.cout, scale, shift, SLOPE, self.dtype, space_to_depth=nxt_k4, want16=not last,
                                               want_nchw=last)
            if save:
                saved.append((cur, y32, mean, invstd, scale, shift, use_batch))
            if last:
                feat = out_nchw
            else:
                cur = out16
        flat = feat.reshape(n, -1)
        if flat.shape[1] != self.fc1.in_features:
            raise ops.L.EsrError('Discriminator_VGG_128: classifier expects %d features, the image gives %d (input_patch_size mismatch)'
                                 % (self.fc1.in_features, flat.shape[1]))
        h1 = ops.linear_fwd(flat, self.fc1.weight.detach(), self.fc1.bias.detach(), lrelu=True, slope=SLOPE)
        out = ops.linear_fwd(h1, self.fc2.weight.detach(), self.fc2.bias.detach())
        if save:
            return out, (saved, feat, h1, (h, w))
        return out

    # ---- backward ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def backward(self, g_out, sv, want_input=True, want_params=True):
        """g_out: dL/dlogits [N,1] -> (dL/dx [N,C,H,W] | None, [gradient per parameter in module.parameters() order] | None)"""
        saved, feat, h1, (H, W) = sv
        _, pkt = self._packed()
        n = feat.shape[0]
        dev = feat.device
        g_out = g_out.float().contiguous()
        grads = {}
        pending = []

        def target(*params):
            """the parameters' flat gradient views when every one of them is registered and its .grad is unset or already that view:
            (views, accumulate) - else (None, False)"""
            views = [flat.grad_view(p) for p in params]
            if not want_params or any(v is None for v in views):
                return None, False
            if not all(p.grad is None or p.grad is v or p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, views)):
                return None, False
            states = {p.grad is None for p in params}
            if len(states) != 1:
                return None, False
            pending.extend(zip(params, views))
            return views, not states.pop()

        v2, acc2 = target(self.fc2.weight, self.fc2.bias)
        g_h1, dw2, db2 = ops.linear_bwd(g_out, None, h1, self.fc2.weight.detach(), want_w=want_params, dw=v2[0] if v2 else None,
                                        db=v2[1] if v2 else None, accumulate=acc2)
        flat_in = feat.reshape(n, -1)
        v1, acc1 = target(self.fc1.weight, self.fc1.bias)
        g_flat, dw1, db1 = ops.linear_bwd(g_h1, h1, flat_in, self.fc1.weight.detach(), slope=SLOPE, want_w=want_params, dw=v1[0] if v1 else None,
                                          db=v1[1] if v1 else None, accumulate=acc1)
        if want_params:
            grads[id(self.fc1.weight)], grads[id(self.fc1.bias)] = (DIRECT, DIRECT) if v1 else (dw1, db1)
            grads[id(self.fc2.weight)], grads[id(self.fc2.bias)] = (DIRECT, DIRECT) if v2 else (dw2, db2)
        g, layout = g_flat.reshape(feat.shape), 2
        gx = None
        split = ops.is_split(self.dtype)
        for li in range(len(self.layers) - 1, -1, -1):
            L = self.layers[li]
            cur, y32, mean, invstd, scale, shift, use_batch = saved[li]
            dgamma = dbeta = None
            vbn, accbn = None, False
            if L.bn is not None and want_params:
                vbn, accbn = target(L.bn.weight, L.bn.bias)
                if vbn:
                    dgamma, dbeta = vbn
                else:
                    dgamma = torch.empty(L.cout, dtype=torch.float32, device=dev)
                    dbeta = torch.empty(L.cout, dtype=torch.float32, device=dev)
            gy16 = ops.bn_lrelu_bwd(g, layout, y32, L.cout, scale, shift, mean, invstd, SLOPE, self.dtype, has_bn=L.bn is not None,
                                    train=use_batch, dgamma=dgamma, dbeta=dbeta, accumulate=accbn)
            if want_params:
                if L.bn is not None:
                    grads[id(L.bn.weight)], grads[id(L.bn.bias)] = (DIRECT, DIRECT) if vbn else (dgamma, dbeta)
                cin3 = 4 * L.cin if L.k4 else L.cin
                vc, accc = target(L.conv.weight, L.conv.bias)
                if vc and not L.k4:
                    ops.conv3x3_wgrad(cur, gy16, L.cout, cin3, split=split, dw=vc[0], db=vc[1], accumulate=accc)
                    grads[id(L.conv.weight)], grads[id(L.conv.bias)] = DIRECT, DIRECT
                elif vc:       # 4x4 stride-2 conv: the tensor cores produce dW of its 3x3 space-to-depth form; one gather puts it in place
                    dw, db = ops.conv3x3_wgrad(cur, gy16, L.cout, cin3, split=split)
                    if accc:
                        vc[0].add_(k3x3_to_k4s2(dw))
                        vc[1].add_(db)
                    else:
                        vc[0].copy_(k3x3_to_k4s2(dw))
                        vc[1].copy_(db)
                    grads[id(L.conv.weight)], grads[id(L.conv.bias)] = DIRECT, DIRECT
                else:
                    dw, db = ops.conv3x3_wgrad(cur, gy16, L.cout, cin3, split=split)
                    grads[id(L.conv.weight)] = k3x3_to_k4s2(dw) if L.k4 else dw
                    grads[id(L.conv.bias)] = db
            if li == 0:
                if want_input:
                    gx = torch.zeros((n, L.cin, H, W), dtype=torch.float32, device=dev)
                    ops.conv3x3(gy16, pkt[0], out_nchw=gx)
                break
            g = torch.empty((n, ops.logical_planes(cur, self.dtype), cur.shape[2], cur.shape[3], 8), dtype=torch.float32, device=dev)
            ops.conv3x3(gy16, pkt[li], out32=g)
            layout = 1 if L.k4 else 0
        for prm, view in pending:
            if prm.grad is None:
                prm.grad = view
        plist = [grads.get(id(p)) for p in self.params()] if want_params else None
        return gx, plist


    # ---- second-order pass of the gradient penalty (WGAN-GP) ----------------------------------------------------------------------
    def _zero_bias(self, n, dev):
        key = ('zb', str(dev))
        if key not in self._const or self._const[key].numel() < n:
            self._const[key] = torch.zeros(max(n, 1024), dtype=torch.float32, device=dev)
        return self._const[key]

    @torch.no_grad()
    def input_gradient(self, sv):
        """g = d(sum of the logits) / d(image): what GradientPenaltyLoss takes the norm of (models/modules/loss.py:271-273)"""
        saved, feat, h1, _ = sv
        ones = torch.ones((feat.shape[0], 1), dtype=torch.float32, device=feat.device)
        return self.backward(ones, sv, want_input=True, want_params=False)[0]

    @torch.no_grad()
    def second_order_param_grads(self, v, sv, gscale=1.0):
        """Parameter gradient of  s(theta) = sum_b D'(x; theta)[v]_b,  the critic's directional derivative along v (an image-shaped tensor,
        constant here).  With v = dL_gp/dg this is d L_gp / d theta, the double backward of models/modules/loss.py:271-278.  Two passes
        over the layers: the tangent forward (same conv launches without bias, BatchNorm's Jacobian, the primal LeakyReLU masks) and the
        backward over the (primal, tangent) pair (esr_bn_double_bwd + the dgrad / wgrad launches, each conv's dW = wgrad(a, yb) + wgrad(u, tb)).
        Returns the gradients in module.parameters() order, scaled by gscale."""
        saved, feat, h1, (H, W) = sv
        pk, pkt = self._packed()
        n, dev = feat.shape[0], feat.device
        split = ops.is_split(self.dtype)
        for L, (_, _, _, _, _, _, use_batch) in zip(self.layers, saved):
            if L.bn is not None and not use_batch:
                raise NotImplementedError('esr_b200: the gradient penalty through BatchNorm in eval mode (running statistics) is not built')
        # ---- tangent forward
        cur_u, _ = ops.pack_nchw(v.float().contiguous(), dtype=self.dtype)
        tang, fdot = [], None
        for li, L in enumerate(self.layers):
            last = li == len(self.layers) - 1
            cur_a, y32, mean, invstd, scale, shift, use_batch = saved[li]
            t32 = torch.empty_like(y32)
            ops.conv3x3(cur_u, pk[li], out32=t32, bias=self._zero_bias(pk[li].cout_pad, dev))
            nxt_k4 = (not last) and self.layers[li + 1].k4
            w16, w_nchw, c1, c2 = ops.bn_tangent_fwd(t32, y32, L.cout, scale, shift, mean, invstd, SLOPE, self.dtype, has_bn=L.bn is not None,
                                                     space_to_depth=nxt_k4, want16=not last, want_nchw=last)
            tang.append((cur_u, t32, c1, c2))
            cur_u, fdot = w16, w_nchw
        fdot = fdot.reshape(n, -1)
        W1, W2 = self.fc1.weight.detach(), self.fc2.weight.detach()
        t1 = ops.linear_fwd(fdot, W1, None)
        hdot = torch.where(h1 > 0, t1, t1 * SLOPE)          # tangent of the hidden LeakyReLU ([B, 100] glue)
        # ---- backward of s = sum_b (W2 hdot)_b over the (primal, tangent) pair; the primal adjoint is zero until the first BatchNorm
        grads = {}
        ones = torch.ones((n, 1), dtype=torch.float32, device=dev)
        hbar, dw2, _ = ops.linear_bwd(ones, None, hdot, W2, gscale=gscale)
        fbar, dw1, _ = ops.linear_bwd(hbar, h1, fdot, W1, slope=SLOPE, gscale=gscale)
        grads[id(self.fc2.weight)], grads[id(self.fc2.bias)] = dw2, torch.zeros_like(self.fc2.bias)
        grads[id(self.fc1.weight)], grads[id(self.fc1.bias)] = dw1, torch.zeros_like(self.fc1.bias)
        wb, zb, layout = fbar.reshape(feat.shape), None, 2
        for li in range(len(self.layers) - 1, -1, -1):
            L = self.layers[li]
            cur_a, y32, mean, invstd, scale, shift, use_batch = saved[li]
            cur_u, t32, c1, c2 = tang[li]
            dgamma = dbeta = None
            if L.bn is not None:
                dgamma, dbeta = torch.empty(L.cout, dtype=torch.float32, device=dev), torch.empty(L.cout, dtype=torch.float32, device=dev)
            tb16, yb16 = ops.bn_double_bwd(zb, wb, layout, y32, t32, L.cout, scale, shift, mean, invstd, c1, c2, SLOPE, self.dtype,
                                           has_bn=L.bn is not None, gscale=gscale, dgamma=dgamma, dbeta=dbeta)
            if L.bn is not None:
                grads[id(L.bn.weight)], grads[id(L.bn.bias)] = dgamma, dbeta
            cin3 = 4 * L.cin if L.k4 else L.cin
            dw, db = ops.conv3x3_wgrad(cur_a, yb16, L.cout, cin3, split=split, scale=gscale)
            ops.conv3x3_wgrad(cur_u, tb16, L.cout, cin3, split=split, scale=gscale, dw=dw, db=False, accumulate=True)
            grads[id(L.conv.weight)] = k3x3_to_k4s2(dw) if L.k4 else dw
            grads[id(L.conv.bias)] = db
            if li == 0:
                break
            lp = ops.logical_planes(cur_a, self.dtype)
            zb = torch.empty((n, lp, cur_a.shape[2], cur_a.shape[3], 8), dtype=torch.float32, device=dev)
            wb = torch.empty_like(zb)
            ops.conv3x3(yb16, pkt[li], out32=zb)
            ops.conv3x3(tb16, pkt[li], out32=wb)
            layout = 1 if L.k4 else 0
        return [grads.get(id(p)) for p in self.params()]


class _GradPenaltyFn(torch.autograd.Function):
    """L_gp = mean_b (||d(sum D(x)) / dx_b||_2 - 1)^2 and its gradient with respect to the critic's parameters, for a critic forward
    that ran on this engine (models/modules/loss.py:260-279; models/SRRaGAN_model.py:362-371).  `crit` ties the node behind the critic's
    forward; the saved activations of that forward are passed as `sv`."""

    @staticmethod
    def forward(ctx, crit, eng, sv, *params):
        gx = eng.input_gradient(sv)
        norms = gx.reshape(gx.shape[0], -1).norm(2, dim=1)
        ctx.eng, ctx.sv, ctx.n_params = eng, sv, len(params)
        ctx.save_for_backward(gx, norms)
        return ((norms - 1) ** 2).mean()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        gx, norms = ctx.saved_tensors
        b = gx.shape[0]
        v = gx * ((2.0 / b) * (norms - 1) / norms.clamp_min(1e-30)).view(b, 1, 1, 1)        # dL_gp / dg
        plist = ctx.eng.second_order_param_grads(v, ctx.sv)
        g = g.float()
        pg = tuple((plist[k] * g if (ctx.needs_input_grad[3 + k] and plist[k] is not None) else None) for k in range(ctx.n_params))
        return (None, None, None) + pg


def gradient_penalty(interp_crit):
    """fused WGAN-GP penalty for logits produced by this engine (None if `interp_crit` did not come from it)"""
    node = interp_crit.grad_fn
    if node is None or not hasattr(node, 'eng') or not hasattr(node, 'sv'):
        return None
    return _GradPenaltyFn.apply(interp_crit, node.eng, node.sv, *node.eng.params())


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng, *params):
        out, sv = eng.forward(x, save=True)
        ctx.eng, ctx.sv, ctx.n_params = eng, sv, len(params)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable      # a second differentiation (WGAN-GP) must fail loudly, not return zeros
    def backward(ctx, g):
        want_params = any(ctx.needs_input_grad[2:])
        gx, plist = ctx.eng.backward(g, ctx.sv, want_input=ctx.needs_input_grad[0], want_params=want_params)
        pg = tuple((plist[k] if (want_params and ctx.needs_input_grad[2 + k] and not isinstance(plist[k], str)) else None) for k in range(ctx.n_params))
        return (gx, None) + pg


def disc_forward(module, x):
    eng = module.engine()
    params = eng.params()
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
        return _DiscFn.apply(x, eng, *params)
    return eng.forward(x)
