"""RRDB generator engine: turns an `RRDBNet` parameter container into a sequence of fused conv launches.

Reference control flow being replaced: architecture.py:278-302 (RRDBNet.forward), block.py:85-97
(ShortcutBlock), block.py:262-270 (RRDB), block.py:230-235 (ResidualDenseBlock_5C), block.py:299-300 (Upsampler).

Data layout in HBM (all planar-8, see include/esr_b200.h):
  * "dense" operand buffers of z+nf+4*gc channels: conv i of a dense block reads the prefix [z | x | x1..x_i] and writes
    x_{i+1} into its own plane range, so block.py:234's torch.cat never happens.  Inference rotates three of them;
    when a backward pass will follow, every dense block keeps its own (the saved activations ARE these buffers);
  * fp32 "trunk" buffers carry the residual stream x in full precision at RRDB boundaries; inside an RRDB the
    0.2-scaled dense-block residuals use the 16-bit copy of x (the tensor-core operand);
  * conv5 of each dense block fuses `x5*0.2 + x`; the third block of an RRDB also fuses the RRDB residual:
        out = (0.2*acc + x_rdb3)*0.2 + x_rrdb = 0.04*acc + 0.2*x_rdb3 + x_rrdb
  * LR_conv fuses the ShortcutBlock add and writes its output already nearest-x2 replicated; every upconv does the
    same for the next one, so the up-sampled tensor is written once and never re-read for resizing.

Backward (input gradient: what Z_optimization.py:747 needs; plus weight gradients for training): every conv's dgrad is the same tcgen05
kernel with transposed/rotated weights (`transpose_flip`), the dense-block gradient accumulates in one fp32 buffer
(in-place `res2 == out32`), LeakyReLU derivatives come from the saved 16-bit activations (`mask16`), and the latent
channels' gradient is accumulated by every launch into one plane (`lead_acc`).
"""
import os

import torch

from . import ops
from . import optim as flat

SLOPE = 0.2
DIRECT = 'direct'      # marker: this gradient was written straight into the parameter's registered flat-buffer view


def _param_version(mods):
    return tuple((m.weight._version, m.bias._version, m.weight.data_ptr()) for m in mods)


def combine_dense_backward_weights(W, a5, z, nf, gc):
    """Weights of the dense-block backward written as a dense block (see RRDBEngine.packed_bwd_dense).

    W[jj]: [R, cout_j, cin_j, 3, 3] weights of conv_{jj+1} of R dense blocks (block.py:196-242; input channels
    [z latent | nf | gc*(jj)]), a5: [R] output scale of each block.  Returns a list over i = 0..4 of [R, rows_i, K_i, 3, 3]
    tensors: the conv that maps the concatenated gradients [g_5 | g_4 | ... | g_{i+1}] (K_i channels) to the gradient of
    slice i (i = 0: [z rows padded to 8 | nf rows], i >= 1: the gc rows of x_i), taps rotated by 180 degrees, the block scale
    folded into the g_5 columns.  Pure tensor re-layout (no arithmetic except the scale): tested on CPU against autograd."""
    R = W[0].shape[0]
    tr = lambda t: t.permute(0, 2, 1, 3, 4).flip(3, 4)    # [R, o, c] -> [R, c, o], taps rotated by 180 degrees
    W = list(W)
    W[4] = W[4] * a5.view(R, 1, 1, 1, 1)
    out = []
    for i in range(5):
        cols = slice(0, z + nf) if i == 0 else slice(z + nf + (i - 1) * gc, z + nf + i * gc)
        wc = torch.cat([tr(W[jj][:, :, cols]) for jj in range(4, i - 1 if i > 0 else -1, -1)], dim=2)   # convs 5 .. i+1
        if i == 0 and z:
            wc = torch.cat([wc[:, :z], torch.zeros((R, 8 - z) + tuple(wc.shape[2:]), device=wc.device, dtype=wc.dtype), wc[:, z:]], dim=1)
        out.append(wc.contiguous())
    return out


class _Saved:
    """activations kept by a forward pass that a backward pass will use"""
    pass


class RRDBEngine:
    def __init__(self, net, dtype=torch.float16):
        self.net = net
        self.dtype = dtype
        self._packed = None
        self._packed_t = None
        self._packed_version = None
        self._packed_bd = None
        self._stale = {'p': False, 't': False, 'bd': False}
        self._bufs = {}
        self._plans = {}

    # ---------------------------------------------------------------- weights
    def _convs(self):
        net = self.net
        convs = [net.model[0]]
        shortcut = net.model[1]
        for blk in list(shortcut.sub)[:-1]:
            for rdb in (blk.RDB1, blk.RDB2, blk.RDB3):
                convs += [seq[0] for seq in rdb.convs]
        convs.append(shortcut.sub[-1])
        for up in net.upsamplers():
            convs.append(up[1] if net.upsample_mode == 'upconv' else up[0])
        convs += [net.model[-3], net.model[-1]]
        return convs

    def _leads(self, convs):
        z, n_up = self.net.z_lead, len(self.net.upsamplers())
        return [0 if (len(convs) - 2 - n_up <= i < len(convs) - 2) else z for i in range(len(convs))]

    def _check_version(self):
        convs = self._convs()
        ver = _param_version(convs)
        if ver != self._packed_version:
            # the packed buffers persist (recorded launch plans keep pointing at them): they are refreshed in place, every conv
            # of the network in one host call, the next time they are asked for
            self._packed_version = ver
            self._stale = {'p': True, 't': True, 'bd': True}
        return convs

    def _refresh(self, kind, existing, build, sources):
        """existing packed objects (flat list, None entries allowed) refreshed in place from `sources` [(weight, bias)], or built
        by `build(queue)` the first time / after the parameters moved to another device"""
        q = []
        flat = [pc for pc in (existing or []) if pc is not None]
        dev = next(w for w, _ in sources if w is not None).device
        if not flat or flat[0].wpacked.device != dev:
            existing = build(q)
        elif self._stale[kind]:
            for pc, (w, b) in zip(existing, sources):
                if pc is not None:
                    pc.repack(w, b, q)
        ops.run_pack_queue(q)
        self._stale[kind] = False
        return existing

    def packed(self):
        convs = self._check_version()
        leads = self._leads(convs)
        self._packed = self._refresh('p', self._packed,
                                     lambda q: [ops.PackedConv(c.weight, c.bias, dtype=self.dtype, lead=l, queue=q) for c, l in zip(convs, leads)],
                                     [(c.weight, c.bias) for c in convs])
        return self._packed

    def packed_t(self):
        """dgrad operands of the convs OUTSIDE the dense blocks: I/O swapped, taps rotated by 180 degrees, no bias
        (the dense blocks use `packed_bwd_dense`)"""
        convs = self._check_version()
        leads = self._leads(convs)
        nb = len(self.net.model[1].sub) - 1
        inside = lambda i: 1 <= i <= 15 * nb
        self._packed_t = self._refresh('t', self._packed_t,
                                       lambda q: [None if inside(i) else ops.PackedConv(c.weight, None, dtype=self.dtype, lead=l, transpose_flip=True, queue=q)
                                                  for i, (c, l) in enumerate(zip(convs, leads))],
                                       [(c.weight, None) for c in convs])
        return self._packed_t

    def packed_bwd_dense(self):
        """Backward of a dense block as a dense block.  With g_i the gradient of conv_i's pre-activation output
        (block.py:230-235: x_i = lrelu(conv_i(cat(x, x_1..x_{i-1})))),

            g_i = lrelu'(x_i) * sum_{j>i} conv_j^T(g_j)[slice x_i]        i = 4..1
            g_x =               sum_{j>=1} conv_j^T(g_j)[slice x (and z)]

        i.e. every gradient slice is ONE 3x3 conv over the concatenation [g_5 | g_4 | ... | g_{i+1}] with the combined
        weight  Wc_i[c, (j, o), ky, kx] = W_j[o, col_i(c), 2-ky, 2-kx]: the same growing-prefix structure as the forward
        (no fp32 partial-gradient buffer, FLOPs identical).  The block's output scale (0.2, or 0.04 for the third block of
        an RRDB) is folded into the W_5 rows.  Returns [block][i] -> PackedConv, i = 0 (g_x [+ latent rows in front, padded
        to one plane]) .. 4, blocks in forward order.  Built with a handful of batched tensor ops over all blocks."""
        convs = self._check_version()
        if self._packed_bd is not None and not self._stale['bd'] and self._packed_bd[0][0].wpacked.device == convs[0].weight.device:
            return self._packed_bd
        net = self.net
        z, nf, gc = net.z_lead, net.nf, net.gc
        nb = len(net.model[1].sub) - 1
        R = 3 * nb
        with torch.no_grad():
            W = [torch.stack([convs[1 + r * 5 + jj].weight.detach().float() for r in range(R)]) for jj in range(5)]  # [R, cout_j, cin_j, 3, 3]
            a5 = torch.tensor([0.04 if r % 3 == 2 else 0.2 for r in range(R)], device=W[0].device)
            combined = combine_dense_backward_weights(W, a5, z, nf, gc)
            flat_old = [pc for row in self._packed_bd for pc in row] if self._packed_bd is not None else None
            flat = self._refresh('bd', flat_old,
                                 lambda q: [ops.PackedConv(combined[i][r], None, dtype=self.dtype, queue=q) for r in range(R) for i in range(5)],
                                 [(combined[i][r], None) for r in range(R) for i in range(5)])
            out = [flat[r * 5:(r + 1) * 5] for r in range(R)]
        self._packed_bd = out
        return out

    # ---------------------------------------------------------------- buffers
    def buffers(self, n, h, w, dev, save):
        key = (n, h, w, str(dev), bool(save))
        b = self._bufs.get(key)
        if b is not None:
            return b
        net = self.net
        zp = 1 if net.z_lead else 0
        nfp, gcp = net.nf // 8, net.gc // 8
        nb = len(net.model[1].sub) - 1
        dense_planes = zp + nfp + 4 * gcp
        z16 = lambda n_, planes, hh, ww, _8: ops.alloc16(self.dtype, n_, planes, hh, ww, dev, zero=True)
        z32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        b = {
            'in16': z16(n, zp + 1, h, w, 8),
            'D': [z16(n, dense_planes, h, w, 8) for _ in range(3 * nb + 1 if save else 3)],
            'T': [z32(n, nfp, h, w, 8) for _ in range(2)],
            'F': z32(n, nfp, h, w, 8),
            'up': [],
        }
        s = 1
        for r in net.up_factors():
            s *= r
            b['up'].append(z16(n, nfp, h * s, w * s, 8))
        if net.upsample_mode == 'pixelshuffle' and save:      # LR_conv + skip, the first shuffle stage's input (kept for its wgrad)
            b['ps_in'] = z16(n, nfp, h, w, 8)
        b['hr_a'] = z16(n, nfp + zp, h * s, w * s, 8)
        b['hr_b'] = z16(n, nfp + zp, h * s, w * s, 8)
        if len(self._bufs) >= 2:  # keep at most two shapes alive (train + eval sizes)
            self._bufs.pop(next(iter(self._bufs)))
        self._bufs[key] = b
        return b

    def _dense(self, B, save, k, j):
        """operand buffer holding the input of dense block j (0..2) of RRDB k; j == 3 -> where the RRDB output goes"""
        if save:
            return B['D'][3 * k + j]
        a, b_ = (0, 1) if k % 2 == 0 else (1, 0)
        return B['D'][(a, b_, 2, b_)[j]]

    def _conv_sequence(self, B, pk, save, out, plan):
        """every conv launch of one forward pass, in order (recorded into `plan` when given, launched otherwise)"""
        net = self.net
        z = net.z_lead
        zp = 1 if z else 0
        nfp, gcp = net.nf // 8, net.gc // 8
        T, F = B['T'], B['F']
        nb = len(net.model[1].sub) - 1
        conv = (lambda *a, **k: ops.conv3x3(*a, plan=plan, **k)) if plan is not None else ops.conv3x3
        it = iter(pk)
        # fea_conv: no activation; fp32 copy kept for the ShortcutBlock add
        conv(B['in16'], next(it), out16=self._dense(B, save, 0, 0), out16_off=zp, out32=F)
        for k in range(nb):
            src_T = F if k == 0 else T[k % 2]
            dst_T = T[(k + 1) % 2]
            for j in range(3):
                Di, Do = self._dense(B, save, k, j), self._dense(B, save, k, j + 1)
                for i in range(4):
                    conv(Di, next(it), cin_planes=zp + nfp + i * gcp, lrelu=True, slope=SLOPE,
                                out16=Di, out16_off=zp + nfp + i * gcp)
                if j < 2:   # x5*0.2 + x, residual from the 16-bit copy of x in the block's own buffer
                    conv(Di, next(it), alpha=0.2, res1=Di, res1_off=zp, beta1=1.0, out16=Do, out16_off=zp)
                else:       # ... and the RRDB residual from the fp32 trunk
                    conv(Di, next(it), alpha=0.04, res1=Di, res1_off=zp, beta1=0.2, res2=src_T, beta2=1.0,
                                out16=Do, out16_off=zp, out32=dst_T)
        Dlast = self._dense(B, save, nb - 1, 3) if nb > 0 else self._dense(B, save, 0, 0)
        ups = B['up']
        factors = net.up_factors()
        if any(r != 2 for r in factors):
            raise NotImplementedError('esr_b200: only x2 up-sampling stages are built (scale 2/4/8)')
        if net.upsample_mode == 'upconv':
            # LR_conv + ShortcutBlock add, stored nearest-x2 replicated for the first upconv (block.py:299-300)
            conv(Dlast, next(it), cin_planes=zp + nfp, res1=F, beta1=1.0, out16=ups[0], up2=True)
            for k in range(len(factors)):
                if k < len(factors) - 1:
                    conv(ups[k], next(it), cin_planes=nfp, lrelu=True, slope=SLOPE, out16=ups[k + 1], up2=True)
                else:  # the last upconv feeds HR_conv0, which sees the HR latent in plane 0 of hr_a
                    conv(ups[k], next(it), cin_planes=nfp, lrelu=True, slope=SLOPE, out16=B['hr_a'], out16_off=zp)
        else:
            # pixelshuffle_block (block.py:278-291): conv(nf -> 4nf) -> PixelShuffle(2) -> act; the shuffle is the
            # store addressing of the conv epilogue, the (elementwise) activation is applied before it
            if save:
                tmp, tmp_off = B['ps_in'], 0
            else:
                tmp, tmp_off = (B['D'][2] if Dlast is not B['D'][2] else B['D'][0]), zp
            conv(Dlast, next(it), cin_planes=zp + nfp, res1=F, beta1=1.0, out16=tmp, out16_off=tmp_off)
            src, src_off = tmp, tmp_off
            for k in range(len(factors)):
                dst, dst_off = (ups[k], 0) if k < len(factors) - 1 else (B['hr_a'], zp)
                conv(src, next(it), in_plane_off=src_off, cin_planes=nfp, lrelu=True, slope=SLOPE, out16=dst, out16_off=dst_off,
                            pixel_shuffle=2)
                src, src_off = dst, dst_off
        conv(B['hr_a'], next(it), lrelu=True, slope=SLOPE, out16=B['hr_b'], out16_off=zp)
        conv(B['hr_b'], next(it), out_nchw=out)

    # ---------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, x, pad=0, save=False):
        """x: [N, z*s^2 + 3, h, w] fp32 NCHW (reference layout).  Returns G(x): [N, out_nc, S*(h+2pad), S*(w+2pad)]
        (and the saved-activation record when `save`)."""
        net = self.net
        ops.require_cuda(x)
        if net.norm_type is not None:
            raise NotImplementedError('esr_b200: norm layers inside RRDBNet are not supported')
        n, cin, h0, w0 = x.shape
        h, w = h0 + 2 * pad, w0 + 2 * pad
        dev = x.device
        pk = self.packed()
        B = self.buffers(n, h, w, dev, save)
        z = net.z_lead
        zp = 1 if z else 0
        nfp, gcp = net.nf // 8, net.gc // 8
        T, F = B['T'], B['F']
        S = net.upscale
        nb = len(net.model[1].sub) - 1

        x = x.float().contiguous()
        if z:
            # latent: [N, z*S^2, h, w] is a raw view of [N, z, S*h, S*w] (SRRaGAN_model.py:233, architecture.py:281-283)
            zc = cin - 3
            assert zc == z * S * S, 'latent channels do not match upscale^2 * num_latent_channels'
            z_hr = x[:, :zc].contiguous().view(n, z, S * h0, S * w0)
            # eval mode replicate-pads Z in the HR domain BEFORE the bilinear 1/S resize (CEMnet.py:290-292,
            # architecture.py:284), so the padding is folded into the resize kernel, not applied after it
            z_lr = ops.latent_downscale(z_hr, S, pad_hr=pad * S)
            img = x[:, zc:].contiguous()
            dt = self.dtype
            ops.pack_nchw(z_lr, dst16=B['in16'], plane_off=0, dtype=dt)
            ops.pack_nchw(img, pad=pad, dst16=B['in16'], plane_off=1, dtype=dt)
            for d in B['D']:
                ops.pack_nchw(z_lr, dst16=d, plane_off=0, dtype=dt)
            ops.pack_nchw(z_hr, pad=pad * S, dst16=B['hr_a'], plane_off=0, dtype=dt)
            ops.pack_nchw(z_hr, pad=pad * S, dst16=B['hr_b'], plane_off=0, dtype=dt)
        else:
            ops.pack_nchw(x, pad=pad, dst16=B['in16'], plane_off=0, dtype=self.dtype)

        out = torch.empty((n, net.out_nc, h * S, w * S), dtype=torch.float32, device=dev)
        if ops.PLAN_REPLAY:
            # the ~350 launches have fixed arguments (cached buffers, packed weights): replay the recorded structs with one
            # host call; only the output image is new
            key = (n, h, w, str(dev), bool(save))
            plan = self._plans.get(key)
            if plan is None or plan.bufs is not B or plan.pk is not pk:   # weights are re-packed in place: the plan survives
                rec = []
                self._conv_sequence(B, pk, save, out, rec)
                plan = ops.LaunchPlan(rec)
                plan.bufs, plan.pk = B, pk
                if len(self._plans) >= 2:
                    self._plans.pop(next(iter(self._plans)))
                self._plans[key] = plan
            plan.array[plan.n - 1].out_nchw = out.data_ptr()
            plan.run()
        else:
            self._conv_sequence(B, pk, save, out, None)
        if not save:
            return out
        sv = _Saved()
        sv.B, sv.n, sv.h, sv.w, sv.h0, sv.w0, sv.pad, sv.cin = B, n, h, w, h0, w0, pad, cin
        # the saved activations ARE the cached buffers: a later grad-enabled forward of the same shape overwrites them, so
        # each one takes a generation number and `backward` refuses a stale record instead of returning wrong gradients
        B['gen'] = sv.gen = B.get('gen', 0) + 1
        return out, sv

    # ---------------------------------------------------------------- backward
    def backward_input(self, g_out, sv):
        return self.backward(g_out, sv, wgrad=False)[0]

    @torch.no_grad()
    def backward(self, g_out, sv, wgrad=False):
        """g_out: dL/dG on the padded HR domain [N, out_nc, S*h, S*w].  Returns (dL/dx, param_grads): dL/dx with x's
        layout [N, z*S^2 + 3, h0, w0] (latent part exact; the LR-image part is returned only when pad == 0) and, when
        `wgrad`, a list of (dW [cout,cin,3,3], db [cout]) fp32 per conv in `_convs()` order (else None).

        Weight gradients (training, models/SRRaGAN_model.py:436-500): every conv's wgrad launch pairs the activation
        buffer its forward launch read (kept by `save=True`) with the 16-bit gradient of its pre-activation output,
        which the dgrad chain produces anyway as the operand of the next transposed conv."""
        net = self.net
        convs = self._convs()
        leads = self._leads(convs)
        grads = [None] * len(convs) if wgrad else None

        pending = []      # (parameter, flat view) pairs whose .grad must point at the view once the launches are issued

        def wg(idx, x16, gy16, gy_off=0, scale=1.0, bias_from_nchw=None):
            """weight / bias gradient of conv `idx`.  With a FlatAdam-registered parameter pair the launches write (or accumulate, when
            .grad already is that view: gradient accumulation, SRRaGAN_model.py:392-396) straight into the flat gradient buffer."""
            if not wgrad:
                return
            conv = convs[idx]
            cout, cin = int(conv.weight.shape[0]), int(conv.weight.shape[1])
            gw, gb = flat.grad_view(conv.weight), flat.grad_view(conv.bias)
            split = ops.is_split(self.dtype)
            own = lambda p, v: p.grad is None or p.grad is v or p.grad.data_ptr() == v.data_ptr()
            if gw is not None and gb is not None and own(conv.weight, gw) and own(conv.bias, gb) and (conv.weight.grad is None) == (conv.bias.grad is None):
                acc = conv.weight.grad is not None
                if bias_from_nchw is None:
                    ops.conv3x3_wgrad(x16, gy16, cout, cin, lead=leads[idx], gy_off=gy_off, scale=scale, split=split, dw=gw, db=gb, accumulate=acc)
                else:
                    ops.conv3x3_wgrad(x16, gy16, cout, cin, lead=leads[idx], gy_off=gy_off, scale=scale, split=split, dw=gw, db=False, accumulate=acc)
                    ops.sum_nchw(bias_from_nchw, out=gb, accumulate=acc)
                pending.extend([(conv.weight, gw), (conv.bias, gb)])
                grads[idx] = (DIRECT, DIRECT)
                return
            grads[idx] = ops.conv3x3_wgrad(x16, gy16, cout, cin, lead=leads[idx], gy_off=gy_off, scale=scale, split=split)
            if bias_from_nchw is not None:
                grads[idx] = (grads[idx][0], ops.sum_nchw(bias_from_nchw))

        B, n, h, w, pad = sv.B, sv.n, sv.h, sv.w, sv.pad
        if B.get('gen') != sv.gen:
            raise ops.L.EsrError('esr_b200: the activations saved for this backward were overwritten by a later grad-enabled forward of the '
                                 'same shape (the engine keeps ONE saved set per shape: run backward before the next forward)')
        dev = g_out.device
        wt = self.packed_t()
        z = net.z_lead
        zp = 1 if z else 0
        nfp, gcp = net.nf // 8, net.gc // 8
        S = net.upscale
        nb = len(net.model[1].sub) - 1
        n_up = len(net.up_factors())
        H, W = h * S, w * S
        # gradient operands share the activations' format (tcgen05 kind::f16 traps on f16 x bf16): training runs the whole
        # engine in bf16 because fp16 gradients underflow (a dense block's inner gradients sit 3-4 decades below the trunk's)
        gdt = self.dtype
        # every gradient buffer below is fully written by the launch that produces it before anything reads it: no zero-fill (at C2 the
        # memsets of the HR-resolution buffers alone were 10 GB per step).  ESR_POISON=1 fills them with NaN instead (tests).
        poison = os.environ.get('ESR_POISON', '0') == '1'

        zero = os.environ.get('ESR_ZERO_SCRATCH', '0') == '1'

        def f32(*s):
            t = (torch.zeros if zero else torch.empty)(s, dtype=torch.float32, device=dev)
            return t.fill_(float('nan')) if poison else t

        def f16(n_, planes, hh, ww, _8):
            t = ops.alloc16(gdt, n_, planes, hh, ww, dev, zero=zero)
            return t.fill_(float('nan')) if poison else t
        gz_hr = torch.zeros((n, 1, H, W, 8), dtype=torch.float32, device=dev) if z else None      # accumulated into by every launch that
        gz_lr = torch.zeros((n, 1, h, w, 8), dtype=torch.float32, device=dev) if z else None      # reads the latent plane
        lead = dict(lead_planes=zp, lead_acc=gz_hr) if z else {}
        idx_lr = 1 + 15 * nb
        idx_hr0 = idx_lr + 1 + n_up

        # HR_conv1^T, HR_conv0^T (LeakyReLU derivative of the conv below comes from its saved output)
        g_out = g_out.float().contiguous()
        g16, _ = ops.pack_nchw(g_out, dtype=gdt)
        wg(idx_hr0 + 1, B['hr_b'], g16, bias_from_nchw=g_out)    # the last conv's bias gradient from the unrounded dL/dG (heavy cancellation)
        g_b = f16(n, nfp, H, W, 8)
        ops.conv3x3(g16, wt[idx_hr0 + 1], mask16=B['hr_b'], mask_off=zp, mask_slope=SLOPE, out16=g_b, **lead)
        wg(idx_hr0, B['hr_a'], g_b)
        g_a = f16(n, nfp, H, W, 8)
        ops.conv3x3(g_b, wt[idx_hr0], mask16=B['hr_a'], mask_off=zp, mask_slope=SLOPE, out16=g_a, **lead)
        del g_b, g16
        cur = g_a
        g_t32 = None
        if net.upsample_mode == 'pixelshuffle':
            # pixelshuffle_block (block.py:278-291) in reverse: `cur` is the gradient of the stage's pre-activation output in the
            # shuffled layout (the consumer's launch applied the LeakyReLU mask, which commutes with the permutation): un-shuffle it to
            # the conv's 4nf output channels, then the ordinary transposed conv / weight gradient at the stage's input resolution
            for k in range(n_up - 1, -1, -1):
                hk, wk = h * 2 ** k, w * 2 ** k
                gyu = ops.pixel_unshuffle2(cur, dtype=gdt)
                src = B['ps_in'] if k == 0 else B['up'][k - 1]
                wg(idx_lr + 1 + k, src, gyu)
                if k > 0:   # the stage's input is the previous stage's activated output
                    cur = f16(n, nfp, hk, wk, 8)
                    ops.conv3x3(gyu, wt[idx_lr + 1 + k], mask16=src, mask_off=0, mask_slope=SLOPE, out16=cur)
                else:       # ... or LR_conv + fea, no activation
                    g_t32, cur = f32(n, nfp, h, w, 8), f16(n, nfp, h, w, 8)
                    ops.conv3x3(gyu, wt[idx_lr + 1], out32=g_t32, out16=cur)
                del gyu
        # upconvs: conv^T at the high resolution, then the adjoint of nearest x2 (2x2 sum)
        for k in (range(n_up - 1, -1, -1) if net.upsample_mode != 'pixelshuffle' else ()):
            hk, wk = h * 2 ** (k + 1), w * 2 ** (k + 1)
            wg(idx_lr + 1 + k, B['up'][k], cur)
            gu = f32(n, nfp, hk, wk, 8)
            ops.conv3x3(cur, wt[idx_lr + 1 + k], out32=gu)
            if k > 0:   # B['up'][k] is the (replicated) LeakyReLU output of upconv k-1
                _, cur = ops.downsum2x(gu, act16_hi=B['up'][k], slope=SLOPE, dtype=gdt, want32=False)
            else:       # B['up'][0] is LR_conv + fea, no activation
                g_t32, cur = ops.downsum2x(gu, dtype=gdt)
            del gu
        lead = dict(lead_planes=zp, lead_acc=gz_lr) if z else {}
        # LR_conv^T -> gradient w.r.t. the last RRDB's output
        wg(idx_lr, self._dense(B, True, nb - 1, 3) if nb > 0 else self._dense(B, True, 0, 0), cur)
        # dense blocks, in reverse.  Gd[t % 2] = [g_5 | g_4 | g_3 | g_2 | g_1] of block t; a block's closing launch writes the
        # next block's g_5 (the gradient arriving at that block's output) straight into the other buffer.
        wbd = self.packed_bwd_dense() if nb > 0 else None
        Gd = [f16(n, nfp + 4 * gcp, h, w, 8) for _ in range(2)]
        go32 = f32(n, nfp, h, w, 8)
        ops.conv3x3(cur, wt[idx_lr], out32=go32, out16=Gd[0], **lead)
        gy32 = [f32(n, nfp, h, w, 8) for _ in range(2)]
        gi32 = f32(n, nfp, h, w, 8)
        t = 0
        for k in range(nb - 1, -1, -1):
            for j in (2, 1, 0):
                Sb = self._dense(B, True, k, j)
                r = 3 * k + j
                base = 1 + r * 5
                G, Gn = Gd[t % 2], Gd[(t + 1) % 2]
                wg(base + 4, Sb, G, scale=0.04 if j == 2 else 0.2)          # conv5: its output gradient is scale * g_5
                for i in (4, 3, 2, 1):
                    off = nfp + (4 - i) * gcp
                    ops.conv3x3(G, wbd[r][i], cin_planes=off, mask16=Sb, mask_off=zp + nfp + (i - 1) * gcp, mask_slope=SLOPE,
                                out16=G, out16_off=off)
                    wg(base + i - 1, Sb, G, gy_off=off)
                # closing launch: g_x (+ the latent rows) = conv over all five gradients + what arrives at the block's output
                if j == 2:
                    ops.conv3x3(G, wbd[r][0], res3=go32, beta3=0.2, out32=gy32[1], out16=Gn, **lead)
                elif j == 1:
                    ops.conv3x3(G, wbd[r][0], res3=gy32[1], beta3=1.0, out32=gy32[0], out16=Gn, **lead)
                else:   # ... plus the RRDB skip connection
                    ops.conv3x3(G, wbd[r][0], res3=gy32[0], beta3=1.0, res1=go32, beta1=1.0, out32=gi32, out16=Gn, **lead)
                t += 1
            go32, gi32 = gi32, go32
        # ShortcutBlock: the fea_conv output feeds the first RRDB and the skip
        _, gf16 = ops.planes_add(go32, g_t32, dtype=gdt, want32=False)
        wg(0, B['in16'], gf16)
        g_img = torch.zeros((n, 3, h, w), dtype=torch.float32, device=dev)
        ops.conv3x3(gf16, wt[0], out_nchw=g_img, **lead)
        gx = torch.zeros((n, sv.cin, sv.h0, sv.w0), dtype=torch.float32, device=dev)
        if z:
            gz = ops.latent_grad(gz_hr, gz_lr, n, z, S * sv.h0, S * sv.w0, S, pad * S)
            gx[:, :sv.cin - 3] = gz.view(n, z * S * S, sv.h0, sv.w0)
        if pad == 0:
            gx[:, sv.cin - 3:] = g_img
        for prm, view in pending:
            if prm.grad is None:
                prm.grad = view
        return gx, grads
