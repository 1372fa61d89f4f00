"""VGG19 feature extractor engine (models/modules/architecture.py:658-724): the perceptual-loss network of the
training step (models/SRRaGAN_model.py:448-451, `netF(var_H).detach()` and `netF(fake_H)`).

Forward: input normalisation fused into the layout conversion, every `Conv2d(3x3) + ReLU` is one fused conv launch
(LeakyReLU with slope 0) on 16-bit planes, `MaxPool2d(2,2)` its own small kernel; the feature map (conv5_4 before its
ReLU for feature_layer=34) leaves as NCHW fp32.  Backward (w.r.t. the input image only - the extractor is frozen): the
same conv kernels with transposed weights, ReLU derivatives from the saved activations in the epilogue, pooling
backward fused with the ReLU in front of it, 1/std folded into the first conv's transposed weights."""
import torch

from . import ops

VGG19_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512, 'M']


def vgg19_layers(feature_layer):
    """[(kind, index_in_features, cin, cout)] of torchvision's vgg19().features[:feature_layer + 1]"""
    out, idx, cin = [], 0, 3
    for v in VGG19_CFG:
        if v == 'M':
            out.append(('pool', idx, cin, cin))
            idx += 1
        else:
            out.append(('conv', idx, cin, v))
            out.append(('relu', idx + 1, v, v))
            idx += 2
            cin = v
    return [l for l in out if l[1] <= feature_layer]


class VGGEngine:
    def __init__(self, module, dtype=torch.float16):
        self.m = module
        self.dtype = dtype
        self.layers = vgg19_layers(module.feature_layer)
        self._ver = None
        self._pk = self._pkt = None

    def _convs(self):
        return [self.m.features[i] for kind, i, _, _ in self.layers if kind == 'conv']

    def _packed(self):
        convs = self._convs()
        ver = tuple((c.weight._version, c.weight.data_ptr(), c.bias._version) for c in convs) + (self.m.std.data_ptr(),)
        if ver != self._ver:
            self._pk = [ops.PackedConv(c.weight, c.bias, dtype=self.dtype) for c in convs]
            self._pkt = []
            for k, c in enumerate(convs):
                w = c.weight.detach().float()
                if k == 0 and self.m.use_input_norm:   # d/dx of (x - mean) / std
                    w = w / self.m.std.view(1, -1, 1, 1).to(w.device)
                self._pkt.append(ops.PackedConv(w, None, dtype=self.dtype, transpose_flip=True))
            self._ver = ver
        return self._pk, self._pkt

    @torch.no_grad()
    def forward(self, x, save=False):
        x = x.float().contiguous()      # callers pass HR_unpadder crops (views)
        ops.require_cuda(x)
        m = self.m
        n, c, h, w = x.shape
        if c != 3:
            raise ops.L.EsrError('VGGFeatureExtractor expects 3-channel images')
        pk, _ = self._packed()
        dev = x.device
        if m.use_input_norm:
            scale, shift = (1.0 / m.std).reshape(-1).float().to(dev), (-m.mean / m.std).reshape(-1).float().to(dev)
        else:
            scale, shift = torch.ones(3, device=dev), torch.zeros(3, device=dev)
        cur = ops.pack_nchw_affine(x, scale.contiguous(), shift.contiguous(), dtype=self.dtype)
        saved = []          # per layer: the layer's input planes
        out = None
        k = 0
        for li, (kind, idx, cin, cout) in enumerate(self.layers):
            saved.append(cur)
            if kind == 'conv':
                hh, ww = cur.shape[2], cur.shape[3]
                relu_next = li + 1 < len(self.layers) and self.layers[li + 1][0] == 'relu'
                if relu_next:
                    o = ops.alloc16(self.dtype, n, ops.planes_for(cout), hh, ww, dev)
                    ops.conv3x3(cur, pk[k], lrelu=True, slope=0.0, out16=o)
                    cur = o
                else:   # the feature map itself: pre-activation, fp32 NCHW
                    out = torch.empty((n, cout, hh, ww), dtype=torch.float32, device=dev)
                    ops.conv3x3(cur, pk[k], out_nchw=out)
                    cur = None
                k += 1
            elif kind == 'pool':
                if cur.shape[2] % 2 or cur.shape[3] % 2:
                    raise ops.L.EsrError('VGGFeatureExtractor: image size must be divisible by 2 at every pooling stage')
                cur = ops.maxpool2x2(cur, split=ops.is_split(self.dtype))
            # 'relu' is fused into the conv in front of it
        if out is None:   # feature_layer ends on a ReLU / pooling layer
            out = ops.unpack_planes(cur, self.layers[-1][3], split=ops.is_split(self.dtype))
        return (out, saved) if save else out

    @torch.no_grad()
    def backward_input(self, g_feat, saved):
        """d(loss)/d(image) given d(loss)/d(features)"""
        _, pkt = self._packed()
        layers = self.layers
        n = g_feat.shape[0]
        dev = g_feat.device
        # loss scaling: the gradient of a mean-reduced feature loss is ~1e-6 per element, below fp16's range.  The chain is
        # linear in g, so scale it to a peak of 1024 on the way in and undo it on the image gradient (device-side, no sync).
        g_feat = g_feat.float()
        scale = 1024.0 / g_feat.abs().max().clamp_min(1e-30)
        g, _ = ops.pack_nchw((g_feat * scale).contiguous(), dtype=self.dtype)
        li = len(layers) - 1
        k = len(pkt) - 1
        if layers[li][0] != 'conv':
            raise NotImplementedError('esr_b200: backward through a VGG feature map taken after ReLU / pooling is not built')
        while li >= 0:
            kind, idx, cin, cout = layers[li]
            assert kind == 'conv'
            inp = saved[li]
            hh, ww = inp.shape[2], inp.shape[3]
            if li == 0:
                gx = torch.zeros((n, 3, hh, ww), dtype=torch.float32, device=dev)
                ops.conv3x3(g, pkt[k], out_nchw=gx)
                return gx / scale
            prev = layers[li - 1][0]
            if prev == 'relu':      # conv <- relu <- conv: mask with the (post-ReLU) activation this conv read
                o = ops.alloc16(self.dtype, n, ops.planes_for(cin), hh, ww, dev)
                ops.conv3x3(g, pkt[k], mask16=inp, mask_slope=0.0, out16=o)
                g = o
                li -= 2
            else:                   # conv <- pool <- relu <- conv
                o = ops.alloc16(self.dtype, n, ops.planes_for(cin), hh, ww, dev)
                ops.conv3x3(g, pkt[k], out16=o)
                g = ops.maxpool2x2_bwd(o, saved[li - 1], split=ops.is_split(self.dtype))
                li -= 3
            k -= 1


class _VGGFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng):
        out, saved = eng.forward(x, save=True)
        ctx.eng, ctx.saved = eng, saved
        return out

    @staticmethod
    def backward(ctx, g):
        return ctx.eng.backward_input(g.contiguous(), ctx.saved), None


def vgg_forward(module, x):
    eng = module.engine()
    if torch.is_grad_enabled() and x.requires_grad:
        return _VGGFn.apply(x, eng)
    return eng.forward(x)
