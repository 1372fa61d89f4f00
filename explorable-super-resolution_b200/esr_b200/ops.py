"""Tensor-level wrappers over the C-ABI (esr_b200.lib).  PyTorch is plumbing here: it owns device memory
and the stream; every computation is a call into libesr_b200.so.

Planar-8 activation tensors are torch tensors of shape [N, planes, H, W, 8] (float16 / bfloat16 operands,
float32 residual trunk)."""
import ctypes as C

import torch

from . import lib as L

_TORCH2ESR = {torch.float16: L.ESR_F16, torch.bfloat16: L.ESR_BF16}

# Split precision ("parity mode", esr_dtype ESR_BF16X3): a 16-bit tensor [N, 2P, H, W, 8] holds P bf16 hi planes and P bf16 lo planes
# (lo = bf16(v - hi)); convs run x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with fp32 accumulation.  Engines select it by passing SPLIT
# where they would pass torch.float16 / torch.bfloat16.
SPLIT = 'bf16x3'


def is_split(dt):
    return dt == SPLIT


def elem_dtype(dt):
    return torch.bfloat16 if dt == SPLIT else dt


def esr_dtype(dt):
    return L.ESR_BF16X3 if dt == SPLIT else _TORCH2ESR[dt]


def alloc16(dt, n, planes, h, w, dev, zero=False):
    """16-bit planes tensor of `planes` logical planes in precision `dt` (split tensors hold twice the planes)"""
    shape = (n, planes * (2 if dt == SPLIT else 1), h, w, 8)
    return (torch.zeros if zero else torch.empty)(shape, dtype=elem_dtype(dt), device=dev)


def logical_planes(t, dt):
    return t.shape[1] // 2 if dt == SPLIT else t.shape[1]

# engines replay recorded launch sequences (LaunchPlan) when True; False launches every conv through its own host call
PLAN_REPLAY = True


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise L.EsrError("esr_b200 ops need CUDA tensors (there is no CPU fallback)")
        if t is not None and not t.is_contiguous():
            raise L.EsrError("esr_b200 ops need contiguous tensors")


def device_check():
    L.check(L.load().esr_device_check())


_cfg_cache = {}
_pack_scratch = {}


def _pack_config(cin_planes, cout, kcp, want_rows, edt=L.ESR_F16):
    """(packed bytes, cout_pad, row-image n-block, row-image bytes) - pure functions of the shape, memoised"""
    key = (cin_planes, cout, kcp, want_rows, edt == L.ESR_BF16X3)
    if key not in _cfg_cache:
        lib = L.load()
        cp = C.c_int(0)
        nbytes = int(lib.esr_conv3x3_packed_bytes_ex(cin_planes, cout, kcp, edt, C.byref(cp)))
        nbn, nb = C.c_int(0), C.c_size_t(0)
        ok = want_rows and lib.esr_conv3x3_rows_config_ex(cin_planes, cout, edt, C.byref(nbn), C.byref(nb)) == 0
        _cfg_cache[key] = (nbytes, cp.value, nbn.value if ok else 0, nb.value if ok else 0)
    return _cfg_cache[key]


def run_pack_queue(queue):
    """launch every queued weight re-packing with ONE host call (esr_pack_conv3x3_weights_batch)"""
    if not queue:
        return
    arr = (L.PackItem * len(queue))(*[q[0] for q in queue])
    failed = C.c_int(-1)
    lib = L.load()
    dev = queue[0][1].device
    need = int(lib.esr_pack_batch_scratch_bytes(len(queue)))
    key = (str(dev), torch.cuda.current_stream().cuda_stream)
    scratch = _pack_scratch.get(key)
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=dev)
        _pack_scratch[key] = scratch
    rc = lib.esr_pack_conv3x3_weights_batch(arr, len(queue), _ptr(scratch), scratch.numel(), _stream(), C.byref(failed))
    if rc != 0:
        raise L.EsrError('esr_b200 error %d packing conv %d: %s' % (rc, failed.value, L.load().esr_last_error().decode()))
    del queue[:]


class PackedConv:
    """Tensor-core image of one 3x3 conv's weights (+ padded fp32 bias).  `lead` = latent channels in front.  The buffers are
    allocated once; `repack` refreshes them in place after an optimizer step (recorded launch arguments stay valid).  With
    `queue` (a list) the packing launches are deferred to `run_pack_queue`."""

    def __init__(self, weight, bias, dtype=torch.float16, lead=0, transpose_flip=False, kcp=None, rows=True, queue=None):
        require_cuda(weight, bias)
        cout, cin = int(weight.shape[0]), int(weight.shape[1])
        assert tuple(weight.shape[2:]) == (3, 3), "only 3x3 kernels"
        self.split = dtype == SPLIT
        self.dtype = elem_dtype(dtype)
        self.esr_dtype = esr_dtype(dtype)
        self.lead = lead
        self.transpose_flip = bool(transpose_flip)
        self._w_shape = (cout, cin, 3, 3)
        lead_planes = (lead + 7) // 8 + (cin - lead + 7) // 8     # esr_conv3x3_cin_planes
        if transpose_flip:
            # dgrad: outputs are the forward inputs in plane space ([latent plane | rest]), inputs the forward outputs
            self.cout, self.cin = lead_planes * 8, cout
            self.cin_planes = (cout + 7) // 8
        else:
            self.cout, self.cin = cout, cin
            self.cin_planes = lead_planes
        if kcp is None:
            kcp = 4 if self.cin_planes >= 4 else 2
        self.kcp = kcp
        nbytes, self.cout_pad, self.rows_nbn, rows_bytes = _pack_config(self.cin_planes, self.cout, kcp, bool(rows), self.esr_dtype)
        self.wpacked = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
        self.bias = torch.empty(self.cout_pad, dtype=torch.float32, device=weight.device)
        # second image for the row-streaming kernel (used for images wide enough for 128-pixel strips)
        self.wrows = torch.empty(rows_bytes, dtype=torch.uint8, device=weight.device) if self.rows_nbn else None
        self._enqueue(weight, bias, queue)

    def _enqueue(self, weight, bias, queue):
        weight = weight.detach().float().contiguous()
        b = bias.detach().float().contiguous() if (bias is not None and not self.transpose_flip) else None
        it = L.PackItem()
        it.w, it.cout, it.cin, it.lead, it.kcp = weight.data_ptr(), self._w_shape[0], self._w_shape[1], self.lead, self.kcp
        it.dtype, it.transpose_flip = self.esr_dtype, int(self.transpose_flip)
        it.wpacked, it.bias_out, it.bias_in = self.wpacked.data_ptr(), self.bias.data_ptr(), (b.data_ptr() if b is not None else None)
        it.wpacked_rows, it.rows_nbn = (self.wrows.data_ptr() if self.wrows is not None else None), self.rows_nbn
        entry = (it, weight, b)          # the fp32 sources stay referenced until the launches are issued
        if queue is None:
            run_pack_queue([entry])
        else:
            queue.append(entry)

    def repack(self, weight, bias, queue=None):
        """refresh the packed images in place from new weight values (same shape, same device)"""
        require_cuda(weight, bias)
        assert tuple(weight.shape) == self._w_shape and weight.device == self.wpacked.device
        self._enqueue(weight, bias, queue)


def conv3x3(x16, pc, *, in_plane_off=0, cin_planes=None, lrelu=False, slope=0.2, alpha=1.0,
            res1=None, res1_off=0, beta1=1.0, res2=None, res2_off=0, beta2=1.0, res3=None, res3_off=0, beta3=1.0,
            out16=None, out16_off=0, up2=False, pixel_shuffle=0, out32=None, out32_off=0,
            out_nchw=None, lead_planes=0, lead_acc=None, mask16=None, mask_off=0, mask_slope=0.2, tail_first=0,
            tile_p=0, tile_mt=0, rows=True, plan=None, bias=None):
    """One fused conv launch.  x16: [N, planes, H, W, 8] operand tensor.  With `plan` (a list) the filled argument struct is
    appended instead of launched: `LaunchPlan` replays such a recording with one host call."""
    require_cuda(x16, res1, res2, res3, out16, out32, out_nchw, lead_acc, mask16)
    n, pt, h, w, e = x16.shape
    assert e == 8 and x16.dtype == pc.dtype   # tcgen05 kind::f16 wants both operands in one format (f16 x bf16 traps)
    a = L.ConvArgs()
    a.n, a.h, a.w, a.dtype = n, h, w, pc.esr_dtype
    a.in_, a.in_planes_total, a.in_plane_off = x16.data_ptr(), pt, in_plane_off
    a.cin_planes = pc.cin_planes if cin_planes is None else cin_planes
    assert a.cin_planes == pc.cin_planes or not pc.split      # the packed image's chunk structure is that of its own cin
    a.wpacked, a.bias = pc.wpacked.data_ptr(), (pc.bias if bias is None else bias).data_ptr()      # bias: override (e.g. zeros for a tangent)
    assert bias is None or (bias.dtype == torch.float32 and bias.numel() >= pc.cout_pad and bias.is_cuda)
    a.cout, a.cout_pad, a.kcp = pc.cout, pc.cout_pad, pc.kcp
    a.lrelu, a.slope, a.alpha = int(lrelu), slope, alpha
    if res1 is not None:
        assert res1.dtype in (torch.float32, x16.dtype) and tuple(res1.shape[2:]) == (h, w, 8) and res1.shape[0] == n
        a.res1, a.res1_planes_total, a.res1_plane_off, a.beta1 = res1.data_ptr(), res1.shape[1], res1_off, beta1
        a.res1_is16 = int(res1.dtype != torch.float32)
    if res2 is not None:
        assert res2.dtype == torch.float32 and tuple(res2.shape[2:]) == (h, w, 8) and res2.shape[0] == n
        a.res2, a.res2_planes_total, a.res2_plane_off, a.beta2 = res2.data_ptr(), res2.shape[1], res2_off, beta2
    if res3 is not None:
        assert res3.dtype == torch.float32 and tuple(res3.shape[2:]) == (h, w, 8) and res3.shape[0] == n
        a.res3, a.res3_planes_total, a.res3_plane_off, a.beta3 = res3.data_ptr(), res3.shape[1], res3_off, beta3
    if lead_planes:
        assert lead_acc is not None and lead_acc.dtype == torch.float32 and tuple(lead_acc.shape[2:]) == (h, w, 8)
        a.lead_planes, a.lead_acc, a.lead_planes_total = lead_planes, lead_acc.data_ptr(), lead_acc.shape[1]
    if mask16 is not None:
        assert mask16.dtype in _TORCH2ESR and tuple(mask16.shape[2:]) == (h, w, 8)   # only the sign is read
        a.mask16, a.mask_planes_total, a.mask_plane_off, a.mask_slope = mask16.data_ptr(), mask16.shape[1], mask_off, mask_slope
    a.tail_first_plane = tail_first
    if out16 is not None:
        f = 2 if up2 else (pixel_shuffle if pixel_shuffle else 1)
        assert out16.dtype == x16.dtype and tuple(out16.shape[2:]) == (f * h, f * w, 8) and out16.shape[0] == n
        a.out16, a.out16_planes_total, a.out16_plane_off = out16.data_ptr(), out16.shape[1], out16_off
        a.out16_up2, a.out16_pixel_shuffle = int(up2), int(pixel_shuffle)
    if out32 is not None:
        assert out32.dtype == torch.float32 and tuple(out32.shape[2:]) == (h, w, 8) and out32.shape[0] == n
        a.out32, a.out32_planes_total, a.out32_plane_off = out32.data_ptr(), out32.shape[1], out32_off
    if out_nchw is not None:
        assert out_nchw.dtype == torch.float32 and out_nchw.shape[0] == n and tuple(out_nchw.shape[2:]) == (h, w)
        a.out_nchw, a.out_nchw_c = out_nchw.data_ptr(), out_nchw.shape[1]
    a.tile_p, a.tile_mt = tile_p, tile_mt
    if pc.wrows is not None and rows:   # rows: True = library decides by width, 'force' = always, False = tile kernel
        a.wpacked_rows, a.rows_nbn, a.rows_mode = pc.wrows.data_ptr(), pc.rows_nbn, (1 if rows == 'force' else 0)
    if plan is not None:
        plan.append(a)
        return
    L.check(L.load().esr_conv3x3_fwd(C.byref(a), _stream()))


class LaunchPlan:
    """A recorded sequence of conv launches with fixed arguments (cached activation buffers, packed weights) replayed by ONE
    host call (esr_conv3x3_fwd_batch).  The caller keeps every tensor the structs point at alive and drops the plan when any
    of them is re-allocated."""

    def __init__(self, recorded):
        self.n = len(recorded)
        self.array = (L.ConvArgs * self.n)(*recorded)
        self._failed = C.c_int(-1)

    def run(self):
        rc = L.load().esr_conv3x3_fwd_batch(self.array, self.n, _stream(), C.byref(self._failed))
        if rc != 0:
            raise L.EsrError('esr_b200 error %d in planned launch %d: %s' % (rc, self._failed.value, L.load().esr_last_error().decode()))


_wgrad_ws = {}
_wgrad_ws_size = {}


def _cin_planes(cin, lead):
    return (lead + 7) // 8 + (cin - lead + 7) // 8     # esr_conv3x3_cin_planes


def _wgrad_ws_bytes(cp, cout):
    key = (cp, cout)
    if key not in _wgrad_ws_size:
        _wgrad_ws_size[key] = int(L.load().esr_conv3x3_wgrad_workspace(cp, cout))
    return _wgrad_ws_size[key]


def conv3x3_wgrad(x16, gy16, cout, cin, *, lead=0, x_off=0, gy_off=0, dw=None, db=None, scale=1.0, accumulate=False, split=False):
    """Weight / bias gradient of a 3x3 conv: x16 = the conv's input planes, gy16 = gradient of its (pre-activation)
    output.  Returns (dw [cout,cin,3,3] fp32, db [cout] fp32); pass dw/db tensors to accumulate into them."""
    require_cuda(x16, gy16, dw, db if db is not False else None)
    lib = L.load()
    n, xpt, h, w, e = x16.shape
    assert e == 8 and gy16.dtype == x16.dtype and tuple(gy16.shape[2:]) == (h, w, 8) and gy16.shape[0] == n
    dev = x16.device
    if dw is None:
        dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=dev)
        accumulate = False
    if db is None:
        db = torch.empty((cout,), dtype=torch.float32, device=dev)
    elif db is False:      # no bias gradient wanted
        db = None
    assert dw.dtype == torch.float32 and tuple(dw.shape) == (cout, cin, 3, 3) and (db is None or db.dtype == torch.float32)
    assert dw.is_contiguous() and (db is None or db.is_contiguous())
    # wide convs are covered in input-channel slices of at most 208 channels (5 M chunks of (row, plane) groups)
    cp_all = _cin_planes(cin, lead)
    if cp_all <= 26:
        slices = [(0, cin, lead, x_off)]
    else:
        if lead:
            raise L.EsrError('conv3x3_wgrad: latent lead channels with more than 208 input channels are not supported')
        slices = [(c0, min(192, cin - c0), 0, x_off + c0 // 8) for c0 in range(0, cin, 192)]
    key = (str(dev), torch.cuda.current_stream().cuda_stream)
    for si, (c0, cs, ld, xo) in enumerate(slices):
        cp = _cin_planes(cs, ld)
        nbytes = _wgrad_ws_bytes(cp, cout)
        if nbytes == 0:
            raise L.EsrError('conv3x3_wgrad: (cin %d, cout %d) is not supported by the tensor-core tiling' % (cs, cout))
        ws = _wgrad_ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _wgrad_ws[key] = ws
        a = L.WgradArgs()
        a.n, a.h, a.w, a.dtype = n, h, w, (L.ESR_BF16X3 if split else _TORCH2ESR[x16.dtype])
        a.x, a.x_planes_total, a.x_plane_off = x16.data_ptr(), xpt, xo
        a.gy, a.gy_planes_total, a.gy_plane_off = gy16.data_ptr(), gy16.shape[1], gy_off
        a.cout, a.cin, a.lead = cout, cs, ld
        a.cin_total, a.cin_off = cin, c0
        a.dw, a.db, a.scale, a.accumulate = dw.data_ptr(), (db.data_ptr() if (si == 0 and db is not None) else None), scale, int(accumulate)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        L.check(lib.esr_conv3x3_wgrad(C.byref(a), _stream()))
    return dw, db


def sum_nchw(src, scale=1.0, out=None, accumulate=False):
    """per-channel sum of an NCHW fp32 tensor -> [C] fp32 (written or accumulated into `out` when given)"""
    require_cuda(src, out)
    n, c, h, w = src.shape
    if out is None:
        out, accumulate = torch.empty((c,), dtype=torch.float32, device=src.device), False
    L.check(L.load().esr_sum_nchw(_ptr(src), n, c, h, w, scale, int(accumulate), _ptr(out), _stream()))
    return out


def planes_for(c):
    return (c + 7) // 8


def pack_nchw(src, *, pad=0, dtype=torch.float16, dst16=None, dst32=None, plane_off=0, want16=True, want32=False):
    """NCHW fp32 -> planar-8 (optionally replicate padded).  Returns (dst16, dst32)."""
    require_cuda(src, dst16, dst32)
    src = src.float()
    n, c, h, w = src.shape
    planes = planes_for(c)
    ho, wo = h + 2 * pad, w + 2 * pad
    split = dtype == SPLIT
    if dst16 is None and want16:
        dst16 = alloc16(dtype, n, planes, ho, wo, src.device)
    if dst32 is None and want32:
        dst32 = torch.empty((n, planes, ho, wo, 8), dtype=torch.float32, device=src.device)
    pt = (dst16 if dst16 is not None else dst32).shape[1]
    if dst16 is not None and not split:
        dtype = dst16.dtype
    L.check(L.load().esr_pack_nchw(_ptr(src), n, c, h, w, pad, esr_dtype(dtype), _ptr(dst16), _ptr(dst32), pt, plane_off,
                                   _stream()))
    return dst16, dst32


def pack_nchw_affine(src, scale, shift, dtype=torch.float16):
    """NCHW fp32 -> 16-bit planes of src * scale[c] + shift[c] (input normalisation of the VGG feature extractor)"""
    require_cuda(src, scale, shift)
    src = src.float().contiguous()
    n, c, h, w = src.shape
    dst = alloc16(dtype, n, planes_for(c), h, w, src.device)
    L.check(L.load().esr_pack_nchw_affine(_ptr(src), n, c, h, w, _ptr(scale), _ptr(shift), esr_dtype(dtype), _ptr(dst), dst.shape[1], 0,
                                          _stream()))
    return dst


def maxpool2x2(src16, split=False):
    require_cuda(src16)
    n, pt, h, w, _ = src16.shape
    dst = torch.empty((n, pt, h // 2, w // 2, 8), dtype=src16.dtype, device=src16.device)
    L.check(L.load().esr_maxpool2x2_planes16(_ptr(src16), L.ESR_BF16X3 if split else _TORCH2ESR[src16.dtype], n, pt, h, w, _ptr(dst), _stream()))
    return dst


def maxpool2x2_bwd(gout16, act16, split=False):
    """gradient w.r.t. the pre-activation that fed ReLU -> MaxPool2d(2,2); act16 = the pooling's (post-ReLU) input"""
    require_cuda(gout16, act16)
    n, pt, h, w, _ = act16.shape
    assert gout16.dtype == act16.dtype and tuple(gout16.shape) == (n, pt, h // 2, w // 2, 8)
    gin = torch.empty_like(act16)
    L.check(L.load().esr_maxpool2x2_bwd_planes16(_ptr(gout16), _ptr(act16), L.ESR_BF16X3 if split else _TORCH2ESR[act16.dtype], n, pt, h, w,
                                                 _ptr(gin), _stream()))
    return gin


def unpack_planes(src, c, plane_off=0, split=False):
    """planar-8 -> NCHW fp32 (first c channels starting at plane_off; split tensors return hi + lo)."""
    require_cuda(src)
    n, pt, h, w, _ = src.shape
    dst = torch.empty((n, c, h, w), dtype=torch.float32, device=src.device)
    if src.dtype == torch.float32:
        L.check(L.load().esr_unpack_planes32(_ptr(src), n, c, h, w, pt, plane_off, _ptr(dst), _stream()))
    else:
        L.check(L.load().esr_unpack_planes16(_ptr(src), L.ESR_BF16X3 if split else _TORCH2ESR[src.dtype], n, c, h, w, pt, plane_off, _ptr(dst),
                                             _stream()))
    return dst


def upsample2x(src16):
    require_cuda(src16)
    n, pt, h, w, _ = src16.shape
    dst = torch.empty((n, pt, 2 * h, 2 * w, 8), dtype=src16.dtype, device=src16.device)
    L.check(L.load().esr_upsample2x_planes16(_ptr(src16), n, pt, h, w, _ptr(dst), _stream()))
    return dst


def pixel_unshuffle2(src16, dtype=None):
    """adjoint of nn.PixelShuffle(2) on 16-bit planes [N, P, 2h, 2w, 8] -> [N, 4P, h, w, 8] (split tensors: both halves)"""
    require_cuda(src16)
    n, pt, h2, w2, _ = src16.shape
    assert h2 % 2 == 0 and w2 % 2 == 0
    h, w = h2 // 2, w2 // 2
    split = dtype == SPLIT
    planes = pt // 2 if split else pt
    dst = torch.empty((n, 4 * pt, h, w, 8), dtype=src16.dtype, device=src16.device)
    for half in range(2 if split else 1):
        L.check(L.load().esr_pixel_unshuffle2_planes16(_ptr(src16), n, planes, h, w, pt, half * planes, _ptr(dst), 4 * pt, half * 4 * planes, _stream()))
    return dst


def latent_downscale(z_hr, s, pad_hr=0):
    require_cuda(z_hr)
    n, c, hh, wh = z_hr.shape
    out = torch.empty((n, c, (hh + 2 * pad_hr) // s, (wh + 2 * pad_hr) // s), dtype=torch.float32, device=z_hr.device)
    L.check(L.load().esr_latent_downscale(_ptr(z_hr), n, c, hh, wh, s, pad_hr, _ptr(out), _stream()))
    return out


def cem_down(g, s, phase, kv, kh, sub_from=None):
    """DownscaleOP (optionally fused `sub_from - Down(g)`).  kv/kh: [rank, len] fp32 device tensors."""
    require_cuda(g, kv, kh, sub_from)
    n, c, hh, wh = g.shape
    out = torch.empty((n, c, hh // s, wh // s), dtype=torch.float32, device=g.device)
    L.check(L.load().esr_cem_down(_ptr(g), n, c, hh, wh, s, phase, _ptr(kv), _ptr(kh), kv.shape[1], kv.shape[0],
                                  _ptr(sub_from), _ptr(out), _stream()))
    return out


def cem_inv(e, kv, kh):
    require_cuda(e, kv, kh)
    n, c, hl, wl = e.shape
    out = torch.empty_like(e)
    L.check(L.load().esr_cem_inv(_ptr(e), n, c, hl, wl, _ptr(kv), _ptr(kh), kv.shape[1], kv.shape[0], _ptr(out), _stream()))
    return out


def cem_up_add(f, g, s, phase, kv, kh, crop=0):
    require_cuda(f, g, kv, kh)
    n, c, hl, wl = f.shape
    out = torch.empty((n, c, hl * s - 2 * crop, wl * s - 2 * crop), dtype=torch.float32, device=f.device)
    L.check(L.load().esr_cem_up_add(_ptr(f), _ptr(g), n, c, hl, wl, s, phase, _ptr(kv), _ptr(kh), kv.shape[1], kv.shape[0],
                                    crop, _ptr(out), _stream()))
    return out


def downsum2x(src32, act16_hi=None, slope=0.2, dtype=torch.float16, want32=True, want16=True):
    """adjoint of nearest x2 on fp32 planes [N,P,2h,2w,8] (+ LeakyReLU derivative from the hi-res saved activation)."""
    require_cuda(src32, act16_hi)
    n, pt, h2, w2, _ = src32.shape
    h, w = h2 // 2, w2 // 2
    d32 = torch.empty((n, pt, h, w, 8), dtype=torch.float32, device=src32.device) if want32 else None
    d16 = alloc16(dtype, n, pt, h, w, src32.device) if want16 else None
    if act16_hi is not None:   # only the sign is read (the hi half of a split tensor)
        assert tuple(act16_hi.shape[2:]) == tuple(src32.shape[2:]) and act16_hi.shape[1] == pt * (2 if dtype == SPLIT else 1)
    L.check(L.load().esr_downsum2x_planes(_ptr(src32), n, pt, h, w, _ptr(act16_hi), slope, esr_dtype(dtype), _ptr(d32), _ptr(d16),
                                          _stream()))
    return d32, d16


def planes_add(a32, b32, dtype=torch.float16, want32=True, want16=True):
    require_cuda(a32, b32)
    assert a32.shape == b32.shape and a32.dtype == b32.dtype == torch.float32
    o32 = torch.empty_like(a32) if want32 else None
    o16 = alloc16(dtype, a32.shape[0], a32.shape[1], a32.shape[2], a32.shape[3], a32.device) if want16 else None
    L.check(L.load().esr_planes_add_ex(_ptr(a32), _ptr(b32), a32.numel() // 8, a32[0].numel() // 8, esr_dtype(dtype), _ptr(o32), _ptr(o16),
                                       _stream()))
    return o32, o16


def sep_adjoint_2d(gout, taps_v, taps_h, *, full_out, a_stride, c_off, n_in, n_store, m_stride=1, m_phase=0, crop=0,
                   sub_from=None, sub_crop=0):
    """Adjoint of a separable, clamp-addressed, strided 2-D filter (sum over `rank` separable terms).
    gout: [N,C,Ho_store,Wo_store] (logical output size full_out=(Ho,Wo); stored = logical cropped by `crop`).
    Returns the gradient at the stored input positions [N,C,n_store[0],n_store[1]]."""
    require_cuda(gout, taps_v, taps_h, sub_from)
    gout = gout.float().contiguous()
    n, c, hs, ws = gout.shape
    Ho, Wo = full_out
    lib = L.load()
    rank, ln = taps_v.shape
    total = None
    for r in range(rank):
        tmp = torch.empty((n, c, n_store[0], ws), dtype=torch.float32, device=gout.device)
        L.check(lib.esr_sep_adjoint_1d(_ptr(gout), n * c, 1, ws, Ho, crop, hs, n_in[0], n_store[0], a_stride, c_off, m_stride, m_phase,
                                       _ptr(taps_v[r]), ln, None, 0, _ptr(tmp), _stream()))
        out = torch.empty((n, c, n_store[0], n_store[1]), dtype=torch.float32, device=gout.device)
        last = r == rank - 1 and total is None
        L.check(lib.esr_sep_adjoint_1d(_ptr(tmp), n * c, n_store[0], 1, Wo, crop, ws, n_in[1], n_store[1], a_stride, c_off, m_stride,
                                       m_phase, _ptr(taps_h[r]), ln, _ptr(sub_from) if (sub_from is not None and rank == 1) else None,
                                       sub_crop, _ptr(out), _stream()))
        total = out if total is None else total + out
    if sub_from is not None and rank > 1:
        raise L.EsrError('sep_adjoint_2d: fused subtraction supports rank-1 filters only')
    return total


def latent_grad(gz_hr, gz_lr, n, c, hh, wh, s, pad_hr):
    require_cuda(gz_hr, gz_lr)
    dev = (gz_hr if gz_hr is not None else gz_lr).device
    dst = torch.empty((n, c, hh, wh), dtype=torch.float32, device=dev)
    L.check(L.load().esr_latent_grad(_ptr(gz_hr), _ptr(gz_lr), n, c, hh, wh, s, pad_hr, _ptr(dst), _stream()))
    return dst


# ---- discriminator pieces (BatchNorm2d + LeakyReLU on planes, space-to-depth, fully connected layers) ----------------------------
_bn_ws = {}


def _bn_workspace(dev, planes):
    nbytes = int(L.load().esr_bn_workspace_bytes(planes))
    key = (str(dev), torch.cuda.current_stream().cuda_stream)
    ws = _bn_ws.get(key)
    if ws is None or ws.numel() * 4 < nbytes:
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        _bn_ws[key] = ws
    return ws


def bn_stats(y32, c, gamma, beta, eps, momentum, train, running_mean, running_var):
    """batch (train) or running (eval) statistics of fp32 planes -> (save_mean, save_invstd, scale, shift), each [c] fp32;
    train=True also updates running_mean / running_var in place like nn.BatchNorm2d"""
    require_cuda(y32, gamma, beta, running_mean, running_var)
    n, p, h, w, _ = y32.shape
    assert y32.dtype == torch.float32
    out = torch.empty((4, c), dtype=torch.float32, device=y32.device)
    ws = _bn_workspace(y32.device, p)
    L.check(L.load().esr_bn_stats(_ptr(y32), n, p, h, w, c, _ptr(gamma), _ptr(beta), eps, momentum, int(train), _ptr(running_mean),
                                  _ptr(running_var), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), _ptr(ws), ws.numel() * 4, _stream()))
    return out[0], out[1], out[2], out[3]


def bn_lrelu_fwd(y32, c, scale, shift, slope, dtype, space_to_depth=False, want16=True, want_nchw=False):
    """LeakyReLU(y * scale[c] + shift[c]) -> (16-bit planes [plain | space-to-depth], NCHW fp32)"""
    require_cuda(y32, scale, shift)
    n, p, h, w, _ = y32.shape
    d16 = None
    if want16:
        d16 = alloc16(dtype, n, 4 * p, h // 2, w // 2, y32.device) if space_to_depth else alloc16(dtype, n, p, h, w, y32.device)
    dn = torch.empty((n, c, h, w), dtype=torch.float32, device=y32.device) if want_nchw else None
    L.check(L.load().esr_bn_lrelu_fwd(_ptr(y32), n, p, h, w, c, _ptr(scale), _ptr(shift), slope, esr_dtype(dtype), _ptr(d16),
                                      int(space_to_depth), _ptr(dn), _stream()))
    return d16, dn


def space_to_depth(src16):
    require_cuda(src16)
    n, p, h, w, _ = src16.shape
    dst = torch.empty((n, 4 * p, h // 2, w // 2, 8), dtype=src16.dtype, device=src16.device)
    L.check(L.load().esr_space_to_depth_planes16(_ptr(src16), n, p, h, w, _ptr(dst), _stream()))
    return dst


def bn_lrelu_bwd(g, g_layout, y32, c, scale, shift, mean, invstd, slope, dtype, *, has_bn, train=True, gscale=1.0, dgamma=None, dbeta=None,
                 accumulate=False, scratch=None):
    """gradient of the conv output y from the gradient g of LeakyReLU(BN(y)) -> gy16 planes; fills dgamma / dbeta when given.
    g_layout 0: fp32 planes like y, 1: fp32 planes of the space-to-depth image, 2: NCHW fp32."""
    require_cuda(g, y32, scale, shift, mean, invstd, dgamma, dbeta)
    n, p, h, w, _ = y32.shape
    assert g.dtype == torch.float32 and g.numel() == (n * c * h * w if g_layout == 2 else y32.numel())
    gy16 = alloc16(dtype, n, p, h, w, y32.device)
    if scratch is None:
        scratch = torch.zeros((2, c), dtype=torch.float32, device=y32.device)
    ws = _bn_workspace(y32.device, p) if has_bn else None
    L.check(L.load().esr_bn_lrelu_bwd(_ptr(g), g_layout, _ptr(y32), n, p, h, w, c, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), slope,
                                      int(has_bn), int(train), gscale, int(accumulate), _ptr(dgamma), _ptr(dbeta), _ptr(scratch[0]),
                                      _ptr(scratch[1]), esr_dtype(dtype), _ptr(gy16), _ptr(ws), (ws.numel() * 4 if ws is not None else 0),
                                      _stream()))
    return gy16


_bn_dbl_ws = {}


def bn_tangent_fwd(t32, y32, c, scale, shift, mean, invstd, slope, dtype, *, has_bn, space_to_depth=False, want16=True, want_nchw=False):
    """tangent of LeakyReLU(BatchNorm(y)) along the conv tangent t (WGAN-GP second-order pass) -> (w16 | None, w_nchw | None, c1, c2)"""
    require_cuda(t32, y32, scale, shift, mean, invstd)
    n, p, h, w, _ = y32.shape
    assert t32.shape == y32.shape and t32.dtype == y32.dtype == torch.float32
    d16 = None
    if want16:
        d16 = alloc16(dtype, n, 4 * p, h // 2, w // 2, y32.device) if space_to_depth else alloc16(dtype, n, p, h, w, y32.device)
    dn = torch.empty((n, c, h, w), dtype=torch.float32, device=y32.device) if want_nchw else None
    cc = torch.empty((2, c), dtype=torch.float32, device=y32.device)
    ws = _bn_workspace(y32.device, p) if has_bn else None
    L.check(L.load().esr_bn_tangent_fwd(_ptr(t32), _ptr(y32), n, p, h, w, c, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), slope, int(has_bn),
                                        _ptr(cc[0]), _ptr(cc[1]), esr_dtype(dtype), _ptr(d16), int(space_to_depth), _ptr(dn), _ptr(ws),
                                        (ws.numel() * 4 if ws is not None else 0), _stream()))
    return d16, dn, cc[0], cc[1]


def bn_double_bwd(zb, wb, g_layout, y32, t32, c, scale, shift, mean, invstd, c1, c2, slope, dtype, *, has_bn, gscale=1.0, dgamma=None, dbeta=None,
                  accumulate=False):
    """adjoints (tb16, yb16) of the conv's tangent and output from the adjoints of the tangent / primal activations; fills dgamma / dbeta"""
    require_cuda(zb, wb, y32, t32, scale, shift, mean, invstd, c1, c2, dgamma, dbeta)
    n, p, h, w, _ = y32.shape
    dev = y32.device
    tb16, yb16 = alloc16(dtype, n, p, h, w, dev), alloc16(dtype, n, p, h, w, dev)
    coef = torch.empty((5, c), dtype=torch.float32, device=dev)
    ws = None
    if has_bn:
        nbytes = int(L.load().esr_bn_dbl_workspace_bytes(p))
        key = (str(dev), torch.cuda.current_stream().cuda_stream)
        ws = _bn_dbl_ws.get(key)
        if ws is None or ws.numel() * 4 < nbytes:
            ws = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
            _bn_dbl_ws[key] = ws
    L.check(L.load().esr_bn_double_bwd(_ptr(zb), _ptr(wb), g_layout, _ptr(y32), _ptr(t32), n, p, h, w, c, _ptr(scale), _ptr(shift), _ptr(mean),
                                       _ptr(invstd), _ptr(c1), _ptr(c2), slope, int(has_bn), gscale, int(accumulate), _ptr(dgamma), _ptr(dbeta),
                                       _ptr(coef), esr_dtype(dtype), _ptr(tb16), _ptr(yb16), _ptr(ws), (ws.numel() * 4 if ws is not None else 0),
                                       _stream()))
    return tb16, yb16


def linear_fwd(x, weight, bias, lrelu=False, slope=0.2):
    require_cuda(x, weight, bias)
    b, k = x.shape
    j = weight.shape[0]
    assert x.dtype == weight.dtype == torch.float32 and weight.shape[1] == k
    out = torch.empty((b, j), dtype=torch.float32, device=x.device)
    L.check(L.load().esr_linear_fwd(_ptr(x), _ptr(weight), _ptr(bias), b, k, j, int(lrelu), slope, _ptr(out), _stream()))
    return out


def linear_bwd(g, act, x, weight, slope=0.2, want_gx=True, want_w=True, gscale=1.0, dw=None, db=None, accumulate=False):
    """-> (gx [B,K] | None, dW [J,K] | None, db [J] | None); act = the layer's LeakyReLU output (None for a linear layer).
    dw / db given: written (or accumulated) in place."""
    require_cuda(g, act, x, weight, dw, db)
    b, k = x.shape
    j = weight.shape[0]
    assert g.dtype == torch.float32 and tuple(g.shape) == (b, j)
    gx = torch.empty((b, k), dtype=torch.float32, device=x.device) if want_gx else None
    if want_w and dw is None:
        dw, db, accumulate = torch.empty((j, k), dtype=torch.float32, device=x.device), torch.empty((j,), dtype=torch.float32, device=x.device), False
    if not want_w:
        dw = db = None
    L.check(L.load().esr_linear_bwd(_ptr(g), _ptr(act), slope, _ptr(x), _ptr(weight), b, k, j, gscale, int(accumulate), _ptr(gx), _ptr(dw), _ptr(db),
                                    _stream()))
    return gx, dw, db


# ---- structure-tensor statistics of the latent-control loss ----------------------------------------------------------------------
def structure_tensor(img):
    """[N,C,H,W] fp32 -> [N,3] per-image means of (dx^2, dy^2, dx*dy) of the 2x2 finite differences"""
    require_cuda(img)
    n, c, h, w = img.shape
    assert img.dtype == torch.float32
    out = torch.empty((n, 3), dtype=torch.float32, device=img.device)
    ws = torch.empty(int(L.load().esr_structure_tensor_workspace_bytes(n)) // 4, dtype=torch.float32, device=img.device)
    L.check(L.load().esr_structure_tensor_fwd(_ptr(img), n, c, h, w, _ptr(out), _ptr(ws), ws.numel() * 4, _stream()))
    return out


def structure_tensor_bwd(img, g):
    require_cuda(img, g)
    n, c, h, w = img.shape
    assert g.dtype == torch.float32 and tuple(g.shape) == (n, 3)
    grad = torch.empty_like(img)
    L.check(L.load().esr_structure_tensor_bwd(_ptr(img), _ptr(g), n, c, h, w, _ptr(grad), _stream()))
    return grad
