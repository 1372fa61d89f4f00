"""Consistency-Enforcing Module — host side.

Mirror of the reference's CEM/CEMnet.py surface: `CEMnet` (numpy filter design + projection utilities),
`CEM_PyTorch` / `Filter_Layer` (the nn.Module that wraps a generator), `CEM_downsampler`, `Get_CEM_Conf`,
`Adjust_State_Dict_Keys`, `Return_kernel`.  Filter design is float64 numpy at construction time
(restating CEMnet.py:22-49,186-206); every tensor op at run time is a CUDA kernel behind the C-ABI
(esr_cem_down / esr_cem_inv / esr_cem_up_add) — there is no torch/cuDNN compute path and no CPU fallback.

Algebra used by the fused projection (CEMnet.py:303-310):
    out = Up(Inv(x)) + G - Up(Inv(Down(G)))  =  G + Up(Inv(x - Down(G)))      (all three ops are linear)
"""
import collections

import numpy as np
import torch
import torch.nn as nn
from scipy.signal import convolve2d as conv2

from CEM.imresize_CEM import imresize, calc_strides


def _cuda_or_cpu():
    return torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')


def _ops():
    from esr_b200 import ops
    return ops


def _separable_terms(k2d, tol=1e-7):
    """K = sum_r outer(v_r, h_r) (float64 SVD, terms kept until the residual is below tol * max|K|).
    The reference's bicubic ds_kernel and inv_hTh are rank 1 (SURVEY key fact 5)."""
    k2d = np.asarray(k2d, dtype=np.float64)
    u, s, vt = np.linalg.svd(k2d)
    v_terms, h_terms = [], []
    approx = np.zeros_like(k2d)
    for r in range(len(s)):
        v_terms.append(u[:, r] * np.sqrt(s[r]))
        h_terms.append(vt[r, :] * np.sqrt(s[r]))
        approx += np.outer(v_terms[-1], h_terms[-1])
        if np.abs(k2d - approx).max() <= tol * np.abs(k2d).max():
            break
    return np.stack(v_terms).astype(np.float32), np.stack(h_terms).astype(np.float32)


class CEMnet:
    NFFT_add = 36

    def __init__(self, conf, upscale_kernel=None):
        self.conf = conf
        self.ds_factor = np.array(conf.scale_factor, dtype=np.int32)
        assert np.round(self.ds_factor) == self.ds_factor, 'Currently only supporting integer scale factors'
        assert upscale_kernel is None or isinstance(upscale_kernel, str) or isinstance(upscale_kernel, np.ndarray), \
            'To support given kernels, change the Return_Invalid_Margin_Size_in_LR function and make sure everything else works'
        self.ds_kernel = Return_kernel(self.ds_factor, upscale_kernel=upscale_kernel)
        self.ds_kernel_invalidity_half_size_LR = self.Return_Invalid_Margin_Size_in_LR('ds_kernel', self.conf.filter_pertubation_limit)
        self.compute_inv_hTh()
        self.invalidity_margins_LR = 2 * self.ds_kernel_invalidity_half_size_LR + self.inv_hTh_invalidity_half_size
        self.invalidity_margins_HR = self.ds_factor * self.invalidity_margins_LR

    # ---- filter design (CEMnet.py:35-49,186-206) -------------------------------------------------
    def Return_Invalid_Margin_Size_in_LR(self, filter, max_allowed_perturbation):
        """How deep (in LR pixels) the response to a constant image is perturbed by the border by more than
        the allowed factor; searched along the central column and the central row."""
        TEST_IM_SIZE = 100
        assert filter in ['ds_kernel', 'inv_hTh']
        s = int(self.ds_factor)
        if filter == 'ds_kernel':
            resp = imresize(np.ones([s * TEST_IM_SIZE, s * TEST_IM_SIZE]), [1 / s], use_zero_padding=True)
        else:
            resp = conv2(np.ones([TEST_IM_SIZE, TEST_IM_SIZE]), self.inv_hTh, mode='same')
        mid = TEST_IM_SIZE // 2
        resp = resp / resp[mid, mid]
        resp[resp <= 0] = max_allowed_perturbation / 2
        invalid = np.exp(-np.abs(np.log(resp))) < max_allowed_perturbation
        deepest = [np.argwhere(invalid[:mid, mid])[-1][0] + 1, np.argwhere(invalid[mid, :mid])[-1][0] + 1]
        return np.max(deepest)

    def compute_inv_hTh(self):
        s = int(self.ds_factor)
        hTh = conv2(self.ds_kernel, np.rot90(self.ds_kernel, 2)) * s ** 2
        hTh = Aliased_Down_Sampling(hTh, s)
        pad = int(self.NFFT_add / 2)
        H = np.fft.fft2(np.pad(hTh, ((pad, pad), (pad, pad)), mode='constant', constant_values=0))
        # frequencies the kernel wipes out would blow up in the inverse: bound |H| from below
        H = H * np.maximum(1, self.conf.lower_magnitude_bound / np.abs(H))
        inv = np.real(np.fft.ifft2(1 / H))
        # re-centre on the maximum
        n = inv.shape[0]
        r, c = np.argmax(inv) // n, np.mod(np.argmax(inv), n)
        if not np.all(np.equal(np.ceil(np.array(inv.shape) / 2), np.array([r, c]) - 1)):
            half = np.min([n - r - 1, n - c - 1, r, c])
            inv = inv[r - half:r + half + 1, c - half:c + half + 1]
        self.inv_hTh = inv
        self.inv_hTh_invalidity_half_size = self.Return_Invalid_Margin_Size_in_LR('inv_hTh', self.conf.filter_pertubation_limit)
        drop = self.inv_hTh.shape[0] // 2 - self.Return_Invalid_Margin_Size_in_LR('inv_hTh', self.conf.desired_inv_hTh_energy_portion)
        if drop > 0:
            self.inv_hTh = self.inv_hTh[drop:-drop, drop:-drop]

    # ---- numpy utilities used by callers (GUI.py:1392,1401) ---------------------------------------
    def Pad_LR_Batch(self, batch, num_recursion=1):
        m = self.invalidity_margins_LR
        for _ in range(num_recursion):
            batch = 1.0 * np.pad(batch, pad_width=((0, 0), (m, m), (m, m), (0, 0)), mode='edge')
        return batch

    def Unpad_HR_Batch(self, batch, num_recursion=1):
        m = (self.ds_factor ** num_recursion) * self.invalidity_margins_LR * num_recursion
        return batch[:, m:-m, m:-m, :]

    def DT_Satisfying_Upscale(self, LR_image):
        margin = 2 * self.inv_hTh_invalidity_half_size + self.ds_kernel_invalidity_half_size_LR
        LR_image = Pad_Image(LR_image, margin)
        filtered = np.stack([conv2(LR_image[:, :, c], self.inv_hTh, mode='same') for c in range(LR_image.shape[-1])], -1)
        HR_image = imresize(filtered, scale_factor=[self.ds_factor])
        return Unpad_Image(HR_image, self.ds_factor * margin)

    def Enforce_DT_on_Image_Pair(self, LR_source, HR_input):
        same = [LR_source.shape[i] == HR_input.shape[i] for i in range(LR_source.ndim)]
        scaled = [self.ds_factor * LR_source.shape[i] == HR_input.shape[i] for i in range(LR_source.ndim)]
        assert np.all(np.logical_or(same, scaled))
        if len(same) == 2:
            LR_source, HR_input = np.expand_dims(LR_source, -1), np.expand_dims(HR_input, -1)
        LR_source = self.DT_Satisfying_Upscale(LR_source) if np.any(scaled) else self.Project_2_ortho_2_NS(LR_source)
        return HR_input - self.Project_2_ortho_2_NS(HR_input) + LR_source

    def Project_2_ortho_2_NS(self, HR_input):
        down = imresize(HR_input, scale_factor=[1 / self.ds_factor])
        if down.ndim < HR_input.ndim:
            down = np.reshape(down, list(HR_input.shape[:2] // self.ds_factor) + ([HR_input.shape[2]] if HR_input.ndim > 2 else []))
        return self.DT_Satisfying_Upscale(down)

    # ---- torch wrapping (CEMnet.py:66-91) ---------------------------------------------------------
    def WrapArchitecture_PyTorch(self, generated_image=None, training_patch_size=None, only_padders=False, grayscale=False):
        mLR = int(self.invalidity_margins_LR)
        mHR = int(self.ds_factor * mLR)
        self.LR_padder = torch.nn.ReplicationPad2d((mLR, mLR, mLR, mLR))
        self.HR_padder = torch.nn.ReplicationPad2d((mHR, mHR, mHR, mHR))
        self.HR_unpadder = lambda x: x[:, :, mHR:-mHR, mHR:-mHR]
        self.LR_unpadder = lambda x: x[:, :, mLR:-mLR, mLR:-mLR]
        self.loss_mask = None
        if training_patch_size is not None:
            mask = np.zeros([1, 1, training_patch_size, training_patch_size])
            m = int(self.invalidity_margins_HR)
            mask[:, :, m:-m, m:-m] = 1
            assert np.mean(mask) > 0, 'Loss mask completely nullifies image.'
            print('Using only only %.3f of patch area for learning. The rest is considered to have boundary effects' % (np.mean(mask)))
            self.loss_mask = torch.from_numpy(mask).float().to(_cuda_or_cpu())
        if only_padders:
            return
        wrapped = CEM_PyTorch(self, generated_image, grayscale=grayscale)
        self.OP_names = [m[0] for m in wrapped.named_modules() if 'Filter_OP' in m[0]]
        return wrapped

    def Mask_Invalid_Regions_PyTorch(self, im1, im2):
        assert self.loss_mask is not None, 'Mask not defined, probably didn''t pass patch size'
        return self.loss_mask * im1, self.loss_mask * im2


class Filter_Layer(nn.Module):
    """One fixed CEM filter.  Keeps the reference's parameter (`Filter_OP.weight`, [C,1,k,k], frozen, tagged
    `filter_layer`) so state dicts and `weights_init_kaiming`'s skip rule are unchanged, and runs the
    matching CUDA kernel.  `kind` in {'inv', 'up', 'down'}."""

    def __init__(self, filter, kind, ds_factor=1, phase=0, num_channels=3):
        super(Filter_Layer, self).__init__()
        filter = np.ascontiguousarray(filter)
        self.Filter_OP = nn.Conv2d(in_channels=num_channels, out_channels=num_channels, kernel_size=filter.shape, bias=False,
                                   groups=num_channels)
        w = torch.from_numpy(np.tile(filter[None, None], reps=[num_channels, 1, 1, 1])).float().to(_cuda_or_cpu())
        self.Filter_OP.weight = nn.Parameter(data=w, requires_grad=False)
        self.Filter_OP.filter_layer = True
        self.kind, self.ds_factor, self.phase = kind, int(ds_factor), int(phase)
        kv, kh = _separable_terms(filter)
        self.register_buffer('_kv', torch.from_numpy(kv).contiguous(), persistent=False)
        self.register_buffer('_kh', torch.from_numpy(kh).contiguous(), persistent=False)

    def _taps(self, dev):
        if self._kv.device != dev:
            self._kv, self._kh = self._kv.to(dev), self._kh.to(dev)
        return self._kv, self._kh

    def forward(self, x, sub_from=None, add_to=None, crop=0):
        ops = _ops()
        x = x.float().contiguous()
        kv, kh = self._taps(x.device)
        if self.kind == 'down':
            return ops.cem_down(x, self.ds_factor, self.phase, kv, kh, sub_from=sub_from)
        if self.kind == 'inv':
            return ops.cem_inv(x, kv, kh)
        return ops.cem_up_add(x, add_to, self.ds_factor, self.phase, kv, kh, crop=crop)


class CEM_PyTorch(nn.Module):
    def __init__(self, CEMnet, generated_image, grayscale=False):
        super(CEM_PyTorch, self).__init__()
        num_channels = 1 if grayscale else 3
        self.ds_factor = CEMnet.ds_factor
        self.conf = CEMnet.conf
        self.using_SR_model = generated_image is not None
        if self.using_SR_model:
            self.generated_image_model = generated_image
        s = int(CEMnet.ds_factor)
        pre_stride, post_stride = calc_strides(None, s)
        assert pre_stride[0] == pre_stride[1]
        self.Conv_LR_with_Inv_hTh_OP = Filter_Layer(CEMnet.inv_hTh, 'inv', num_channels=num_channels)
        self.Upscale_OP = Filter_Layer(CEMnet.ds_kernel * s ** 2, 'up', ds_factor=s, phase=pre_stride[0], num_channels=num_channels)
        self.DownscaleOP = Filter_Layer(np.rot90(CEMnet.ds_kernel, 2), 'down', ds_factor=s, phase=pre_stride[0], num_channels=num_channels)
        self.LR_padder = CEMnet.LR_padder
        self.HR_padder = CEMnet.HR_padder
        self.HR_unpadder = CEMnet.HR_unpadder
        self.LR_unpadder = CEMnet.LR_unpadder
        self.invalidity_margins_LR = int(CEMnet.invalidity_margins_LR)
        self.invalidity_margins_HR = int(CEMnet.invalidity_margins_HR)
        self.pre_pad = False  # flag instead of a forward() argument, as in the reference (DataParallel)
        self.return_2_components = 'decomposed_output' in self.conf.__dict__ and self.conf.decomposed_output

    def project(self, x_lr, generated_image, crop=0):
        """out = G + Up(Inv(x - Down(G))), optionally cropped by `crop` HR pixels per side (3 launches)."""
        e = self.DownscaleOP(generated_image, sub_from=x_lr.float().contiguous())
        f = self.Conv_LR_with_Inv_hTh_OP(e)
        return self.Upscale_OP(f, add_to=generated_image.float().contiguous(), crop=crop)

    def project_backward(self, g_out, hr_full, crop=0):
        """Gradient of `project` w.r.t. (generated_image, x_lr):  with P = Up.Inv (linear),
              g_x = P^T g,   g_G = g - Down^T g_x,      g = crop-adjoint of g_out
        Every factor is the exact adjoint of the clamp-addressed (replicate padded) forward filter, built from the
        1-D adjoint kernel esr_sep_adjoint_1d.  hr_full = (H, W) of the un-cropped HR domain."""
        ops = _ops()
        s = int(self.ds_factor)
        Hh, Wh = hr_full
        hl, wl = Hh // s, Wh // s
        dev = g_out.device
        up, inv, down = self.Upscale_OP, self.Conv_LR_with_Inv_hTh_OP, self.DownscaleOP
        kv, kh = up._taps(dev)
        r = kv.shape[1] // 2
        t = ops.sep_adjoint_2d(g_out, kv, kh, full_out=(Hh, Wh), a_stride=1, c_off=-r, n_in=(Hh, Wh), n_store=(hl, wl),
                               m_stride=s, m_phase=up.phase, crop=crop)                       # Up^T
        kv, kh = inv._taps(dev)
        r = kv.shape[1] // 2
        g_x = ops.sep_adjoint_2d(t, kv, kh, full_out=(hl, wl), a_stride=1, c_off=-r, n_in=(hl, wl), n_store=(hl, wl))   # Inv^T
        kv, kh = down._taps(dev)
        r = kv.shape[1] // 2
        g_G = ops.sep_adjoint_2d(g_x, kv, kh, full_out=(hl, wl), a_stride=s, c_off=down.phase - r, n_in=(Hh, Wh), n_store=(Hh, Wh),
                                 sub_from=g_out.float().contiguous(), sub_crop=crop)       # g - Down^T g_x
        return g_G, g_x

    def forward(self, x):
        return_2_components = self.return_2_components and not self.pre_pad
        if torch.is_grad_enabled() and self._needs_grad(x):
            if (not self.using_SR_model) or self.conf.sigmoid_range_limit or return_2_components:
                raise NotImplementedError('esr_b200: backward is built for the fused CEM(G(x)) projection only')
            from esr_b200.autograd import cem_generator_forward_with_grad
            return cem_generator_forward_with_grad(self, x)
        mLR = self.invalidity_margins_LR
        if self.using_SR_model:
            # eval mode pads LR by the invalidity margin before G (CEMnet.py:286-295); the generator mirror
            # folds the replicate padding into its input packing kernel.
            generated_image = self.generated_image_model(x, pad=mLR if self.pre_pad else 0)
        else:
            generated_image, x = x[1], x[0]
            if self.pre_pad:
                generated_image = self.HR_padder(generated_image)
        x = x[:, -3:, :, :]
        if self.pre_pad:
            x = self.LR_padder(x)
        assert np.all(np.mod(generated_image.size()[2:], int(self.ds_factor)) == 0)
        if self.conf.sigmoid_range_limit or return_2_components:
            ortho = self.Upscale_OP(self.Conv_LR_with_Inv_hTh_OP(x))
            NS = generated_image - self.Upscale_OP(self.Conv_LR_with_Inv_hTh_OP(self.DownscaleOP(generated_image)))
            if self.conf.sigmoid_range_limit:
                NS = torch.tanh(NS) * (self.conf.input_range[1] - self.conf.input_range[0])
            output = [ortho, NS] if return_2_components else ortho + NS
            return self.HR_unpadder(output) if self.pre_pad else output
        return self.project(x, generated_image, crop=self.invalidity_margins_HR if self.pre_pad else 0)

    def _needs_grad(self, x):
        xs = x if isinstance(x, (list, tuple)) else [x]
        if any(isinstance(t, torch.Tensor) and t.requires_grad for t in xs):
            return True
        return self.using_SR_model and any(p.requires_grad for p in self.generated_image_model.parameters())

    def train(self, mode=True):
        super(CEM_PyTorch, self).train(mode=mode)
        self.pre_pad = not mode
        return self

    def Image_2_Sigmoid_Range_Converter(self, images, opposite_direction=False):
        lo, hi = self.conf.input_range[0], self.conf.input_range[1]
        if opposite_direction:
            return images * (hi - lo) + lo
        return (torch.clamp(images, min=lo, max=hi) - lo) / (hi - lo)

    def Inverse_Sigmoid(self, images):
        p = self.Image_2_Sigmoid_Range_Converter(images)
        return torch.log(p / (1. - p))


def Aliased_Down_Sampling(array, factor):
    pre_stride, post_stride = calc_strides(array, 1 / factor, align_center=True)
    if array.ndim > 2:
        return array[pre_stride[0]::factor, pre_stride[1]::factor, :]
    return array[pre_stride[0]::factor, pre_stride[1]::factor]


def Return_kernel(ds_factor, upscale_kernel=None):
    up = imresize(None, [ds_factor, ds_factor], return_upscale_kernel=True, kernel=upscale_kernel)
    return np.rot90(up, 2).astype(np.float32) / (ds_factor ** 2)


def Pad_Image(image, margin_size):
    padding = ((margin_size, margin_size), (margin_size, margin_size)) + (((0, 0),) if image.ndim == 3 else ())
    return np.pad(image, pad_width=padding, mode='edge')


def Unpad_Image(image, margin_size):
    return image[margin_size:-margin_size, margin_size:-margin_size, :]


def Get_CEM_Conf(sf):
    class conf:
        scale_factor = sf
        avoid_skip_connections = False
        generate_HR_image = False
        pseudo_CEM_supplement = False
        desired_inv_hTh_energy_portion = 1 - 1e-6
        filter_pertubation_limit = 0.999
        sigmoid_range_limit = False
        lower_magnitude_bound = 0.01  # lower bound on |FFT(hTh)|
    return conf


def Adjust_State_Dict_Keys(loaded_state_dict, current_state_dict):
    """Checkpoints of a bare generator get the `generated_image_model.` prefix when the current network is
    CEM-wrapped; the CEM's own filter tensors are taken from the current network (CEMnet.py:403-412)."""
    wrapped_now = all(('generated_image_model' in k or 'Filter' in k) for k in current_state_dict.keys())
    wrapped_then = any('generated_image_model' in k for k in loaded_state_dict.keys())
    if wrapped_now and not wrapped_then:
        out = collections.OrderedDict()
        for k in loaded_state_dict:
            out['generated_image_model.' + k] = loaded_state_dict[k]
        for k in [k for k in current_state_dict.keys() if 'Filter' in k]:
            out[k] = current_state_dict[k]
        return out
    return loaded_state_dict


class CEM_downsampler(nn.Module):
    """HR -> LR with the CEM's down-sampling kernel, replicate padded against border artefacts
    (CEMnet.py:414-428).  Input [N,C,H,W] float tensor on the GPU."""

    def __init__(self, ds_factor, grayscale=False, differentiable=False):
        super(CEM_downsampler, self).__init__()
        self.CEM = CEMnet(Get_CEM_Conf(ds_factor))
        self.CEM.invalidity_margins_LR = 1 * self.CEM.ds_kernel_invalidity_half_size_LR
        self.CEM = self.CEM.WrapArchitecture_PyTorch(grayscale=grayscale)
        if not differentiable:
            self.CEM.eval()

    def forward(self, input):
        padded_HR = self.CEM.HR_padder(input)
        return self.CEM.LR_unpadder(self.CEM.DownscaleOP(padded_HR))
