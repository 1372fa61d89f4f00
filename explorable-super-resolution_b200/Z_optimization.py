"""Latent-exploration inner loop with the reference's surface (Z_optimization.py): `Optimizable_Z`, `ArcTanH`,
`TV_Loss`, `Z_optimizer`.

The loop itself (Z_optimization.py:647-797) is host code, as in the reference: Adam over a tanh-parametrised latent map,
one generator+CEM forward WITH gradient per iteration, loss on the clamped output, backward to Z only.  What changes is
underneath `model.test(prevent_grads_calc=False)` / `Z_loss.backward()`: a single autograd node that runs the fused
tcgen05 forward launches and, in backward, the dgrad launches + the exact CEM adjoint (esr_b200.autograd) instead of
~1800 autograd nodes over cuDNN calls.

Objectives built: 'l1' (optionally masked), 'TV', 'max_STD' / 'min_STD' / 'STD_increase' / 'STD_decrease' (global, or 'local_' over
7x7 patches through ReturnPatchExtractionMat), 'Mag' (local magnitude), 'hist' / 'dict' (+ 'patch', 'noDC', 'no_localSTD', 'localSTD':
SoftHistogramLoss on the esr_soft_hist kernels), 'VGG' (perceptual distance through the VGG19 engine), 'Adversarial' (the critic
engine), 'random_l1' (+ '_limited'), 'periodicity' (integer or 'nonInt' sub-pixel periods, optionally 'Plus' an STD increase), 'scribble'
(masked l1 towards the scribbled colours, brightened / darkened regions through the V channel, total variation inside the smoothing
scribbles).  desired_SVD and digit raise NotImplementedError."""
import time

import numpy as np
import os

import torch
import torch.nn.functional as F


def _dev():
    return torch.device('cuda')


class Optimizable_Z(torch.nn.Module):
    """Z = Z_range * tanh(pre_tanh_Z), optionally frozen to its initial value outside Z_mask (Z_optimization.py:273-319)."""

    def __init__(self, Z_shape, Z_range=None, initial_pre_tanh_Z=None, Z_mask=None, random_perturbations=False):
        super(Optimizable_Z, self).__init__()
        self.Z = torch.nn.Parameter(data=torch.zeros(Z_shape, dtype=torch.float32, device=_dev()))
        self.mask = None
        if Z_mask is not None and not np.all(Z_mask):
            self.mask = torch.from_numpy(np.asarray(Z_mask)).float().to(self.Z.device)
            self.initial_pre_tanh_Z = (1 * initial_pre_tanh_Z).float().to(self.Z.device)
        if initial_pre_tanh_Z is not None:
            assert initial_pre_tanh_Z.size()[1:] == self.Z.data.size()[1:] and (initial_pre_tanh_Z.size(0) in [1, self.Z.data.size(0)]), \
                'Initilizer size does not match desired Z size'
            if random_perturbations:
                initial_pre_tanh_Z = initial_pre_tanh_Z + 0.001 * torch.randn_like(initial_pre_tanh_Z)
            self.Z.data[:initial_pre_tanh_Z.size(0), ...] = initial_pre_tanh_Z.to(self.Z.device)
        self.Z_range = Z_range

    def forward(self):
        if self.Z_range is not None:
            big = torch.finfo(self.Z.dtype).max
            self.Z.data.clamp_(-big, big)     # in place (same values as the reference's re-assignment): the storage must stay
        if self.mask is not None:             # put for the CUDA-graph replay of the iteration
            self.Z.data.copy_(self.mask * self.Z.data + (1 - self.mask) * self.initial_pre_tanh_Z)
        return self.Z_range * torch.tanh(self.Z) if self.Z_range is not None else self.Z

    def PreTanhZ(self):
        if self.mask is not None:
            return self.mask * self.Z.data + (1 - self.mask) * self.initial_pre_tanh_Z
        return self.Z.data

    def Randomize_Z(self, what_2_shuffle):
        assert what_2_shuffle in ['all', 'allButFirst']
        torch.nn.init.xavier_uniform_(self.Z.data if what_2_shuffle == 'all' else self.Z.data[1:], gain=100)

    def Return_Detached_Z(self):
        return self.forward().detach()

    def Assign_Z(self, Z):
        self.Z.data = 1 * Z


def ArcTanH(input_tensor):
    eps = torch.finfo(input_tensor.dtype).eps
    return 0.5 * torch.log((1 + input_tensor + eps) / (1 - input_tensor + eps))


def TV_Loss(image):
    return (image[:, :, :, :-1] - image[:, :, :, 1:]).abs().mean(dim=(1, 2, 3)) + (image[:, :, :-1, :] - image[:, :, 1:, :]).abs().mean(dim=(1, 2, 3))


def _dilate_rect(mask, k):
    """grey-scale dilation with a k x k rectangle, as cv2.dilate(mask, np.ones([k, k])) computes it (Z_optimization.py:358, the editable
    region of the non-local mode): anchor at the centre (k // 2), i.e. out[y, x] = max over dy, dx in [-(k // 2), k - 1 - k // 2] of
    mask[y + dy, x + dx], pixels outside the image ignored."""
    mask = np.asarray(mask)
    a = k // 2
    padded = np.full((mask.shape[0] + k - 1, mask.shape[1] + k - 1), -np.inf, dtype=np.float64)
    padded[a:a + mask.shape[0], a:a + mask.shape[1]] = mask
    out = np.full(mask.shape, -np.inf, dtype=np.float64)
    for dy in range(k):
        rows = padded[dy:dy + mask.shape[0]]
        for dx in range(k):
            out = np.maximum(out, rows[:, dx:dx + mask.shape[1]])
    return out.astype(mask.dtype)


def _half_open(shift, negative=False):
    """slice bound that drops |shift| rows / columns from one side (utils/util.py:260-264)"""
    if negative:
        return shift if shift < 0 else None
    return shift if shift > 0 else None


def Return_Translated_SubImage(image, translation):
    """the part of `image` that overlaps its own copy translated by (dy, dx) (utils/util.py:266-273)"""
    dy, dx = translation[0], translation[1]
    return image[:, :, _half_open(dy):_half_open(dy, True), _half_open(dx):_half_open(dx, True)]


def Return_Interpolated_SubImage(image, grid):
    """bilinear samples of `image` on a normalised grid, for sub-pixel periods (utils/util.py:276-277)"""
    return F.grid_sample(image, grid.repeat([image.size(0), 1, 1, 1]))


_UNBUILT = ['desired_SVD', 'digit']


class _SoftHistFn(torch.autograd.Function):
    """esr_soft_hist_fwd / _bwd: x [D, P] fp64 -> hist [B] (mean_p E[p][b]) or, dictionary mode, [P] (-log mean_b E[p][b])"""

    @staticmethod
    def forward(ctx, x, bins, vmax, eps, temperature, dictionary):
        import ctypes as C
        from esr_b200 import lib as L
        x, bins = x.double().contiguous(), bins.double().contiguous()
        if not x.is_cuda:
            raise L.EsrError('esr_b200: SoftHistogramLoss needs CUDA tensors (there is no CPU fallback)')
        D, P = x.shape
        B = bins.shape[1]
        out = torch.empty(P if dictionary else B, dtype=torch.float64, device=x.device)
        sum_e = torch.empty(P, dtype=torch.float64, device=x.device) if dictionary else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        L.check(L.load().esr_soft_hist_fwd(ptr(x), D, P, ptr(bins), B, float(vmax), float(eps), float(temperature), int(dictionary), ptr(out), ptr(sum_e),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        ctx.save_for_backward(x, bins, sum_e if dictionary else out)
        ctx.cfg = (float(vmax), float(eps), float(temperature), bool(dictionary))
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes as C
        from esr_b200 import lib as L
        x, bins, aux = ctx.saved_tensors
        vmax, eps, temperature, dictionary = ctx.cfg
        D, P = x.shape
        g = g.double().contiguous()
        dx = torch.empty_like(x)
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        L.check(L.load().esr_soft_hist_bwd(ptr(x), D, P, ptr(bins), bins.shape[1], vmax, eps, temperature, ptr(None if dictionary else g),
                                           ptr(g if dictionary else None), ptr(aux if dictionary else None), ptr(dx),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return dx, None, None, None, None, None


class SoftHistogramLoss(torch.nn.Module):
    """Kernel-density histogram / patch-dictionary objective of the GUI's imprinting tools (Z_optimization.py:24-230): the grey-level
    (or patch) distribution of the edited region is pulled towards that of a desired image region - KL divergence between soft
    histograms (`hist`), or mean distance to the nearest desired patch (`dict`).  Same constructor and semantics as the reference for
    grey-scale inputs (the only mode Z_optimizer uses, :536-539); the O(dims x pixels x bins) distance tensor the reference
    materialises in double precision is one CUDA kernel here (esr_soft_hist_fwd / _bwd, fp64 as well).  The automatic temperature
    search differentiates through the generator's backward (a double backward) and is not built."""

    def __init__(self, bins, min, max, desired_hist_image_mask=None, desired_hist_image=None, gray_scale=True, input_im_HR_mask=None, patch_size=1,
                 automatic_temperature=False, image_Z=None, temperature=0.05, dictionary_not_histogram=False, no_patch_DC=False, no_patch_STD=False):
        super(SoftHistogramLoss, self).__init__()
        if automatic_temperature:
            raise NotImplementedError('esr_b200 SoftHistogramLoss: the automatic temperature search needs a double backward through the generator')
        if not gray_scale:
            raise NotImplementedError('esr_b200 SoftHistogramLoss: colour histograms are not built (Z_optimizer only asks for gray_scale=True)')
        assert no_patch_DC or not no_patch_STD, 'Not supporting removing of only patch STD without DC'
        self.exp_power, self.SQRT_EPSILON = 2, 1e-7
        self.device = _dev()
        self.bin_width = (max - min) / (bins - 1)          # min / max are the CENTRES of the first / last bin
        self.max = max
        self.no_patch_DC, self.no_patch_STD = no_patch_DC, no_patch_STD
        self.temperature = float(temperature)
        self.gray_scale, self.patch_size = gray_scale, patch_size
        self.dictionary_not_histogram = dictionary_not_histogram
        self.num_dims = 1
        self.KDE = patch_size > 1           # kernel density estimation over the desired samples instead of a fixed-bin histogram
        self.bins = torch.linspace(min, max, bins).view(1, -1).double().to(self.device)        # [dims, bins]
        if desired_hist_image is not None:
            desired_hist_image = [im.mean(1, keepdim=True).view([-1, 1]) for im in desired_hist_image]
        if patch_size > 1:
            assert desired_hist_image is not None, 'Not supporting patch histograms for model training loss for now'
            self.num_dims = patch_size ** 2
            overlap = (self.num_dims - patch_size) / self.num_dims        # an entire patch but one row / column
            mats = [ReturnPatchExtractionMat(m, patch_size=patch_size, device=self.device, patches_overlap=overlap) for m in desired_hist_image_mask]
            desired = torch.cat([torch.sparse.mm(mats[i], desired_hist_image[i].to(self.device)).view([self.num_dims, -1, 1])
                                 for i in range(len(desired_hist_image))], 1)
            if no_patch_DC:
                desired = desired - torch.mean(desired, dim=0, keepdim=True)
                if no_patch_STD:
                    std = torch.max(torch.std(desired, dim=0, keepdim=True), other=torch.tensor(1 / 255).to(self.device))
                    self.mean_patches_STD = 1 * std.mean().item()
                    desired = desired / std * self.mean_patches_STD      # keeps the dynamic range, hence the kernel support
            self.desired_hist_image_mask = None
            desired_hist_image = desired
        else:
            if desired_hist_image is not None and len(desired_hist_image) > 1:
                print('Not supproting multiple hist image versions for non-patch histogram/dictionary. Removing extra image versions.')
            if desired_hist_image is not None:
                desired_hist_image = 1 * desired_hist_image[0].view([self.num_dims, -1, 1]).to(self.device)
            m = desired_hist_image_mask[0] if desired_hist_image_mask is not None else None
            self.desired_hist_image_mask = torch.from_numpy(np.asarray(m)).view([-1]).bool().to(self.device) if m is not None else None
        if self.KDE:
            if self.desired_hist_image_mask is not None:
                desired_hist_image = desired_hist_image[:, self.desired_hist_image_mask, :]
            self.bins = self.Desired_Im_2_Bins(desired_hist_image)[:, 0, :].contiguous()        # the desired samples are the bins
        if not dictionary_not_histogram:
            self.loss = torch.nn.KLDivLoss()
        if patch_size > 1:
            self.patch_extraction_mat = ReturnPatchExtractionMat(input_im_HR_mask.data.cpu().numpy(), patch_size=patch_size, device=self.device,
                                                                 patches_overlap=0.5)
            self.image_mask = None
        else:
            self.image_mask = input_im_HR_mask.view([-1]).bool().to(self.device) if input_im_HR_mask is not None else None
        if not dictionary_not_histogram:
            if desired_hist_image is not None:
                with torch.no_grad():
                    self.desired_hists_list = [self.ComputeSoftHistogram(desired_hist_image, image_mask=self.desired_hist_image_mask, return_log_hist=False,
                                                                         reshape_image=False, compute_hist_normalizer=True).detach()]
            else:
                self.desired_hist_image = desired_hist_image

    def Feed_Desired_Hist_Im(self, desired_hist_image):
        self.desired_hists_list = []
        for desired_im in desired_hist_image:
            desired_im = desired_im.mean(0, keepdim=True).view([1, -1, 1])
            with torch.no_grad():
                self.desired_hists_list.append(self.ComputeSoftHistogram(desired_im, image_mask=self.desired_hist_image_mask, return_log_hist=False,
                                                                         reshape_image=False, compute_hist_normalizer=True).detach())

    def Desired_Im_2_Bins(self, desired_im):
        """the desired samples [dims, n, 1] with near-duplicates (every coordinate within half a bin of a LATER sample) removed
        (Z_optimization.py:105-133) -> [dims, 1, n_kept] fp64"""
        im = desired_im.view([self.num_dims, -1])
        n = im.size(1)
        keep = torch.ones(n, dtype=torch.bool, device=im.device)
        step = 2048
        for a in range(0, n, step):        # row blocks instead of the reference's all-at-once n x n matrix (and its retry on out-of-memory)
            close = ((im[:, a:a + step].unsqueeze(2) - im.unsqueeze(1)).abs() < self.bin_width / 2).all(0)      # [block, n]
            idx = torch.arange(a, min(a + step, n), device=im.device).unsqueeze(1)
            later = torch.arange(n, device=im.device).unsqueeze(0) > idx
            keep[a:a + step] = ~(close & later).any(1)
        return im[:, keep].view([self.num_dims, 1, -1]).double()

    def ComputeSoftHistogram(self, image, image_mask, return_log_hist, reshape_image, compute_hist_normalizer, temperature=None):
        if temperature is None:
            temperature = 1 * self.temperature
        if reshape_image:
            if self.patch_size > 1:
                image = torch.sparse.mm(self.patch_extraction_mat, image.view([-1, 1])).view([self.num_dims, -1])
                if self.no_patch_DC:
                    image = image - torch.mean(image, dim=0, keepdim=True)
                    if self.no_patch_STD:
                        image = image / torch.max(torch.std(image, dim=0, keepdim=True), other=torch.tensor(1 / 255).to(self.device)) * self.mean_patches_STD
            else:
                image = image.contiguous().view([self.num_dims, -1])
                if image_mask is not None:
                    image = image[:, image_mask]
        else:
            image = image.view([self.num_dims, -1])
            if image_mask is not None and not self.KDE:
                pass        # (the reference applies the desired image's mask only in the KDE branch of the constructor, :90-91)
        n_samples = image.size(1)
        res = _SoftHistFn.apply(image.to(self.device), self.bins, self.max, self.SQRT_EPSILON, float(temperature), self.dictionary_not_histogram)
        if self.dictionary_not_histogram:
            return res.view([1, -1])
        hist = res
        if compute_hist_normalizer or not self.KDE:
            self.normalizer = hist.sum() / n_samples
        hist = (hist / self.normalizer / n_samples).float()
        if self.KDE:        # another "bin" accounts for everything the desired samples do not cover
            hist = torch.cat([hist, (1 - torch.min(torch.tensor(1, dtype=hist.dtype, device=hist.device), hist.sum())).view([1])])
        if return_log_hist:
            return torch.log(hist + torch.finfo(hist.dtype).eps).view([1, -1])
        return hist.view([1, -1])

    def forward(self, cur_images):
        hists = []
        for cur_image in cur_images:
            cur_image = cur_image.mean(0, keepdim=True)
            hists.append(self.ComputeSoftHistogram(cur_image, self.image_mask, return_log_hist=True, reshape_image=True, compute_hist_normalizer=False,
                                                   temperature=self.temperature))
        if self.dictionary_not_histogram:
            return torch.cat(hists, 0).mean(1).float()
        return self.loss(torch.cat(hists, 0), torch.cat(self.desired_hists_list, 0)).float()


def Patch_Indexes_2_Sparse_Mat(patches_indexes, mask_size, device):
    """[n_patches, p*p] pixel indexes -> sparse 0/1 matrix [n_patches*p*p, mask_size]; row order: position-in-patch major, patch minor
    (Z_optimization.py:267-271), so `mm(mat, image.view(-1,1)).view(p*p, n_patches)` lists every patch as a column"""
    rows = np.arange(patches_indexes.size).reshape([-1])
    cols = patches_indexes.transpose().reshape([-1])
    return torch.sparse_coo_tensor(torch.from_numpy(np.stack([rows, cols]).astype(np.int64)), torch.ones(rows.size, dtype=torch.float32),
                                   (patches_indexes.size, mask_size)).to(device)


def ReturnPatchExtractionMat(mask, patch_size, device, patches_overlap=1, return_non_covered=False):
    """All patch_size x patch_size patches lying inside `mask` as a sparse pixel-gathering matrix (Z_optimization.py:232-265; the GUI's
    Estimate_DerivedControlIndicator, GUI.py:2408-2412).  patches_overlap < 1 keeps, scanning in raster order, only patches whose
    share of already covered pixels does not exceed it (0: no shared pixel at all) and can return the uncovered pixels as a
    second matrix.  Host-side index construction; the products with it are sparse mm."""
    from scipy.ndimage import binary_opening
    from sklearn.feature_extraction.image import extract_patches_2d
    mask = binary_opening(mask, np.ones([patch_size, patch_size]).astype(bool))
    numbered = np.multiply(mask, 1 + np.arange(mask.size).reshape(mask.shape))
    patches_indexes = extract_patches_2d(numbered, (patch_size, patch_size)).reshape([-1, patch_size ** 2])
    patches_indexes = patches_indexes[np.all(patches_indexes > 0, 1), :] - 1
    non_covered_mat = None
    if patches_overlap < 1:
        unique_indexes = list(set(list(patches_indexes.reshape([-1]))))
        lo = min(unique_indexes)
        # (the reference sizes this one short and addresses it with "- lo - 1": the smallest index lands on the LAST slot; kept as is)
        taken = np.zeros([max(unique_indexes) - lo]).astype(bool)
        keep = np.ones([patches_indexes.shape[0]]).astype(bool)
        for k in range(patches_indexes.shape[0]):
            slots = patches_indexes[k, :] - lo - 1
            if (patches_overlap == 0 and np.any(taken[slots])) or np.mean(taken[slots]) > patches_overlap:
                keep[k] = False
                continue
            taken[slots] = True
        patches_indexes = patches_indexes[keep]
        print('%.3f of desired pixels are covered by assigned patches' % (taken[np.array(unique_indexes) - lo - 1].mean()))
        if return_non_covered:
            rest = np.array(unique_indexes)
            rest = rest[np.logical_not(taken[rest - lo - 1])]
            non_covered_mat = Patch_Indexes_2_Sparse_Mat(rest, mask.size, device)
    mat = Patch_Indexes_2_Sparse_Mat(patches_indexes, mask.size, device)
    return (mat, non_covered_mat) if return_non_covered else mat


class Z_optimizer():
    MIN_LR = 1e-5
    PATCH_SIZE_4_STD = 7

    def __init__(self, objective, Z_size, model, Z_range, max_iters, data=None, loggers=None, image_mask=None, Z_mask=None, initial_Z=None,
                 initial_LR=None, existing_optimizer=None, batch_size=1, HR_unpadder=None, auto_set_hist_temperature=False, random_Z_inits=False,
                 jpeg_extractor=None, non_local_Z_optimization=False):
        if jpeg_extractor is not None:
            raise NotImplementedError('esr_b200: the JPEG sibling project is out of scope')
        for word in _UNBUILT:
            if word in objective:
                raise NotImplementedError('esr_b200: Z_optimizer objective [%s] is not built yet (SURVEY 8f-1)' % objective)
        self.jpeg_mode = False
        self.data_keys = {'reconstructed': 'SR'}
        if initial_Z is not None or 'cur_Z' in model.__dict__.keys():
            if initial_Z is None:
                initial_Z = 1 * model.GetLatent()
            eps = torch.finfo(initial_Z.dtype).eps
            initial_pre_tanh_Z = ArcTanH(torch.clamp(initial_Z / Z_range, min=-1 + eps, max=1. - eps))
        else:
            initial_pre_tanh_Z = None
        self.non_local_Z_optimization = non_local_Z_optimization and image_mask is not None and image_mask.mean() < 1
        self.model_training = HR_unpadder is not None
        assert not (self.non_local_Z_optimization and self.model_training), 'Shouldn''t happen...'
        if not self.model_training:
            self.initial_output = model.Output_Batch(within_0_1=True)
        if self.non_local_Z_optimization:
            NON_EDIT_MARGINS = 24
            new_Z_mask = np.zeros_like(Z_mask)
            new_Z_mask[NON_EDIT_MARGINS:-NON_EDIT_MARGINS, NON_EDIT_MARGINS:-NON_EDIT_MARGINS] = 1
            Z_mask = np.minimum(1, new_Z_mask + _dilate_rect(image_mask, 16))
        self.Z_model = Optimizable_Z(Z_shape=[batch_size, model.num_latent_channels] + list(Z_size), Z_range=Z_range,
                                     initial_pre_tanh_Z=initial_pre_tanh_Z, Z_mask=Z_mask,
                                     random_perturbations=(random_Z_inits and 'random' not in objective) or ('random' in objective and 'limited' in objective))
        assert (initial_LR is not None) or (existing_optimizer is not None), \
            'Should either supply optimizer from previous iterations or initial LR for new optimizer'
        self.objective = objective
        self.data = data
        self.device = _dev()
        self.model = model
        if image_mask is None:
            self.image_mask = torch.ones(list(model.fake_H.size()[2:]), dtype=model.fake_H.dtype, device=self.device) \
                if 'fake_H' in model.__dict__.keys() else None
            self.Z_mask = None
        else:
            assert Z_mask is not None, 'Should either supply both masks or niether'
            self.image_mask = torch.from_numpy(image_mask).type(model.fake_H.dtype).to(self.device)
            self.Z_mask = torch.from_numpy(Z_mask).type(model.fake_H.dtype).to(self.device)
            self.initial_Z = 1. * model.GetLatent()
            if self.non_local_Z_optimization:
                self.constraining_mask = 1 - (self.image_mask > 0).type(self.image_mask.dtype)
                self.constraining_loss = lambda produced_im: F.l1_loss(input=produced_im * self.constraining_mask,
                                                                       target=self.initial_output * self.constraining_mask)
                self.constraining_loss_weight = 0.1
        if 'local' in objective:      # relative STD change over patches (Z_optimization.py:391-397)
            desired_overlap = 1 if 'STD' in objective else 0.5
            self.patch_extraction_map, self.non_covered_indexes_extraction_mat = ReturnPatchExtractionMat(
                mask=image_mask, patch_size=self.PATCH_SIZE_4_STD, device=model.fake_H.device, patches_overlap=desired_overlap, return_non_covered=True)
        if not self.model_training:
            self.initial_STD = self.Masked_STD(first_image_only=True)
            print('Initial STD: %.3e' % (self.initial_STD.mean().item()))
        if existing_optimizer is None:
            if any(p in objective for p in ['l1', 'scribble']) and 'random' not in objective:
                if data is not None and 'desired' in data.keys():
                    self.desired_im = data['desired']
                if self.image_mask is None:
                    self.loss = torch.nn.L1Loss()
                elif 'scribble' in objective:
                    self._init_scribble(data)
                    self.constraining_loss_weight = 1
                else:
                    loss_mask = (self.image_mask > 0).type(self.image_mask.dtype)
                    self.loss = lambda produced_im, GT_im: torch.stack(
                        [F.l1_loss(input=produced_im[i].unsqueeze(0) * loss_mask, target=GT_im * loss_mask) for i in range(produced_im.size(0))], 0)
                    self.constraining_loss_weight = 1
            elif 'Mag' in objective:      # local magnitude: patches keep their mean, their STD moves by STD_increment (:450-455)
                self.desired_patches = torch.sparse.mm(self.patch_extraction_map, self.initial_output.mean(dim=1).view([-1, 1])).view(
                    [self.PATCH_SIZE_4_STD ** 2, -1])
                desired_STD = torch.max(torch.std(self.desired_patches, dim=0, keepdim=True), torch.tensor(1 / 255).to(self.device))
                mean = torch.mean(self.desired_patches, dim=0, keepdim=True)
                self.desired_patches = (self.desired_patches - mean) / desired_STD * \
                    (desired_STD + data['STD_increment'] * (1 if 'increase' in objective else -1)) + mean
                self.constraining_loss_weight = 255 / 10 * data['STD_increment'] ** 2
            elif 'VGG' in objective and 'random' not in objective:      # perceptual distance to the desired image (:505-509)
                self.desired_im = data['desired']
                self.GT_HR_VGG = model.netF(self.desired_im.to(self.device)).detach().to(self.device)
                from esr_b200.losses import L1Loss
                self.loss = L1Loss().to(self.device)
            elif any(p in objective for p in ['hist', 'dict']):       # imprinting tools (:510-542)
                if auto_set_hist_temperature:
                    raise NotImplementedError('esr_b200: auto_set_hist_temperature needs a double backward through the generator')
                self.STD_PRESERVING_WEIGHT = 1e4
                optimal_temperature = 5e-4 if 'hist' in objective else 1e-3
                self.loss = SoftHistogramLoss(bins=256, min=0, max=1, desired_hist_image=self.data['desired'] if self.data is not None else None,
                                              desired_hist_image_mask=data['Desired_Im_Mask'] if self.data is not None else None,
                                              input_im_HR_mask=self.image_mask, gray_scale=True, patch_size=6 if 'patch' in objective else 1,
                                              temperature=optimal_temperature, dictionary_not_histogram='dict' in objective,
                                              no_patch_DC='noDC' in objective, no_patch_STD='no_localSTD' in objective)
                self.constraining_loss_weight = 10
            elif 'Adversarial' in objective:      # fool the critic (:543-545)
                from models.modules.loss import GANLoss
                self.netD = model.netD
                self.loss = GANLoss('wgan-gp', 1.0, 0.0).to(self.device)
            elif 'STD' in objective and not any(p in objective for p in ['periodicity', 'TV']):
                assert self.objective.replace('local_', '') in ['max_STD', 'min_STD', 'STD_increase', 'STD_decrease']
                if any(p in objective for p in ['increase', 'decrease']):
                    STD_CHANGE_FACTOR = 1.05
                    self.desired_STD = self.initial_STD
                    if data['STD_increment'] is None:
                        self.desired_STD = self.desired_STD * (STD_CHANGE_FACTOR if 'increase' in objective else 1 / STD_CHANGE_FACTOR)
                    else:
                        self.desired_STD = self.desired_STD + (data['STD_increment'] if 'increase' in objective else -data['STD_increment'])
                        self.constraining_loss_weight = 255 / 10 * data['STD_increment'] ** 2
            elif 'periodicity' in objective:      # the image should repeat itself at the given displacements (:470-504)
                self.STD_PRESERVING_WEIGHT = 20
                self.PLUS_MEANS_STD_INCREASE = True
                if 'nonInt' in objective:
                    image_size = list(self.initial_output.size()[2:])
                    self.periodicity_points, self.half_period_points = [], []
                    if 'Plus' in objective and self.PLUS_MEANS_STD_INCREASE:
                        self.desired_STD = self.initial_STD + data['STD_increment']
                    for point in data['periodicity_points']:
                        point = np.array(point)
                        self.periodicity_points.append([])
                        self.half_period_points.append([])
                        for half_period_round in range(1 + ('Plus' in objective and not self.PLUS_MEANS_STD_INCREASE)):
                            for minus_point in range(2):
                                cur_point = 1 * point
                                if half_period_round:
                                    cur_point = cur_point * 0.5
                                if minus_point:
                                    cur_point = cur_point * -1
                                # sampling grid of the overlap between the image and its copy displaced by cur_point (x first, as grid_sample expects)
                                y_range = [_half_open(cur_point[0]), _half_open(cur_point[0], True)]
                                x_range = [_half_open(cur_point[1]), _half_open(cur_point[1], True)]
                                ranges = []
                                for axis, cur_range in enumerate([x_range, y_range]):
                                    cur_range = [cur_range[0] if cur_range[0] is not None else 0,
                                                 image_size[axis] + cur_range[1] if cur_range[1] is not None else image_size[axis]]
                                    num = image_size[axis] - np.ceil(np.abs(np.array([0, image_size[axis]]) - cur_range)).astype(np.int16).max()
                                    ranges.append(np.linspace(start=cur_range[0], stop=cur_range[1], num=num) / image_size[axis] * 2 - 1)
                                grid = np.meshgrid(*ranges)
                                grid = torch.from_numpy(np.stack(grid, -1)).view([1] + list(grid[0].shape) + [2]).type(
                                    self.initial_output.dtype).to(self.initial_output.device)
                                (self.half_period_points if half_period_round else self.periodicity_points)[-1].append(grid)
                else:
                    self.periodicity_points = [np.array(point) for point in data['periodicity_points']]
            elif 'TV' in objective:
                self.STD_PRESERVING_WEIGHT = 100
            elif 'limited' in objective:
                self.initial_image = 1 * model.output_image.detach()
                self.rmse_weight = data['rmse_weight']
            elif 'random' in objective:
                self.STD_PRESERVING_WEIGHT = 1e3
            # capturable: the whole iteration (generator forward, backward to Z, Adam) can then be replayed as one CUDA graph
            self.optimizer = torch.optim.Adam(self.Z_model.parameters(), lr=initial_LR, capturable=torch.cuda.is_available())
            self._own_optimizer = True
        else:
            self.optimizer = existing_optimizer
            self._own_optimizer = False
        self.LR = initial_LR
        self.scheduler = None
        self.loggers = loggers
        self.cur_iter = 0
        self.max_iters = max_iters
        self.random_Z_inits = 'all' if (random_Z_inits or self.model_training) \
            else 'allButFirst' if (initial_pre_tanh_Z is not None and initial_pre_tanh_Z.size(0) < batch_size) else False
        self.HR_unpadder = HR_unpadder

    def _init_scribble(self, data):
        """The scribble tool (Z_optimization.py:408-449).  data['scribble_mask'] labels every pixel: 0 untouched, 1 colour scribble (l1 towards
        data['desired']), 2 / 3 brighten / darken (the current output's V channel times 1 +- data['brightness_factor'], smoothed over a 3x3
        neighbourhood, becomes the target there), > 3 smoothing scribbles (one label per region: total variation between neighbours that both
        carry the label)."""
        from scipy.signal import convolve2d
        from esr_b200.colors import hsv2rgb, rgb2hsv
        loss_mask = (self.image_mask > 0).type(self.image_mask.dtype)
        SMOOTHING_MARGIN = 1
        labels = np.asarray(data['scribble_mask'])
        scribble_mask_tensor = torch.from_numpy(labels).type(loss_mask.dtype).to(loss_mask.device)
        scribble_multiplier = np.ones_like(labels).astype(np.float32)
        scribble_multiplier += data['brightness_factor'] * (labels == 2) - data['brightness_factor'] * (labels == 3)
        if SMOOTHING_MARGIN > 0:
            k = 2 * SMOOTHING_MARGIN + 1
            scribble_multiplier = convolve2d(np.pad(scribble_multiplier, ((SMOOTHING_MARGIN,) * 2,) * 2, mode='edge'), np.ones([k, k]) / k ** 2, mode='valid')
        L1_loss_mask = loss_mask * ((scribble_mask_tensor > 0) * (scribble_mask_tensor < 4)).float()
        TV_loss_masks = [loss_mask * (scribble_mask_tensor == i).float().unsqueeze(0).unsqueeze(0) for i in torch.unique(scribble_mask_tensor * loss_mask) if i > 3]
        cur_HSV = rgb2hsv(np.clip(255 * self.initial_output[0].data.cpu().numpy().transpose((1, 2, 0)).copy(), 0, 255))
        cur_HSV[:, :, 2] = cur_HSV[:, :, 2] * scribble_multiplier
        desired_RGB = np.expand_dims(hsv2rgb(cur_HSV).transpose((2, 0, 1)), 0) / 255
        desired_RGB_mask = ((scribble_mask_tensor == 2) | (scribble_mask_tensor == 3)).float()
        self.desired_im = self.desired_im.to(loss_mask.device) * (1 - desired_RGB_mask) + \
            desired_RGB_mask * torch.from_numpy(desired_RGB).type(loss_mask.dtype).to(loss_mask.device)

        def Scribble_TV_Loss(produced_im):
            loss = 0
            for TV_loss_mask in TV_loss_masks:
                # differences to the 8 neighbours, each unordered pair once: (dy, dx) in {(-1,-1), (-1,0), (0,-1), (1,-1)}
                for y_shift in [-1, 0, 1]:
                    for x_shift in [-1, 0]:
                        if y_shift in [0, 1] and x_shift == 0:
                            continue
                        point = np.array([y_shift, x_shift])
                        cur_mask = Return_Translated_SubImage(TV_loss_mask, point) * Return_Translated_SubImage(TV_loss_mask, -point)
                        loss = loss + (cur_mask * (Return_Translated_SubImage(produced_im, point) -
                                                   Return_Translated_SubImage(produced_im, -point)).abs()).mean(dim=(1, 2, 3))
            return loss

        def Scribble_Loss(produced_im, GT_im):
            loss_per_im = []
            for im_num in range(produced_im.size(0)):
                loss_per_im.append(F.l1_loss(input=produced_im[im_num].unsqueeze(0) * L1_loss_mask, target=GT_im * L1_loss_mask))
                if len(TV_loss_masks) > 0:
                    loss_per_im[-1] = loss_per_im[-1] + Scribble_TV_Loss(produced_im[im_num].unsqueeze(0))
            return torch.stack(loss_per_im, 0)
        self.loss = Scribble_Loss

    def Masked_STD(self, first_image_only=False):
        model_output = self.model.Output_Batch(within_0_1=True)
        if 'local' in self.objective:      # STD of every 7x7 patch inside the mask (+ of the pixels no patch covers), per image (:618-625)
            values = []
            for im_num in range(1 if first_image_only else model_output.size(0)):
                gray = model_output[im_num].mean(dim=0).view([-1, 1])
                values.append(torch.sparse.mm(self.patch_extraction_map, gray).view([self.PATCH_SIZE_4_STD ** 2, -1]).std(dim=0))
                if self.non_covered_indexes_extraction_mat is not None:
                    values[-1] = torch.cat([values[-1], torch.sparse.mm(self.non_covered_indexes_extraction_mat, gray).std(dim=0)], 0)
            return torch.stack(values, 1)
        return torch.std(model_output * self.image_mask, dim=(1, 2, 3)).view(1, -1)

    def feed_data(self, data):
        self.data = data
        self.cur_iter = 0
        if 'l1' in self.objective:
            self.desired_im = data['desired'].to(self.device)

    def Manage_Model_Grad_Requirements(self, verify_disabled):
        if verify_disabled:
            self.original_requires_grad_status = []
            for p in self.model.netG.parameters():
                self.original_requires_grad_status.append(p.requires_grad)
                p.requires_grad = False
        else:
            for i, p in enumerate(self.model.netG.parameters()):
                p.requires_grad = self.original_requires_grad_status[i]

    GRAPH_WARMUP_ITERS = 3

    def _iteration(self, keep_pre_tanh):
        """device work of one iteration of Z_optimization.py:663-754: Z -> generator (+CEM) with graph -> objective -> backward to
        Z -> Adam step.  Returns (per-image loss values, scalar loss, copy of the pre-tanh Z the loss belongs to); no host reads."""
        self.optimizer.zero_grad()
        self.data['Z'] = self.Z_model()
        pre_tanh = 1 * self.Z_model.PreTanhZ() if keep_pre_tanh else None
        self.model.feed_data(self.data, need_GT=False)
        self.model.test(prevent_grads_calc=False)
        self.output_image = self.model.Output_Batch(within_0_1=True)
        if self.model_training:
            self.output_image = self.HR_unpadder(self.output_image)
        if 'random' in self.objective:
            dom = self.output_image
            Z_loss = torch.min((dom.unsqueeze(0) - dom.unsqueeze(1)).abs() +
                               torch.eye(dom.size(0), device=dom.device).unsqueeze(2).unsqueeze(3).unsqueeze(4), dim=0)[0]
            if 'limited' in self.objective:
                Z_loss = Z_loss - self.rmse_weight * (dom - self.initial_image).abs()
            if self.Z_mask is not None:
                Z_loss = Z_loss * self.image_mask
            Z_loss = -1 * Z_loss.mean(dim=(1, 2, 3))
            if 'local' in self.objective:
                Z_loss = Z_loss + self.STD_PRESERVING_WEIGHT * ((self.Masked_STD(first_image_only=False) - self.initial_STD) ** 2).mean()
        elif any(p in self.objective for p in ['l1', 'scribble']):
            Z_loss = self.loss(self.output_image.to(self.device), self.desired_im.to(self.device))
        elif any(p in self.objective for p in ['hist', 'dict']):
            Z_loss = self.loss(self.output_image.to(self.device))
            if 'localSTD' in self.objective:
                Z_loss = Z_loss + (self.STD_PRESERVING_WEIGHT * (self.Masked_STD(first_image_only=False) - self.initial_STD) ** 2).mean(0).to(self.device)
        elif 'Adversarial' in self.objective:
            Z_loss = self.loss(self.netD(self.model.CEM_net.HR_unpadder(self.output_image).to(self.device)), True)
        elif 'Mag' in self.objective:
            values = []
            for im_num in range(self.output_image.size(0)):
                patches = torch.sparse.mm(self.patch_extraction_map, self.output_image[im_num].mean(dim=0).view([-1, 1])).view([self.PATCH_SIZE_4_STD ** 2, -1])
                values.append(((patches - self.desired_patches) ** 2).mean())
            Z_loss = torch.stack(values, 0)
        elif 'VGG' in self.objective:
            Z_loss = self.loss(self.model.netF(self.output_image).to(self.device), self.GT_HR_VGG)
        elif 'periodicity' in self.objective:
            Z_loss = self.PeriodicityLoss().to(self.device)
            if 'Plus' in self.objective and self.PLUS_MEANS_STD_INCREASE:
                Z_loss = Z_loss + self.STD_PRESERVING_WEIGHT * ((self.Masked_STD(first_image_only=False) - self.desired_STD) ** 2).mean()
        elif 'STD' in self.objective and 'TV' not in self.objective:
            Z_loss = self.Masked_STD(first_image_only=False)
            if any(p in self.objective for p in ['increase', 'decrease']):
                Z_loss = (Z_loss - self.desired_STD) ** 2
            Z_loss = Z_loss.mean(0)
        elif 'TV' in self.objective:
            Z_loss = (self.STD_PRESERVING_WEIGHT * (self.Masked_STD(first_image_only=False) - self.initial_STD) ** 2).mean(0) + \
                TV_Loss(self.output_image * self.image_mask)
        if 'max' in self.objective:
            Z_loss = -1 * Z_loss
        Z_loss_items = Z_loss.detach()
        Z_loss = Z_loss.mean()
        if self.non_local_Z_optimization:
            Z_loss = Z_loss + self.constraining_loss_weight * self.constraining_loss(self.output_image.to(self.device))
        Z_loss.backward()
        self.optimizer.step()
        return Z_loss_items, Z_loss.detach(), pre_tanh

    def _capture_iteration(self, keep_pre_tanh):
        """One iteration is ~800 kernel launches of a few microseconds each at GUI region sizes: launch-bound.  After a few
        eager iterations the iteration is captured once into a CUDA graph and replayed (same launches, same order, same
        numerics).  Any capture problem falls back to eager iterations."""
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._side_stream):
                static = self._iteration(keep_pre_tanh)
            # the capture itself executes nothing: the first replay performs this iteration
            return g, static
        except Exception as e:  # pragma: no cover - depends on driver / allocator state
            print('Z_optimizer: CUDA graph capture failed (%s); continuing with eager iterations.' % repr(e)[:200])
            self._graph_ok = False
            torch.cuda.synchronize()
            return None, None

    def optimize(self):
        USE_MIN_LOSS_Z = not self.model_training
        graph, static = None, None
        self._graph_ok = (self._own_optimizer and torch.cuda.is_available() and os.environ.get('ESR_ZOPT_GRAPH', '1') != '0'
                          and not self.model_training and self.loggers is None
                          and not any(p in self.objective for p in ['local', 'Mag', 'hist', 'dict', 'periodicity', 'scribble'])      # (sparse products, grid samples: eager iterations)
                          and (self.max_iters < 0 or self.max_iters >= self.GRAPH_WARMUP_ITERS + 4))
        if self._graph_ok:
            self._side_stream = torch.cuda.Stream()
            self._side_stream.wait_stream(torch.cuda.current_stream())
        self.Manage_Model_Grad_Requirements(verify_disabled=True)
        self.loss_values, per_iter_pre_tanh_Z = [], []
        if self.random_Z_inits and self.cur_iter == 0:
            self.Z_model.Randomize_Z(what_2_shuffle=self.random_Z_inits)
        z_iter = self.cur_iter
        while True:
            if self.max_iters > 0:
                if z_iter == (self.cur_iter + self.max_iters):
                    break
            elif len(self.loss_values) >= -self.max_iters:  # stop when the loss stops decreasing, or after 5*|max_iters|
                if z_iter == (self.cur_iter - 5 * self.max_iters):
                    break
                if (self.loss_values[self.max_iters] - self.loss_values[-1]) / np.abs(self.loss_values[self.max_iters]) < 1e-2 * self.LR:
                    break
            if graph is None and self._graph_ok and len(self.loss_values) == self.GRAPH_WARMUP_ITERS:
                graph, static = self._capture_iteration(USE_MIN_LOSS_Z)
            if graph is not None:
                graph.replay()
                Z_loss_items, Z_loss, pre_tanh = static
            elif self._graph_ok:
                # warm-up iterations run on the stream the capture will use: autograd's AccumulateGrad node of Z must not have
                # been created on another stream (torch invalidates the capture otherwise)
                with torch.cuda.stream(self._side_stream):
                    Z_loss_items, Z_loss, pre_tanh = self._iteration(USE_MIN_LOSS_Z)
                torch.cuda.current_stream().wait_stream(self._side_stream)
            else:
                Z_loss_items, Z_loss, pre_tanh = self._iteration(USE_MIN_LOSS_Z)
            if USE_MIN_LOSS_Z:
                per_iter_pre_tanh_Z.append(pre_tanh.clone() if graph is not None else pre_tanh)
            cur_LR = self.optimizer.param_groups[0]['lr']
            if self.loggers is not None:
                for logger_num, logger in enumerate(self.loggers):
                    cur_value = Z_loss_items[logger_num].mean().item() if Z_loss_items.dim() > 0 else Z_loss_items.mean().item()
                    logger.print_format_results('val', {'epoch': 0, 'iters': z_iter, 'time': time.time(), 'model': '', 'lr': cur_LR,
                                                        'Z_loss': cur_value}, dont_print=True)
            if not self.model_training:
                self.latest_Z_loss_values = [val.mean().item() for val in Z_loss_items] if Z_loss_items.dim() > 0 else [Z_loss_items.item()]
            self.loss_values.append(Z_loss.item())
            z_iter += 1
        if USE_MIN_LOSS_Z:
            if np.min(self.loss_values) != self.loss_values[-1]:
                min_loss_iter = int(np.argmin(self.loss_values))
                print('Minimum loss observed in %d/%d iteration, discarding subsequent iterations.' % (min_loss_iter + 1, len(self.loss_values)))
                self.Z_model.Z.data = 1 * per_iter_pre_tanh_Z[min_loss_iter]
                self.loss_values = self.loss_values[:min_loss_iter + 1]
        if 'random' in self.objective and 'limited' in self.objective:
            self.loss_values[0] = self.loss_values[1]
        self.cur_iter = z_iter + 1
        Z_2_return = self.Z_model.Return_Detached_Z()
        self.Manage_Model_Grad_Requirements(verify_disabled=False)
        if not self.model_training:
            print('Final STDs: ', ['%.3e' % (val.item()) for val in self.Masked_STD(first_image_only=False).mean(0)])
        if self.model_training:
            self.data['Z'] = Z_2_return
            self.model.feed_data(self.data, need_GT=False)
            self.model.fake_H = self.model.netG(self.model.model_input)
        return Z_2_return

    def PeriodicityLoss(self):
        """mean |I(p + d/2) - I(p - d/2)| over the masked overlap, per requested displacement d (Z_optimization.py:799-814), plus the
        term that keeps the region's STD where it started (dropped when the 'Plus' variant asks for an STD increase instead)"""
        if 'Plus' in self.objective and self.PLUS_MEANS_STD_INCREASE:
            loss = 0
        else:
            loss = (self.STD_PRESERVING_WEIGHT * (self.Masked_STD(first_image_only=False) - self.initial_STD) ** 2).mean()
        image = self.output_image
        mask = self.image_mask.unsqueeze(0).unsqueeze(0)
        for point_num, point in enumerate(self.periodicity_points):
            if 'nonInt' in self.objective:
                cur_mask = Return_Interpolated_SubImage(mask, point[0]) * Return_Interpolated_SubImage(mask, point[1])
                loss = loss + (cur_mask * (Return_Interpolated_SubImage(image, point[0]) - Return_Interpolated_SubImage(image, point[1])).abs()).mean(dim=(1, 2, 3))
                if 'Plus' in self.objective and not self.PLUS_MEANS_STD_INCREASE:
                    half = self.half_period_points[point_num]
                    half_mask = Return_Interpolated_SubImage(mask, half[0]) * Return_Interpolated_SubImage(mask, half[1])
                    loss = loss - (half_mask * (Return_Interpolated_SubImage(image, half[0]) - Return_Interpolated_SubImage(image, half[1])).abs()).mean(dim=(1, 2, 3))
            else:
                cur_mask = Return_Translated_SubImage(mask, point) * Return_Translated_SubImage(mask, -point)
                loss = loss + (cur_mask * (Return_Translated_SubImage(image, point) - Return_Translated_SubImage(image, -point)).abs()).mean(dim=(1, 2, 3))
        return loss

    def ReturnStatus(self):
        return self.cur_iter, self.Z_model.PreTanhZ()
