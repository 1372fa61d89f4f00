"""The training input pipeline after the image decode, on the GPU.

The reference's `LRHRDataset.__getitem__` (data/LRHR_dataset.py:48-131) does, per image and on the CPU: float conversion of the decoded
BGR image (`util.read_img`, data/util.py:95-109), `imresize(img_HR, 1/scale, kernel)` with the CEM's own down-sampling kernel
(CEM/imresize_CEM.py:8-87) to make the LR image, a random `LR_size` crop with the matching `HR_size` crop, random flips / transposition
(`util.augment`, data/util.py:118-130), BGR -> RGB and HWC -> CHW.  `GpuLRHRBatcher` takes a batch of decoded HR images (uint8, HWC, BGR,
equal sizes - e.g. the 480x480 DIV2K sub-images of the reference's training set; lmdb / PNG decoding stays on the CPU) and does the rest on
the device: one `esr_cem_down` launch (CEM.CEMnet.CEM_downsampler: the same replicate-padded antialiasing filter) for the whole batch instead of
a numpy convolution per image, crops / flips as index operations.  The random draws come from a `numpy.random.RandomState` in the
reference's order (crop row, crop column, then hflip, vflip, rot90 per image), so a seeded CPU restatement reproduces the batch."""
import numpy as np
import torch


class GpuLRHRBatcher:
    def __init__(self, scale, patch_size, use_flip=True, use_rot=True, device='cuda', seed=None):
        from CEM.CEMnet import CEM_downsampler
        self.scale, self.HR_size, self.LR_size = int(scale), int(patch_size), int(patch_size) // int(scale)
        self.use_flip, self.use_rot = bool(use_flip), bool(use_rot)
        self.device = torch.device(device)
        self.down = CEM_downsampler(self.scale).to(self.device)
        self.rng = np.random.RandomState(seed)
        self._copy_stream = None

    def draw(self, n, H, W):
        """the per-image random decisions of LRHR_dataset.py:107-116 + util.augment: (rnd_h, rnd_w, hflip, vflip, rot90), LR coordinates"""
        out = []
        for _ in range(n):
            rnd_h = int(self.rng.randint(0, max(0, H - self.LR_size) + 1))
            rnd_w = int(self.rng.randint(0, max(0, W - self.LR_size) + 1))
            hflip = self.use_flip and self.rng.random_sample() < 0.5
            vflip = self.use_rot and self.rng.random_sample() < 0.5
            rot90 = self.use_rot and self.rng.random_sample() < 0.5
            out.append((rnd_h, rnd_w, bool(hflip), bool(vflip), bool(rot90)))
        return out

    @torch.no_grad()
    def __call__(self, hr_u8, decisions=None):
        """hr_u8: [N, H, W, 3] uint8 (BGR, as cv2 decodes), host (pinned for an asynchronous copy) or device; H, W multiples of `scale`.
        Returns {'LR': [N,3,LR_size,LR_size], 'HR': [N,3,HR_size,HR_size]} float32 RGB in [0,1] on the device."""
        hr_u8 = torch.as_tensor(hr_u8)
        assert hr_u8.dtype == torch.uint8 and hr_u8.dim() == 4 and hr_u8.size(3) == 3
        n, H, W, _ = hr_u8.shape
        s = self.scale
        assert H % s == 0 and W % s == 0 and H >= self.HR_size and W >= self.HR_size, 'HR images must be modcropped and at least patch-sized'
        x = hr_u8.to(self.device, non_blocking=True)
        # read_img: float32 / 255 (a true division: dividing by a Python scalar would multiply by the rounded reciprocal), still BGR
        hr = x.permute(0, 3, 1, 2).float() / torch.full((), 255.0, dtype=torch.float32, device=self.device)
        lr = self.down(hr)                                          # imresize(img_HR, 1/scale, kernel) for the whole batch
        if decisions is None:
            decisions = self.draw(n, H // s, W // s)
        lr_out = torch.empty((n, 3, self.LR_size, self.LR_size), dtype=torch.float32, device=self.device)
        hr_out = torch.empty((n, 3, self.HR_size, self.HR_size), dtype=torch.float32, device=self.device)
        for i, (rh, rw, hflip, vflip, rot90) in enumerate(decisions):
            a = lr[i, :, rh:rh + self.LR_size, rw:rw + self.LR_size]
            b = hr[i, :, rh * s:rh * s + self.HR_size, rw * s:rw * s + self.HR_size]
            if hflip:
                a, b = a.flip(2), b.flip(2)
            if vflip:
                a, b = a.flip(1), b.flip(1)
            if rot90:
                a, b = a.transpose(1, 2), b.transpose(1, 2)
            lr_out[i] = a.flip(0)                                   # BGR -> RGB
            hr_out[i] = b.flip(0)
        return {'LR': lr_out, 'HR': hr_out}
