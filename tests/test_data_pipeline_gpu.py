"""GPU input pipeline (esr_b200.data.GpuLRHRBatcher) against a CPU restatement of data/LRHR_dataset.py:48-131 built on the
numpy `imresize` mirror (itself bit-identical to the reference's, tests/test_cem_design.py): same seeded crop / flip decisions, LR from the
CEM kernel down-sampling of the whole image, BGR -> RGB, HWC -> CHW."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _restate(img_u8, scale, HR_size, decision):
    from CEM.imresize_CEM import imresize
    rh, rw, hflip, vflip, rot90 = decision
    img_HR = img_u8.astype(np.float32) / 255.                               # util.read_img (data/util.py:103)
    img_LR = imresize(img_HR, scale_factor=[1 / float(scale)], kernel=None)  # LRHR_dataset.py:88
    LR_size = HR_size // scale
    img_LR = img_LR[rh:rh + LR_size, rw:rw + LR_size, :]
    img_HR = img_HR[rh * scale:rh * scale + HR_size, rw * scale:rw * scale + HR_size, :]
    out = []
    for img in (img_LR, img_HR):                                            # util.augment (data/util.py:118-130)
        if hflip:
            img = img[:, ::-1, :]
        if vflip:
            img = img[::-1, :, :]
        if rot90:
            img = img.transpose(1, 0, 2)
        img = img[:, :, [2, 1, 0]]                                          # BGR -> RGB, HWC -> CHW (:123-127)
        out.append(np.ascontiguousarray(np.transpose(img, (2, 0, 1))).astype(np.float32))
    return out


def test_gpu_batcher_matches_the_dataset_restatement():
    from esr_b200 import ops
    from esr_b200.data import GpuLRHRBatcher
    ops.device_check()
    rng = np.random.RandomState(3)
    scale, patch = 4, 64
    imgs = rng.randint(0, 256, size=(5, 96, 128, 3), dtype=np.uint8)
    b = GpuLRHRBatcher(scale, patch, use_flip=True, use_rot=True, seed=11)
    decisions = b.draw(len(imgs), imgs.shape[1] // scale, imgs.shape[2] // scale)
    assert any(d[2] for d in decisions) or any(d[3] for d in decisions) or any(d[4] for d in decisions)
    # border crops included: force one crop into a corner
    decisions[0] = (0, 0) + decisions[0][2:]
    decisions[1] = (imgs.shape[1] // scale - patch // scale, imgs.shape[2] // scale - patch // scale) + decisions[1][2:]
    batch = b(torch.from_numpy(imgs).pin_memory(), decisions=decisions)
    assert batch['LR'].shape == (5, 3, 16, 16) and batch['HR'].shape == (5, 3, 64, 64) and batch['LR'].is_cuda
    worst = 0.0
    for i in range(len(imgs)):
        lr_ref, hr_ref = _restate(imgs[i], scale, patch, decisions[i])
        assert np.array_equal(batch['HR'][i].cpu().numpy(), hr_ref)
        err = np.abs(batch['LR'][i].cpu().numpy() - lr_ref).max()
        worst = max(worst, float(err))
    print('GPU batcher LR vs numpy imresize restatement: max abs diff %.2e' % worst)
    assert worst < 2e-6
    # the batch goes straight into the model API: feed_data accepts device tensors
    assert batch['LR'].dtype == torch.float32 and float(batch['LR'].min()) >= -0.2 and float(batch['LR'].max()) <= 1.2
