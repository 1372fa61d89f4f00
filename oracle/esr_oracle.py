"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (functional PyTorch fp32 on explicit state-dict tensors) of the reference's hot path:
RRDBNet forward, the CEM projection, the latent packing, the Z-optimisation l1 loop, Discriminator_VGG_128, the (relativistic)
GAN losses and the structure-tensor latent-control loss.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this file; the product path
(explorable-super-resolution_b200/) never does and has no CPU fallback.

Parity status: PINNED.  The reference has no tests or golden vectors of its own, so the pins are
outputs of the unmodified reference executed in the build container by oracle/make_golden*.py
(tests/golden/*.npz); tests/test_oracle.py, tests/test_discriminator_cpu.py and tests/test_filterloss_cpu.py check this file
against every one of them.  The orchestration layers (SRRaGANModel.optimize_parameters, Z_optimizer.optimize, checkpoint
loading) are pinned directly: oracle/make_golden_trainstep.py / _zopt.py / _ckpt.py run the reference's own classes on the CPU
and tests/test_trainstep_orchestration.py / test_zopt_orchestration.py / test_checkpoint_compat.py hold the product's to them.

Every function cites the reference lines it restates (paths relative to /root/reference/codes)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2  # models/modules/block.py:10 (act: neg_slope=0.2)


def _conv(x, sd, key, act):
    """conv_block, CNA order: Conv2d(k=3, s=1, zero pad 1, bias) [+ LeakyReLU(0.2)]  (block.py:129-146)."""
    y = F.conv2d(x, sd[key + '.weight'], sd[key + '.bias'], stride=1, padding=1)
    return F.leaky_relu(y, LRELU_SLOPE) if act else y


def rdb_forward(x, sd, prefix, nf):
    """ResidualDenseBlock_5C.forward (block.py:230-235): each conv sees the concatenation of the block
    input (latent channels, if any, in front) and all previous outputs; the 5th conv has no activation in
    CNA mode (block.py:208-209); result x5*0.2 + x[:, -nf:]."""
    outs = [x]
    for i in range(5):
        outs.append(_conv(torch.cat(outs, 1), sd, '%s.convs.%d.0' % (prefix, i), act=i < 4))
    return outs[-1] * 0.2 + outs[0][:, -nf:]


def rrdb_forward(x, sd, prefix, nf, z):
    """RRDB.forward (block.py:262-270): the latent channels x[:, :z] are re-concatenated in front of the
    input of RDB2 and RDB3."""
    out = rdb_forward(x, sd, prefix + '.RDB1', nf)
    if z > 0:
        out = torch.cat([x[:, :z], out], 1)
    out = rdb_forward(out, sd, prefix + '.RDB2', nf)
    if z > 0:
        out = torch.cat([x[:, :z], out], 1)
    out = rdb_forward(out, sd, prefix + '.RDB3', nf)
    return out * 0.2 + x[:, -nf:]


def rrdbnet_forward(x, sd, nf, nb, upscale=4, z=0, upsample_mode='upconv', prefix=''):
    """RRDBNet.forward (architecture.py:278-302) for latent_input in {None, 'all_layers_HR_downscaled'}.

    x: [N, z*upscale^2 + 3, h, w].  With a latent, the first z*s^2 channels are a raw memory view of
    Z[N, z, s*h, s*w] (architecture.py:281-283), bilinearly resized by 1/s (align_corners=False, :284) for the
    LR layers and used at full resolution by the two HR convs (:297-298); z goes IN FRONT of every conv
    input except the up-sampling convs (:290-300).  Module indices: 0 fea_conv, 1 ShortcutBlock(nb RRDBs +
    LR_conv), 2.. upsamplers, then HR_conv0, LeakyReLU, HR_conv1 (:271-273)."""
    p = prefix + 'model.'
    n_up = 1 if upscale == 3 else int(math.log(upscale, 2))
    if z > 0:
        lat, x = torch.split(x, [x.size(1) - 3, 3], dim=1)
        z_hr = lat.reshape(lat.size(0), -1, upscale * lat.size(2), upscale * lat.size(3))
        z_lr = F.interpolate(z_hr, scale_factor=1 / upscale, mode='bilinear', align_corners=False, recompute_scale_factor=False)
        x = torch.cat([z_lr, x], 1)
    fea = _conv(x, sd, p + '0', act=False)
    # ShortcutBlock (block.py:85-97): input is cat([z, fea]); z re-concatenated before every sub-module but the first
    out = torch.cat([z_lr, fea], 1) if z > 0 else fea
    for b in range(nb):
        if b > 0 and z > 0:
            out = torch.cat([z_lr, out], 1)
        out = rrdb_forward(out, sd, p + '1.sub.%d' % b, nf, z)
    if z > 0:
        out = torch.cat([z_lr, out], 1)
    out = _conv(out, sd, p + '1.sub.%d' % nb, act=False)  # LR_conv
    out = fea + out
    idx = 2
    for _ in range(n_up):
        if upsample_mode == 'upconv':  # block.py:299-309: nearest x2 -> conv -> LeakyReLU
            out = F.interpolate(out, scale_factor=3 if upscale == 3 else 2, mode='nearest')
            out = _conv(out, sd, p + '%d.1' % idx, act=True)
        else:  # block.py:278-291: conv -> PixelShuffle -> LeakyReLU
            out = F.leaky_relu(F.pixel_shuffle(_conv(out, sd, p + '%d.0' % idx, act=False), 2), LRELU_SLOPE)
        idx += 1
    if z > 0:
        out = torch.cat([z_hr, out], 1)
    out = _conv(out, sd, p + '%d' % idx, act=True)  # HR_conv0
    if z > 0:
        out = torch.cat([z_hr, out], 1)
    return _conv(out, sd, p + '%d' % (idx + 2), act=False)  # HR_conv1


# ------------------------------------------------------------------------------------------------ CEM
def _depthwise(x, k2d, pad):
    """Filter_Layer (CEMnet.py:243-252): replicate-pad then depth-wise correlation with a fixed 2-D filter."""
    c = x.size(1)
    w = torch.as_tensor(np.ascontiguousarray(k2d), dtype=torch.float32).view(1, 1, *k2d.shape).repeat(c, 1, 1, 1)
    return F.conv2d(F.pad(x, (pad[1], pad[1], pad[0], pad[0]), mode='replicate'), w, groups=c)


def cem_down(g, ds_kernel, s, pre):
    """DownscaleOP (CEMnet.py:265,270-275): replicate-pad floor(k/2), correlate with rot90(ds_kernel, 2), keep
    sample `pre` of every s x s cell."""
    k = np.rot90(ds_kernel, 2)
    y = _depthwise(g, k, [k.shape[0] // 2, k.shape[1] // 2])
    return y[:, :, pre::s, pre::s]


def cem_inv(e, inv_hTh):
    """Conv_LR_with_Inv_hTh_OP (CEMnet.py:262-264)."""
    return _depthwise(e, inv_hTh, [inv_hTh.shape[0] // 2, inv_hTh.shape[1] // 2])


def cem_up(f, ds_kernel, s, pre):
    """Upscale_OP (CEMnet.py:266-272): zero-stuff (value at offset `pre` of each cell), replicate-pad
    floor(k/2), correlate with ds_kernel * s^2."""
    n, c, h, w = f.shape
    stuffed = torch.zeros(n, c, h * s, w * s, dtype=f.dtype)
    stuffed[:, :, pre::s, pre::s] = f
    k = ds_kernel * s ** 2
    return _depthwise(stuffed, k, [k.shape[0] // 2, k.shape[1] // 2])


def cem_project(x_lr, g, ds_kernel, inv_hTh, s, pre):
    """CEM_PyTorch.forward, consistency part (CEMnet.py:303-310):
    ortho = Up(Inv(x));  NS = G - Up(Inv(Down(G)));  out = ortho + NS."""
    ortho = cem_up(cem_inv(x_lr, inv_hTh), ds_kernel, s, pre)
    ns = g - cem_up(cem_inv(cem_down(g, ds_kernel, s, pre), inv_hTh), ds_kernel, s, pre)
    return ortho + ns


def cem_wrapped_forward(x, sd, ds_kernel, inv_hTh, s, pre, margin_lr, eval_mode, nf, nb, z=0, upsample_mode='upconv'):
    """CEM_PyTorch.forward around an RRDBNet (CEMnet.py:283-311).  eval mode (`pre_pad`): replicate-pad the LR
    image by margin_lr and the HR view of the latent by s*margin_lr (:286-295), crop s*margin_lr at the end (:311)."""
    gp = 'generated_image_model.'
    if eval_mode:
        if z > 0:
            lat, img = torch.split(x, [x.size(1) - 3, 3], dim=1)
            z_hr = lat.reshape(lat.size(0), -1, s * lat.size(2), s * lat.size(3))
            img = F.pad(img, (margin_lr,) * 4, mode='replicate')
            z_hr = F.pad(z_hr, (s * margin_lr,) * 4, mode='replicate')
            lat = z_hr.reshape(z_hr.size(0), z_hr.size(1) * s * s, img.size(2), img.size(3))
            x = torch.cat([lat, img], 1)
        else:
            x = F.pad(x, (margin_lr,) * 4, mode='replicate')
    g = rrdbnet_forward(x, sd, nf, nb, upscale=s, z=z, upsample_mode=upsample_mode, prefix=gp)
    out = cem_project(x[:, -3:], g, ds_kernel, inv_hTh, s, pre)
    if eval_mode:
        m = s * margin_lr
        out = out[:, :, m:-m, m:-m]
    return out


def pack_latent(z_hr, x_lr, s):
    """SRRaGANModel.Prepare_Input (SRRaGAN_model.py:230-236): Z[B,c,sH,sW] is re-viewed (raw memory) as
    [B, c*s^2, H, W] and concatenated in front of the LR image."""
    b, c, hh, wh = z_hr.shape
    return torch.cat([z_hr.contiguous().view(b, c * s * s, hh // s, wh // s), x_lr], 1)


def z_optimize_l1(sd, ds_kernel, inv_hTh, s, pre, margin_lr, nf, nb, z, x_lr, desired, Z_range, lr, iters, batch=1):
    """Z_optimizer.optimize for objective 'l1' with all-ones masks (Z_optimization.py:647-797 with :292-300, :683-734):
    Adam on the pre-tanh latent map (init 0), Z = Z_range*tanh(.), eval-mode CEM(G([Z | LR])), output clamped to [0,1]
    (SRRaGAN_model.py:224-228), per-image L1 to `desired`, mean over the batch.  The iterate with the minimum loss is kept
    (:755-762).  Returns (loss values, final Z)."""
    zp = torch.zeros(batch, z, s * x_lr.size(2), s * x_lr.size(3), requires_grad=True)
    opt = torch.optim.Adam([zp], lr=lr)
    losses, iterates = [], []
    lr_b = x_lr.expand(batch, -1, -1, -1)
    for _ in range(iters):
        opt.zero_grad()
        iterates.append(zp.detach().clone())
        Z = Z_range * torch.tanh(zp)
        out = cem_wrapped_forward(pack_latent(Z, lr_b, s), sd, ds_kernel, inv_hTh, s, pre, margin_lr, True, nf, nb, z=z)
        out = torch.clamp(out, 0, 1)
        loss = torch.stack([F.l1_loss(out[i:i + 1], desired) for i in range(batch)], 0).mean()
        loss.backward()
        losses.append(loss.item())
        opt.step()
    best = int(np.argmin(losses))
    final = zp.detach() if losses[best] == losses[-1] else iterates[best]
    return losses if losses[best] == losses[-1] else losses[:best + 1], Z_range * torch.tanh(final)


# ---- discriminator and GAN losses (training step) ---------------------------------------------------------------------------------
def discriminator_vgg128_forward(x, sd, training=True, eps=1e-5, momentum=0.1, update_running=False):
    """Discriminator_VGG_128.forward (models/modules/architecture.py:446-508): conv0 3x3 + LeakyReLU (no norm), then nine
    conv_block(CNA) = Conv2d (4x4 stride 2 pad 1 / 3x3 stride 1 pad 1, bias) -> BatchNorm2d(affine) -> LeakyReLU(0.2)
    (block.py:129-146), flatten (:505), Linear -> LeakyReLU(0.2) -> Linear (:493).  Module indices in `features`: conv0 at 0,
    layer k >= 1 has its conv at 3k-1 and its norm at 3k.  training=True normalises with batch statistics (biased variance);
    update_running=True also applies nn.BatchNorm2d's running-statistics update to the tensors in `sd` (in place)."""
    y = F.leaky_relu(F.conv2d(x, sd['features.0.weight'], sd['features.0.bias'], stride=1, padding=1), LRELU_SLOPE)
    k = 1
    while 'features.%d.weight' % (3 * k - 1) in sd:
        w = sd['features.%d.weight' % (3 * k - 1)]
        y = F.conv2d(y, w, sd['features.%d.bias' % (3 * k - 1)], stride=2 if w.shape[-1] == 4 else 1, padding=1)
        p = 'features.%d.' % (3 * k)
        rm, rv = sd[p + 'running_mean'], sd[p + 'running_var']
        if training and not update_running:
            rm, rv = rm.clone(), rv.clone()
        y = F.leaky_relu(F.batch_norm(y, rm, rv, sd[p + 'weight'], sd[p + 'bias'], training, momentum, eps), LRELU_SLOPE)
        k += 1
    y = y.reshape(y.shape[0], -1)
    y = F.leaky_relu(F.linear(y, sd['classifier.0.weight'], sd['classifier.0.bias']), LRELU_SLOPE)
    return F.linear(y, sd['classifier.2.weight'], sd['classifier.2.bias'])


def gan_loss_vanilla(logits, target_is_real):
    """GANLoss('vanilla') = BCEWithLogitsLoss against constant 1 / 0 labels (models/modules/loss.py:212-246)"""
    return F.binary_cross_entropy_with_logits(logits, torch.full_like(logits, 1.0 if target_is_real else 0.0))


def relativistic_d_loss(pred_real, pred_fake):
    """discriminator step (models/SRRaGAN_model.py:353-354,360): (BCE(real - mean(fake), 1) + BCE(fake - mean(real), 0)) / 2"""
    return (gan_loss_vanilla(pred_real - pred_fake.mean(), True) + gan_loss_vanilla(pred_fake - pred_real.mean(), False)) / 2


def relativistic_g_loss(pred_real, pred_fake):
    """generator's GAN term (models/SRRaGAN_model.py:475-476), before the gan_weight factor; pred_real is detached there"""
    return (gan_loss_vanilla(pred_real - pred_fake.mean(), False) + gan_loss_vanilla(pred_fake - pred_real.mean(), True)) / 2


# ---- latent-control loss L_struct -------------------------------------------------------------------------------------------------
def structure_tensor_means(img):
    """per-image means of (Ix^2, Iy^2, Ix*Iy) with the 2x2 depth-wise filters [[-1,1],[0,0]] and [[-1,0],[1,0]] applied without
    padding (models/modules/loss.py:51-62, 140-147): Ix = x[i][j+1]-x[i][j], Iy = x[i+1][j]-x[i][j] on the (H-1)x(W-1) grid"""
    ix = (img[:, :, :, 1:] - img[:, :, :, :-1])[:, :, :-1, :]
    iy = (img[:, :, 1:, :] - img[:, :, :-1, :])[:, :, :, :-1]
    return torch.stack([(ix ** 2).mean(dim=(1, 2, 3)), (iy ** 2).mean(dim=(1, 2, 3)), (ix * iy).mean(dim=(1, 2, 3))], 1)


def filter_loss_structure_tensor(sr, hr, z, history, latent_channels='SVDinNormedOut_structure_tensor', noise_std=1 / 255):
    """FilterLoss.forward in model-training mode for the structure-tensor descriptors (loss.py:133-178,208): measured values
    normalised by the ground truth's structure tensor, targets = spatial mean of Z mapped onto the running 5-95 percentile
    range of everything measured so far (`history`: one list per channel, extended in place); returns |measured - target| [B,3]"""
    cur_z = z.mean(dim=(2, 3))
    d_sr, d_hr = structure_tensor_means(sr), structure_tensor_means(hr)
    if latent_channels == 'SVDinNormedOut_structure_tensor':
        norm = torch.sqrt(d_hr[:, 0]) * torch.sqrt(d_hr[:, 1])
        measured = [d_sr[:, i] / (norm + noise_std) for i in range(3)]
    else:
        measured = [d_sr[:, i] / (d_hr[:, i] + torch.sign(d_sr[:, i]) * noise_std) if i < 2 else d_sr[:, i] for i in range(3)]
    target = []
    for i in range(3):
        history[i] += [v.item() for v in measured[i]]
        ub, lb = np.percentile(history[i], 95), np.percentile(history[i], 5)
        target.append(cur_z[:, i] / 2 * (ub - lb) + np.mean([ub, lb]))
    return (torch.stack(measured, 1) - torch.stack(target, 1)).abs()


# ---- soft histogram / dictionary objective (latent-exploration tools) -------------------------------------------------------------
def soft_histogram(x, bins, vmax, temperature, dictionary, eps=1e-7):
    """SoftHistogramLoss.ComputeSoftHistogram's kernel (Z_optimization.py:196-210) with the [dims, samples, bins] tensor materialised in
    double precision like the reference: cyclic distance min(|d|, |d - max|, |d + max|), -(dist + eps)^2 / T, mean over dims, exp; then
    mean over the samples (histogram, -> [bins]) or -log of the mean over the bins (dictionary, -> [samples]).  x: [dims, P], bins: [dims, B]."""
    x, bins = x.double().unsqueeze(-1), bins.double().unsqueeze(1)
    d = (x - bins).abs()
    d = torch.min(d, (x - bins - vmax).abs())
    d = torch.min(d, (x - bins + vmax).abs())
    e = torch.exp((-((d + eps) ** 2) / temperature).mean(0))          # [P, B]
    return -torch.log(e.mean(1)) if dictionary else e.mean(0)


def soft_histogram_loss_gray(cur_images, desired, image_mask, temperature, dictionary, n_bins=256):
    """SoftHistogramLoss(bins=256, min=0, max=1, gray_scale=True, patch_size=1).forward (Z_optimization.py:218-230 with :36-101,180-216):
    grey levels of every image inside image_mask against the grey levels of the WHOLE desired image (its mask is only applied in the
    kernel-density branch, :90-91): KL divergence between the normalised soft histograms, or the mean dictionary distance per image."""
    bins = torch.linspace(0, 1, n_bins).view(1, -1)
    mask = image_mask.reshape(-1) > 0
    des = desired.mean(1, keepdim=True).reshape(1, -1)
    h_des = soft_histogram(des, bins, 1.0, temperature, False)
    normalizer = h_des.sum() / des.shape[1]
    h_des = (h_des / normalizer / des.shape[1]).float()
    out = []
    for im in cur_images:
        xs = im.mean(0, keepdim=True).reshape(1, -1)[:, mask]
        if dictionary:
            out.append(soft_histogram(xs, bins, 1.0, temperature, True).mean().float())
        else:
            h = soft_histogram(xs, bins, 1.0, temperature, False)
            h = (h / (h.sum() / xs.shape[1]) / xs.shape[1]).float()        # fixed-bin mode re-computes the normaliser on every call (:211-212)
            out.append(torch.log(h + torch.finfo(h.dtype).eps).view(1, -1))
    if dictionary:
        return torch.stack(out)
    return F.kl_div(torch.cat(out, 0), h_des.view(1, -1).expand(len(out), -1), reduction='mean')
