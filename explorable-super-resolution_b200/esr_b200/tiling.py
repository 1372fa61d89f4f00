"""Tiled inference of the CEM-wrapped generator for images that do not fit one pass.

The reference super-resolves a whole image in one `netG(model_input)` call (models/SRRaGAN_model.py:530-538, test.py); its memory grows
with the image (every dense block keeps a 192-channel buffer) and the GUI therefore limits the image size.  Here the LR image is cut
into tiles that are extended by `overlap` LR pixels on every interior side; each extended tile goes through the ordinary eval-mode
forward (which replicate-pads by the CEM's invalidity margin and crops it again, CEM/CEMnet.py:286-311) and only its core is kept.
A pixel of the stitched result therefore equals the one-pass result wherever the tile's overlap covers the generator's receptive field
(1 + 15 pixels per RRDB + the CEM filters' support): exactly for shallow generators, to within the decay of the receptive field for
the 23-block one.  Image borders are handled by the same replicate padding as in the one-pass forward, because border tiles are not
extended beyond the image."""
import torch


def _tile_ranges(n, tile, overlap):
    """[(core_lo, core_hi, ext_lo, ext_hi)] covering 0..n with cores of at most `tile` pixels"""
    out = []
    lo = 0
    while lo < n:
        hi = min(n, lo + tile)
        out.append((lo, hi, max(0, lo - overlap), min(n, hi + overlap)))
        lo = hi
    return out


@torch.no_grad()
def tiled_forward(netG, lr, z_hr=None, tile=128, overlap=32, scale=None):
    """netG: the generator as models.networks.define_G returns it (CEM_PyTorch around RRDBNet, or a bare RRDBNet), in eval mode.
    lr: [N, 3, h, w] LR images; z_hr: [N, z, S*h, S*w] latent maps or None.  Returns the [N, out_nc, S*h, S*w] output, stitched from
    tiles of at most `tile` x `tile` LR pixels extended by `overlap` LR pixels towards their neighbours."""
    mod = netG.module if hasattr(netG, 'module') else netG
    if scale is None:
        scale = int(getattr(mod, 'ds_factor', 0) or getattr(mod, 'upscale', 0) or mod.generated_image_model.upscale)
    S = int(scale)
    n, _, h, w = lr.shape
    if z_hr is not None:
        assert tuple(z_hr.shape[2:]) == (S * h, S * w), 'latent map must cover the HR image'
    out = None
    for (y0, y1, ya, yb) in _tile_ranges(h, tile, overlap):
        for (x0, x1, xa, xb) in _tile_ranges(w, tile, overlap):
            lr_t = lr[:, :, ya:yb, xa:xb]
            if z_hr is not None:
                # the latent is fed as a RAW re-view of the HR map at LR resolution (SRRaGAN_model.Prepare_Input, architecture.py:281-283):
                # cut it in the HR domain, then re-view the tile's own map
                z_t = z_hr[:, :, S * ya:S * yb, S * xa:S * xb].contiguous()
                z_t = z_t.view(n, z_t.size(1) * S * S, yb - ya, xb - xa)
                x_t = torch.cat([z_t, lr_t], dim=1)
            else:
                x_t = lr_t.contiguous()
            y_t = netG(x_t)
            if out is None:
                out = torch.empty((n, y_t.size(1), S * h, S * w), dtype=y_t.dtype, device=y_t.device)
            out[:, :, S * y0:S * y1, S * x0:S * x1] = y_t[:, :, S * (y0 - ya):S * (y1 - ya), S * (x0 - xa):S * (x1 - xa)]
    return out
