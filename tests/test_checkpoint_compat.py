"""Published-checkpoint compatibility against the UNMODIFIED reference (oracle/make_golden_ckpt.py): a plain RRDBNet checkpoint -
bare state dict or {'model_state_dict','optimizer_state_dict'} - loaded into a CEM-wrapped generator with latent inputs gives
the same state dict as the reference's BaseModel.load_network (base_model.py:132-190): keys re-prefixed for the CEM wrapper,
matched positionally, zero weights for the latent input channels IN FRONT of every conv that sees them, the CEM's designed
filters untouched."""
import contextlib
import io

import numpy as np
import pytest
import torch

from util import golden


class ND(dict):
    def __missing__(self, k):
        return None


@pytest.mark.parametrize('fmt', ['bare', 'wrapped'])
def test_plain_checkpoint_into_latent_cem_model(tmp_path, fmt):
    if torch.cuda.is_available():
        pytest.skip('CPU-suite test')
    from models.SRRaGAN_model import SRRaGANModel
    g = golden('checkpoint_into_latent_cem')
    sd = {k[3:]: torch.from_numpy(g[k].astype(np.float32)) for k in g.files if k.startswith('ck:')}
    path = str(tmp_path / 'plain.pth')
    torch.save(sd if fmt == 'bare' else {'model_state_dict': sd, 'optimizer_state_dict': {}}, path)
    opt = ND(model='srragan', scale=4, gpu_ids=None, is_train=False, range=[0, 1],
             path=ND(models=str(tmp_path / 'models'), pretrained_model_G=path, log=str(tmp_path)),
             network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input='all_layers', latent_input_domain='HR_downscaled', latent_channels=3,
                          norm_type=None, mode='CNA', nf=8, nb=1, in_nc=3, out_nc=3, gc=32, scale=4))
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        model = SRRaGANModel(opt)
    own = model.netG.state_dict()
    ref_keys = [k[4:] for k in g.files if k.startswith('out:')]
    assert list(own.keys()) == ref_keys
    for k in ref_keys:
        ref = g['out:' + k].astype(np.float32)
        if 'Filter_OP' in k:
            assert np.allclose(own[k].numpy(), ref, rtol=0, atol=1e-7), k       # designed, never loaded (bit-identity: test_cem_design)
        else:
            assert np.array_equal(own[k].numpy(), ref), k
    # the latent channels are the first three input channels of the first conv and carry zero weights
    w0 = own['generated_image_model.model.0.weight']
    assert w0.shape[1] == 6 and float(w0[:, :3].abs().max()) == 0.0 and float(w0[:, 3:].abs().max()) > 0
