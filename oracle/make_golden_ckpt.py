"""Checkpoint-compatibility golden fixture: the UNMODIFIED reference SRRaGANModel (test mode, CPU) loads (a) a bare state dict of
a plain RRDBNet and (b) a {'model_state_dict','optimizer_state_dict'} file of the same net into a CEM-wrapped generator WITH
latent inputs (base_model.py:132-190: key adjustment for the CEM wrapper, positional key matching, zero weights for the extra
latent input channels in front, CEM filters never loaded).  Stores the checkpoint tensors and the resulting state dict."""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import torch  # noqa: E402
from make_golden import save  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self      # base_model.py:172-173 moves the extended weights with .cuda(); CPU here


class ND(dict):
    def __missing__(self, k):
        return None


def make_opt(tmp, ckpt):
    return ND(model='srragan', scale=4, gpu_ids=None, is_train=False, range=[0, 1],
              path=ND(models=os.path.join(tmp, 'models'), pretrained_model_G=ckpt, log=tmp),
              network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input='all_layers', latent_input_domain='HR_downscaled', latent_channels=3,
                           norm_type=None, mode='CNA', nf=8, nb=1, in_nc=3, out_nc=3, gc=32, scale=4))


def main():
    import models.modules.architecture as arch
    from models.SRRaGAN_model import SRRaGANModel
    torch.manual_seed(77)
    plain = arch.RRDBNet(in_nc=3, out_nc=3, nf=8, nb=1, upscale=4, num_latent_channels=0)
    with torch.no_grad():
        for p in plain.parameters():
            p.copy_((torch.randn(p.shape) * 0.1).half().float())      # fp16-exact: stored in 16 bits
    sd = {k: v.detach().clone() for k, v in plain.state_dict().items()}
    arrays = {'ck:' + k: v.numpy().astype(np.float16) for k, v in sd.items()}
    results = {}
    with tempfile.TemporaryDirectory() as tmp:
        for tag, blob in (('bare', sd), ('wrapped', {'model_state_dict': sd, 'optimizer_state_dict': {}})):
            path = os.path.join(tmp, tag + '.pth')
            torch.save(blob, path)
            torch.manual_seed(5)
            with contextlib.redirect_stdout(io.StringIO()):
                model = SRRaGANModel(make_opt(tmp, path))
            out = model.netG.state_dict()
            results[tag] = {k: v.detach().clone() for k, v in out.items()}
            print(tag, len(out), list(out.keys())[:2], list(out.keys())[-3:])
    assert all(torch.equal(results['bare'][k], results['wrapped'][k]) for k in results['bare'])      # both file formats give the same model
    arrays.update({'out:' + k: (v.numpy() if 'Filter_OP' in k else v.numpy().astype(np.float16)) for k, v in results['bare'].items()})
    assert all(np.array_equal(arrays['out:' + k].astype(np.float32), v.numpy()) for k, v in results['bare'].items())
    save('checkpoint_into_latent_cem', **arrays)


if __name__ == '__main__':
    main()
