"""Drop-in boundary: module tree / state-dict keys and shapes identical to the reference's (golden key lists
come from the reference's own RRDBNet.state_dict())."""
import pytest
import torch

from util import golden, golden_state_dict, mirror_rrdb


@pytest.mark.parametrize('name,extra', [('rrdb_plain_x4', {}), ('rrdb_latent_x4', {}), ('rrdb_plain_x2', {}), ('rrdb_plain_x8', {}),
                                        ('rrdb_pixelshuffle_x4', {'upsample_mode': 'pixelshuffle'})])
def test_keys_and_shapes(name, extra):
    g = golden(name)
    ref = golden_state_dict(g)
    net = mirror_rrdb(g, **extra)
    mine = net.state_dict()
    assert list(mine.keys()) == list(ref.keys())
    assert [tuple(v.shape) for v in mine.values()] == [tuple(v.shape) for v in ref.values()]


def test_cem_wrapper_keys_and_adjust():
    from CEM.CEMnet import CEMnet, Get_CEM_Conf, Adjust_State_Dict_Keys
    g = golden('rrdb_plain_x4')
    net = mirror_rrdb(g)
    cem = CEMnet(Get_CEM_Conf(4))
    wrapped = cem.WrapArchitecture_PyTorch(net, None)
    keys = list(wrapped.state_dict().keys())
    assert keys[0] == 'generated_image_model.model.0.weight'
    assert keys[-3:] == ['Conv_LR_with_Inv_hTh_OP.Filter_OP.weight', 'Upscale_OP.Filter_OP.weight', 'DownscaleOP.Filter_OP.weight']
    assert tuple(wrapped.state_dict()['DownscaleOP.Filter_OP.weight'].shape) == (3, 1, 17, 17)
    assert tuple(wrapped.state_dict()['Conv_LR_with_Inv_hTh_OP.Filter_OP.weight'].shape) == (3, 1, 27, 27)
    assert cem.OP_names == ['Conv_LR_with_Inv_hTh_OP.Filter_OP', 'Upscale_OP.Filter_OP', 'DownscaleOP.Filter_OP']
    adjusted = Adjust_State_Dict_Keys(golden_state_dict(g), wrapped.state_dict())
    assert list(adjusted.keys()) == keys
    wrapped.load_state_dict(adjusted, strict=True)
    # eval()/train() toggles the padding flag like the reference (CEMnet.py:313-315)
    wrapped.eval()
    assert wrapped.pre_pad is True
    wrapped.train()
    assert wrapped.pre_pad is False
    assert all(not p.requires_grad for n, p in wrapped.named_parameters() if 'Filter_OP' in n)


def test_kaiming_init_skips_cem_filters():
    import models.networks as networks
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    g = golden('rrdb_plain_x4')
    wrapped = CEMnet(Get_CEM_Conf(4)).WrapArchitecture_PyTorch(mirror_rrdb(g), None)
    before = wrapped.DownscaleOP.Filter_OP.weight.clone()
    networks.init_weights(wrapped, 'kaiming', scale=0.1)
    assert torch.equal(before, wrapped.DownscaleOP.Filter_OP.weight)
    assert float(wrapped.generated_image_model.model[0].bias.abs().max()) == 0.0


def test_no_eager_fallback():
    import models.modules.block as B
    blk = B.RRDB(32)
    with pytest.raises(NotImplementedError):
        blk(torch.zeros(1, 32, 4, 4))
