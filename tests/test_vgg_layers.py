"""Host logic of the VGG feature extractor on CPU: the layer table and the parameter container reproduce
torchvision.models.vgg19().features[:35] (layer kinds, indices, shapes, state-dict keys) — checked against torchvision
itself when it is importable, and against the published VGG19 configuration otherwise."""
import os
import sys

import pytest
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'explorable-super-resolution_b200'))


def test_layer_table_is_vgg19_conv5_4_before_relu():
    from esr_b200.vgg import vgg19_layers
    layers = vgg19_layers(34)
    assert len(layers) == 35 and layers[-1] == ('conv', 34, 512, 512)          # conv5_4, its ReLU (35) is cut off
    assert [i for k, i, _, _ in layers if k == 'pool'] == [4, 9, 18, 27]
    assert [(ci, co) for k, _, ci, co in layers if k == 'conv'] == [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256),
                                                                    (256, 256), (256, 512), (512, 512), (512, 512), (512, 512), (512, 512), (512, 512),
                                                                    (512, 512), (512, 512)]


def test_container_matches_torchvision_when_available():
    tv = pytest.importorskip('torchvision')
    import models.modules.architecture as arch
    ref = tv.models.vgg19(weights=None).features[:35]
    mine = arch.VGGFeatureExtractor(feature_layer=34, arch_config='untrained').features
    assert len(ref) == len(mine)
    for a, b in zip(ref, mine):
        assert type(a) is type(b)
        if isinstance(a, nn.Conv2d):
            assert a.weight.shape == b.weight.shape and a.kernel_size == b.kernel_size and a.padding == b.padding
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    # a torchvision-layout checkpoint (optionally saved from nn.DataParallel: 'module.' prefix) loads
    sd = {'module.features.' + k: v for k, v in ref.state_dict().items()}
    loaded = arch.VGGFeatureExtractor(feature_layer=34, state_dict=sd)
    assert all((loaded.features.state_dict()[k] == v).all() for k, v in ref.state_dict().items())
