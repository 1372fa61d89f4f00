"""TEST INFRASTRUCTURE - seeded constructions shared by the golden-fixture generators (which run the unmodified reference in the build
container) and the tests that re-create the same networks on the GPU box.  No reference imports here."""
import torch


def kink_free_biases(net, seed):
    """LeakyReLU's derivative jumps at 0: wherever a pre-activation is smaller than an implementation's forward rounding error the
    two may legitimately pick different slopes, and every such flip moves the gradient by a full component - the gradient error of
    ANY two non-bit-identical implementations is ~sqrt(fraction of flipped units), not their arithmetic error.  Per-channel biases
    of magnitude 2..3 with random signs keep every LeakyReLU input away from 0 (both slopes stay exercised, channel by channel),
    so gradients are comparable at arithmetic precision (same construction as make_golden.py E3).  Seeded: the test re-creates it."""
    import re
    g = torch.Generator().manual_seed(seed + 300)
    n_up = len([k for k in net.state_dict() if re.match(r'model\.\d+\.1\.weight$', k)])
    with torch.no_grad():
        for name, p in net.named_parameters():
            if not name.endswith('bias'):
                continue
            # convs followed by a LeakyReLU: the four growth convs of every dense block, the up-convs, HR_conv0
            activated = bool(re.search(r'convs\.[0-3]\.0\.bias$', name) or re.match(r'model\.\d+\.1\.bias$', name) or name == 'model.%d.bias' % (2 + n_up))
            if activated:
                sign = torch.where(torch.rand(p.shape, generator=g) < 0.5, -1.0, 1.0)
                p.copy_(sign * (2.0 + torch.rand(p.shape, generator=g)))
            else:       # linear outputs (fea_conv, conv5 of the dense blocks, LR_conv, HR_conv1): small biases keep the trunk bounded
                p.copy_(0.05 * torch.randn(p.shape, generator=g))


