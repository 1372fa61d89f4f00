"""Gradient path of the generator (dgrad for Z-optimisation, dgrad+wgrad for training).  Not built yet in
this round: fail loudly rather than fall back to an eager PyTorch graph."""


def rrdb_forward_with_grad(net, x, pad):
    raise NotImplementedError(
        'esr_b200: backward through RRDBNet (dgrad/wgrad tcgen05 kernels) is not built yet; '
        'call the generator under torch.no_grad(). There is deliberately no PyTorch/cuDNN fallback.')
