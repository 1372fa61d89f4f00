"""`create_model(opt)` — same dispatch as the reference's models/__init__.py:1-13."""


def create_model(opt, *args, **kwargs):
    model = opt['model']
    if model == 'srragan':
        from .SRRaGAN_model import SRRaGANModel as M
    else:
        raise NotImplementedError('Model [{:s}] not recognized (esr_b200 builds the srragan hot path only).'.format(model))
    m = M(opt, *args, **kwargs)
    print('Model [{:s}] is created.'.format(m.__class__.__name__))
    return m
