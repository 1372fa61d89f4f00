"""Arithmetic mode of the engines.

  'throughput' (default): fp16 operands for the frozen generator / VGG19 (10-bit mantissa, the class of the TF32 cuDNN path the
                reference runs by default on a GPU), bf16 operands for training and the critic (range); fp32 accumulation.
  'parity'    : split precision (esr_dtype ESR_BF16X3): every 16-bit tensor carries bf16 hi + lo halves and every convolution
                runs x_hi*w_hi + x_lo*w_hi + x_hi*w_lo on the same tcgen05 kernels - 16 mantissa bits per operand with fp32's
                exponent range, i.e. the arithmetic class of the reference's fp32 convs (models/modules/block.py:141-146), at
                three times the MMA work.  This is the mode the 1e-3 parity tests of training, critic and VGG19 run in.

Select with ESR_PRECISION=parity in the environment or `esr_b200.precision.set_precision('parity')` before the first forward."""
import os

_MODES = ('throughput', 'parity')
_mode = os.environ.get('ESR_PRECISION', 'throughput')
if _mode not in _MODES:
    raise ValueError('ESR_PRECISION must be one of %r' % (_MODES,))


def set_precision(mode):
    global _mode
    if mode not in _MODES:
        raise ValueError('precision must be one of %r' % (_MODES,))
    _mode = mode


def get_precision():
    return _mode


def parity():
    return _mode == 'parity'


class use:
    """context manager: `with precision.use('parity'): ...`"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = get_precision()
        set_precision(self.mode)

    def __exit__(self, *a):
        set_precision(self.prev)
