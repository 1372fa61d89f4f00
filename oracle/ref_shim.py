"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference (/root/reference/codes) in this
container with in-memory shims.  Never imported by the product path; never
available on the GPU box (/root/reference does not exist there).  Used by
oracle/make_golden.py to generate tests/golden/*.npz and to validate oracle/esr_oracle.py.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("ESR_REFERENCE_ROOT", "/root/reference/codes")


def available():
    return os.path.isdir(REF_ROOT)


def install():
    """Monkey-patch sys.modules so the reference imports; returns nothing."""
    import scipy.signal
    import scipy.signal.windows

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    scipy.signal.gaussian = scipy.signal.windows.gaussian
    for n in ['GPUtil', 'lmdb', 'deepdiff', 'skimage', 'skimage.transform', 'skimage.color',
              'skimage.io', 'matplotlib', 'matplotlib.pyplot', 'imagesize']:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sk = sys.modules['skimage']
    sk.io = sys.modules['skimage.io']
    sk.transform = sys.modules['skimage.transform']
    sk.color = sys.modules['skimage.color']
    sys.modules['skimage.transform'].resize = None
    sys.modules['skimage.color'].rgb2hsv = None
    sys.modules['skimage.color'].hsv2rgb = None
    sys.modules['deepdiff'].DeepDiff = None
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    import torch
    if not torch.cuda.is_available():
        torch.cuda.FloatTensor = torch.FloatTensor
        torch.cuda.DoubleTensor = torch.DoubleTensor
