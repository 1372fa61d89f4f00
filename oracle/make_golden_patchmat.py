"""Patch-extraction-matrix golden fixture: the UNMODIFIED reference ReturnPatchExtractionMat (Z_optimization.py:232-271; imported by
GUI.py:14) on three masks / overlap settings.  numpy's removed `np.bool` alias is restored in memory for the run.  Stores the
coalesced indices of the sparse matrices."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import torch  # noqa: E402
import Z_optimization as Zmod  # noqa: E402
from make_golden import save  # noqa: E402


def masks():
    m = np.zeros([20, 24])
    m[3:15, 4:20] = 1
    m[7:9, 9:12] = 0
    m[16:19, 1:3] = 1           # too small for a 3x3 patch... removed by the opening
    return m


def main():
    np.bool = np.bool_          # the alias the reference's code uses (removed from numpy 1.24 on)
    m = masks()
    arrays = {'mask': m}
    cases = [('full', dict(patch_size=3, patches_overlap=1)), ('half', dict(patch_size=3, patches_overlap=0.5, return_non_covered=True)),
             ('none', dict(patch_size=4, patches_overlap=0, return_non_covered=True))]
    for tag, kw in cases:
        out = Zmod.ReturnPatchExtractionMat(m, device=torch.device('cpu'), **kw)
        mats = out if isinstance(out, tuple) else (out, None)
        for name, mat in zip(('mat', 'rest'), mats):
            if mat is None:
                continue
            mat = mat.coalesce()
            arrays['%s:%s_idx' % (tag, name)] = mat.indices().numpy()
            arrays['%s:%s_shape' % (tag, name)] = np.array(mat.shape)
        print(tag, [None if t is None else tuple(t.shape) for t in mats])
    save('patch_extraction_mat', **arrays)


if __name__ == '__main__':
    main()
