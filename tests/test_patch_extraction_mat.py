"""ReturnPatchExtractionMat (imported by the reference's GUI.py:14) against the unmodified reference's output
(oracle/make_golden_patchmat.py): same sparse patch-gathering matrices for full overlap, half overlap with the uncovered-pixel
matrix, and no overlap with 4x4 patches; and its use as in GUI.py:2408-2412 (per-patch STD through one sparse mm)."""
import numpy as np
import pytest
import torch

from util import golden

CASES = [('full', dict(patch_size=3, patches_overlap=1)), ('half', dict(patch_size=3, patches_overlap=0.5, return_non_covered=True)),
         ('none', dict(patch_size=4, patches_overlap=0, return_non_covered=True))]


@pytest.mark.parametrize('tag,kw', CASES)
def test_patch_extraction_matrices_match_reference(tag, kw):
    from Z_optimization import ReturnPatchExtractionMat
    g = golden('patch_extraction_mat')
    out = ReturnPatchExtractionMat(g['mask'], device=torch.device('cpu'), **kw)
    mats = out if isinstance(out, tuple) else (out, None)
    for name, mat in zip(('mat', 'rest'), mats):
        key = '%s:%s_idx' % (tag, name)
        if mat is None:
            assert key not in g.files
            continue
        mat = mat.coalesce()
        assert list(mat.shape) == list(g['%s:%s_shape' % (tag, name)])
        assert np.array_equal(mat.indices().numpy(), g[key])
        assert float(mat.values().min()) == 1.0 == float(mat.values().max())


def test_patch_std_map_as_the_gui_computes_it():
    from Z_optimization import ReturnPatchExtractionMat
    g = golden('patch_extraction_mat')
    mask = np.ones([10, 12])
    mat = ReturnPatchExtractionMat(mask, 3, device=torch.device('cpu'), patches_overlap=1)
    z = torch.rand(1, 1, 10, 12, generator=torch.Generator().manual_seed(3))
    std_map = torch.sparse.mm(mat, z.mean(dim=1).view([-1, 1])).view([9, -1]).std(dim=0).view([8, 10])
    ref = z[0, 0].unfold(0, 3, 1).unfold(1, 3, 1).reshape(8, 10, 9).std(dim=-1)
    assert torch.allclose(std_map, ref, atol=1e-6)
    assert g['mask'].shape == (20, 24)
