"""imresize golden fixture: the UNMODIFIED reference CEM/imresize_CEM.imresize on small random images - down / up by 2, 3, 4, colour
and grey, edge and zero padding, centre alignment (used by the data pipeline, data/LRHR_dataset.py:8, and the GUI, GUI.py:908,2291)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from CEM.imresize_CEM import imresize  # noqa: E402
from make_golden import save  # noqa: E402

CASES = [('down4', (24, 32, 3), dict(scale_factor=1 / 4)), ('up4', (6, 7, 3), dict(scale_factor=4)), ('down2', (20, 26), dict(scale_factor=[0.5])),
         ('up3', (5, 6, 3), dict(scale_factor=3)), ('down3', (18, 21, 3), dict(scale_factor=1 / 3)),
         ('down4_zero', (24, 32, 3), dict(scale_factor=1 / 4, use_zero_padding=True)), ('up2_center', (7, 9, 3), dict(scale_factor=2, align_center=True)),
         ('down2_center', (14, 18, 3), dict(scale_factor=0.5, align_center=True)), ('up4_shape', (6, 7, 3), dict(output_shape=[24, 28]))]


def main():
    rng = np.random.RandomState(12)
    arrays = {}
    for tag, shape, kw in CASES:
        im = rng.rand(*shape)
        arrays[tag + ':in'] = im
        arrays[tag + ':out'] = imresize(im, **kw)
        print(tag, shape, arrays[tag + ':out'].shape)
    save('imresize_cases', **arrays)


if __name__ == '__main__':
    main()
