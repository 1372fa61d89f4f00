"""models.SRRaGAN_model._save_png without OpenCV: a valid 8-bit PNG whose pixels are the BGR input in RGB order."""
import struct
import zlib

import numpy as np


def _decode(path):
    data = open(path, 'rb').read()
    assert data[:8] == b'\x89PNG\r\n\x1a\n'
    pos, chunks = 8, []
    while pos < len(data):
        n, tag = struct.unpack('>I', data[pos:pos + 4])[0], data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack('>I', data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xffffffff
        chunks.append((tag, body))
        pos += 12 + n
    assert [t for t, _ in chunks] == [b'IHDR', b'IDAT', b'IEND']
    w, h, depth, ctype = struct.unpack('>IIBB', chunks[0][1][:10])
    ch = 3 if ctype == 2 else 1
    raw = zlib.decompress(chunks[1][1])
    rows = np.frombuffer(raw, dtype=np.uint8).reshape(h, 1 + w * ch)
    assert depth == 8 and not rows[:, 0].any()
    return rows[:, 1:].reshape(h, w, ch)


def test_png_fallback_roundtrip(tmp_path, monkeypatch):
    import builtins
    real_import = builtins.__import__

    def no_cv2(name, *a, **k):
        if name == 'cv2':
            raise ImportError('no cv2')
        return real_import(name, *a, **k)
    monkeypatch.setattr(builtins, '__import__', no_cv2)
    from models.SRRaGAN_model import _save_png
    rng = np.random.RandomState(0)
    bgr = rng.randint(0, 256, size=(13, 21, 3)).astype(np.uint8)
    _save_png(bgr, str(tmp_path / 'val' / 'a.png'))
    assert np.array_equal(_decode(str(tmp_path / 'val' / 'a.png')), bgr[:, :, ::-1])
    grey = rng.randint(0, 256, size=(7, 9)).astype(np.uint8)
    _save_png(grey, str(tmp_path / 'val' / 'g.png'))
    assert np.array_equal(_decode(str(tmp_path / 'val' / 'g.png'))[:, :, 0], grey)
