"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/esr_b200.h declares.
No compute call is made here (no GPU in the build container)."""
import os
import re
import subprocess

from conftest import REPO


def _declared():
    txt = open(os.path.join(REPO, 'include', 'esr_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(esr_[a-z0-9_]+)\s*\(', txt)))


def test_header_symbols_exported_and_bound():
    from esr_b200 import lib
    L = lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), 'libesr_b200.so does not export %s' % n
        assert n in lib.SIGNATURES, 'ctypes binding is missing %s' % n
    assert set(lib.SIGNATURES) <= set(names)
    assert L.esr_version() >= 100
    assert L.esr_launch_count() == 0 or L.esr_launch_count() > 0


def test_sass_is_blackwell_native():
    from esr_b200 import lib
    lib.load()
    if not os.path.exists('/usr/local/cuda/bin/cuobjdump'):
        return
    sass = subprocess.run(['/usr/local/cuda/bin/cuobjdump', '-sass', lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'UBLKCP', 'LDTM'):
        assert mnemonic in sass, mnemonic


def test_conv_args_struct_matches_header_order():
    from esr_b200 import lib
    txt = open(os.path.join(REPO, 'include', 'esr_b200.h')).read()
    body = txt[txt.index('typedef struct {'):txt.index('} esr_conv3x3_args;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in body.split(';'):
        decl = decl.replace('typedef struct {', '').strip()
        if not decl:
            continue
        for part in decl.split(','):
            fields.append(re.findall(r'([A-Za-z_][A-Za-z0-9_]*)\s*$', part.strip())[0])
    bound = [f[0].rstrip('_') for f in lib.ConvArgs._fields_]
    assert fields == bound


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch
    from esr_b200 import ops, lib
    if torch.cuda.is_available():
        return
    with pytest.raises(lib.EsrError):
        ops.pack_nchw(torch.zeros(1, 3, 4, 4))


def test_integration_stub_matches_binding():
    """the ctypes stub shown to maintainers in INTEGRATION.md lists the same fields, in the same order, as the shipped binding"""
    from esr_b200 import lib
    txt = open(os.path.join(REPO, 'INTEGRATION.md')).read()
    body = txt[txt.index('class ConvArgs(C.Structure):'):txt.index('lib.esr_conv3x3_fwd.argtypes')]
    shown = re.findall(r'\("([a-z_0-9]+)", C\.c_[a-z_]+\)', body)
    assert shown == [f[0] for f in lib.ConvArgs._fields_]


def test_pack_item_struct_matches_header_order():
    from esr_b200 import lib
    txt = open(os.path.join(REPO, 'include', 'esr_b200.h')).read()
    end = txt.index('} esr_pack_item;')
    body = re.sub(r'/\*.*?\*/', '', txt[txt.rindex('typedef struct {', 0, end):end], flags=re.S)
    fields = []
    for decl in body.split(';'):
        decl = decl.replace('typedef struct {', '').strip()
        if not decl:
            continue
        for part in decl.split(','):
            fields.append(re.findall(r'([A-Za-z_][A-Za-z0-9_]*)\s*$', part.strip())[0])
    assert fields == ['w_oihw' if f[0] == 'w' else f[0] for f in lib.PackItem._fields_]
