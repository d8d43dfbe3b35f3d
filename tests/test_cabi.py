"""CPU-side checks of the C-ABI boundary: the shared library loads and exports every symbol that
include/dasr_b200.h declares (no compute calls — there is no GPU in this suite)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'dasr_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(dasr_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_header_symbol():
    from dasr_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), 'library does not export %s' % s
    # and the ctypes table binds exactly the header's functions
    assert sorted(_lib.SYMBOLS.keys()) == syms


def test_param_structs_match_header_layout():
    """sizeof of the ctypes mirrors == what the C compiler lays out (compiled check via the library's own
    behaviour is GPU-only; here: field order/count against the header text)."""
    from dasr_b200 import _lib
    txt = open(os.path.join(ROOT, 'include', 'dasr_b200.h')).read()
    for name, struct in (('DasrConvF32Params', _lib.ConvF32Params), ('DasrConvTcParams', _lib.ConvTcParams), ('DasrPackJob', _lib.PackJob)):
        end = txt.index('} %s;' % name)
        body = txt[txt.rindex('typedef struct {', 0, end) + len('typedef struct {'):end]
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        fields = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            for f in decl.split(','):
                fields.append(re.split(r'[\s\*]+', re.sub(r'\[.*', '', f.strip()))[-1])
        assert fields == [f[0] for f in struct._fields_], name


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from dasr_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.DasrError):
        _lib.load()
